"""make_syn with the reference's command line (src/make_syn.f90:28-166):  python -m rf_inv_b200.make_syn [params.in]

Draws one random model per chain exactly as the reference does (the stream ``sgrnd(iseed)``, init_model, init_sig),
computes the synthetic receiver functions of chain 1 on the GPU, adds Gaussian noise shaped by each trace's filter,
and writes what the reference writes into the current directory:

  * ``test_vel``           alpha, beta, rho, h per layer of chain 1 (list-directed, make_syn.f90:66-75)
  * ``test_traceNN``       noise-free trace NN as a SAC file, first nsmp samples (make_syn.f90:140-158)
  * ``test_traceNNwn``     the same with noise (make_syn.f90:113-138)

(The reference builds the file names with an ``A10`` edit descriptor from the 11-character literal "test_trace.", which
drops the dot.)  Like the reference, the observed-data files named in params.in must exist: nsmp and delta come from them.
Everything numerical runs through the C-ABI: rfinv_pt_init / rfinv_pt_get_state (model), rfinv_eval_batch (RF),
rfinv_pt_draw (the noise deviates, same stream position as the reference), rfinv_filter_traces (r2c, filter, c2r).
"""
from __future__ import annotations

import argparse
import copy
import ctypes as C
import os
from typing import Dict

import numpy as np

from . import capi
from . import io as rio
from .config import RFConfig
from .pt import ParallelTempering


def _ld_real(x: float) -> str:
    """gfortran list-directed real(8): 26 columns (src/make_syn.f90:73 writes four of them per line)."""
    ax = abs(x)
    if x == 0.0:
        body, fixed = "0.0000000000000000", True
    elif 0.1 <= ax < 1e17:
        e = int(np.floor(np.log10(ax))) + 1
        body, fixed = f"{x:.{max(17 - e if e >= 1 else 17, 0)}f}", True
    else:
        m, ex = f"{x:.16E}".split("E")
        body, fixed = f"{m}E{ex[0]}{abs(int(ex)):03d}", False
    width = 21 if fixed else 26
    out = body.rjust(width) if len(body) < width else " " + body
    return out + ("     " if fixed else "")


def synthesize(cfg: RFConfig, device: int = 0) -> Dict[str, np.ndarray]:
    """The numbers make_syn produces (no files): model of chain 1, its layer stack, noise levels, filtered noise,
    noise-free and noisy traces (first nsmp samples)."""
    c = copy.copy(cfg)
    c.ncool = c.nchains                      # make_syn never calls init_pt_mcmc: no temperature draws on the stream
    if c.r_inv is None:
        c.r_inv = np.zeros((c.ntrc, c.nsmp, c.nsmp))   # the likelihood value is not used by make_syn
    pt = ParallelTempering(c, 1, device=device)        # virtual rank 0: sgrnd(iseed), init_model, init_sig, init_rft
    try:
        lib, h = pt._lib, pt.ev.handle
        st = pt.state()
        k, z, dvp, dvs, sig = st["k"][:1], st["z"][:1], st["dvp"][:1], st["dvs"][:1], st["sig"][:1]
        _, rft, _ = pt.ev.calc_likelihood(k, z, dvp, dvs, sig, want_rft=True)
        nlay, alpha, beta, rho, hth, _ = pt.ev.format_model(k, z, dvp, dvs)
        n, T, S = c.nfft, c.ntrc, c.nsmp

        def draw(kind: int, count: int) -> np.ndarray:
            out = np.empty(count)
            capi.check(lib.rfinv_pt_draw(h, 0, kind, count, out.ctypes.data_as(capi.dp)))
            return out

        white = np.zeros((T, n)); noise_sigma = np.zeros(T)
        if c.is_ray_common:                  # one noise series for all traces (make_syn.f90:81-93)
            noise_sigma[:] = draw(0, 1)[0] * (c.sig_max[0] - c.sig_min[0]) + c.sig_min[0]
            white[:] = draw(1, n) * noise_sigma[0]
        else:                                # make_syn.f90:95-110
            for t in range(T):
                noise_sigma[t] = draw(0, 1)[0] * (c.sig_max[t] - c.sig_min[t]) + c.sig_min[t]
                white[t] = draw(1, n) * noise_sigma[t]
        noise = np.empty((T, n))
        trace_of = np.arange(T, dtype=np.int32)
        capi.check(lib.rfinv_filter_traces(h, T, trace_of.ctypes.data_as(capi.i32p), white.ctypes.data_as(capi.dp),
                                           noise.ctypes.data_as(capi.dp)))
        if c.is_ray_common and T > 1:
            # make_syn.f90:86-92 as written: every pass of the loop loads rx from noise(:,1), which the first pass has
            # already replaced by its filtered version -- traces 2.. are shaped by flt(:,1) and then by their own filter
            again = np.ascontiguousarray(np.repeat(noise[:1], T - 1, axis=0))
            out2 = np.empty((T - 1, n))
            capi.check(lib.rfinv_filter_traces(h, T - 1, trace_of[1:].copy().ctypes.data_as(capi.i32p), again.ctypes.data_as(capi.dp),
                                               out2.ctypes.data_as(capi.dp)))
            noise[1:] = out2
    finally:
        pt.close()
    L = int(nlay[0])
    clean = rft[0, :, :S].copy()
    return dict(k=int(k[0]), z=z[0].copy(), dvp=dvp[0].copy(), dvs=dvs[0].copy(), sig=sig[0].copy(), nlay=L,
                alpha=alpha[0, :L].copy(), beta=beta[0, :L].copy(), rho=rho[0, :L].copy(), h=hth[0, :L].copy(),
                noise_sigma=noise_sigma, noise=noise, rft=clean, noisy=clean + noise[:, :S])


def write_sac_make_syn(path: str, data: np.ndarray, delta: float, t_start: float, t_end: float) -> None:
    """SAC file with exactly the header words make_syn sets (src/make_syn.f90:121-137); native endianness."""
    data = np.asarray(data, dtype=np.float32)
    raw = np.zeros(158 + data.size, dtype=np.float32)
    ints = raw.view(np.int32)
    raw[0] = np.float32(delta)       # record 1   delta
    raw[5] = np.float32(t_start)     # record 6   b
    raw[6] = np.float32(t_end)       # record 7   e
    ints[76] = 6                     # record 77  nvhdr
    ints[85] = 1                     # record 86  iftype
    ints[79] = data.size             # record 80  npts
    ints[105] = 1                    # record 106 leven
    raw[158:] = data
    raw.tofile(path)


def run(params_path: str, out_dir: str = ".", device: int = 0, verbose: bool = True) -> Dict[str, np.ndarray]:
    cfg = rio.load_problem(params_path)
    res = synthesize(cfg, device=device)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "test_vel"), "w") as f:
        for i in range(res["nlay"]):
            f.write("".join(_ld_real(float(v)) for v in (res["alpha"][i], res["beta"][i], res["rho"][i], res["h"][i])) + "\n")
    t_end = getattr(cfg, "t_end", cfg.t_start + (cfg.nsmp - 1) * cfg.delta)
    for t in range(cfg.ntrc):
        if verbose:
            if cfg.is_ray_common:
                if t == 0:
                    print(f" Noise level of all traces:  {res['noise_sigma'][0]}")
            else:
                print(f" Noise level of trace {t + 1:12d} : {res['noise_sigma'][t]}")
        write_sac_make_syn(os.path.join(out_dir, f"test_trace{t + 1:02d}wn"), res["noisy"][t], cfg.delta, cfg.t_start, t_end)
        write_sac_make_syn(os.path.join(out_dir, f"test_trace{t + 1:02d}"), res["rft"][t], cfg.delta, cfg.t_start, t_end)
    return res


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("params", nargs="?", default="params.in")
    ap.add_argument("--out-dir", default=".", help="where test_vel / test_traceNN[wn] go (the reference: current directory)")
    a = ap.parse_args(argv)
    run(a.params, a.out_dir)


if __name__ == "__main__":
    main()
