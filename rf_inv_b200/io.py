"""File formats either side of the hot path, through the C-ABI host services (rf_inv_b200/csrc/host_io.cu):
``params.in`` + SAC traces + reference velocity model in (src/params.f90, src/model.f90:109-171), the ``mcmc_out``
file set out (src/mcmc_out.f90:97-318).  ``write_sac`` follows make_syn's writer (src/make_syn.f90:127-137) and exists
for tests and for making synthetic inputs."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import capi
from .config import RFConfig
from .evaluator import _p


def _arr(ptr, n, dtype=np.float64):
    if not ptr or n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype).copy()


def load_problem(params_path: str, base_dir: Optional[str] = None) -> RFConfig:
    """get_params + read_obs + read_ref_model.  Relative paths inside params.in are resolved against ``base_dir``
    (default: the directory of params.in; the reference resolves them against the current directory)."""
    lib = capi.load()
    base = base_dir if base_dir is not None else os.path.dirname(os.path.abspath(params_path))
    h = C.c_void_p()
    capi.check(lib.rfinv_problem_load(params_path.encode(), base.encode(), C.byref(h)))
    try:
        c = lib.rfinv_problem_config(h).contents
        T, S = c.ntrc, c.nsmp
        cfg = RFConfig(
            ntrc=T, nfft=c.nfft, nsmp=S, delta=c.delta, t_start=c.t_start, rayps=list(_arr(c.rayps, T)),
            a_gus=list(_arr(c.a_gus, T)), ipha=[int(x) for x in _arr(c.ipha, T, np.int32)], deconv_mode=c.deconv_mode,
            sdep=c.sdep, bdep=c.bdep, obs=_arr(c.obs, T * S).reshape(T, S), r_inv=None, vp_ref=_arr(c.vp_ref, c.nref),
            vs_ref=_arr(c.vs_ref, c.nref), z_ref_min=c.z_ref_min, dz_ref=c.dz_ref, vp_mode=c.vp_mode, k_min=c.k_min,
            k_max=c.k_max, prior_mode=c.prior_mode, z_min=c.z_min, z_max=c.z_max, h_min=c.h_min, dvs_prior=c.dvs_prior,
            dvp_prior=c.dvp_prior, sig_min=list(_arr(c.sig_min, T)), sig_max=list(_arr(c.sig_max, T)), vp_min=c.vp_min,
            vp_max=c.vp_max, vs_min=c.vs_min, vs_max=c.vs_max, vpvs_min=c.vpvs_min, vpvs_max=c.vpvs_max, dev_z=c.dev_z,
            dev_dvs=c.dev_dvs, dev_dvp=c.dev_dvp, dev_sig=c.dev_sig, nburn=c.nburn, niter=c.niter, ncorr=c.ncorr,
            nchains=c.nchains, ncool=c.ncool, iseed=c.iseed, t_high=c.t_high, nbin_z=c.nbin_z, nbin_vs=c.nbin_vs,
            nbin_vp=c.nbin_vp, nbin_vpvs=c.nbin_vpvs, nbin_sig=c.nbin_sig, nbin_amp=c.nbin_amp, amp_min=c.amp_min,
            amp_max=c.amp_max)
        out_dir = lib.rfinv_problem_out_dir(h).decode()
        cfg.out_dir = out_dir if os.path.isabs(out_dir) else os.path.normpath(os.path.join(base, out_dir))
        cfg.t_end = float(lib.rfinv_problem_t_end(h))
    finally:
        lib.rfinv_problem_free(h)
    return cfg


def write_side_copies(params_path: str, out_dir: str, input_dir: str, base_dir: Optional[str] = None) -> None:
    """<out_dir>/params.in.copy and <input_dir>/inputNN (src/params.f90:113-330, 462-468)."""
    lib = capi.load()
    base = base_dir if base_dir is not None else os.path.dirname(os.path.abspath(params_path))
    h = C.c_void_p()
    capi.check(lib.rfinv_problem_load(params_path.encode(), base.encode(), C.byref(h)))
    try:
        capi.check(lib.rfinv_problem_write_copies(h, out_dir.encode(), input_dir.encode()))
    finally:
        lib.rfinv_problem_free(h)


def write_outputs(cfg: RFConfig, out_dir: str, nproc_total: int, hist: Dict[str, np.ndarray], likelihood_hist: np.ndarray,
                  vp_model: Optional[np.ndarray] = None, vs_model: Optional[np.ndarray] = None) -> None:
    """output_results (src/mcmc_out.f90:97-318) from job-wide sums (see ParallelTempering.hist)."""
    lib = capi.load()
    c = cfg.to_c()
    a64 = lambda k: np.ascontiguousarray(hist[k], dtype=np.int64)
    f64 = lambda k: np.ascontiguousarray(hist[k], dtype=np.float64)
    nk, nz, nsig, namp, nvpz, nvsz, nvpvsz = (a64(k) for k in ("nk", "nz", "nsig", "namp", "nvpz", "nvsz", "nvpvsz"))
    vpm, vsm, vpvsm = f64("vp_mean"), f64("vs_mean"), f64("vpvs_mean")
    lh = np.ascontiguousarray(likelihood_hist, dtype=np.float64)
    n_models = 0 if vp_model is None else int(vp_model.shape[0])
    vpmod = None if vp_model is None else np.ascontiguousarray(vp_model, dtype=np.float64)
    vsmod = None if vs_model is None else np.ascontiguousarray(vs_model, dtype=np.float64)
    os.makedirs(out_dir, exist_ok=True)
    capi.check(lib.rfinv_write_outputs(C.byref(c), out_dir.encode(), int(nproc_total), int(hist["nmod"]),
                                       _p(nk, capi.i64p), _p(nz, capi.i64p), _p(nsig, capi.i64p), _p(namp, capi.i64p),
                                       _p(nvpz, capi.i64p), _p(nvsz, capi.i64p), _p(nvpvsz, capi.i64p), _p(vpm, capi.dp),
                                       _p(vsm, capi.dp), _p(vpvsm, capi.dp), _p(lh, capi.dp), int(lh.shape[0]),
                                       _p(vpmod, capi.dp), _p(vsmod, capi.dp), n_models))


def write_sac(path: str, data: np.ndarray, delta: float, b: float) -> None:
    """Minimal SAC file with the header words make_syn sets (src/make_syn.f90:127-137); native endianness."""
    data = np.asarray(data, dtype=np.float32)
    raw = np.zeros(158 + data.size, dtype=np.float32)
    ints = raw.view(np.int32)
    raw[0] = np.float32(delta)                      # record 1  delta
    raw[5] = np.float32(b)                          # record 6  b
    raw[6] = np.float32(b + (data.size - 1) * delta)  # record 7  e
    ints[76] = 6                                    # record 77 nvhdr
    ints[79] = data.size                            # record 80 npts
    ints[85] = 1                                    # record 86 iftype
    ints[105] = 1                                   # record 106 leven
    raw[158:] = data
    raw.tofile(path)
