"""Named synthetic workloads (SURVEY.md section 8d / BASELINE.json configs) and deterministic inputs.

Only builds *inputs* (configurations and random velocity models); it computes no receiver functions.
Observed traces are attached by the caller: tests use the oracle, bench.py uses the CUDA path itself.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .config import RFConfig

DELTA_F32 = float(np.float32(0.05))  # the reference takes delta from the SAC header, a float32 (src/params.f90:443-449)


def reference_velmod(z_max_km: float = 60.0, dz: float = 0.5, vp: float = 5.0, vs: float = 2.89):
    """sample_syn/model/sample.velmod extended in depth: constant Vp 5.00, Vs 2.89."""
    n = int(round(z_max_km / dz)) + 1
    return np.full(n, vp), np.full(n, vs)


def _base(**kw) -> RFConfig:
    vp_ref, vs_ref = reference_velmod()
    d = dict(delta=DELTA_F32, t_start=0.0, deconv_mode=0, sdep=0.0, vp_ref=vp_ref, vs_ref=vs_ref, z_ref_min=0.0,
             dz_ref=0.5, vp_mode=0, k_min=1, k_max=10, prior_mode=2, z_min=0.0, z_max=40.0, h_min=0.05, dvs_prior=2.0,
             dvp_prior=0.2, vp_min=0.1, vp_max=8.6, vs_min=0.001, vs_max=5.0, vpvs_min=0.0, vpvs_max=5.0, dev_z=0.02,
             dev_dvs=0.02, dev_dvp=0.02, dev_sig=0.002, nburn=0, niter=0, ncorr=10, nchains=1, ncool=1, iseed=12345678,
             t_high=15.0)
    d.update(kw)
    T = d["ntrc"]
    d.setdefault("sig_min", [0.01] * T)
    d.setdefault("sig_max", [0.01] * T)
    return RFConfig(**d)


def make_config(name: str) -> RFConfig:
    """Workload names: sample (C1), c2, c3, c3_buried, c3_deconv, c3_common, c4, c4_laplace, c5, target."""
    if name == "sample":  # sample_syn/params.in as shipped (incl. SEA_DEP 2.0), 20 ranks x 5 chains
        vp_ref, vs_ref = reference_velmod(30.0)
        return _base(ntrc=2, nfft=256, nsmp=101, rayps=[0.06, 0.08], a_gus=[4.0, 4.0], ipha=[1, 1], sdep=2.0,
                     vp_ref=vp_ref, vs_ref=vs_ref, k_max=10, z_max=20.0, nburn=3000, niter=8000, ncorr=10, nchains=5,
                     ncool=1, t_high=15.0)
    if name == "c2":      # single P-RF trace, npts=1024, k_max=20, 4096 chains
        return _base(ntrc=1, nfft=1024, nsmp=512, rayps=[0.06], a_gus=[8.0], ipha=[1], k_max=20, nchains=1)
    if name == "c3":      # OBS: sea-water layer, 3 traces with distinct rays and Gaussian widths, sigma solved for trace 3
        return _base(ntrc=3, nfft=1024, nsmp=512, rayps=[0.05, 0.06, 0.07], a_gus=[2.0, 4.0, 8.0], ipha=[1, 1, 1],
                     sdep=2.0, k_max=20, sig_min=[0.01, 0.01, 0.005], sig_max=[0.01, 0.01, 0.05], nchains=16)
    if name == "c3_buried":  # BASELINE.json config 3 as worded: sea-water top layer AND buried receiver (1 km below the sea floor)
        cfg = make_config("c3")
        cfg.bdep = 1.0
        return cfg
    if name == "c3_deconv":  # c3 with water-level deconvolution (DECONV_MODE 1)
        cfg = make_config("c3")
        cfg.deconv_mode = 1
        return cfg
    if name == "c3_common":  # c3 with one ray for all traces: one propagator pass per model (check_ray, src/forward.f90:59-76)
        cfg = make_config("c3")
        cfg.rayps = [0.06, 0.06, 0.06]
        return cfg
    if name in ("c4", "c4_laplace"):  # joint P and S with velocity-perturbation prior
        return _base(ntrc=4, nfft=1024, nsmp=512, rayps=[0.05, 0.07, 0.10, 0.12], a_gus=[4.0, 4.0, 4.0, 4.0],
                     ipha=[1, 1, -1, -1], vp_mode=1, k_max=20, nchains=16, prior_mode=1 if name == "c4_laplace" else 2)
    if name == "c5":      # large sweep: npts=2048, k_max=30, 5 traces
        return _base(ntrc=5, nfft=2048, nsmp=1000, rayps=[0.04, 0.05, 0.06, 0.07, 0.08], a_gus=[4.0] * 5, ipha=[1] * 5,
                     k_max=30, nchains=16)
    if name == "target":  # north-star shape: npts=1024, k_max=30, 3 traces per model (distinct rays)
        return _base(ntrc=3, nfft=1024, nsmp=512, rayps=[0.05, 0.06, 0.07], a_gus=[4.0, 4.0, 4.0], ipha=[1, 1, 1],
                     k_max=30, nchains=16)
    raise KeyError(name)


def true_model(cfg: RFConfig) -> Dict[str, np.ndarray]:
    """Data-generating model: sample_syn/true/true.velmod (interfaces at 3.14 and 7.75 km below the sea floor)."""
    km = cfg.k_max
    z = np.zeros(km - 1); dvp = np.zeros(km); dvs = np.zeros(km)
    z[0] = cfg.sdep + 3.1415985198691487
    z[1] = z[0] + 4.6098561491817236
    dvs[0] = 4.7560963020861591 - 2.89
    dvs[1] = 3.5364953336536549 - 2.89
    dvs[km - 1] = 4.6845348442560484 - 2.89
    return dict(k=np.array([2], dtype=np.int32), z=z[None, :], dvp=dvp[None, :], dvs=dvs[None, :],
                sig=np.asarray(cfg.sig_min, dtype=np.float64)[None, :])


def models_valid(cfg: RFConfig, k, z, dvp, dvs) -> np.ndarray:
    """Vectorised restatement of format_model's validity flag for *input generation* (rejection sampling)."""
    n = k.shape[0]
    ok = np.ones(n, dtype=bool)
    for c in range(n):
        kk = int(k[c])
        order = np.argsort(z[c, :kk], kind="stable")
        tz = z[c, :kk][order]; ts = dvs[c, :kk][order]; tp = dvp[c, :kk][order]
        tops = np.concatenate(([cfg.sdep], tz))
        bots = np.concatenate((tz, [cfg.z_max]))
        zc = 0.5 * (tops + bots)
        iz = np.floor((zc - cfg.z_ref_min) / cfg.dz_ref + 0.5).astype(int)
        iz = np.clip(iz, 0, len(cfg.vp_ref) - 1)
        b = np.asarray(cfg.vs_ref)[iz] + np.concatenate((ts, [dvs[c, cfg.k_max - 1]]))
        a = np.asarray(cfg.vp_ref)[iz] + (np.concatenate((tp, [dvp[c, cfg.k_max - 1]])) if cfg.vp_mode == 1 else 0.0)
        h = bots - tops
        bad = (a < cfg.vp_min) | (a > cfg.vp_max) | (b < cfg.vs_min) | (b > cfg.vs_max)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = a / b
        bad |= (r < cfg.vpvs_min) | (r > cfg.vpvs_max)
        if bad.any() or h[0] < 0.125 * a[0] or (kk > 1 and (h[1:kk] < cfg.h_min).any()):
            ok[c] = False
    return ok


def draw_models(cfg: RFConfig, n: int, seed: int = 1, dvs_scale: float = None, k_fixed: int = None) -> Dict[str, np.ndarray]:
    """n random valid models drawn like init_model (src/model.f90:62-95): k uniform on [k_min, k_max) -- or k_fixed for
    every model --, interfaces uniform on [z_min, z_max], velocity perturbations from the prior, rejection until valid.
    (numpy's generator, not the reference's mt19937 stream: these are benchmark / parity inputs.)"""
    rng = np.random.default_rng(seed)
    km, T = cfg.k_max, cfg.ntrc
    k = np.zeros(n, dtype=np.int32); z = np.zeros((n, km - 1)); dvp = np.zeros((n, km)); dvs = np.zeros((n, km))
    todo = np.arange(n)
    s_vs = cfg.dvs_prior if dvs_scale is None else dvs_scale
    while todo.size:
        m = todo.size
        kk = rng.integers(cfg.k_min, cfg.k_max, m).astype(np.int32)
        if k_fixed is not None:
            kk[:] = k_fixed
        zz = np.zeros((m, km - 1)); pp = np.zeros((m, km)); ss = np.zeros((m, km))
        for i in range(m):
            zz[i, :kk[i]] = rng.uniform(cfg.z_min, cfg.z_max, kk[i])
            if cfg.prior_mode == 1:
                ss[i, :kk[i]] = rng.laplace(0.0, s_vs, kk[i]); pp[i, :kk[i]] = rng.laplace(0.0, cfg.dvp_prior, kk[i])
                ss[i, km - 1] = rng.laplace(0.0, s_vs); pp[i, km - 1] = rng.laplace(0.0, cfg.dvp_prior)
            else:
                ss[i, :kk[i]] = rng.normal(0.0, s_vs, kk[i]); pp[i, :kk[i]] = rng.normal(0.0, cfg.dvp_prior, kk[i])
                ss[i, km - 1] = rng.normal(0.0, s_vs); pp[i, km - 1] = rng.normal(0.0, cfg.dvp_prior)
        ok = models_valid(cfg, kk, zz, pp, ss)
        good = todo[ok]
        k[good] = kk[ok]; z[good] = zz[ok]; dvp[good] = pp[ok]; dvs[good] = ss[ok]
        todo = todo[~ok]
    sig = np.empty((n, T))
    for t in range(T):
        lo, hi = cfg.sig_min[t], cfg.sig_max[t]
        sig[:, t] = lo if cfg.sig_mode[t] == 0 else rng.uniform(lo, hi, n)
    return dict(k=k, z=z, dvp=dvp, dvs=dvs, sig=sig)


def to_soa(models: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Chain-major host layout -> the library's chain-fastest device layout."""
    return dict(k=np.ascontiguousarray(models["k"], dtype=np.int32),
                z=np.ascontiguousarray(models["z"].T), dvp=np.ascontiguousarray(models["dvp"].T),
                dvs=np.ascontiguousarray(models["dvs"].T), sig=np.ascontiguousarray(models["sig"].T))


def band_bins(cfg: RFConfig):
    """Frequency bins forward_kernel propagates per trace: mirror of band_limits() in csrc/capi.cu (bin groups whose
    Gaussian-filter weight is below 2^-52 of the filter mass are zero filled; RFINV_FULL_BAND=1 or the water-level
    deconvolution keep the full band).  Accounting only -- the library decides for itself."""
    import os
    n, nh = cfg.nfft, cfg.nh
    jfull = 1 if n <= 64 else (2 if n <= 256 else (4 if n <= 2048 else 8))
    nthr = (n // 2) // jfull
    allowed = [1, 2, 3, 4, 6, 8]
    full = os.environ.get("RFINV_FULL_BAND", "0") not in ("", "0") or cfg.deconv_mode == 1
    groups = []
    for t in range(cfg.ntrc):
        g = jfull
        if not full:
            omega = np.arange(nh) * 2.0 * np.pi / (cfg.delta * n)
            f = np.exp(-(omega / (2.0 * cfg.a_gus[t])) ** 2)
            total = f.sum()
            for cand in range(jfull - 1, 0, -1):
                if f[cand * nthr:].sum() <= total * 2.220446049250313e-16:
                    g = cand
                else:
                    break
        groups.append(min(a for a in allowed if a >= g and a <= jfull) if any(a >= g and a <= jfull for a in allowed) else jfull)
    if cfg.is_ray_common:
        groups = [max(groups)] * cfg.ntrc
    return [min(gr * nthr, n // 2) + 1 for gr in groups], groups


def flops_per_eval(cfg: RFConfig, k_mean: float) -> Dict[str, float]:
    """Algorithmic fp64 flop per forward+likelihood evaluation: an exact count of the algorithm the
    kernels execute (an FMA is 2 flop, a multiply / add / divide 1), not of the reference's dense complex form:

      W = sum_rays nb_ray*(56*k + 56) + T*5*n*log2(n) + T*(S^2 + 3*S),   nb_ray = bins inside the band limit (band_bins)

    per layer and frequency bin (wave coordinates, rf_inv_b200/csrc/forward.cu): two plane rotations for each of the
    two propagated vectors (2 x (4 mul + 4 fma) = 24 flop), the 2x2 interface blocks with unit {1,4} diagonal
    (2 x (2 fma + 2 mul + 2 fma) = 20), one rotation of each (cos, sin) pair to the bin (2 x (2 mul + 2 fma) = 12):
    56 flop in 36 FP64 instructions; the last solid layer has no interface (-20).  Per bin: rows 3,4 of E^-1, the
    boundary condition with one complex division, sign conventions, Gaussian filter, Hermitian packing: 76.
    One packed complex inverse FFT per trace (the customary 5 n log2 n); symmetric quadratic form S(S+1)/2 MAC + row dot.
    SURVEY.md 8d's W_min (110 k + 100 per bin) and the reference's W_ref (570 k + 550) are reported next to it."""
    Tf = 1 if cfg.is_ray_common else cfg.ntrc
    n, nh, S, T = cfg.nfft, cfg.nh, cfg.nsmp, cfg.ntrc
    bins, groups = band_bins(cfg)
    nb = float(bins[0]) if cfg.is_ray_common else float(sum(bins))   # bins propagated per evaluation (all rays)
    prop = nb * (56.0 * k_mean + 56.0)
    fft = T * 5.0 * n * np.log2(n)
    quad = T * (1.0 * S * S + 3.0 * S)
    w_min_survey = Tf * nh * (110.0 * k_mean + 100.0) + fft + T * (2.0 * S * S + 3.0 * S)
    w_ref = Tf * nh * (570.0 * k_mean + 550.0) + fft + T * (2.0 * S * S + 3.0 * S)
    # FP64 instructions the forward path issues per evaluation (each occupies one issue slot of the FP64 pipe, FMA or not)
    fwd_instr = nb * (36.0 * k_mean + 40.0) + T * (3.0 * (n / 8) * 84.0 + (n / 2) * 4.0)
    return dict(propagator=prop, fft=fft, quadform=quad, total=prop + fft + quad, survey_w_min=w_min_survey,
                reference_as_written=w_ref, forward_fp64_instructions=fwd_instr, bins_propagated_per_eval=nb,
                bins_full_band_per_eval=float(Tf * nh), band_groups=groups)


def lapack_r_inv(cfg: RFConfig) -> np.ndarray:
    """init_r_inv the way the reference does it (src/likelihood.f90:168-241): LAPACK dgesvd of
    R_ij = r^((i-j)^2), truncated at s > 1e-3 -- here through scipy's gesvd driver.  Returns the array in
    the memory order of the Fortran r_inv(i,j,t): [ntrc][nsmp][nsmp] with element [t][j][i]."""
    import scipy.linalg

    S = cfg.nsmp
    out = np.empty((cfg.ntrc, S, S))
    idx = np.arange(S)
    d2 = ((idx[:, None] - idx[None, :]) ** 2).astype(np.float64)
    cache = {}
    for t in range(cfg.ntrc):
        a = float(cfg.a_gus[t])
        if a not in cache:
            r = np.exp(-a ** 2 * cfg.delta ** 2)
            u, s, vt = scipy.linalg.svd(np.power(r, d2), full_matrices=True, lapack_driver="gesvd")
            dinv = np.where(s > 1.0e-3, 1.0 / s, 0.0)
            cache[a] = np.ascontiguousarray(((vt.T * dinv[None, :]) @ u.T).T)
        out[t] = cache[a]
    return out
