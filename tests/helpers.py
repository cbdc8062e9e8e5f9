"""Shared test helpers: configurations with oracle-made observations, oracle adapters."""
from __future__ import annotations

import os

import numpy as np

import oracle_c
import rfinv_oracle as pyo
from rf_inv_b200 import workloads
from rf_inv_b200.config import RFConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def py_config(c: RFConfig) -> pyo.Config:
    """RFConfig -> the numpy oracle's Config (obs is (nsmp, ntrc) there, like the Fortran array)."""
    return pyo.Config(
        ntrc=c.ntrc, nfft=c.nfft, nsmp=c.nsmp, delta=c.delta, t_start=c.t_start, rayps=list(c.rayps),
        a_gus=list(c.a_gus), ipha=list(c.ipha), deconv_mode=c.deconv_mode, sdep=c.sdep, bdep=c.bdep,
        obs=None if c.obs is None else np.asarray(c.obs).T.copy(), vp_ref=np.asarray(c.vp_ref),
        vs_ref=np.asarray(c.vs_ref), z_ref_min=c.z_ref_min, dz_ref=c.dz_ref, vp_mode=c.vp_mode, k_min=c.k_min,
        k_max=c.k_max, z_min=c.z_min, z_max=c.z_max, h_min=c.h_min, prior_mode=c.prior_mode, dvs_prior=c.dvs_prior,
        dvp_prior=c.dvp_prior, sig_min=list(c.sig_min), sig_max=list(c.sig_max), vp_min=c.vp_min, vp_max=c.vp_max,
        vs_min=c.vs_min, vs_max=c.vs_max, vpvs_min=c.vpvs_min, vpvs_max=c.vpvs_max, dev_z=c.dev_z, dev_dvs=c.dev_dvs,
        dev_dvp=c.dev_dvp, dev_sig=c.dev_sig, nburn=c.nburn, niter=c.niter, ncorr=c.ncorr, nchains=c.nchains,
        ncool=c.ncool, t_high=c.t_high, iseed=c.iseed)


_RINV_CACHE = {}


def scipy_r_inv(cfg: RFConfig) -> np.ndarray:
    """init_r_inv with scipy's gesvd (the oracle's dgesvd stand-in) -> [ntrc][nsmp][nsmp] (memory order of
    the Fortran r_inv(i,j,t): element [t][j][i])."""
    out = np.empty((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    for t in range(cfg.ntrc):
        key = (cfg.nsmp, float(cfg.a_gus[t]), float(cfg.delta))
        if key not in _RINV_CACHE:
            one = pyo.Config(ntrc=1, nfft=cfg.nfft, nsmp=cfg.nsmp, delta=cfg.delta, t_start=0.0, rayps=[0.0],
                             a_gus=[cfg.a_gus[t]], ipha=[1])
            _RINV_CACHE[key] = np.ascontiguousarray(pyo.init_r_inv(one)[:, :, 0].T)
        out[t] = _RINV_CACHE[key]
    return out


def attach_obs_and_rinv(cfg: RFConfig, noise_seed: int = 7, noise: float = 0.0) -> RFConfig:
    """Observed traces = oracle forward of the data-generating model, stored as float32 like a SAC file."""
    tm = workloads.true_model(cfg)
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp))
    cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    _, rft, _ = oracle_c.eval_batch(cfg, tm["k"], tm["z"], tm["dvp"], tm["dvs"], tm["sig"])
    obs = rft[0, :, :cfg.nsmp]
    if noise > 0:
        obs = obs + np.random.default_rng(noise_seed).normal(0.0, noise, obs.shape)
    cfg.obs = obs.astype(np.float32).astype(np.float64)
    cfg.r_inv = scipy_r_inv(cfg)
    return cfg


def small_config(**kw) -> RFConfig:
    """A 2-trace, nfft=256 configuration in the style of sample_syn (fast on the CPU oracle)."""
    vp_ref, vs_ref = workloads.reference_velmod(60.0)
    d = dict(ntrc=2, nfft=256, nsmp=101, delta=workloads.DELTA_F32, t_start=0.0, rayps=[0.06, 0.08],
             a_gus=[4.0, 4.0], ipha=[1, 1], vp_ref=vp_ref, vs_ref=vs_ref, k_max=10, z_max=20.0,
             sig_min=[0.01, 0.01], sig_max=[0.01, 0.01])
    d.update(kw)
    return RFConfig(**d)


def rel_err_rft(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| relative to max |b| per (model, trace) -- the bar of SURVEY.md 8d."""
    scale = np.max(np.abs(b), axis=-1, keepdims=True)
    return float(np.nanmax(np.abs(a - b) / scale))


def logl_err(cfg: RFConfig, a: np.ndarray, b: np.ndarray, sig: np.ndarray) -> float:
    """|a-b| relative to the magnitude of the terms logL is summed from (src/likelihood.f90:94-96):
    logL = sum_t -phi/(2 sig^2) - nsmp log(sig) can cancel to ~0, so |logL| alone is not a fair scale."""
    scale = np.abs(b) + cfg.nsmp * np.sum(np.abs(np.log(sig)), axis=1)
    return float(np.nanmax(np.abs(a - b) / scale))
