"""Host-side mirror of the reference's evaluation boundary, on top of the C-ABI.

``Evaluator.calc_likelihood`` is the batched form of ``calc_likelihood`` (src/likelihood.f90:56-101) with
the same argument meaning; ``Evaluator.format_model`` is ``format_model`` (src/model.f90:175-290).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import capi
from .config import RFConfig


def _p(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(t) if a is not None else C.cast(None, t)


class Evaluator:
    """One handle = one GPU + one stream (the reference is one MPI rank = one process, not re-entrant)."""

    def __init__(self, cfg: RFConfig, device: int = 0):
        self.cfg = cfg
        self._lib = capi.load()
        self._c = cfg.to_c()
        h = C.c_void_p()
        capi.check(self._lib.rfinv_create(C.byref(self._c), int(device), C.byref(h)))
        self._h = h
        self.device = device

    # -- life cycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rfinv_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("Evaluator is closed")
        return self._h

    def set_stream(self, cuda_stream: int) -> None:
        capi.check(self._lib.rfinv_set_stream(self.handle, int(cuda_stream)))

    def synchronize(self) -> None:
        capi.check(self._lib.rfinv_synchronize(self.handle))

    def quadform_form(self):
        """Per trace: (rank, rank_s, split) of the quadratic form chosen at create (rank 0 = dense m^T R^-1 m)."""
        T = self.cfg.ntrc
        r, rs, sp = (C.c_int32 * T)(), (C.c_int32 * T)(), (C.c_int32 * T)()
        capi.check(self._lib.rfinv_get_quadform_form(self.handle, r, rs, sp))
        return [(int(r[t]), int(rs[t]), int(sp[t])) for t in range(T)]

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.rfinv_last_launch_count(self.handle))

    # -- the boundary ---------------------------------------------------------------------------
    def _models(self, k, z, dvp, dvs):
        cfg = self.cfg
        k = np.ascontiguousarray(k, dtype=np.int32)
        n = k.shape[0]
        z = np.ascontiguousarray(z, dtype=np.float64)
        dvp = np.ascontiguousarray(dvp, dtype=np.float64)
        dvs = np.ascontiguousarray(dvs, dtype=np.float64)
        if z.shape != (n, cfg.k_max - 1) or dvp.shape != (n, cfg.k_max) or dvs.shape != (n, cfg.k_max):
            raise ValueError("expected z[C][k_max-1], dvp[C][k_max], dvs[C][k_max]")
        return n, k, z, dvp, dvs

    def calc_likelihood(self, k, z, dvp, dvs, sig, want_rft: bool = False, want_valid: bool = False
                        ) -> Tuple[np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]:
        """calc_likelihood(fwd_flag=.true.) for C models.  Host arrays in, host arrays out.

        Returns (log_likelihood[C], rft[C][ntrc][nfft] or None, is_valid[C] or None)."""
        n, k, z, dvp, dvs = self._models(k, z, dvp, dvs)
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        if sig.shape != (n, self.cfg.ntrc):
            raise ValueError("expected sig[C][ntrc]")
        logl = np.empty(n)
        rft = np.empty((n, self.cfg.ntrc, self.cfg.nfft)) if want_rft else None
        valid = np.empty(n, dtype=np.uint8) if want_valid else None
        capi.check(self._lib.rfinv_eval_batch(self.handle, n, _p(k, capi.i32p), _p(z, capi.dp), _p(dvp, capi.dp),
                                              _p(dvs, capi.dp), _p(sig, capi.dp), _p(logl, capi.dp), _p(rft, capi.dp),
                                              _p(valid, capi.u8p)))
        return logl, rft, (valid.astype(bool) if valid is not None else None)

    def calc_likelihood_flags(self, fwd_flag, k, z, dvp, dvs, sig, phi: np.ndarray, want_valid: bool = False):
        """calc_likelihood with a per-model fwd_flag (src/likelihood.f90:74-82).  `phi[C][ntrc]` is updated in place: written
        for the models with fwd_flag, read (the cached value of the chain's current RF) for the others.
        Returns (log_likelihood[C], is_valid[C] or None)."""
        n, k, z, dvp, dvs = self._models(k, z, dvp, dvs)
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        flags = np.ascontiguousarray(fwd_flag, dtype=np.uint8)
        if sig.shape != (n, self.cfg.ntrc) or flags.shape != (n,):
            raise ValueError("expected sig[C][ntrc], fwd_flag[C]")
        if phi.dtype != np.float64 or phi.shape != (n, self.cfg.ntrc) or not phi.flags["C_CONTIGUOUS"]:
            raise ValueError("phi must be a C-contiguous float64 array [C][ntrc] (updated in place)")
        logl = np.empty(n)
        valid = np.empty(n, dtype=np.uint8) if want_valid else None
        capi.check(self._lib.rfinv_eval_batch_flags(self.handle, n, _p(flags, capi.u8p), _p(k, capi.i32p), _p(z, capi.dp),
                                                    _p(dvp, capi.dp), _p(dvs, capi.dp), _p(sig, capi.dp), _p(phi, capi.dp),
                                                    _p(logl, capi.dp), _p(valid, capi.u8p)))
        return logl, (valid.astype(bool) if valid is not None else None)

    def calc_likelihood_begin(self, slot: int, k, z, dvp, dvs, sig, logl: np.ndarray, valid: Optional[np.ndarray] = None):
        """Asynchronous calc_likelihood of one group of chains in `slot` (0 or 1): returns at once, `logl` (and `valid`) are
        filled when calc_likelihood_end(slot) returns.  The arrays are used in place (no copies: pass C-contiguous arrays of the
        right dtype, page-locked for real overlap) and must stay alive and untouched until then."""
        n = k.shape[0]
        for a, dt, shape in ((k, np.int32, (n,)), (z, np.float64, (n, self.cfg.k_max - 1)), (dvp, np.float64, (n, self.cfg.k_max)),
                             (dvs, np.float64, (n, self.cfg.k_max)), (sig, np.float64, (n, self.cfg.ntrc)), (logl, np.float64, (n,))):
            if a.dtype != dt or a.shape != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("calc_likelihood_begin needs C-contiguous arrays of the boundary's dtypes and shapes")
        capi.check(self._lib.rfinv_eval_batch_begin(self.handle, int(slot), n, _p(k, capi.i32p), _p(z, capi.dp), _p(dvp, capi.dp),
                                                    _p(dvs, capi.dp), _p(sig, capi.dp), _p(logl, capi.dp), _p(valid, capi.u8p)))

    def calc_likelihood_end(self, slot: int) -> None:
        capi.check(self._lib.rfinv_eval_batch_end(self.handle, int(slot)))

    def calc_likelihood_device(self, C_models: int, d_k: int, d_z: int, d_dvp: int, d_dvs: int, d_sig: int, d_logl: int,
                               d_rft_smp: int = 0, d_is_valid: int = 0) -> None:
        """Same evaluation on device-resident chain-fastest arrays (raw device pointers); asynchronous."""
        capi.check(self._lib.rfinv_eval_batch_device(self.handle, int(C_models), d_k, d_z, d_dvp, d_dvs, d_sig, d_logl,
                                                     d_rft_smp, d_is_valid))

    def format_model(self, k, z, dvp, dvs):
        """format_model for C models -> (nlay[C], alpha, beta, rho, h [C][k_max+1], is_valid[C])."""
        n, k, z, dvp, dvs = self._models(k, z, dvp, dvs)
        stride = self.cfg.k_max + 1
        nlay = np.zeros(n, dtype=np.int32)
        out = [np.zeros((n, stride)) for _ in range(4)]
        valid = np.zeros(n, dtype=np.uint8)
        capi.check(self._lib.rfinv_format_model_batch(self.handle, n, _p(k, capi.i32p), _p(z, capi.dp), _p(dvp, capi.dp),
                                                      _p(dvs, capi.dp), _p(nlay, capi.i32p), *[_p(o, capi.dp) for o in out],
                                                      _p(valid, capi.u8p)))
        return nlay, out[0], out[1], out[2], out[3], valid.astype(bool)

    def r_inv(self) -> np.ndarray:
        """The inverse data covariance in use, [ntrc][nsmp][nsmp] (src/likelihood.f90:222)."""
        out = np.empty((self.cfg.ntrc, self.cfg.nsmp, self.cfg.nsmp))
        capi.check(self._lib.rfinv_get_r_inv(self.handle, _p(out, capi.dp)))
        return out
