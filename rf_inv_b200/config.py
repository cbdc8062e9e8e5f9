"""Host-side mirror of ``rfinv_config`` (include/rfinv_b200.h).

The reference keeps this state in Fortran module globals filled by ``get_params`` / ``read_obs``
(src/params.f90:101-476) and ``read_ref_model`` (src/model.f90:109-171).  ``RFConfig`` carries the same
fields with the same names; ``to_c()`` produces the ctypes struct that crosses the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class RfinvConfigC(C.Structure):
    """Field order must match ``struct rfinv_config`` exactly."""

    _fields_ = [
        ("ntrc", C.c_int32), ("nfft", C.c_int32), ("nsmp", C.c_int32), ("deconv_mode", C.c_int32),
        ("delta", C.c_double), ("t_start", C.c_double), ("sdep", C.c_double),
        ("rayps", c_double_p), ("a_gus", c_double_p), ("ipha", c_int32_p),
        ("obs", c_double_p), ("r_inv", c_double_p),
        ("nref", C.c_int32), ("pad0_", C.c_int32),
        ("z_ref_min", C.c_double), ("dz_ref", C.c_double),
        ("vp_ref", c_double_p), ("vs_ref", c_double_p),
        ("vp_mode", C.c_int32), ("k_min", C.c_int32), ("k_max", C.c_int32), ("prior_mode", C.c_int32),
        ("z_min", C.c_double), ("z_max", C.c_double), ("h_min", C.c_double),
        ("dvs_prior", C.c_double), ("dvp_prior", C.c_double),
        ("sig_min", c_double_p), ("sig_max", c_double_p),
        ("vp_min", C.c_double), ("vp_max", C.c_double), ("vs_min", C.c_double), ("vs_max", C.c_double),
        ("vpvs_min", C.c_double), ("vpvs_max", C.c_double),
        ("dev_z", C.c_double), ("dev_dvs", C.c_double), ("dev_dvp", C.c_double), ("dev_sig", C.c_double),
        ("nburn", C.c_int32), ("niter", C.c_int32), ("ncorr", C.c_int32),
        ("nchains", C.c_int32), ("ncool", C.c_int32), ("iseed", C.c_int32),
        ("t_high", C.c_double),
        ("nbin_z", C.c_int32), ("nbin_vs", C.c_int32), ("nbin_vp", C.c_int32), ("nbin_vpvs", C.c_int32),
        ("nbin_sig", C.c_int32), ("nbin_amp", C.c_int32),
        ("amp_min", C.c_double), ("amp_max", C.c_double),
        ("bdep", C.c_double),
    ]


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class RFConfig:
    """Same names as the reference's ``params`` / ``model`` module variables."""

    # observation / RF synthesis (src/params.f90:60-74)
    ntrc: int
    nfft: int
    nsmp: int
    delta: float
    t_start: float
    rayps: Sequence[float]
    a_gus: Sequence[float]
    ipha: Sequence[int]
    deconv_mode: int = 0
    sdep: float = 0.0
    bdep: float = 0.0                     # BOREHOLE_DEP: receiver depth below the surface / sea floor (commented out in the
                                          # reference, src/params.f90:67,203-224); 0 = at the surface
    obs: Optional[np.ndarray] = None      # [ntrc][nsmp]
    r_inv: Optional[np.ndarray] = None    # [ntrc][nsmp][nsmp]; None -> library builds it
    # reference velocity model (src/model.f90:35-36)
    vp_ref: Optional[np.ndarray] = None
    vs_ref: Optional[np.ndarray] = None
    z_ref_min: float = 0.0
    dz_ref: float = 0.5
    # prior / validity (src/params.f90:83-90)
    vp_mode: int = 0
    k_min: int = 1
    k_max: int = 10
    prior_mode: int = 2
    z_min: float = 0.0
    z_max: float = 20.0
    h_min: float = 0.05
    dvs_prior: float = 2.0
    dvp_prior: float = 0.2
    sig_min: Sequence[float] = ()
    sig_max: Sequence[float] = ()
    vp_min: float = 0.1
    vp_max: float = 8.6
    vs_min: float = 0.001
    vs_max: float = 5.0
    vpvs_min: float = 0.0
    vpvs_max: float = 5.0
    # proposal widths (src/params.f90:92)
    dev_z: float = 0.02
    dev_dvs: float = 0.02
    dev_dvp: float = 0.02
    dev_sig: float = 0.002
    # parallel tempering (src/params.f90:49-56)
    nburn: int = 0
    niter: int = 0
    ncorr: int = 1
    nchains: int = 1
    ncool: int = 1
    iseed: int = 12345678
    t_high: float = 1.0
    # posterior histograms (src/params.f90:94-96)
    nbin_z: int = 100
    nbin_vs: int = 50
    nbin_vp: int = 50
    nbin_vpvs: int = 100
    nbin_sig: int = 50
    nbin_amp: int = 100
    amp_min: float = -0.8
    amp_max: float = 0.8
    out_dir: str = "."
    _keep: List[np.ndarray] = field(default_factory=list, repr=False, compare=False)

    @property
    def nh(self) -> int:
        return self.nfft // 2 + 1

    @property
    def sig_mode(self) -> List[int]:
        """src/params.f90:262 -- the threshold is the single-precision literal 1.0e-5."""
        eps = float(np.float32(1.0e-5))
        return [1 if (self.sig_max[t] - self.sig_min[t]) > eps else 0 for t in range(self.ntrc)]

    @property
    def is_ray_common(self) -> bool:
        """src/forward.f90:59-76."""
        return all(self.rayps[t] == self.rayps[0] and self.ipha[t] == self.ipha[0] for t in range(self.ntrc))

    def validate(self) -> None:
        if self.ntrc < 1:
            raise ValueError("ntrc must be >= 1")
        for name in ("rayps", "a_gus", "ipha", "sig_min", "sig_max"):
            if len(getattr(self, name)) != self.ntrc:
                raise ValueError(f"{name} must have ntrc={self.ntrc} entries")
        if not (64 <= self.nfft <= 4096) or (self.nfft & (self.nfft - 1) and self.nfft > 2048):
            raise ValueError("nfft must be a power of two in [64, 4096] or any other length in [64, 2048]")
        if not (1 <= self.nsmp <= self.nfft):
            raise ValueError("nsmp must be in [1, nfft]")
        if self.bdep < 0.0:
            raise ValueError("BOREHOLE_DEP must be positive")          # src/params.f90:214-217 (commented out there)
        if self.deconv_mode not in (0, 1):
            raise ValueError("deconv_mode must be either 0 or 1")  # src/params.f90:195-199
        if self.vp_mode not in (0, 1):
            raise ValueError("vp_mode should be 0 or 1")             # src/pt_mcmc.f90:329-335
        if not (2 <= self.k_max <= 64) or not (1 <= self.k_min < self.k_max):
            raise ValueError("need 1 <= k_min < k_max <= 64")
        if self.vp_ref is None or self.vs_ref is None or len(self.vp_ref) != len(self.vs_ref):
            raise ValueError("vp_ref / vs_ref missing or of different length")
        if self.obs is not None and tuple(np.shape(self.obs)) != (self.ntrc, self.nsmp):
            raise ValueError("obs must have shape [ntrc][nsmp]")
        if self.r_inv is not None and tuple(np.shape(self.r_inv)) != (self.ntrc, self.nsmp, self.nsmp):
            raise ValueError("r_inv must have shape [ntrc][nsmp][nsmp]")

    def to_c(self) -> RfinvConfigC:
        """Builds the C struct; the numpy buffers it points to are kept alive on ``self``."""
        self.validate()
        keep = []

        def arr(a, dtype=np.float64):
            b = np.ascontiguousarray(np.asarray(a, dtype=dtype))
            keep.append(b)
            return b

        c = RfinvConfigC()
        for name in ("ntrc", "nfft", "nsmp", "deconv_mode", "vp_mode", "k_min", "k_max", "prior_mode", "nburn",
                     "niter", "ncorr", "nchains", "ncool", "nbin_z", "nbin_vs", "nbin_vp", "nbin_vpvs",
                     "nbin_sig", "nbin_amp"):
            setattr(c, name, int(getattr(self, name)))
        c.iseed = C.c_int32(int(self.iseed) & 0xFFFFFFFF).value
        for name in ("delta", "t_start", "sdep", "bdep", "z_ref_min", "dz_ref", "z_min", "z_max", "h_min", "dvs_prior",
                     "dvp_prior", "vp_min", "vp_max", "vs_min", "vs_max", "vpvs_min", "vpvs_max", "dev_z",
                     "dev_dvs", "dev_dvp", "dev_sig", "t_high", "amp_min", "amp_max"):
            setattr(c, name, float(getattr(self, name)))
        c.rayps = _ptr(arr(self.rayps), C.c_double)
        c.a_gus = _ptr(arr(self.a_gus), C.c_double)
        c.ipha = _ptr(arr(self.ipha, np.int32), C.c_int32)
        c.obs = _ptr(arr(self.obs) if self.obs is not None else None, C.c_double)
        c.r_inv = _ptr(arr(self.r_inv) if self.r_inv is not None else None, C.c_double)
        vp = arr(self.vp_ref)
        c.nref = int(vp.shape[0])
        c.vp_ref = _ptr(vp, C.c_double)
        c.vs_ref = _ptr(arr(self.vs_ref), C.c_double)
        c.sig_min = _ptr(arr(self.sig_min), C.c_double)
        c.sig_max = _ptr(arr(self.sig_max), C.c_double)
        self._keep = keep
        return c
