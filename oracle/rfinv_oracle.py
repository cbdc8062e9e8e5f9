"""CPU oracle (numpy) for RF_INV's forward-model + likelihood path and its PT-MCMC caller.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rf_inv_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU arm.

It restates the reference's algorithm *as written* (dense complex 4x4 propagators, libm sin/cos per
layer-frequency, dense S x S quadratic form), including the single-precision literals that leak
into the reference's fp64 results.  Every function cites the reference file:line it follows
(paths relative to the reference checkout).

Third-party arithmetic the reference links but does not vendor (Makefile:18-19, versions unpinned):
  * FFTW3 ``dfftw_plan_dft_c2r_1d`` / ``dfftw_execute``  -> ``numpy.fft.irfft(x, n) * n``
    (unnormalised inverse real FFT, sign +i; imaginary parts of DC and Nyquist ignored).
  * LAPACK ``dgesvd('A','A')``                           -> ``scipy.linalg.svd(lapack_driver='gesvd')``.

Parity status: the forward path (land, P phase, deconv_mode 0) is pinned by the reference's own
fixtures ``sample_syn/data/sample_{1,2}.trc`` (float32, ~6e-8 abs) and ``vp_to_rho(5.0)`` is pinned
bit-exactly by ``sample_syn/true/true.velmod`` (see tests/test_oracle_golden.py).  Everything else
(sea layer, S phase, deconvolution, likelihood value, RNG output, accept/reject sequences) has no
fixture in the reference and no Fortran compiler exists here: for those **parity unpinned** -- they
are checked numpy-oracle vs C-oracle vs CUDA plus analytic consistency tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

PI = 3.1415926535897931  # forward.f90:33
F32 = np.float32

# single-precision literals that leak into fp64 arithmetic in the reference
OMG_DC = float(F32(1.0e-5))          # forward.f90:247
V1_TINY = float(F32(1.0e-16))        # math.f90:44
COLD_EPS = float(F32(1.0e-6))        # pt_mcmc.f90:196,204
SIG_MODE_EPS = float(F32(1.0e-5))    # params.f90:262
BROCHER = [float(F32(c)) for c in (1.6612, 0.4721, 0.0671, 0.0043, 0.000106)]  # model.f90:308-309


def nint(x: float) -> int:
    """Fortran NINT: round half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0.0 else -int(math.floor(-x + 0.5))


# --------------------------------------------------------------------------------------
# configuration (the reference's module globals: params.f90:50-96, model.f90:35-36)
# --------------------------------------------------------------------------------------
@dataclass
class Config:
    # observation / RF synthesis
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
    bdep: float = 0.0                          # station depth below the surface / sea floor (BOREHOLE_DEP, commented out
                                               # in the reference: params.f90:67,203-224); 0 = station at the surface
    bdep_literal: bool = False                 # transcribe the commented block as it stands (see calc_seis)
    obs: Optional[np.ndarray] = None          # (nsmp, ntrc) float64 (float32 values promoted)
    # reference velocity model
    vp_ref: Optional[np.ndarray] = None
    vs_ref: Optional[np.ndarray] = None
    z_ref_min: float = 0.0
    dz_ref: float = 0.5
    # prior / validity
    vp_mode: int = 0
    k_min: int = 1
    k_max: int = 10
    z_min: float = 0.0
    z_max: float = 20.0
    h_min: float = 0.05
    prior_mode: int = 2
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
    # proposals
    dev_z: float = 0.02
    dev_dvs: float = 0.02
    dev_dvp: float = 0.02
    dev_sig: float = 0.002
    # PT
    nburn: int = 0
    niter: int = 0
    ncorr: int = 1
    nchains: int = 1
    ncool: int = 1
    t_high: float = 1.0
    iseed: int = 1

    @property
    def nh(self) -> int:
        return self.nfft // 2 + 1

    @property
    def sig_mode(self) -> List[int]:
        # params.f90:262
        return [1 if (self.sig_max[t] - self.sig_min[t]) > SIG_MODE_EPS else 0 for t in range(self.ntrc)]

    @property
    def is_ray_common(self) -> bool:
        # forward.f90:59-76
        return all(self.rayps[t] == self.rayps[0] and self.ipha[t] == self.ipha[0] for t in range(self.ntrc))


# --------------------------------------------------------------------------------------
# model.f90
# --------------------------------------------------------------------------------------
def vp_to_rho(a1: float) -> float:
    """model.f90:298-314 (Brocher 2005 quintic, float32 coefficients promoted to fp64)."""
    a2 = a1 * a1
    a3 = a2 * a1
    a4 = a3 * a1
    a5 = a4 * a1
    return BROCHER[0] * a1 - BROCHER[1] * a2 + BROCHER[2] * a3 - BROCHER[3] * a4 + BROCHER[4] * a5


def format_model(cfg: Config, k: int, z: np.ndarray, dvp: np.ndarray, dvs: np.ndarray):
    """model.f90:175-290.  Returns (nlay, alpha, beta, rho, h, is_valid)."""
    is_valid = True
    order = np.argsort(np.asarray(z[:k]), kind="stable")           # sort.f90:34-68 (keys distinct)
    tz = np.asarray(z[:k], dtype=np.float64)[order]
    tdvp = np.asarray(dvp[:k], dtype=np.float64)[order]
    tdvs = np.asarray(dvs[:k], dtype=np.float64)[order]
    alpha: List[float] = []
    beta: List[float] = []
    rho: List[float] = []
    h: List[float] = []
    sdep = cfg.sdep
    if sdep > 0.0:                                                  # model.f90:201-207
        alpha.append(1.5); beta.append(-999.0); rho.append(1.0); h.append(sdep)

    def layer(zc, d_vs, d_vp):
        nonlocal is_valid
        iz = nint((zc - cfg.z_ref_min) / cfg.dz_ref) + 1            # model.f90:212
        b = cfg.vs_ref[iz - 1] + d_vs
        a = cfg.vp_ref[iz - 1] + d_vp if cfg.vp_mode == 1 else cfg.vp_ref[iz - 1]
        if (a < cfg.vp_min or a > cfg.vp_max or b < cfg.vs_min or b > cfg.vs_max
                or a / b < cfg.vpvs_min or a / b > cfg.vpvs_max):   # model.f90:219-224
            is_valid = False
        return a, b

    # top layer, model.f90:210-231
    a, b = layer(0.5 * (sdep + tz[0]), tdvs[0], tdvp[0])
    alpha.append(a); beta.append(b); rho.append(vp_to_rho(a)); h.append(tz[0] - sdep)
    if h[-1] < 0.125 * a:                                           # model.f90:229 (not h_min)
        is_valid = False
    # middle layers, model.f90:235-262
    for j in range(1, k):
        a, b = layer(0.5 * (tz[j] + tz[j - 1]), tdvs[j], tdvp[j])
        alpha.append(a); beta.append(b); rho.append(vp_to_rho(a)); h.append(tz[j] - tz[j - 1])
        if h[-1] < cfg.h_min:
            is_valid = False
    # half space, model.f90:264-283
    a, b = layer(0.5 * (cfg.z_max + tz[k - 1]), dvs[cfg.k_max - 1], dvp[cfg.k_max - 1])
    alpha.append(a); beta.append(b); rho.append(vp_to_rho(a)); h.append(999.0)
    return (len(alpha), np.array(alpha), np.array(beta), np.array(rho), np.array(h), is_valid)


# --------------------------------------------------------------------------------------
# forward.f90
# --------------------------------------------------------------------------------------
def init_filter(cfg: Config) -> np.ndarray:
    """forward.f90:95-119 -> flt(nh, ntrc)."""
    nh = cfg.nh
    df = 1.0 / (cfg.delta * cfg.nfft)
    flt = np.empty((nh, cfg.ntrc))
    for t in range(cfg.ntrc):
        fac_norm = cfg.nfft * cfg.a_gus[t] * cfg.delta / math.sqrt(PI)
        for i in range(nh):
            omega = i * 2.0 * PI * df
            flt[i, t] = math.exp(-(omega / (2.0 * cfg.a_gus[t])) ** 2) / fac_norm
    return flt


def e_inverse(omega: np.ndarray, rho, alpha, beta, p) -> np.ndarray:
    """forward.f90:350-380, vectorised over omega -> (nw,4,4) complex."""
    ei = 1j
    nw = omega.shape[0]
    e = np.zeros((nw, 4, 4), dtype=np.complex128)
    with np.errstate(invalid="ignore"):
        eta = np.sqrt(np.float64(1.0 / (beta * beta) - p * p))
        xi = np.sqrt(np.float64(1.0 / (alpha * alpha) - p * p))
    bp = 1.0 - 2.0 * beta * beta * p * p
    e[:, 0, 0] = beta * beta * p / alpha
    e[:, 0, 1] = bp / (2.0 * alpha * xi)
    e[:, 0, 2] = -p / (2.0 * omega * rho * alpha * xi) * ei
    e[:, 0, 3] = -1.0 / (2.0 * omega * rho * alpha) * ei
    e[:, 1, 0] = bp / (2.0 * beta * eta)
    e[:, 1, 1] = -beta * p
    e[:, 1, 2] = -1.0 / (2.0 * omega * rho * beta) * ei
    e[:, 1, 3] = p / (2.0 * omega * rho * beta * eta) * ei
    e[:, 2, 0] = e[:, 0, 0]
    e[:, 2, 1] = -e[:, 0, 1]
    e[:, 2, 2] = -e[:, 0, 2]
    e[:, 2, 3] = e[:, 0, 3]
    e[:, 3, 0] = e[:, 1, 0]
    e[:, 3, 1] = -e[:, 1, 1]
    e[:, 3, 2] = -e[:, 1, 2]
    e[:, 3, 3] = e[:, 1, 3]
    return e


def layer_matrix_sol(omega: np.ndarray, rho, alpha, beta, p, z) -> np.ndarray:
    """forward.f90:385-421, vectorised over omega -> (nw,4,4) complex."""
    ei = 1j
    nw = omega.shape[0]
    pm = np.empty((nw, 4, 4), dtype=np.complex128)
    beta2 = beta * beta
    p2 = p * p
    bp = 1.0 - 2.0 * beta2 * p2
    with np.errstate(invalid="ignore"):
        eta = np.sqrt(np.float64(1.0 / beta2 - p2))
        xi = np.sqrt(np.float64(1.0 / (alpha * alpha) - p2))
    cos_xi = np.cos(omega * xi * z)
    cos_eta = np.cos(omega * eta * z)
    sin_xi = np.sin(omega * xi * z)
    sin_eta = np.sin(omega * eta * z)
    pm[:, 0, 0] = 2.0 * beta2 * p2 * cos_xi + bp * cos_eta
    pm[:, 1, 0] = p * (2.0 * beta2 * xi * sin_xi - bp / eta * sin_eta) * ei
    pm[:, 2, 0] = omega * rho * (-4.0 * beta2 * beta2 * p2 * xi * sin_xi - bp * bp / eta * sin_eta)
    pm[:, 3, 0] = 2.0 * omega * beta2 * rho * p * bp * (cos_xi - cos_eta) * ei
    pm[:, 0, 1] = p * (bp / xi * sin_xi - 2.0 * beta2 * eta * sin_eta) * ei
    pm[:, 1, 1] = bp * cos_xi + 2.0 * beta2 * p2 * cos_eta
    pm[:, 2, 1] = pm[:, 3, 0]
    pm[:, 3, 1] = -omega * rho * (bp * bp / xi * sin_xi + 4.0 * beta2 * beta2 * p2 * eta * sin_eta)
    pm[:, 0, 2] = (p2 / xi * sin_xi + eta * sin_eta) / (omega * rho)
    pm[:, 1, 2] = p * (-cos_xi + cos_eta) / (omega * rho) * ei
    pm[:, 2, 2] = pm[:, 0, 0]
    pm[:, 3, 2] = pm[:, 0, 1]
    pm[:, 0, 3] = pm[:, 1, 2]
    pm[:, 1, 3] = (xi * sin_xi + p2 / eta * sin_eta) / (omega * rho)
    pm[:, 2, 3] = pm[:, 1, 0]
    pm[:, 3, 3] = pm[:, 1, 1]
    return pm


def layer_matrix_liq(omega: np.ndarray, rho, alpha, p, z) -> np.ndarray:
    """forward.f90:424-442 -> (nw,2,2) complex."""
    nw = omega.shape[0]
    lq = np.empty((nw, 2, 2), dtype=np.complex128)
    xi = math.sqrt(1.0 / (alpha * alpha) - p * p)
    cos_xi = np.cos(omega * xi * z)
    sin_xi = np.sin(omega * xi * z)
    g = rho * omega / xi
    lq[:, 0, 0] = cos_xi
    lq[:, 0, 1] = sin_xi / g
    lq[:, 1, 0] = -g * sin_xi
    lq[:, 1, 1] = cos_xi
    return lq


def calc_seis(cfg: Config, nlay, rayp, ipha, alpha, beta, rho, h):
    """forward.f90:212-344 -> (ur_freq[nh], uz_freq[nh]) (elements nh+1..n are never written)."""
    npts = cfg.nfft
    nh = npts // 2 + 1
    sea_flag = beta[0] < 0
    ilay0 = 1 if sea_flag else 0
    domg = 2.0 * PI / (npts * cfg.delta)
    omg = np.arange(nh, dtype=np.float64) * domg
    omg[0] = OMG_DC                                                   # forward.f90:246-248
    e_inv = e_inverse(omg, rho[nlay - 1], alpha[nlay - 1], beta[nlay - 1], rayp)
    p_prod = np.zeros((nh, 4, 4), dtype=np.complex128)
    for j in range(4):
        p_prod[:, j, j] = 1.0
    for ilay in range(ilay0, nlay - 1):
        p_mat = layer_matrix_sol(omg, rho[ilay], alpha[ilay], beta[ilay], rayp, h[ilay])
        p_prod = np.matmul(p_mat, p_prod)                             # forward.f90:262
    sl = np.matmul(e_inv, p_prod)                                     # forward.f90:264
    with np.errstate(invalid="ignore", divide="ignore"):
        if not sea_flag:                                              # forward.f90:267-275
            denom = sl[:, 2, 0] * sl[:, 3, 1] - sl[:, 2, 1] * sl[:, 3, 0]
            if ipha >= 0:
                ur = sl[:, 3, 1] / denom
                uz = -sl[:, 3, 0] / denom
            else:
                ur = -sl[:, 2, 1] / denom
                uz = sl[:, 2, 0] / denom
        else:                                                         # forward.f90:276-287
            lq = layer_matrix_liq(omg, rho[0], alpha[0], rayp, h[0])
            a = sl[:, 3, 1] * lq[:, 0, 0] + sl[:, 3, 3] * lq[:, 1, 0]
            b = sl[:, 2, 1] * lq[:, 0, 0] + sl[:, 2, 3] * lq[:, 1, 0]
            if ipha >= 0:
                ur = a / (a * sl[:, 2, 0] - b * sl[:, 3, 0])
                uz = lq[:, 0, 0] * sl[:, 3, 0] / (b * sl[:, 3, 0] - a * sl[:, 2, 0])
            else:
                ur = -b / (a * sl[:, 2, 0] - b * sl[:, 3, 0])
                uz = -lq[:, 0, 0] * sl[:, 2, 0] / (b * sl[:, 3, 0] - a * sl[:, 2, 0])
        if cfg.bdep > 0.0:
            # Buried station ("case of bore hole"), forward.f90:289-338 -- COMMENTED OUT in the reference.  The
            # displacement-stress vector at the surface / sea floor is carried down to the station depth.
            sd = np.zeros((nh, 4, 1), dtype=np.complex128)
            sd[:, 0, 0] = ur
            sd[:, 1, 0] = uz
            if sea_flag:                                              # forward.f90:297-306: normal stress under the water
                den = b * sl[:, 3, 0] - a * sl[:, 2, 0]
                sd[:, 3, 0] = lq[:, 1, 0] * sl[:, 3, 0] / den if ipha >= 0 else -lq[:, 1, 0] * sl[:, 2, 0] / den
            z_tmp = 0.0
            found = False
            for ilay in range(ilay0, nlay - 1):                       # forward.f90:311-327
                z_tmp = z_tmp + h[ilay]
                if z_tmp < cfg.bdep:
                    sd = np.matmul(layer_matrix_sol(omg, rho[ilay], alpha[ilay], beta[ilay], rayp, h[ilay]), sd)
                else:
                    h_tmp = cfg.bdep + h[ilay] - z_tmp
                    sd = np.matmul(layer_matrix_sol(omg, rho[ilay], alpha[ilay], beta[ilay], rayp, h_tmp), sd)
                    found = True
                    if not cfg.bdep_literal:
                        break      # the commented block has no exit here: it goes on through every deeper layer with the
                                   # negative thickness bdep + h - z_tmp
            if not found:                                             # forward.f90:329-334: station in the half space
                # the commented block propagates by bdep; the distance left below the last interface is bdep - z_tmp
                d = cfg.bdep if cfg.bdep_literal else cfg.bdep - z_tmp
                sd = np.matmul(layer_matrix_sol(omg, rho[nlay - 1], alpha[nlay - 1], beta[nlay - 1], rayp, d), sd)
            ur = sd[:, 0, 0]
            uz = sd[:, 1, 0]
    return ur, uz


def water_level_decon(y, x, pcnt):
    """forward.f90:447-470: z = y conj(x) / max(|x|^2, pcnt*max|x|^2)."""
    amp = (x * np.conj(x)).real
    wlvl = pcnt * np.max(amp)
    return y * np.conj(x) / np.maximum(amp, wlvl)


def direct_arrival(cfg: Config, nlay, h, v, rayp) -> float:
    """forward.f90:474-491; buried station: the commented variant forward.f90:493-516 (delay from the station, not
    from the surface, to the top of the half space)."""
    i0 = 1 if cfg.sdep > 0.0 else 0
    t = 0.0
    if cfg.bdep <= 0.0:
        for i in range(i0, nlay - 1):
            t = t + h[i] * math.sqrt(1.0 / (v[i] * v[i]) - rayp * rayp)
        return t
    z_sum = 0.0
    if i0 < nlay - 1:                                                 # forward.f90:493-513 (0-based: nlay-1 is the half space)
        below = 0.0
        while True:                                                   # layer with the station
            if i0 == nlay - 1:
                below = z_sum
                break
            z_sum = z_sum + h[i0]
            if z_sum > cfg.bdep:
                t = t + (z_sum - cfg.bdep) * math.sqrt(1.0 / (v[i0] * v[i0]) - rayp * rayp)
                i0 = i0 + 1
                below = None
                break
            i0 = i0 + 1
        if below is None:
            vi0 = v[i0] if i0 < nlay else v[nlay - 1]
            for i in range(i0, nlay - 1):                             # other layers
                # the commented block takes v(i0) for every one of them (forward.f90:512); the layer's own velocity is meant
                vv = vi0 if cfg.bdep_literal else v[i]
                t = t + h[i] * math.sqrt(1.0 / (vv * vv) - rayp * rayp)
        elif not cfg.bdep_literal:
            # the loop ran out of layers: the station lies in the half space, bdep - z_sum below its top (the commented
            # block adds nothing in this case; its `else` branch below only covers models without any solid layer)
            t = t - (cfg.bdep - below) * math.sqrt(1.0 / (v[nlay - 1] * v[nlay - 1]) - rayp * rayp)
    else:                                                             # forward.f90:514-516
        t = t - cfg.bdep * math.sqrt(1.0 / (v[i0] * v[i0]) - rayp * rayp)
    return t


def c2r(cx_half: np.ndarray, n: int) -> np.ndarray:
    """Stand-in for FFTW c2r (fftw.f90:44, forward.f90:172,200): unnormalised, sign +i."""
    return np.fft.irfft(cx_half, n) * n


def calc_rf(cfg: Config, flt: np.ndarray, nlay, alpha, beta, rho, h) -> np.ndarray:
    """forward.f90:123-208 -> rft(nfft, ntrc)."""
    n = cfg.nfft
    nh = cfg.nh
    rft = np.empty((n, cfg.ntrc))
    common = cfg.is_ray_common
    freq_r = freq_v = rff = None
    tp = 0.0
    for t in range(cfg.ntrc):
        ipha = cfg.ipha[t]
        if t == 0 or not common:
            ur, uz = calc_seis(cfg, nlay, cfg.rayps[t], ipha, alpha, beta, rho, h)
            freq_r = np.conj(ur)
            freq_v = -np.conj(uz)
            if cfg.deconv_mode == 1 and ipha == 1:
                rff = water_level_decon(freq_r, freq_v, 0.001)
                tp = 0.0
            elif cfg.deconv_mode == 1 and ipha == -1:
                rff = water_level_decon(freq_v, freq_r, 0.001)
                tp = 0.0
            else:
                if ipha == 1:
                    rff = freq_r
                    tp = direct_arrival(cfg, nlay, h, alpha, cfg.rayps[t])
                else:
                    rff = freq_v
                    tp = direct_arrival(cfg, nlay, h, beta, cfg.rayps[t])
        rx = c2r(rff[:nh] * flt[:, t], n)                             # forward.f90:168-172
        i1 = np.arange(1, n + 1)
        if ipha == 1:                                                 # forward.f90:176-184
            npre = nint((-cfg.t_start - tp) / cfg.delta)
            j = np.mod(n - npre + i1, n)
            j[j == 0] = n
            rft[:, t] = rx[j - 1]
        else:                                                         # forward.f90:185-194
            npre = nint((-cfg.t_start + tp) / cfg.delta)
            j = np.mod(n + npre - i1 + 1, n)
            j[j == 0] = n
            rft[:, t] = -rx[j - 1]
        if cfg.deconv_mode == 0:                                      # forward.f90:197-203
            rxv = c2r(freq_v[:nh] * flt[:, t], n)
            fac_norm = np.max(rxv)
            rft[:, t] = rft[:, t] / fac_norm
    return rft


# --------------------------------------------------------------------------------------
# likelihood.f90
# --------------------------------------------------------------------------------------
def init_r_inv(cfg: Config) -> np.ndarray:
    """likelihood.f90:168-241 -> r_inv(nsmp, nsmp, ntrc); dgesvd stand-in = scipy gesvd."""
    import scipy.linalg

    s_n = cfg.nsmp
    out = np.empty((s_n, s_n, cfg.ntrc))
    idx = np.arange(s_n)
    d2 = (idx[:, None] - idx[None, :]) ** 2
    for t in range(cfg.ntrc):
        r = math.exp(-cfg.a_gus[t] ** 2 * cfg.delta ** 2)
        r_mat = np.power(r, d2.astype(np.float64))
        u, s, vt = scipy.linalg.svd(r_mat, full_matrices=True, lapack_driver="gesvd")
        dinv = np.where(s > 1.0e-3, 1.0 / s, 0.0)
        out[:, :, t] = (vt.T * dinv[None, :]) @ u.T                   # likelihood.f90:222-223
    return out


def log_likelihood_from_rft(cfg: Config, r_inv: np.ndarray, rft: np.ndarray, sig: Sequence[float]) -> float:
    """likelihood.f90:85-98."""
    s_n = cfg.nsmp
    ll = 0.0
    for t in range(cfg.ntrc):
        misfits = rft[:s_n, t] - cfg.obs[:s_n, t]
        s = sig[t]
        phi1 = misfits @ r_inv[:, :, t]
        phi = float(phi1 @ misfits)
        ll = ll - 0.5 * phi / (s * s) - float(s_n) * math.log(s)
    return ll


def calc_likelihood(cfg: Config, flt, r_inv, k, z, dvp, dvs, sig, fwd_flag=True, cached_rft=None):
    """likelihood.f90:56-101 -> (log_likelihood, rft).  Ignores is_valid like the reference."""
    if fwd_flag:
        nlay, alpha, beta, rho, h, _ = format_model(cfg, k, z, dvp, dvs)
        rft = calc_rf(cfg, flt, nlay, alpha, beta, rho, h)
    else:
        rft = cached_rft.copy()
    return log_likelihood_from_rft(cfg, r_inv, rft, sig), rft


# --------------------------------------------------------------------------------------
# mt19937.f90, math.f90, prior.f90
# --------------------------------------------------------------------------------------
class MT19937:
    """mt19937.f90:78-130: 1997 seeding (mt[i] = 69069*mt[i-1] mod 2^32), output y / 2^32 in [0,1)."""

    N, M = 624, 397

    def __init__(self, seed: int):
        mt = [0] * self.N
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, self.N):
            mt[i] = (69069 * mt[i - 1]) & 0xFFFFFFFF
        self.mt = mt
        self.mti = self.N
        self.ndraw = 0

    def _reload(self):
        mt, n, m = self.mt, self.N, self.M
        for kk in range(n):
            y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % n] & 0x7FFFFFFF)
            mt[kk] = mt[(kk + m) % n] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        self.mti = 0

    def grnd(self) -> float:
        if self.mti >= self.N:
            self._reload()
        y = self.mt[self.mti]
        self.mti += 1
        self.ndraw += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y / 4294967296.0


def gauss(rng: MT19937) -> float:
    """math.f90:34-50: Box-Muller, cosine branch only, exactly two draws."""
    v1 = rng.grnd()
    v2 = rng.grnd()
    if v1 == 0.0:
        v1 = V1_TINY
    return math.sqrt(-2.0 * math.log(v1)) * math.cos(2.0 * PI * v2)


def laplace(rng: MT19937) -> float:
    """prior.f90:57-123 (von Neumann / Forsythe sampler, data-dependent draw count)."""
    d = 0.69314718055994529
    u1 = rng.grnd()
    u1p = 2.0 * u1
    if u1p < 1.0:
        i_sign = 1
        u1pp = 1.0 - u1p
    else:
        i_sign = -1
        u1pp = 2.0 - u1p
    a = 0.0
    while True:
        u1ppp = 2.0 * u1pp
        if u1ppp >= 1.0:
            u1 = u1ppp - 1.0
            break
        a = a + d
        u1pp = u1ppp
    while True:
        w = d * u1
        val = i_sign * (a + w)
        k = 1
        while True:
            u2 = rng.grnd()
            if u2 >= w:
                u1 = (u2 - w) / (1.0 - w)
                break
            w = u2
            k += 1
        if k % 2 == 1:
            break
    return val


def log_prior_ratio(x_new, x_old, dev, prior_mode) -> float:
    """prior.f90:36-53."""
    if prior_mode == 1:
        return -(abs(x_new) - abs(x_old)) / dev
    return -((x_new * x_new) - (x_old * x_old)) / (2.0 * dev * dev)


# --------------------------------------------------------------------------------------
# PT-MCMC with "virtual ranks" (one MT stream per group of nchains chains = one MPI rank)
# --------------------------------------------------------------------------------------
@dataclass
class RankState:
    """State of one MPI rank of the reference (rf_inv.f90:75-91)."""
    rng: MT19937
    k: np.ndarray
    z: np.ndarray        # (k_max-1, nchains)
    dvp: np.ndarray      # (k_max, nchains)
    dvs: np.ndarray
    sig: np.ndarray      # (ntrc, nchains)
    rft: np.ndarray = None   # (nfft, ntrc, nchains)
    logl: np.ndarray = None
    temps: np.ndarray = None


@dataclass
class PTResult:
    nprop: np.ndarray
    naccept: np.ndarray
    accept_bits: List[List[int]] = field(default_factory=list)   # per iteration: per global chain (-1 null, 0 rej, 1 acc)
    itypes: List[List[int]] = field(default_factory=list)
    likelihood_hist: np.ndarray = None
    swaps: List[tuple] = field(default_factory=list)             # (itarget1, itarget2, accepted)


def rank_seed(iseed: int, rank: int) -> int:
    """rf_inv.f90:75 (default-integer arithmetic wraps at 32 bit)."""
    s = (iseed + rank * rank * 10000 + 23 * rank) & 0xFFFFFFFF
    return s


def init_model(cfg: Config, rng: MT19937):
    """model.f90:43-107."""
    nc, km = cfg.nchains, cfg.k_max
    z = np.zeros((km - 1, nc)); dvp = np.zeros((km, nc)); dvs = np.zeros((km, nc))
    k = np.zeros(nc, dtype=np.int32)
    draw = laplace if cfg.prior_mode == 1 else gauss
    for c in range(nc):
        valid = False
        while not valid:
            k[c] = cfg.k_min + int(rng.grnd() * (cfg.k_max - cfg.k_min))
            for i in range(k[c]):
                z[i, c] = cfg.z_min + rng.grnd() * (cfg.z_max - cfg.z_min)
            for i in range(k[c]):
                dvs[i, c] = draw(rng) * cfg.dvs_prior
                dvp[i, c] = draw(rng) * cfg.dvp_prior
            dvs[km - 1, c] = draw(rng) * cfg.dvs_prior
            dvp[km - 1, c] = draw(rng) * cfg.dvp_prior
            valid = format_model(cfg, int(k[c]), z[:, c], dvp[:, c], dvs[:, c])[5]
    return k, z, dvp, dvs


def init_sig(cfg: Config, rng: MT19937) -> np.ndarray:
    """likelihood.f90:107-139."""
    sig = np.empty((cfg.ntrc, cfg.nchains))
    mode = cfg.sig_mode
    for c in range(cfg.nchains):
        for t in range(cfg.ntrc):
            if mode[t] == 1:
                sig[t, c] = cfg.sig_min[t] + rng.grnd() * (cfg.sig_max[t] - cfg.sig_min[t])
            else:
                sig[t, c] = cfg.sig_min[t]
    return sig


def proposal_types(cfg: Config):
    """pt_mcmc.f90:311-365 -> dict of 1-based itype codes (-1 = absent) and isig_trc."""
    t = dict(birth=1, death=2, z=3, dvs=4, dvp=-1, sig=-1)
    ntype = 4
    if cfg.vp_mode == 1:
        ntype += 1
        t["dvp"] = ntype
    isig = [i for i, m in enumerate(cfg.sig_mode) if m == 1]
    if isig:
        ntype += 1
        t["sig"] = ntype
    t["ntype"] = ntype
    t["isig_trc"] = isig
    return t


def init_rank(cfg: Config, flt, r_inv, rank: int) -> RankState:
    """rf_inv.f90:75-91 init order: sgrnd, init_model, init_sig, init_rft, temps."""
    rng = MT19937(rank_seed(cfg.iseed, rank))
    k, z, dvp, dvs = init_model(cfg, rng)
    sig = init_sig(cfg, rng)
    st = RankState(rng=rng, k=k, z=z, dvp=dvp, dvs=dvs, sig=sig)
    st.rft = np.empty((cfg.nfft, cfg.ntrc, cfg.nchains))
    st.logl = np.empty(cfg.nchains)
    for c in range(cfg.nchains):                                     # likelihood.f90:156-160
        st.logl[c], st.rft[:, :, c] = calc_likelihood(cfg, flt, r_inv, int(k[c]), z[:, c], dvp[:, c], dvs[:, c], sig[:, c])
    st.temps = np.ones(cfg.nchains)                                  # pt_mcmc.f90:447-452
    for c in range(cfg.ncool, cfg.nchains):
        st.temps[c] = math.exp(rng.grnd() * math.log(cfg.t_high))
    return st


def mcmc_step(cfg: Config, flt, r_inv, st: RankState, types, c: int, temp: float):
    """pt_mcmc.f90:54-201 for one chain.  Returns (itype, flag) flag: -1 null, 0 rejected, 1 accepted."""
    rng = st.rng
    km = cfg.k_max
    log_prior12 = 0.0
    prop_k = int(st.k[c])
    prop_dvp = st.dvp[:, c].copy()
    prop_dvs = st.dvs[:, c].copy()
    prop_z = np.zeros(km); prop_z[:km - 1] = st.z[:, c]
    prop_sig = st.sig[:, c].copy()
    null_flag = False
    draw = laplace if cfg.prior_mode == 1 else gauss
    itype = int(rng.grnd() * types["ntype"]) + 1
    if itype == types["birth"]:
        prop_k += 1
        if prop_k < km:
            prop_dvp[prop_k - 1] = draw(rng) * cfg.dvp_prior
            prop_dvs[prop_k - 1] = draw(rng) * cfg.dvs_prior
            prop_z[prop_k - 1] = cfg.z_min + rng.grnd() * (cfg.z_max - cfg.z_min)
        else:
            null_flag = True
    elif itype == types["death"]:
        prop_k -= 1
        if prop_k >= cfg.k_min:
            itarget = int(rng.grnd() * (prop_k + 1)) + 1
            for il in range(itarget, prop_k + 1):
                prop_dvp[il - 1] = st.dvp[il, c]
                prop_dvs[il - 1] = st.dvs[il, c]
                prop_z[il - 1] = st.z[il, c]
            prop_dvp[prop_k] = 0.0
            prop_dvs[prop_k] = 0.0
            prop_z[prop_k] = 0.0
        else:
            null_flag = True
    elif itype == types["z"]:
        itarget = int(rng.grnd() * prop_k) + 1
        prop_z[itarget - 1] = prop_z[itarget - 1] + gauss(rng) * cfg.dev_z
        if prop_z[itarget - 1] < cfg.z_min or prop_z[itarget - 1] > cfg.z_max:
            null_flag = True
    elif itype == types["dvs"]:
        itarget = int(rng.grnd() * (prop_k + 1)) + 1
        if itarget == prop_k + 1:
            itarget = km
        prop_dvs[itarget - 1] = prop_dvs[itarget - 1] + gauss(rng) * cfg.dev_dvs
        log_prior12 = log_prior_ratio(prop_dvs[itarget - 1], st.dvs[itarget - 1, c], cfg.dvs_prior, cfg.prior_mode)
    elif itype == types["dvp"]:
        itarget = int(rng.grnd() * (prop_k + 1)) + 1
        if itarget == prop_k + 1:
            itarget = km
        prop_dvp[itarget - 1] = prop_dvp[itarget - 1] + gauss(rng) * cfg.dev_dvp
        log_prior12 = log_prior_ratio(prop_dvp[itarget - 1], st.dvp[itarget - 1, c], cfg.dvp_prior, cfg.prior_mode)
    elif itype == types["sig"]:
        isig = types["isig_trc"]
        itarget = isig[int(rng.grnd() * len(isig))]
        prop_sig[itarget] = prop_sig[itarget] + gauss(rng) * cfg.dev_sig
        if prop_sig[itarget] < cfg.sig_min[itarget] or prop_sig[itarget] > cfg.sig_max[itarget]:
            null_flag = True
    if not null_flag:
        if not format_model(cfg, prop_k, prop_z[:km - 1], prop_dvp, prop_dvs)[5]:
            null_flag = True
    flag = -1
    if not null_flag:
        fwd_flag = itype != types["sig"]
        ll2, prop_rft = calc_likelihood(cfg, flt, r_inv, prop_k, prop_z[:km - 1], prop_dvp, prop_dvs, prop_sig,
                                        fwd_flag=fwd_flag, cached_rft=st.rft[:, :, c])
        # judge_mcmc, pt_mcmc.f90:600-621
        del_s = (ll2 - st.logl[c]) / temp + log_prior12
        while True:
            r = rng.grnd()
            if r >= np.finfo(np.float64).eps:
                break
        yn = math.log(r) <= del_s
        flag = 1 if yn else 0
        if yn:
            st.logl[c] = ll2
            st.k[c] = prop_k
            st.dvp[:, c] = prop_dvp
            st.dvs[:, c] = prop_dvs
            st.z[:, c] = prop_z[:km - 1]
            st.sig[:, c] = prop_sig
            st.rft[:, :, c] = prop_rft
    return itype, flag


def pt_run(cfg: Config, flt, r_inv, nproc: int, n_tot_iter: Optional[int] = None, ranks: Optional[List[RankState]] = None):
    """pt_mcmc.f90:468-576 for ``nproc`` virtual ranks run in one process; returns (ranks, PTResult)."""
    types = proposal_types(cfg)
    if ranks is None:
        ranks = [init_rank(cfg, flt, r_inv, r) for r in range(nproc)]
    n_all = nproc * cfg.nchains
    n_tot = cfg.nburn + cfg.niter if n_tot_iter is None else n_tot_iter
    res = PTResult(nprop=np.zeros(types["ntype"], dtype=np.int64), naccept=np.zeros(types["ntype"], dtype=np.int64))
    res.likelihood_hist = np.zeros(n_tot)
    for it in range(1, n_tot + 1):
        bits, tys = [], []
        for st in ranks:
            for c in range(cfg.nchains):
                temp = st.temps[c]
                itype, flag = mcmc_step(cfg, flt, r_inv, st, types, c, temp)
                bits.append(flag); tys.append(itype)
                if temp <= 1.0 + COLD_EPS:                         # pt_mcmc.f90:196-201
                    res.nprop[itype - 1] += 1
                    if flag == 1:
                        res.naccept[itype - 1] += 1
                    res.likelihood_hist[it - 1] += st.logl[c]
        res.accept_bits.append(bits); res.itypes.append(tys)
        if cfg.nchains < 2:
            continue
        r0 = ranks[0].rng                                            # pt_mcmc.f90:501-516
        itarget1 = int(r0.grnd() * n_all)
        while True:
            itarget2 = int(r0.grnd() * n_all)
            if itarget2 != itarget1:
                break
        rank1, rank2 = itarget1 // cfg.nchains, itarget2 // cfg.nchains
        c1, c2 = itarget1 % cfg.nchains, itarget2 % cfg.nchains
        s1, s2 = ranks[rank1], ranks[rank2]
        temp1, temp2 = s1.temps[c1], s2.temps[c2]
        e1, e2 = s1.logl[c1], s2.logl[c2]
        del_s = (e2 - e1) * (1.0 / temp1 - 1.0 / temp2)              # judge_pt, pt_mcmc.f90:580-595
        u = s1.rng.grnd()
        yn = (math.log(u) if u > 0.0 else -math.inf) <= del_s
        if yn:
            s2.temps[c2] = temp1
            s1.temps[c1] = temp2
        res.swaps.append((itarget1, itarget2, int(yn)))
    return ranks, res


def make_syn(cfg: Config):
    """make_syn.f90:51-161: one random model per chain from the stream sgrnd(iseed) (no rank offset), init_sig, the
    synthetic RF of chain 1, and per trace a Gaussian white-noise series shaped by the trace's filter
    (r2c -> *flt -> c2r, both transforms unnormalised: make_syn.f90:96-100).  Returns a dict with the model of chain 1,
    its layer stack, noise_sigma[T], the filtered noise[T][nfft], rft[T][nsmp] and the noisy traces.
    The observed data of `cfg` are only used through nsmp / delta, like the reference's read_obs call."""
    rng = MT19937(cfg.iseed & 0xFFFFFFFF)
    k, z, dvp, dvs = init_model(cfg, rng)
    sig = init_sig(cfg, rng)
    flt = init_filter(cfg)
    nlay, alpha, beta, rho, h, is_valid = format_model(cfg, int(k[0]), z[:, 0], dvp[:, 0], dvs[:, 0])
    rft = calc_rf(cfg, flt, nlay, alpha, beta, rho, h)              # [nfft][ntrc]
    n, T = cfg.nfft, cfg.ntrc
    noise = np.zeros((T, n)); noise_sigma = np.zeros(T)
    if cfg.is_ray_common:
        noise_sigma[:] = rng.grnd() * (cfg.sig_max[0] - cfg.sig_min[0]) + cfg.sig_min[0]
        white = np.array([gauss(rng) * noise_sigma[0] for _ in range(n)])
        # make_syn.f90:86-92 as written: the loop loads rx from noise(:,1) and stores the result in noise(:,itrc) -- at
        # itrc = 1 that overwrites noise(:,1), so the traces from the second on are shaped by flt(:,1) AND flt(:,itrc)
        src = white
        for t in range(T):
            noise[t] = c2r(np.fft.rfft(src) * flt[:, t], n)
            if t == 0:
                src = noise[0].copy()
    else:
        for t in range(T):
            noise_sigma[t] = rng.grnd() * (cfg.sig_max[t] - cfg.sig_min[t]) + cfg.sig_min[t]
            white = np.array([gauss(rng) * noise_sigma[t] for _ in range(n)])
            noise[t] = c2r(np.fft.rfft(white) * flt[:, t], n)
    S = cfg.nsmp
    clean = rft[:S, :].T.copy()
    return dict(k=int(k[0]), z=z[:, 0].copy(), dvp=dvp[:, 0].copy(), dvs=dvs[:, 0].copy(), sig=sig[:, 0].copy(), nlay=nlay,
                alpha=alpha[:nlay].copy(), beta=beta[:nlay].copy(), rho=rho[:nlay].copy(), h=h[:nlay].copy(),
                noise_sigma=noise_sigma, noise=noise, rft=clean, noisy=clean + noise[:, :S])

