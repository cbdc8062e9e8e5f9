"""Extended-precision arbiter of the forward + likelihood path (TEST INFRASTRUCTURE, like everything under oracle/).

mpmath evaluation (160-bit mantissa: every rounding error below 1e-45) of the reference's formulas, as written:
  calc_seis   src/forward.f90:212-344   (e_inverse :350-380, layer_matrix_sol :385-421, layer_matrix_liq :424-442)
  calc_rf     src/forward.f90:123-208   (water_level_decon :447-470, direct_arrival :474-491, FFTW c2r = inverse DFT)
  likelihood  src/likelihood.f90:85-98
The INPUTS are the double-precision numbers the fp64 implementations start from -- the formatted layer stack of
format_model (src/model.f90:175-290: sums of doubles, taken as data), the Gaussian filter table of init_filter, R^-1, the
observed traces, the float32 literals that leak into the reference (omega = 1.0e-5 at DC) -- converted exactly; everything
between them and (rft, phi, logL) is evaluated without rounding.  So |oracle - arbiter| is the rounding error of the dense
fp64 restatement and |CUDA - arbiter| that of the CUDA path: where two fp64 evaluations of the same formulas disagree
(ill-conditioned normalisation of S receiver functions), the arbiter says who is closer.

The product E^-1 P_n ... P_1 is carried as a 4x3 matrix (columns 1, 2, 4: the only ones calc_seis reads) -- in exact
arithmetic the order of the multiplications is immaterial.  Buried stations (dead code in the reference) are not covered.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import mpmath as mp
import numpy as np

import rfinv_oracle as pyo

PREC = 160


def _mpf(x) -> mp.mpf:
    return mp.mpf(float(x))          # exact: a double is a binary rational


def _layer_cols(omega, rho, alpha, beta, p, z, cols):
    """layer_matrix_sol (src/forward.f90:385-421) times a 4 x len(cols[0]) complex block, in mp."""
    beta2 = beta * beta
    p2 = p * p
    bp = 1 - 2 * beta2 * p2
    eta = mp.sqrt(1 / beta2 - p2)
    xi = mp.sqrt(1 / (alpha * alpha) - p2)
    cx, ce = mp.cos(omega * xi * z), mp.cos(omega * eta * z)
    sx, se = mp.sin(omega * xi * z), mp.sin(omega * eta * z)
    j = mp.mpc(0, 1)
    m = [[None] * 4 for _ in range(4)]
    m[0][0] = 2 * beta2 * p2 * cx + bp * ce
    m[1][0] = p * (2 * beta2 * xi * sx - bp / eta * se) * j
    m[2][0] = omega * rho * (-4 * beta2 * beta2 * p2 * xi * sx - bp * bp / eta * se)
    m[3][0] = 2 * omega * beta2 * rho * p * bp * (cx - ce) * j
    m[0][1] = p * (bp / xi * sx - 2 * beta2 * eta * se) * j
    m[1][1] = bp * cx + 2 * beta2 * p2 * ce
    m[2][1] = m[3][0]
    m[3][1] = -omega * rho * (bp * bp / xi * sx + 4 * beta2 * beta2 * p2 * eta * se)
    m[0][2] = (p2 / xi * sx + eta * se) / (omega * rho)
    m[1][2] = p * (-cx + ce) / (omega * rho) * j
    m[2][2] = m[0][0]
    m[3][2] = m[0][1]
    m[0][3] = m[1][2]
    m[1][3] = (xi * sx + p2 / eta * se) / (omega * rho)
    m[2][3] = m[1][0]
    m[3][3] = m[1][1]
    nc = len(cols[0])
    return [[sum((m[r][q] * cols[q][c] for q in range(4)), mp.mpc(0)) for c in range(nc)] for r in range(4)]


def _e_inverse_rows34(omega, rho, alpha, beta, p):
    """rows 3, 4 of e_inverse (src/forward.f90:350-380)."""
    j = mp.mpc(0, 1)
    eta = mp.sqrt(1 / (beta * beta) - p * p)
    xi = mp.sqrt(1 / (alpha * alpha) - p * p)
    bp = 1 - 2 * beta * beta * p * p
    e11 = beta * beta * p / alpha
    e12 = bp / (2 * alpha * xi)
    e13 = -p / (2 * omega * rho * alpha * xi) * j
    e14 = -1 / (2 * omega * rho * alpha) * j
    e21 = bp / (2 * beta * eta)
    e22 = -beta * p
    e23 = -1 / (2 * omega * rho * beta) * j
    e24 = p / (2 * omega * rho * beta * eta) * j
    return [[e11, -e12, -e13, e14], [e21, -e22, -e23, e24]]


def calc_seis_mp(cfg: pyo.Config, nlay, rayp, ipha, alpha, beta, rho, h):
    """calc_seis (src/forward.f90:212-344) -> lists ur[nh], uz[nh] of mpc."""
    n = cfg.nfft
    nh = n // 2 + 1
    sea = beta[0] < 0
    ilay0 = 1 if sea else 0
    pi_ref = _mpf(pyo.PI)
    domg = 2 * pi_ref / (n * _mpf(cfg.delta))
    p = _mpf(rayp)
    A = [_mpf(x) for x in alpha[:nlay]]
    B = [_mpf(x) for x in beta[:nlay]]
    R = [_mpf(x) for x in rho[:nlay]]
    H = [_mpf(x) for x in h[:nlay]]
    ur, uz = [], []
    for i in range(nh):
        omega = _mpf(pyo.OMG_DC) if i == 0 else i * domg
        # columns 1, 2, 4 of the identity
        cols = [[mp.mpc(1), mp.mpc(0), mp.mpc(0)], [mp.mpc(0), mp.mpc(1), mp.mpc(0)], [mp.mpc(0), mp.mpc(0), mp.mpc(0)],
                [mp.mpc(0), mp.mpc(0), mp.mpc(1)]]
        for il in range(ilay0, nlay - 1):
            cols = _layer_cols(omega, R[il], A[il], B[il], p, H[il], cols)
        e = _e_inverse_rows34(omega, R[nlay - 1], A[nlay - 1], B[nlay - 1], p)
        sl = [[sum((e[r][q] * cols[q][c] for q in range(4)), mp.mpc(0)) for c in range(3)] for r in range(2)]   # rows 3,4 x cols 1,2,4
        s31, s32, s34 = sl[0]
        s41, s42, s44 = sl[1]
        if not sea:
            den = s31 * s42 - s32 * s41
            if ipha >= 0:
                ur.append(s42 / den); uz.append(-s41 / den)
            else:
                ur.append(-s32 / den); uz.append(s31 / den)
        else:
            xiw = mp.sqrt(1 / (A[0] * A[0]) - p * p)
            cw, sw = mp.cos(omega * xiw * H[0]), mp.sin(omega * xiw * H[0])
            g = R[0] * omega / xiw
            lq00, lq10 = cw, -g * sw
            a = s42 * lq00 + s44 * lq10
            b = s32 * lq00 + s34 * lq10
            if ipha >= 0:
                ur.append(a / (a * s31 - b * s41)); uz.append(lq00 * s41 / (b * s41 - a * s31))
            else:
                ur.append(-b / (a * s31 - b * s41)); uz.append(-lq00 * s31 / (b * s41 - a * s31))
    return ur, uz


def _c2r_mp(xh: Sequence, n: int) -> List:
    """Unnormalised inverse real DFT of the half spectrum xh[0..n/2] (FFTW c2r: the imaginary parts of DC and Nyquist are
    ignored): x[t] = sum_f X_f e^{+2 pi i f t / n} with the Hermitian extension.  Radix-2 FFT in mp (n a power of two)."""
    nh = n // 2 + 1
    full = [mp.mpc(0)] * n
    full[0] = mp.mpc(mp.re(xh[0]), 0)
    full[n // 2] = mp.mpc(mp.re(xh[nh - 1]), 0)
    for f in range(1, nh - 1):
        full[f] = xh[f]
        full[n - f] = mp.conj(xh[f])
    bits = n.bit_length() - 1
    a = [full[int(format(i, "0%db" % bits)[::-1], 2)] for i in range(n)]
    size = 2
    while size <= n:
        w = [mp.expjpi(mp.mpf(2 * q) / size) for q in range(size // 2)]
        for start in range(0, n, size):
            for q in range(size // 2):
                u, v = a[start + q], a[start + q + size // 2] * w[q]
                a[start + q], a[start + q + size // 2] = u + v, u - v
        size *= 2
    return [mp.re(x) for x in a]


def calc_rf_mp(cfg: pyo.Config, flt: np.ndarray, nlay, alpha, beta, rho, h):
    """calc_rf (src/forward.f90:123-208) -> (rft[ntrc][nfft] of mpf, cond[ntrc], npre[ntrc]).  cond = max|rxv| / max(rxv): the
    amplification of relative errors by the normalisation (1 for a healthy P receiver function)."""
    n, nh = cfg.nfft, cfg.nh
    out, conds, npres = [], [], []
    common = cfg.is_ray_common
    freq_r = freq_v = rff = None
    tp = mp.mpf(0)
    for t in range(cfg.ntrc):
        ipha = cfg.ipha[t]
        if t == 0 or not common:
            ur, uz = calc_seis_mp(cfg, nlay, cfg.rayps[t], ipha, alpha, beta, rho, h)
            freq_r = [mp.conj(x) for x in ur]
            freq_v = [-mp.conj(x) for x in uz]
            if cfg.deconv_mode == 1:
                y, x = (freq_r, freq_v) if ipha == 1 else (freq_v, freq_r)
                amp = [mp.re(v * mp.conj(v)) for v in x]
                wl = _mpf(0.001) * max(amp)                                  # 0.001d0, src/forward.f90:149
                rff = [y[i] * mp.conj(x[i]) / max(amp[i], wl) for i in range(nh)]
                tp = mp.mpf(0)
            else:
                rff = freq_r if ipha == 1 else freq_v
                v = alpha if ipha == 1 else beta
                i0 = 1 if cfg.sdep > 0.0 else 0
                p = _mpf(cfg.rayps[t])
                tp = sum((_mpf(h[i]) * mp.sqrt(1 / (_mpf(v[i]) * _mpf(v[i])) - p * p) for i in range(i0, nlay - 1)), mp.mpf(0))
        f = [_mpf(flt[i, t]) for i in range(nh)]
        rx = _c2r_mp([rff[i] * f[i] for i in range(nh)], n)
        shift = (-_mpf(cfg.t_start) - tp) / _mpf(cfg.delta) if ipha == 1 else (-_mpf(cfg.t_start) + tp) / _mpf(cfg.delta)
        npre = int(mp.floor(shift + mp.mpf(0.5))) if shift >= 0 else -int(mp.floor(-shift + mp.mpf(0.5)))
        if ipha == 1:
            tr = [rx[(n - npre + i1) % n - 1 if (n - npre + i1) % n else n - 1] for i1 in range(1, n + 1)]
        else:
            tr = [-rx[(n + npre - i1 + 1) % n - 1 if (n + npre - i1 + 1) % n else n - 1] for i1 in range(1, n + 1)]
        cond = mp.mpf(1)
        if cfg.deconv_mode == 0:
            rxv = _c2r_mp([freq_v[i] * f[i] for i in range(nh)], n)
            fac = max(rxv)
            cond = max(abs(x) for x in rxv) / abs(fac)
            tr = [x / fac for x in tr]
        out.append(tr); conds.append(float(cond)); npres.append(npre)
    return out, conds, npres


def evaluate(cfg: pyo.Config, flt: np.ndarray, r_inv: np.ndarray, k, z, dvp, dvs, sig):
    """One model -> dict(rft[ntrc][nfft] float64 (rounded once), phi[ntrc], logl, cond[ntrc], npre[ntrc])."""
    mp.mp.prec = PREC
    nlay, alpha, beta, rho, h, _ = pyo.format_model(cfg, int(k), np.asarray(z), np.asarray(dvp), np.asarray(dvs))
    rft, cond, npre = calc_rf_mp(cfg, flt, nlay, alpha, beta, rho, h)
    S = cfg.nsmp
    phis, ll = [], mp.mpf(0)
    for t in range(cfg.ntrc):
        m = [rft[t][i] - _mpf(cfg.obs[i, t]) for i in range(S)]
        R = r_inv[:, :, t]
        phi1 = [sum((m[i] * _mpf(R[i, j]) for i in range(S)), mp.mpf(0)) for j in range(S)]      # matmul(misfits, r_inv), src/likelihood.f90:88
        phi = sum((phi1[j] * m[j] for j in range(S)), mp.mpf(0))
        s = _mpf(sig[t])
        ll = ll - phi / (2 * s * s) - S * mp.log(s)
        phis.append(float(phi))
    return dict(rft=np.array([[float(x) for x in tr] for tr in rft]), phi=np.array(phis), logl=float(ll),
                cond=np.array(cond), npre=np.array(npre, dtype=np.int64))
