/*
 * rfinv_oracle.c -- CPU restatement (plain C99 + OpenMP) of RF_INV's forward-model + likelihood path
 * and of the PT-MCMC loop that calls it.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in rf_inv_b200/ links, loads or calls this file.  It is used by
 * tests/ (as the checker), by __graft_entry__.smoke() (as the checker) and by bench.py's cpu_baseline /
 * --impl reference legs (as the timed CPU arm: "CPU restatement of the reference algorithm").
 *
 * It follows the reference *as written*: dense complex 4x4 propagator products, four libm sin/cos per
 * layer-frequency, dense S x S quadratic form, and the single-precision literals that leak into the
 * reference's fp64 results.  Each function cites the reference file:line it restates (paths relative
 * to the reference checkout).  Compile with -ffp-contract=off (no FMA contraction), no -ffast-math.
 *
 * Third-party arithmetic not vendored by the reference (Makefile:18-19, versions unpinned):
 *   FFTW3 c2r  -> own radix-2 inverse FFT of the Hermitian-extended spectrum (unnormalised, sign +i,
 *                 imaginary parts of DC and Nyquist ignored, like FFTW's c2r).
 *   LAPACK dgesvd -> not restated here: r_inv is an input (tests build it with scipy's gesvd).
 *
 * Parity status: forward path (land / P / deconv_mode 0) pinned by the reference's fixtures
 * sample_syn/data/sample_{1,2}.trc (float32) and vp_to_rho(5.0) by sample_syn/true/true.velmod;
 * everything else: parity unpinned (no fixture, no Fortran compiler here) -- cross-checked against
 * the independent numpy restatement oracle/rfinv_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rfinv_b200.h"

#define PI 3.1415926535897931 /* forward.f90:33 */

/* single-precision literals promoted to double */
static const double OMG_DC = (double)1.0e-5f;      /* forward.f90:247 */
static const double V1_TINY = (double)1.0e-16f;    /* math.f90:44 */
static const double COLD_EPS = (double)1.0e-6f;    /* pt_mcmc.f90:196,204 */
static const double SIGMODE_EPS = (double)1.0e-5f; /* params.f90:262 */

typedef struct { double re, im; } cplx;

static inline cplx c_make(double re, double im) { cplx z = {re, im}; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
static inline cplx c_neg(cplx a) { return c_make(-a.re, -a.im); }
static inline cplx c_conj(cplx a) { return c_make(a.re, -a.im); }
static inline cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cplx c_scale(cplx a, double s) { return c_make(a.re * s, a.im * s); }
/* complex division with range reduction (Smith), what gfortran emits (-fcx-fortran-rules) */
static inline cplx c_div(cplx a, cplx b) {
  double r, den;
  if (fabs(b.re) < fabs(b.im)) {
    r = b.re / b.im;
    den = b.re * r + b.im;
    return c_make((a.re * r + a.im) / den, (a.im * r - a.re) / den);
  }
  r = b.im / b.re;
  den = b.im * r + b.re;
  return c_make((a.re + a.im * r) / den, (a.im - a.re * r) / den);
}

static inline int f_nint(double x) { /* Fortran NINT */
  return x >= 0.0 ? (int)floor(x + 0.5) : -(int)floor(-x + 0.5);
}

int32_t orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ model.f90 */
/* model.f90:298-314, Brocher (2005) with float32 coefficients */
double orc_vp_to_rho(double a1) {
  double a2 = a1 * a1, a3 = a2 * a1, a4 = a3 * a1, a5 = a4 * a1;
  return (double)1.6612f * a1 - (double)0.4721f * a2 + (double)0.0671f * a3 - (double)0.0043f * a4 +
         (double)0.000106f * a5;
}

static int layer_vel(const rfinv_config* c, double zc, double d_vs, double d_vp, double* a, double* b) {
  int iz = f_nint((zc - c->z_ref_min) / c->dz_ref) + 1; /* model.f90:212 */
  if (iz < 1) iz = 1;                                    /* the reference would read out of bounds */
  if (iz > c->nref) iz = c->nref;
  *b = c->vs_ref[iz - 1] + d_vs;
  *a = (c->vp_mode == 1) ? c->vp_ref[iz - 1] + d_vp : c->vp_ref[iz - 1];
  if (*a < c->vp_min || *a > c->vp_max || *b < c->vs_min || *b > c->vs_max || *a / *b < c->vpvs_min ||
      *a / *b > c->vpvs_max)
    return 0;
  return 1;
}

/* model.f90:175-290.  alpha/beta/rho/h need k_max+1 entries.  Returns is_valid. */
int32_t orc_format_model(const rfinv_config* c, int32_t k, const double* z, const double* dvp, const double* dvs,
                         int32_t* nlay, double* alpha, double* beta, double* rho, double* h) {
  double tz[256], tp[256], ts[256];
  int is_valid = 1, i = 0;
  for (int j = 0; j < k; ++j) { tz[j] = z[j]; tp[j] = dvp[j]; ts[j] = dvs[j]; }
  for (int j = 1; j < k; ++j) { /* sort.f90:34-68: any correct sort, keys are distinct */
    double a = tz[j], b = tp[j], cc = ts[j];
    int m = j - 1;
    while (m >= 0 && tz[m] > a) { tz[m + 1] = tz[m]; tp[m + 1] = tp[m]; ts[m + 1] = ts[m]; --m; }
    tz[m + 1] = a; tp[m + 1] = b; ts[m + 1] = cc;
  }
  if (c->sdep > 0.0) { /* model.f90:201-207 */
    alpha[i] = 1.5; beta[i] = -999.0; rho[i] = 1.0; h[i] = c->sdep; ++i;
  }
  /* top layer, model.f90:210-231 */
  if (!layer_vel(c, 0.5 * (c->sdep + tz[0]), ts[0], tp[0], &alpha[i], &beta[i])) is_valid = 0;
  rho[i] = orc_vp_to_rho(alpha[i]);
  h[i] = tz[0] - c->sdep;
  if (h[i] < 0.125 * alpha[i]) is_valid = 0; /* model.f90:229 */
  ++i;
  for (int j = 1; j < k; ++j) { /* model.f90:235-262 */
    if (!layer_vel(c, 0.5 * (tz[j] + tz[j - 1]), ts[j], tp[j], &alpha[i], &beta[i])) is_valid = 0;
    rho[i] = orc_vp_to_rho(alpha[i]);
    h[i] = tz[j] - tz[j - 1];
    if (h[i] < c->h_min) is_valid = 0;
    ++i;
  }
  /* half space, model.f90:264-283 */
  if (!layer_vel(c, 0.5 * (c->z_max + tz[k - 1]), dvs[c->k_max - 1], dvp[c->k_max - 1], &alpha[i], &beta[i]))
    is_valid = 0;
  rho[i] = orc_vp_to_rho(alpha[i]);
  h[i] = 999.0;
  ++i;
  *nlay = i;
  return is_valid;
}

/* ------------------------------------------------------------------ forward.f90 */
/* forward.f90:95-119 -> flt[ntrc][nh] */
void orc_init_filter(const rfinv_config* c, double* flt) {
  int nh = c->nfft / 2 + 1;
  double df = 1.0 / (c->delta * c->nfft);
  for (int t = 0; t < c->ntrc; ++t) {
    double fac_norm = c->nfft * c->a_gus[t] * c->delta / sqrt(PI);
    for (int i = 0; i < nh; ++i) {
      double omega = i * 2.0 * PI * df;
      double q = omega / (2.0 * c->a_gus[t]);
      flt[(size_t)t * nh + i] = exp(-(q * q)) / fac_norm;
    }
  }
}

/* forward.f90:350-380 */
static void e_inverse(double omega, double rho, double alpha, double beta, double p, cplx e[4][4]) {
  const cplx ei = {0.0, 1.0};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) e[i][j] = c_make(0.0, 0.0);
  double eta = sqrt(1.0 / (beta * beta) - p * p);
  double xi = sqrt(1.0 / (alpha * alpha) - p * p);
  double bp = 1.0 - 2.0 * beta * beta * p * p;
  e[0][0] = c_make(beta * beta * p / alpha, 0.0);
  e[0][1] = c_make(bp / (2.0 * alpha * xi), 0.0);
  e[0][2] = c_mul(c_make(-p / (2.0 * omega * rho * alpha * xi), 0.0), ei);
  e[0][3] = c_mul(c_make(-1.0 / (2.0 * omega * rho * alpha), 0.0), ei);
  e[1][0] = c_make(bp / (2.0 * beta * eta), 0.0);
  e[1][1] = c_make(-beta * p, 0.0);
  e[1][2] = c_mul(c_make(-1.0 / (2.0 * omega * rho * beta), 0.0), ei);
  e[1][3] = c_mul(c_make(p / (2.0 * omega * rho * beta * eta), 0.0), ei);
  e[2][0] = e[0][0];
  e[2][1] = c_neg(e[0][1]);
  e[2][2] = c_neg(e[0][2]);
  e[2][3] = e[0][3];
  e[3][0] = e[1][0];
  e[3][1] = c_neg(e[1][1]);
  e[3][2] = c_neg(e[1][2]);
  e[3][3] = e[1][3];
}

/* forward.f90:385-421 */
static void layer_matrix_sol(double omega, double rho, double alpha, double beta, double p, double z, cplx m[4][4]) {
  const cplx ei = {0.0, 1.0};
  double beta2 = beta * beta, p2 = p * p;
  double bp = 1.0 - 2.0 * beta2 * p2;
  double eta = sqrt(1.0 / beta2 - p2);
  double xi = sqrt(1.0 / (alpha * alpha) - p2);
  double cos_xi = cos(omega * xi * z), cos_eta = cos(omega * eta * z);
  double sin_xi = sin(omega * xi * z), sin_eta = sin(omega * eta * z);
  m[0][0] = c_make(2.0 * beta2 * p2 * cos_xi + bp * cos_eta, 0.0);
  m[1][0] = c_mul(c_make(p * (2.0 * beta2 * xi * sin_xi - bp / eta * sin_eta), 0.0), ei);
  m[2][0] = c_make(omega * rho * (-4.0 * beta2 * beta2 * p2 * xi * sin_xi - bp * bp / eta * sin_eta), 0.0);
  m[3][0] = c_mul(c_make(2.0 * omega * beta2 * rho * p * bp * (cos_xi - cos_eta), 0.0), ei);
  m[0][1] = c_mul(c_make(p * (bp / xi * sin_xi - 2.0 * beta2 * eta * sin_eta), 0.0), ei);
  m[1][1] = c_make(bp * cos_xi + 2.0 * beta2 * p2 * cos_eta, 0.0);
  m[2][1] = m[3][0];
  m[3][1] = c_make(-omega * rho * (bp * bp / xi * sin_xi + 4.0 * beta2 * beta2 * p2 * eta * sin_eta), 0.0);
  m[0][2] = c_make((p2 / xi * sin_xi + eta * sin_eta) / (omega * rho), 0.0);
  m[1][2] = c_mul(c_make(p * (-cos_xi + cos_eta) / (omega * rho), 0.0), ei);
  m[2][2] = m[0][0];
  m[3][2] = m[0][1];
  m[0][3] = m[1][2];
  m[1][3] = c_make((xi * sin_xi + p2 / eta * sin_eta) / (omega * rho), 0.0);
  m[2][3] = m[1][0];
  m[3][3] = m[1][1];
}

/* forward.f90:424-442 */
static void layer_matrix_liq(double omega, double rho, double alpha, double p, double z, cplx m[2][2]) {
  double xi = sqrt(1.0 / (alpha * alpha) - p * p);
  double cos_xi = cos(omega * xi * z), sin_xi = sin(omega * xi * z);
  double g = rho * omega / xi;
  m[0][0] = c_make(cos_xi, 0.0);
  m[0][1] = c_make(sin_xi / g, 0.0);
  m[1][0] = c_make(-g * sin_xi, 0.0);
  m[1][1] = c_make(cos_xi, 0.0);
}

static void matmul4(cplx a[4][4], cplx b[4][4], cplx out[4][4]) {
  cplx t[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      cplx s = c_make(0.0, 0.0);
      for (int k = 0; k < 4; ++k) s = c_add(s, c_mul(a[i][k], b[k][j]));
      t[i][j] = s;
    }
  memcpy(out, t, sizeof(t));
}

/* forward.f90:212-344 -> ur[nh], uz[nh] */
static void calc_seis(const rfinv_config* c, int nlay, double rayp, int ipha, const double* alpha, const double* beta,
                      const double* rho, const double* h, cplx* ur, cplx* uz) {
  int npts = c->nfft, nhalf = npts / 2 + 1;
  int sea_flag = beta[0] < 0;
  int ilay0 = sea_flag ? 1 : 0;
  double domg = 2.0 * PI / (npts * c->delta);
  for (int iomg = 0; iomg < nhalf; ++iomg) {
    double omg = (double)iomg * domg;
    if (iomg == 0) omg = OMG_DC; /* forward.f90:246-248 */
    cplx e_inv[4][4], p_prod[4][4], p_mat[4][4], sl[4][4];
    e_inverse(omg, rho[nlay - 1], alpha[nlay - 1], beta[nlay - 1], rayp, e_inv);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) p_prod[i][j] = c_make(i == j ? 1.0 : 0.0, 0.0);
    for (int il = ilay0; il < nlay - 1; ++il) {
      layer_matrix_sol(omg, rho[il], alpha[il], beta[il], rayp, h[il], p_mat);
      matmul4(p_mat, p_prod, p_prod); /* forward.f90:262 */
    }
    matmul4(e_inv, p_prod, sl); /* forward.f90:264 */
    cplx sd4 = c_make(0.0, 0.0);
    if (!sea_flag) {            /* forward.f90:267-275 */
      cplx denom = c_sub(c_mul(sl[2][0], sl[3][1]), c_mul(sl[2][1], sl[3][0]));
      if (ipha >= 0) {
        ur[iomg] = c_div(sl[3][1], denom);
        uz[iomg] = c_div(c_neg(sl[3][0]), denom);
      } else {
        ur[iomg] = c_div(c_neg(sl[2][1]), denom);
        uz[iomg] = c_div(sl[2][0], denom);
      }
    } else { /* forward.f90:276-287 */
      cplx lq[2][2];
      layer_matrix_liq(omg, rho[0], alpha[0], rayp, h[0], lq);
      cplx a = c_add(c_mul(sl[3][1], lq[0][0]), c_mul(sl[3][3], lq[1][0]));
      cplx b = c_add(c_mul(sl[2][1], lq[0][0]), c_mul(sl[2][3], lq[1][0]));
      cplx d1 = c_sub(c_mul(a, sl[2][0]), c_mul(b, sl[3][0]));
      cplx d2 = c_sub(c_mul(b, sl[3][0]), c_mul(a, sl[2][0]));
      if (ipha >= 0) {
        ur[iomg] = c_div(a, d1);
        uz[iomg] = c_div(c_mul(lq[0][0], sl[3][0]), d2);
      } else {
        ur[iomg] = c_div(c_neg(b), d1);
        uz[iomg] = c_div(c_mul(c_neg(lq[0][0]), sl[2][0]), d2);
      }
      if (c->bdep > 0.0) /* forward.f90:297-306 (commented out): normal stress under the water */
        sd4 = ipha >= 0 ? c_div(c_mul(lq[1][0], sl[3][0]), d2) : c_div(c_mul(c_neg(lq[1][0]), sl[2][0]), d2);
    }
    if (c->bdep > 0.0) {
      /* Buried station, forward.f90:289-338 -- COMMENTED OUT in the reference.  The displacement-stress vector of the
       * surface / sea floor is carried down to the station.  Two defects of the commented block are not restated (the
       * numpy oracle's bdep_literal switch does): the downward loop has no exit after the station's layer, and a station
       * in the half space is moved by bdep instead of its distance to the last interface. */
      cplx sd[4] = {ur[iomg], uz[iomg], c_make(0.0, 0.0), sd4}, t4[4];
      double z_tmp = 0.0;
      int found = 0;
      for (int il = ilay0; il < nlay - 1 && !found; ++il) {
        z_tmp = z_tmp + h[il];
        double h_use = h[il];
        if (!(z_tmp < c->bdep)) { h_use = c->bdep + h[il] - z_tmp; found = 1; }
        layer_matrix_sol(omg, rho[il], alpha[il], beta[il], rayp, h_use, p_mat);
        for (int i = 0; i < 4; ++i) {
          t4[i] = c_make(0.0, 0.0);
          for (int j = 0; j < 4; ++j) t4[i] = c_add(t4[i], c_mul(p_mat[i][j], sd[j]));
        }
        memcpy(sd, t4, sizeof(sd));
      }
      if (!found) {
        layer_matrix_sol(omg, rho[nlay - 1], alpha[nlay - 1], beta[nlay - 1], rayp, c->bdep - z_tmp, p_mat);
        for (int i = 0; i < 4; ++i) {
          t4[i] = c_make(0.0, 0.0);
          for (int j = 0; j < 4; ++j) t4[i] = c_add(t4[i], c_mul(p_mat[i][j], sd[j]));
        }
        memcpy(sd, t4, sizeof(sd));
      }
      ur[iomg] = sd[0];
      uz[iomg] = sd[1];
    }
  }
}

/* forward.f90:474-491 */
static double direct_arrival(const rfinv_config* c, int nlay, const double* h, const double* v, double rayp) {
  int i0 = c->sdep > 0.0 ? 1 : 0;
  double t = 0.0;
  if (!(c->bdep > 0.0)) {
    for (int i = i0; i < nlay - 1; ++i) t = t + h[i] * sqrt(1.0 / (v[i] * v[i]) - rayp * rayp);
    return t;
  }
  /* buried station, forward.f90:493-516 (commented out): delay from the station down to the top of the half space; each
   * layer with its own velocity (the commented block reads v(i0) for all of them), and a station in the half space
   * counts its distance to the last interface negatively */
  double z_sum = 0.0;
  int i = i0;
  for (; i < nlay - 1; ++i) {
    z_sum = z_sum + h[i];
    if (z_sum > c->bdep) {
      t = t + (z_sum - c->bdep) * sqrt(1.0 / (v[i] * v[i]) - rayp * rayp);
      ++i;
      break;
    }
    if (i == nlay - 2) { /* ran out of solid layers */
      t = t - (c->bdep - z_sum) * sqrt(1.0 / (v[nlay - 1] * v[nlay - 1]) - rayp * rayp);
      return t;
    }
  }
  for (; i < nlay - 1; ++i) t = t + h[i] * sqrt(1.0 / (v[i] * v[i]) - rayp * rayp);
  return t;
}

/* forward.f90:447-470: z = y conj(x) / max(|x|^2, pcnt max|x|^2) */
static void water_level_decon(const cplx* y, const cplx* x, cplx* z, int n, double pcnt, double* amp) {
  double mx = -INFINITY;
  for (int i = 0; i < n; ++i) {
    amp[i] = c_mul(x[i], c_conj(x[i])).re;
    if (amp[i] > mx) mx = amp[i];
  }
  double wlvl = pcnt * mx;
  for (int i = 0; i < n; ++i) {
    cplx num = c_mul(y[i], c_conj(x[i]));
    double d = amp[i] > wlvl ? amp[i] : wlvl;
    z[i] = c_make(num.re / d, num.im / d);
  }
}

/* ---- FFTW c2r stand-in (src/fftw.f90:44): radix-2 inverse FFT of the Hermitian extension for powers of two; any other
 * length (FFTW takes any) by the defining sum over the half spectrum -- O(n^2), independent of the product's Bluestein path ---- */
typedef struct {
  int n, pow2;
  cplx* tw;   /* exp(+2 pi i j / n), j < n */
  int* rev;
} fft_plan;

static fft_plan* fft_plan_create(int n) {
  fft_plan* p = (fft_plan*)malloc(sizeof(fft_plan));
  p->n = n;
  p->pow2 = (n & (n - 1)) == 0;
  p->tw = (cplx*)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
  p->rev = (int*)malloc(sizeof(int) * (size_t)n);
  for (int j = 0; j < n; ++j) p->tw[j] = c_make(cos(2.0 * PI * j / n), sin(2.0 * PI * j / n));
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b)
      if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    p->rev[i] = r;
  }
  return p;
}
static void fft_plan_destroy(fft_plan* p) {
  if (!p) return;
  free(p->tw); free(p->rev); free(p);
}
/* x[m] = sum_j X[j] exp(+2 pi i j m / n), X Hermitian from half[0..n/2]; work has n entries */
static void c2r(const fft_plan* p, const cplx* half, double* out, cplx* work) {
  int n = p->n, nh = n / 2;
  if (!p->pow2) {
    /* x[m] = X0 + [n even] (-1)^m X_{n/2} + 2 sum_{0 < j < n/2} Re(X_j e^{+2 pi i j m / n}); imaginary parts of the
     * DC and Nyquist bins are ignored like FFTW's c2r does */
    int jtop = (n & 1) ? nh : nh - 1;   /* last bin with a distinct mirror */
    for (int m = 0; m < n; ++m) {
      double acc = 0.0;
      for (int j = jtop; j >= 1; --j) {   /* small terms first: the Gaussian filter decays with j */
        cplx w = p->tw[(int)(((long long)j * m) % n)];
        acc += half[j].re * w.re - half[j].im * w.im;
      }
      acc = 2.0 * acc + half[0].re;
      if (!(n & 1)) acc += (m & 1) ? -half[nh].re : half[nh].re;
      out[m] = acc;
    }
    (void)work;
    return;
  }
  work[p->rev[0]] = c_make(half[0].re, 0.0);
  work[p->rev[nh]] = c_make(half[nh].re, 0.0);
  for (int j = 1; j < nh; ++j) {
    work[p->rev[j]] = half[j];
    work[p->rev[n - j]] = c_conj(half[j]);
  }
  for (int len = 2; len <= n; len <<= 1) {
    int step = n / len, hl = len / 2;
    for (int s = 0; s < n; s += len)
      for (int j = 0; j < hl; ++j) {
        cplx w = p->tw[j * step];
        cplx a = work[s + j], b = c_mul(work[s + j + hl], w);
        work[s + j] = c_add(a, b);
        work[s + j + hl] = c_sub(a, b);
      }
  }
  for (int i = 0; i < n; ++i) out[i] = work[i].re;
}

typedef struct {
  cplx *ur, *uz, *fr, *fv, *rff, *cx, *work;
  double *rx, *amp;
  double norm_cond[64]; /* per trace: max|rx| / maxval(rx) of the normalising vertical trace (diagnostics, see orc_eval_batch_cond) */
} rf_ws;
static rf_ws* ws_create(int n) {
  rf_ws* w = (rf_ws*)malloc(sizeof(rf_ws));
  size_t nb = sizeof(cplx) * (size_t)n;
  w->ur = malloc(nb); w->uz = malloc(nb); w->fr = malloc(nb); w->fv = malloc(nb);
  w->rff = malloc(nb); w->cx = malloc(nb); w->work = malloc(nb);
  w->rx = malloc(sizeof(double) * (size_t)n); w->amp = malloc(sizeof(double) * (size_t)n);
  return w;
}
static void ws_destroy(rf_ws* w) {
  if (!w) return;
  free(w->ur); free(w->uz); free(w->fr); free(w->fv); free(w->rff); free(w->cx); free(w->work);
  free(w->rx); free(w->amp); free(w);
}

static int ray_common(const rfinv_config* c) { /* forward.f90:59-76 */
  for (int t = 1; t < c->ntrc; ++t)
    if (c->rayps[t] != c->rayps[0] || c->ipha[t] != c->ipha[0]) return 0;
  return 1;
}

/* forward.f90:123-208 -> rft[ntrc][nfft] */
static void calc_rf(const rfinv_config* c, const fft_plan* plan, const double* flt, rf_ws* w, int nlay,
                    const double* alpha, const double* beta, const double* rho, const double* h, double* rft) {
  int n = c->nfft, nh = n / 2 + 1;
  int common = ray_common(c);
  double tp = 0.0;
  for (int t = 0; t < c->ntrc; ++t) {
    int ipha = c->ipha[t];
    if (t == 0 || !common) {
      calc_seis(c, nlay, c->rayps[t], ipha, alpha, beta, rho, h, w->ur, w->uz);
      for (int i = 0; i < nh; ++i) {
        w->fr[i] = c_conj(w->ur[i]);
        w->fv[i] = c_neg(c_conj(w->uz[i])); /* upward positive */
      }
      if (c->deconv_mode == 1 && ipha == 1) {
        water_level_decon(w->fr, w->fv, w->rff, nh, 0.001, w->amp);
        tp = 0.0;
      } else if (c->deconv_mode == 1 && ipha == -1) {
        water_level_decon(w->fv, w->fr, w->rff, nh, 0.001, w->amp);
        tp = 0.0;
      } else if (ipha == 1) {
        memcpy(w->rff, w->fr, sizeof(cplx) * (size_t)nh);
        tp = direct_arrival(c, nlay, h, alpha, c->rayps[t]);
      } else {
        memcpy(w->rff, w->fv, sizeof(cplx) * (size_t)nh);
        tp = direct_arrival(c, nlay, h, beta, c->rayps[t]);
      }
    }
    const double* f = flt + (size_t)t * nh;
    for (int i = 0; i < nh; ++i) w->cx[i] = c_scale(w->rff[i], f[i]); /* forward.f90:168 */
    c2r(plan, w->cx, w->rx, w->work);
    double* out = rft + (size_t)t * n;
    if (ipha == 1) { /* forward.f90:176-184 */
      int npre = f_nint((-c->t_start - tp) / c->delta);
      for (int i = 1; i <= n; ++i) {
        int j = ((n - npre + i) % n + n) % n;
        if (j == 0) j = n;
        out[i - 1] = w->rx[j - 1];
      }
    } else { /* forward.f90:185-194 */
      int npre = f_nint((-c->t_start + tp) / c->delta);
      for (int i = 1; i <= n; ++i) {
        int j = ((n + npre - i + 1) % n + n) % n;
        if (j == 0) j = n;
        out[i - 1] = -w->rx[j - 1];
      }
    }
    if (c->deconv_mode == 0) { /* forward.f90:197-203 */
      for (int i = 0; i < nh; ++i) w->cx[i] = c_scale(w->fv[i], f[i]);
      c2r(plan, w->cx, w->rx, w->work);
      double fac = w->rx[0], amax = fabs(w->rx[0]);
      for (int i = 1; i < n; ++i) {
        if (w->rx[i] > fac) fac = w->rx[i];
        if (fabs(w->rx[i]) > amax) amax = fabs(w->rx[i]);
      }
      if (fac != fac) fac = NAN;
      for (int i = 0; i < n; ++i) out[i] = out[i] / fac;
      if (t < 64) w->norm_cond[t] = amax / fabs(fac);
    } else if (t < 64) {
      w->norm_cond[t] = 1.0;
    }
  }
}

/* ------------------------------------------------------------------ likelihood.f90 */
/* likelihood.f90:85-98; r_inv [ntrc][nsmp][nsmp], obs [ntrc][nsmp] */
static double loglik_from_rft(const rfinv_config* c, const double* rft, const double* sig, double* misfits) {
  int S = c->nsmp, n = c->nfft;
  double ll = 0.0;
  for (int t = 0; t < c->ntrc; ++t) {
    const double* ri = c->r_inv + (size_t)t * S * S;
    for (int i = 0; i < S; ++i) misfits[i] = rft[(size_t)t * n + i] - c->obs[(size_t)t * S + i];
    double s = sig[t], phi = 0.0;
    for (int j = 0; j < S; ++j) { /* phi1(j) = sum_i m(i) r_inv(i,j); r_inv(i,j) at ri[j*S+i] */
      const double* col = ri + (size_t)j * S;
      double p1 = 0.0;
      for (int i = 0; i < S; ++i) p1 += misfits[i] * col[i];
      phi += p1 * misfits[j];
    }
    ll = ll - 0.5 * phi / (s * s) - (double)S * log(s);
  }
  return ll;
}

/* calc_likelihood (likelihood.f90:56-101) over C models; OpenMP over models.
 * Layouts as rfinv_eval_batch (include/rfinv_b200.h).  Returns 0, or 1 when r_inv/obs missing. */
/* Condition number of the normalisation of every (model, trace): the reference divides by maxval(rx) of the filtered
 * vertical trace (forward.f90:197-203) -- the largest POSITIVE sample, not the largest magnitude.  Where the main pulse
 * of that trace is negative (S incidence on a nearly transparent model) the divisor is a small ripple and every
 * rounding error of the trace is amplified by max|rx| / maxval(rx): two correct fp64 evaluations of the reference's own
 * formulas then agree to eps * cond only.  cond[C][ntrc]; 1 with deconv_mode 1. */
static double* g_cond_out = NULL;
int32_t orc_eval_batch(const rfinv_config* c, int32_t C, const int32_t* k, const double* z, const double* dvp,
                       const double* dvs, const double* sig, double* logl, double* rft_out, uint8_t* is_valid,
                       int32_t nthreads);
int32_t orc_eval_batch_cond(const rfinv_config* c, int32_t C, const int32_t* k, const double* z, const double* dvp,
                            const double* dvs, const double* sig, double* logl, double* rft_out, uint8_t* is_valid,
                            int32_t nthreads, double* cond) {
  g_cond_out = cond;
  int32_t st = orc_eval_batch(c, C, k, z, dvp, dvs, sig, logl, rft_out, is_valid, nthreads);
  g_cond_out = NULL;
  return st;
}

int32_t orc_eval_batch(const rfinv_config* c, int32_t C, const int32_t* k, const double* z, const double* dvp,
                       const double* dvs, const double* sig, double* logl, double* rft_out, uint8_t* is_valid,
                       int32_t nthreads) {
  if (!c->r_inv || !c->obs) return 1;
  int n = c->nfft, nh = n / 2 + 1, km = c->k_max, T = c->ntrc;
  double* flt = (double*)malloc(sizeof(double) * (size_t)nh * T);
  orc_init_filter(c, flt);
  fft_plan* plan = fft_plan_create(n);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    rf_ws* w = ws_create(n);
    double* rft = (double*)malloc(sizeof(double) * (size_t)n * T);
    double* mis = (double*)malloc(sizeof(double) * (size_t)c->nsmp);
    double alpha[258], beta[258], rho[258], h[258];
#pragma omp for schedule(dynamic, 4)
    for (int ic = 0; ic < C; ++ic) {
      int nlay;
      int ok = orc_format_model(c, k[ic], z + (size_t)ic * (km - 1), dvp + (size_t)ic * km, dvs + (size_t)ic * km,
                                &nlay, alpha, beta, rho, h);
      calc_rf(c, plan, flt, w, nlay, alpha, beta, rho, h, rft);
      logl[ic] = loglik_from_rft(c, rft, sig + (size_t)ic * T, mis);
      if (rft_out) memcpy(rft_out + (size_t)ic * n * T, rft, sizeof(double) * (size_t)n * T);
      if (is_valid) is_valid[ic] = (uint8_t)ok;
      if (g_cond_out)
        for (int t = 0; t < T && t < 64; ++t) g_cond_out[(size_t)ic * T + t] = w->norm_cond[t];
    }
    free(rft); free(mis); ws_destroy(w);
  }
  fft_plan_destroy(plan);
  free(flt);
  return 0;
}

/* calc_rf for an explicit layer stack (used to check the golden traces): rft[ntrc][nfft] */
int32_t orc_calc_rf_layers(const rfinv_config* c, int32_t nlay, const double* alpha, const double* beta,
                           const double* rho, const double* h, double* rft) {
  int n = c->nfft, nh = n / 2 + 1;
  double* flt = (double*)malloc(sizeof(double) * (size_t)nh * c->ntrc);
  orc_init_filter(c, flt);
  fft_plan* plan = fft_plan_create(n);
  rf_ws* w = ws_create(n);
  calc_rf(c, plan, flt, w, nlay, alpha, beta, rho, h, rft);
  ws_destroy(w); fft_plan_destroy(plan); free(flt);
  return 0;
}

/* ------------------------------------------------------------------ mt19937.f90 */
typedef struct { uint32_t mt[624]; int mti; } mt_state;

void orc_sgrnd(mt_state* s, uint32_t seed) { /* mt19937.f90:78-90 */
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 69069u * s->mt[i - 1];
  s->mti = 624;
}
double orc_grnd(mt_state* s) { /* mt19937.f90:92-130 */
  uint32_t y;
  if (s->mti >= 624) {
    for (int kk = 0; kk < 624; ++kk) {
      y = (s->mt[kk] & 0x80000000u) | (s->mt[(kk + 1) % 624] & 0x7fffffffu);
      s->mt[kk] = s->mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->mti = 0;
  }
  y = s->mt[s->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return (double)y / 4294967296.0;
}
/* test hook: first `count` outputs for `seed` */
void orc_mt_sequence(uint32_t seed, int32_t count, double* out) {
  mt_state s;
  orc_sgrnd(&s, seed);
  for (int i = 0; i < count; ++i) out[i] = orc_grnd(&s);
}

static double gauss(mt_state* s) { /* math.f90:34-50 */
  double v1 = orc_grnd(s), v2 = orc_grnd(s);
  if (v1 == 0.0) v1 = V1_TINY;
  return sqrt(-2.0 * log(v1)) * cos(2.0 * PI * v2);
}

static double laplace(mt_state* s) { /* prior.f90:57-123 */
  const double d = 0.69314718055994529;
  double u1 = orc_grnd(s), u1p = 2.0 * u1, u1pp, u1ppp, u2, a, w, val;
  int i_sign, k;
  if (u1p < 1.0) { i_sign = 1; u1pp = 1.0 - u1p; } else { i_sign = -1; u1pp = 2.0 - u1p; }
  a = 0.0;
  for (;;) {
    u1ppp = 2.0 * u1pp;
    if (u1ppp >= 1.0) { u1 = u1ppp - 1.0; break; }
    a = a + d;
    u1pp = u1ppp;
  }
  for (;;) {
    w = d * u1;
    val = i_sign * (a + w);
    k = 1;
    for (;;) {
      u2 = orc_grnd(s);
      if (u2 >= w) { u1 = (u2 - w) / (1.0 - w); break; }
      w = u2;
      ++k;
    }
    if (k % 2 == 1) break;
  }
  return val;
}
void orc_deviates(uint32_t seed, int32_t kind, int32_t count, double* out) { /* test hook: 0 gauss, 1 laplace */
  mt_state s;
  orc_sgrnd(&s, seed);
  for (int i = 0; i < count; ++i) out[i] = kind ? laplace(&s) : gauss(&s);
}

static double log_prior_ratio(double x_new, double x_old, double dev, int prior_mode) { /* prior.f90:36-53 */
  if (prior_mode == 1) return -(fabs(x_new) - fabs(x_old)) / dev;
  return -((x_new * x_new) - (x_old * x_old)) / (2.0 * dev * dev);
}

/* ------------------------------------------------------------------ pt_mcmc.f90 with virtual ranks */
typedef struct orc_pt {
  rfinv_config c;
  int nproc, nthreads;
  double* flt;
  fft_plan* plan;
  mt_state* rng;                 /* [nproc] */
  int32_t* k;                    /* [G] , G = nproc*nchains, global chain g = rank*nchains + ichain */
  double *z, *dvp, *dvs, *sig;   /* [G][k_max-1], [G][k_max], [G][k_max], [G][ntrc] */
  double* rft;                   /* [G][ntrc][nfft] */
  double *logl, *temps;          /* [G] */
  int ntype, it_birth, it_death, it_z, it_dvs, it_dvp, it_sig, nsig_trc;
  int isig_trc[64];
  int sig_mode[64];
  int64_t *nprop, *naccept;      /* [ntype] */
  double* likelihood_hist;       /* [cap_hist] */
  int cap_hist, it_done;
  /* posterior bookkeeping, pt_mcmc.f90:204-286 */
  int64_t nmod;
  int64_t *nk, *nz, *nsig, *namp, *nvpz, *nvsz, *nvpvsz;
  double *vp_mean, *vs_mean, *vpvs_mean;
  int64_t n_eval;                /* forward+likelihood evaluations executed */
} orc_pt;

static void pt_types(orc_pt* p) { /* pt_mcmc.f90:311-365 */
  const rfinv_config* c = &p->c;
  p->ntype = 4; p->it_birth = 1; p->it_death = 2; p->it_z = 3; p->it_dvs = 4;
  if (c->vp_mode == 1) { p->ntype++; p->it_dvp = p->ntype; } else p->it_dvp = -1;
  p->nsig_trc = 0;
  for (int t = 0; t < c->ntrc; ++t) {
    p->sig_mode[t] = (c->sig_max[t] - c->sig_min[t] > SIGMODE_EPS) ? 1 : 0;
    if (p->sig_mode[t]) p->isig_trc[p->nsig_trc++] = t;
  }
  if (p->nsig_trc > 0) { p->ntype++; p->it_sig = p->ntype; } else p->it_sig = -1;
}

static void eval_chain(orc_pt* p, rf_ws* w, double* mis, int k, const double* z, const double* dvp, const double* dvs,
                       const double* sig, int fwd, const double* cached_rft, double* ll, double* rft) {
  const rfinv_config* c = &p->c;
  double alpha[258], beta[258], rho[258], h[258];
  int nlay;
  if (fwd) {
    orc_format_model(c, k, z, dvp, dvs, &nlay, alpha, beta, rho, h);
    calc_rf(c, p->plan, p->flt, w, nlay, alpha, beta, rho, h, rft);
  } else {
    memcpy(rft, cached_rft, sizeof(double) * (size_t)c->nfft * c->ntrc);
  }
  *ll = loglik_from_rft(c, rft, sig, mis);
}

static void init_rank(orc_pt* p, int rank) { /* rf_inv.f90:75-91 */
  const rfinv_config* c = &p->c;
  int nc = c->nchains, km = c->k_max, T = c->ntrc, n = c->nfft;
  mt_state* s = &p->rng[rank];
  orc_sgrnd(s, (uint32_t)c->iseed + (uint32_t)rank * (uint32_t)rank * 10000u + 23u * (uint32_t)rank);
  double alpha[258], beta[258], rho[258], h[258];
  for (int ic = 0; ic < nc; ++ic) { /* init_model, model.f90:62-95 */
    size_t g = (size_t)rank * nc + ic;
    double *z = p->z + g * (km - 1), *dvp = p->dvp + g * km, *dvs = p->dvs + g * km;
    int valid = 0, nlay;
    while (!valid) {
      int kk = c->k_min + (int)(orc_grnd(s) * (c->k_max - c->k_min));
      p->k[g] = kk;
      for (int i = 0; i < kk; ++i) z[i] = c->z_min + orc_grnd(s) * (c->z_max - c->z_min);
      for (int i = 0; i < kk; ++i) {
        if (c->prior_mode == 1) { dvs[i] = laplace(s) * c->dvs_prior; dvp[i] = laplace(s) * c->dvp_prior; }
        else if (c->prior_mode == 2) { dvs[i] = gauss(s) * c->dvs_prior; dvp[i] = gauss(s) * c->dvp_prior; }
      }
      if (c->prior_mode == 1) { dvs[km - 1] = laplace(s) * c->dvs_prior; dvp[km - 1] = laplace(s) * c->dvp_prior; }
      else if (c->prior_mode == 2) { dvs[km - 1] = gauss(s) * c->dvs_prior; dvp[km - 1] = gauss(s) * c->dvp_prior; }
      valid = orc_format_model(c, kk, z, dvp, dvs, &nlay, alpha, beta, rho, h);
    }
  }
  for (int ic = 0; ic < nc; ++ic) /* init_sig, likelihood.f90:117-126 */
    for (int t = 0; t < T; ++t) {
      size_t g = (size_t)rank * nc + ic;
      p->sig[g * T + t] =
          p->sig_mode[t] ? c->sig_min[t] + orc_grnd(s) * (c->sig_max[t] - c->sig_min[t]) : c->sig_min[t];
    }
  rf_ws* w = ws_create(n);
  double* mis = (double*)malloc(sizeof(double) * (size_t)c->nsmp);
  for (int ic = 0; ic < nc; ++ic) { /* init_rft, likelihood.f90:156-160 */
    size_t g = (size_t)rank * nc + ic;
    eval_chain(p, w, mis, p->k[g], p->z + g * (km - 1), p->dvp + g * km, p->dvs + g * km, p->sig + g * T, 1, NULL,
               &p->logl[g], p->rft + g * (size_t)n * T);
  }
  free(mis); ws_destroy(w);
  for (int ic = 0; ic < nc; ++ic) { /* pt_mcmc.f90:447-452 */
    size_t g = (size_t)rank * nc + ic;
    p->temps[g] = ic < c->ncool ? 1.0 : exp(orc_grnd(s) * log(c->t_high));
  }
}

orc_pt* orc_pt_create(const rfinv_config* cfg, int32_t nproc, int32_t nthreads) {
  if (!cfg->r_inv || !cfg->obs) return NULL;
  orc_pt* p = (orc_pt*)calloc(1, sizeof(orc_pt));
  p->c = *cfg;
  const rfinv_config* c = &p->c;
  p->nproc = nproc;
#ifdef _OPENMP
  p->nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  p->nthreads = 1;
#endif
  int n = c->nfft, nh = n / 2 + 1, km = c->k_max, T = c->ntrc;
  size_t G = (size_t)nproc * c->nchains;
  p->flt = (double*)malloc(sizeof(double) * (size_t)nh * T);
  orc_init_filter(c, p->flt);
  p->plan = fft_plan_create(n);
  p->rng = (mt_state*)malloc(sizeof(mt_state) * (size_t)nproc);
  p->k = (int32_t*)calloc(G, sizeof(int32_t));
  p->z = (double*)calloc(G * (km - 1), sizeof(double));
  p->dvp = (double*)calloc(G * km, sizeof(double));
  p->dvs = (double*)calloc(G * km, sizeof(double));
  p->sig = (double*)calloc(G * T, sizeof(double));
  p->rft = (double*)calloc(G * (size_t)n * T, sizeof(double));
  p->logl = (double*)calloc(G, sizeof(double));
  p->temps = (double*)calloc(G, sizeof(double));
  pt_types(p);
  p->nprop = (int64_t*)calloc((size_t)p->ntype, sizeof(int64_t));
  p->naccept = (int64_t*)calloc((size_t)p->ntype, sizeof(int64_t));
  p->cap_hist = 0;
  p->likelihood_hist = NULL;
  p->nk = (int64_t*)calloc((size_t)km, sizeof(int64_t));
  p->nz = (int64_t*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1), sizeof(int64_t));
  p->nsig = (int64_t*)calloc((size_t)(c->nbin_sig > 0 ? c->nbin_sig : 1) * T, sizeof(int64_t));
  p->namp = (int64_t*)calloc((size_t)(c->nbin_amp > 0 ? c->nbin_amp : 1) * c->nsmp * T, sizeof(int64_t));
  p->nvpz = (int64_t*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1) * (c->nbin_vp > 0 ? c->nbin_vp : 1), sizeof(int64_t));
  p->nvsz = (int64_t*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1) * (c->nbin_vs > 0 ? c->nbin_vs : 1), sizeof(int64_t));
  p->nvpvsz = (int64_t*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1) * (c->nbin_vpvs > 0 ? c->nbin_vpvs : 1), sizeof(int64_t));
  p->vp_mean = (double*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1), sizeof(double));
  p->vs_mean = (double*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1), sizeof(double));
  p->vpvs_mean = (double*)calloc((size_t)(c->nbin_z > 0 ? c->nbin_z : 1), sizeof(double));
#pragma omp parallel for schedule(dynamic, 1) num_threads(p->nthreads)
  for (int r = 0; r < nproc; ++r) init_rank(p, r);
  p->n_eval = (int64_t)G;
  return p;
}

void orc_pt_destroy(orc_pt* p) {
  if (!p) return;
  free(p->flt); fft_plan_destroy(p->plan); free(p->rng); free(p->k); free(p->z); free(p->dvp); free(p->dvs);
  free(p->sig); free(p->rft); free(p->logl); free(p->temps); free(p->nprop); free(p->naccept);
  free(p->likelihood_hist); free(p->nk); free(p->nz); free(p->nsig); free(p->namp); free(p->nvpz); free(p->nvsz);
  free(p->nvpvsz); free(p->vp_mean); free(p->vs_mean); free(p->vpvs_mean); free(p);
}

/* one chain step, pt_mcmc.f90:54-192.  Returns flag: -1 null proposal, 0 rejected, 1 accepted. */
static int mcmc_step(orc_pt* p, rf_ws* w, double* mis, double* prop_rft, int rank, int ic, double temp, int* itype_out,
                     int64_t* n_eval) {
  const rfinv_config* c = &p->c;
  mt_state* s = &p->rng[rank];
  int km = c->k_max, T = c->ntrc, n = c->nfft;
  size_t g = (size_t)rank * c->nchains + ic;
  double *cz = p->z + g * (km - 1), *cdvp = p->dvp + g * km, *cdvs = p->dvs + g * km, *csig = p->sig + g * T;
  double prop_dvp[256], prop_dvs[256], prop_z[256], prop_sig[64];
  double log_prior12 = 0.0;
  int prop_k = p->k[g], null_flag = 0, itarget;
  memcpy(prop_dvp, cdvp, sizeof(double) * (size_t)km);
  memcpy(prop_dvs, cdvs, sizeof(double) * (size_t)km);
  memcpy(prop_z, cz, sizeof(double) * (size_t)(km - 1));
  prop_z[km - 1] = 0.0;
  memcpy(prop_sig, csig, sizeof(double) * (size_t)T);
  int itype = (int)(orc_grnd(s) * p->ntype) + 1;
  *itype_out = itype;
  if (itype == p->it_birth) {
    prop_k = prop_k + 1;
    if (prop_k < km) {
      if (c->prior_mode == 1) { prop_dvp[prop_k - 1] = laplace(s) * c->dvp_prior; prop_dvs[prop_k - 1] = laplace(s) * c->dvs_prior; }
      else if (c->prior_mode == 2) { prop_dvp[prop_k - 1] = gauss(s) * c->dvp_prior; prop_dvs[prop_k - 1] = gauss(s) * c->dvs_prior; }
      prop_z[prop_k - 1] = c->z_min + orc_grnd(s) * (c->z_max - c->z_min);
    } else null_flag = 1;
  } else if (itype == p->it_death) {
    prop_k = prop_k - 1;
    if (prop_k >= c->k_min) {
      itarget = (int)(orc_grnd(s) * (prop_k + 1)) + 1;
      for (int il = itarget; il <= prop_k; ++il) {
        prop_dvp[il - 1] = cdvp[il]; prop_dvs[il - 1] = cdvs[il]; prop_z[il - 1] = cz[il];
      }
      prop_dvp[prop_k] = 0.0; prop_dvs[prop_k] = 0.0; prop_z[prop_k] = 0.0;
    } else null_flag = 1;
  } else if (itype == p->it_z) {
    itarget = (int)(orc_grnd(s) * prop_k) + 1;
    prop_z[itarget - 1] = prop_z[itarget - 1] + gauss(s) * c->dev_z;
    if (prop_z[itarget - 1] < c->z_min || prop_z[itarget - 1] > c->z_max) null_flag = 1;
  } else if (itype == p->it_dvs) {
    itarget = (int)(orc_grnd(s) * (prop_k + 1)) + 1;
    if (itarget == prop_k + 1) itarget = km;
    prop_dvs[itarget - 1] = prop_dvs[itarget - 1] + gauss(s) * c->dev_dvs;
    log_prior12 = log_prior_ratio(prop_dvs[itarget - 1], cdvs[itarget - 1], c->dvs_prior, c->prior_mode);
  } else if (itype == p->it_dvp) {
    itarget = (int)(orc_grnd(s) * (prop_k + 1)) + 1;
    if (itarget == prop_k + 1) itarget = km;
    prop_dvp[itarget - 1] = prop_dvp[itarget - 1] + gauss(s) * c->dev_dvp;
    log_prior12 = log_prior_ratio(prop_dvp[itarget - 1], cdvp[itarget - 1], c->dvp_prior, c->prior_mode);
  } else if (itype == p->it_sig) {
    itarget = p->isig_trc[(int)(orc_grnd(s) * p->nsig_trc)];
    prop_sig[itarget] = prop_sig[itarget] + gauss(s) * c->dev_sig;
    if (prop_sig[itarget] < c->sig_min[itarget] || prop_sig[itarget] > c->sig_max[itarget]) null_flag = 1;
  }
  if (!null_flag) {
    double alpha[258], beta[258], rho[258], h[258];
    int nlay;
    if (!orc_format_model(c, prop_k, prop_z, prop_dvp, prop_dvs, &nlay, alpha, beta, rho, h)) null_flag = 1;
  }
  if (null_flag) return -1;
  int fwd = itype != p->it_sig;
  double ll2;
  eval_chain(p, w, mis, prop_k, prop_z, prop_dvp, prop_dvs, prop_sig, fwd, p->rft + g * (size_t)n * T, &ll2, prop_rft);
  if (fwd) (*n_eval)++;
  /* judge_mcmc, pt_mcmc.f90:600-621 */
  double del_s = (ll2 - p->logl[g]) / temp + log_prior12, r;
  do { r = orc_grnd(s); } while (!(r >= 2.220446049250313e-16));
  int yn = log(r) <= del_s;
  if (yn) {
    p->logl[g] = ll2;
    p->k[g] = prop_k;
    memcpy(cdvp, prop_dvp, sizeof(double) * (size_t)km);
    memcpy(cdvs, prop_dvs, sizeof(double) * (size_t)km);
    memcpy(cz, prop_z, sizeof(double) * (size_t)(km - 1));
    memcpy(csig, prop_sig, sizeof(double) * (size_t)T);
    memcpy(p->rft + g * (size_t)n * T, prop_rft, sizeof(double) * (size_t)n * T);
  }
  return yn;
}

/* posterior bookkeeping of one cold chain, pt_mcmc.f90:204-286 (serial, called in chain order).
 * Bin indices that the reference would write out of bounds are dropped/clamped here. */
static void record_chain(orc_pt* p, size_t g) {
  const rfinv_config* c = &p->c;
  int km = c->k_max, T = c->ntrc, n = c->nfft, S = c->nsmp;
  double dbin_amp = (c->amp_max - c->amp_min) / c->nbin_amp, dbin_vp = (c->vp_max - c->vp_min) / c->nbin_vp;
  double dbin_vs = (c->vs_max - c->vs_min) / c->nbin_vs, dbin_z = (c->z_max - 0.0) / c->nbin_z;
  double dbin_vpvs = (c->vpvs_max - c->vpvs_min) / c->nbin_vpvs;
  const double *z = p->z + g * (km - 1), *sig = p->sig + g * T;
  int k = p->k[g];
  p->nmod++;
  p->nk[k - 1]++;
  for (int t = 0; t < T; ++t)
    if (p->sig_mode[t]) {
      double dbs = (c->sig_max[t] - c->sig_min[t]) / c->nbin_sig;
      int ibin = (int)((sig[t] - c->sig_min[t]) / dbs) + 1;
      if (ibin >= 1 && ibin <= c->nbin_sig) p->nsig[(size_t)t * c->nbin_sig + ibin - 1]++;
    }
  for (int il = 1; il <= k - 1; ++il) { /* pt_mcmc.f90:224-227: first k-1 interfaces only */
    int ibin = (int)((z[il - 1] - c->z_min) / dbin_z) + 1;
    if (ibin >= 1 && ibin <= c->nbin_z) p->nz[ibin - 1]++;
  }
  double alpha[258], beta[258], rho[258], h[258];
  int nlay;
  orc_format_model(c, k, z, p->dvp + g * km, p->dvs + g * km, &nlay, alpha, beta, rho, h);
  double tmpz = 0.0;
  for (int il = 1; il <= nlay; ++il) {
    int iz1 = (int)(tmpz / dbin_z) + 1;
    int iz2 = il < nlay ? (int)((tmpz + h[il - 1]) / dbin_z) + 1 : c->nbin_z + 1;
    int ivp = (int)((alpha[il - 1] - c->vp_min) / dbin_vp) + 1;
    int ivs = (int)((beta[il - 1] - c->vs_min) / dbin_vs) + 1;
    if (ivs < 1) ivs = 1;
    int ivpvs = (int)(((alpha[il - 1] / beta[il - 1]) - c->vpvs_min) / dbin_vpvs) + 1;
    if (ivpvs < 1) ivpvs = 1;
    if (ivpvs > c->nbin_vpvs) ivpvs = c->nbin_vpvs;
    if (ivp > c->nbin_vp) ivp = c->nbin_vp;
    if (ivs > c->nbin_vs) ivs = c->nbin_vs;
    if (iz2 > c->nbin_z + 1) iz2 = c->nbin_z + 1;
    for (int iz = iz1; iz <= iz2 - 1; ++iz) {
      p->nvpz[(size_t)(ivp - 1) * c->nbin_z + iz - 1]++;
      p->vp_mean[iz - 1] += alpha[il - 1];
      p->nvsz[(size_t)(ivs - 1) * c->nbin_z + iz - 1]++;
      p->nvpvsz[(size_t)(ivpvs - 1) * c->nbin_z + iz - 1]++;
      if (beta[il - 1] > 0.0) {
        p->vpvs_mean[iz - 1] += alpha[il - 1] / beta[il - 1];
        p->vs_mean[iz - 1] += beta[il - 1];
      } else { /* pt_mcmc.f90:260-263: assignment, not accumulation */
        p->vpvs_mean[iz - 1] = c->vpvs_min;
        p->vs_mean[iz - 1] = c->vs_min;
      }
    }
    tmpz = tmpz + h[il - 1];
  }
  const double* rft = p->rft + g * (size_t)n * T;
  for (int t = 0; t < T; ++t)
    for (int it = 0; it < S; ++it) {
      int ibin = (int)((rft[(size_t)t * n + it] - c->amp_min) / dbin_amp) + 1;
      if (ibin < 1) ibin = 1; else if (ibin > c->nbin_amp) ibin = c->nbin_amp;
      p->namp[((size_t)t * S + it) * c->nbin_amp + ibin - 1]++;
    }
}

/* pt_control, pt_mcmc.f90:468-576, n_iter more iterations for all virtual ranks.
 * Optional logs: flags[n_iter][G] (-1 null, 0 rejected, 1 accepted), itypes[n_iter][G] (1-based),
 * swaps[n_iter][3] (itarget1, itarget2, accepted).  record != 0 enables the posterior bookkeeping. */
int32_t orc_pt_run(orc_pt* p, int32_t n_iter, int8_t* flags, int8_t* itypes, int32_t* swaps, int32_t record) {
  const rfinv_config* c = &p->c;
  int nc = c->nchains, n = c->nfft, T = c->ntrc;
  size_t G = (size_t)p->nproc * nc;
  int new_cap = p->it_done + n_iter;
  if (new_cap > p->cap_hist) {
    p->likelihood_hist = (double*)realloc(p->likelihood_hist, sizeof(double) * (size_t)new_cap);
    for (int i = p->cap_hist; i < new_cap; ++i) p->likelihood_hist[i] = 0.0;
    p->cap_hist = new_cap;
  }
  int8_t* fl = (int8_t*)malloc(G);
  int8_t* ty = (int8_t*)malloc(G);
  int64_t n_eval_tot = 0;
#pragma omp parallel num_threads(p->nthreads) reduction(+ : n_eval_tot)
  {
    rf_ws* w = ws_create(n);
    double* mis = (double*)malloc(sizeof(double) * (size_t)c->nsmp);
    double* prop_rft = (double*)malloc(sizeof(double) * (size_t)n * T);
    for (int step = 0; step < n_iter; ++step) {
      int it = p->it_done + step + 1; /* 1-based iteration number */
#pragma omp for schedule(dynamic, 1)
      for (int r = 0; r < p->nproc; ++r)
        for (int ic = 0; ic < nc; ++ic) {
          size_t g = (size_t)r * nc + ic;
          int itype;
          fl[g] = (int8_t)mcmc_step(p, w, mis, prop_rft, r, ic, p->temps[g], &itype, &n_eval_tot);
          ty[g] = (int8_t)itype;
        }
#pragma omp single
      {
        for (size_t g = 0; g < G; ++g) { /* counters, pt_mcmc.f90:196-201, in chain order */
          if (p->temps[g] <= 1.0 + COLD_EPS) {
            p->nprop[ty[g] - 1]++;
            if (fl[g] == 1) p->naccept[ty[g] - 1]++;
            p->likelihood_hist[it - 1] += p->logl[g];
            if (record && it > c->nburn && it % c->ncorr == 0) record_chain(p, g);
          }
        }
        if (flags) memcpy(flags + (size_t)step * G, fl, G);
        if (itypes) memcpy(itypes + (size_t)step * G, ty, G);
        if (nc >= 2) { /* pt_mcmc.f90:498-571 */
          int n_all = (int)G;
          mt_state* s0 = &p->rng[0];
          int itarget1 = (int)(orc_grnd(s0) * n_all), itarget2;
          do { itarget2 = (int)(orc_grnd(s0) * n_all); } while (itarget2 == itarget1);
          int rank1 = itarget1 / nc;
          double temp1 = p->temps[itarget1], temp2 = p->temps[itarget2];
          double e1 = p->logl[itarget1], e2 = p->logl[itarget2];
          double del_s = (e2 - e1) * (1.0 / temp1 - 1.0 / temp2); /* judge_pt, pt_mcmc.f90:580-595 */
          int yn = log(orc_grnd(&p->rng[rank1])) <= del_s;
          if (yn) { p->temps[itarget2] = temp1; p->temps[itarget1] = temp2; }
          if (swaps) { swaps[(size_t)step * 3] = itarget1; swaps[(size_t)step * 3 + 1] = itarget2; swaps[(size_t)step * 3 + 2] = yn; }
        }
      }
    }
    free(prop_rft); free(mis); ws_destroy(w);
  }
  p->it_done += n_iter;
  p->n_eval += n_eval_tot;
  free(fl); free(ty);
  return 0;
}

/* ---- getters (any pointer may be NULL) ---- */
int32_t orc_pt_ntype(const orc_pt* p) { return p->ntype; }
int64_t orc_pt_n_eval(const orc_pt* p) { return p->n_eval; }
void orc_pt_get_state(const orc_pt* p, int32_t* k, double* z, double* dvp, double* dvs, double* sig, double* logl,
                      double* temps, double* rft) {
  const rfinv_config* c = &p->c;
  size_t G = (size_t)p->nproc * c->nchains;
  if (k) memcpy(k, p->k, sizeof(int32_t) * G);
  if (z) memcpy(z, p->z, sizeof(double) * G * (c->k_max - 1));
  if (dvp) memcpy(dvp, p->dvp, sizeof(double) * G * c->k_max);
  if (dvs) memcpy(dvs, p->dvs, sizeof(double) * G * c->k_max);
  if (sig) memcpy(sig, p->sig, sizeof(double) * G * c->ntrc);
  if (logl) memcpy(logl, p->logl, sizeof(double) * G);
  if (temps) memcpy(temps, p->temps, sizeof(double) * G);
  if (rft) memcpy(rft, p->rft, sizeof(double) * G * (size_t)c->nfft * c->ntrc);
}
void orc_pt_get_counters(const orc_pt* p, int64_t* nprop, int64_t* naccept, double* likelihood_hist, int64_t* nmod) {
  if (nprop) memcpy(nprop, p->nprop, sizeof(int64_t) * (size_t)p->ntype);
  if (naccept) memcpy(naccept, p->naccept, sizeof(int64_t) * (size_t)p->ntype);
  if (likelihood_hist) memcpy(likelihood_hist, p->likelihood_hist, sizeof(double) * (size_t)p->it_done);
  if (nmod) *nmod = p->nmod;
}
/* histograms, layouts: nk[k_max], nz[nbin_z], nsig[ntrc][nbin_sig], namp[ntrc][nsmp][nbin_amp],
 * nvpz[nbin_vp][nbin_z], nvsz[nbin_vs][nbin_z], nvpvsz[nbin_vpvs][nbin_z], *_mean[nbin_z] */
void orc_pt_get_hist(const orc_pt* p, int64_t* nk, int64_t* nz, int64_t* nsig, int64_t* namp, int64_t* nvpz,
                     int64_t* nvsz, int64_t* nvpvsz, double* vp_mean, double* vs_mean, double* vpvs_mean) {
  const rfinv_config* c = &p->c;
  if (nk) memcpy(nk, p->nk, sizeof(int64_t) * (size_t)c->k_max);
  if (nz) memcpy(nz, p->nz, sizeof(int64_t) * (size_t)c->nbin_z);
  if (nsig) memcpy(nsig, p->nsig, sizeof(int64_t) * (size_t)c->nbin_sig * c->ntrc);
  if (namp) memcpy(namp, p->namp, sizeof(int64_t) * (size_t)c->nbin_amp * c->nsmp * c->ntrc);
  if (nvpz) memcpy(nvpz, p->nvpz, sizeof(int64_t) * (size_t)c->nbin_z * c->nbin_vp);
  if (nvsz) memcpy(nvsz, p->nvsz, sizeof(int64_t) * (size_t)c->nbin_z * c->nbin_vs);
  if (nvpvsz) memcpy(nvpvsz, p->nvpvsz, sizeof(int64_t) * (size_t)c->nbin_z * c->nbin_vpvs);
  if (vp_mean) memcpy(vp_mean, p->vp_mean, sizeof(double) * (size_t)c->nbin_z);
  if (vs_mean) memcpy(vs_mean, p->vs_mean, sizeof(double) * (size_t)c->nbin_z);
  if (vpvs_mean) memcpy(vpvs_mean, p->vpvs_mean, sizeof(double) * (size_t)c->nbin_z);
}
