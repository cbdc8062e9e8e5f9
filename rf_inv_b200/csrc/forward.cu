// Fused forward kernel: format_model -> layered-medium propagator response -> Gaussian filter ->
// shared-memory inverse FFT -> time shift / normalisation -> misfit.  One CTA per (model, ray).
//
// Replaces, for a whole batch of chains at once (paths relative to the reference checkout):
//   format_model      src/model.f90:175-290
//   calc_seis         src/forward.f90:212-344   (e_inverse :350-380, layer_matrix_sol :385-421,
//                                                layer_matrix_liq :424-442)
//   calc_rf           src/forward.f90:123-208   (water_level_decon :447-470, direct_arrival :474-491)
//   FFTW c2r          src/fftw.f90:44, executed at src/forward.f90:172,200
//   misfit            src/likelihood.f90:87
//
// Arithmetic re-design.  The reference multiplies dense complex 4x4 layer
// matrices P_l and then E^-1 * prod(P_l); only rows 3,4 x columns 1,2 (and 4 with a water layer) of the
// result are used.  P_l has a fixed checkerboard real/imaginary structure: P = D S B S^-1 D^-1 with
// D = diag(1,i,1,i), S = diag(1,1,w,w) (w = angular frequency) and B REAL and free of explicit w.  So we
// propagate two REAL 4-vectors (columns e1 and the free-surface / water-layer combination of e2,e4)
// from the top layer down, y <- B_l y: 32 FMA per layer-frequency instead of 64 complex MACs, and the 10
// distinct entries of B_l are two-term combinations of cos/sin of (w xi h, w eta h).  The four
// trigonometric values per layer-frequency come from one sincos pair per thread and layer, advanced
// across the thread's J frequencies by a fixed rotation (frequencies of a thread are equally spaced).
#include <cstdlib>
#include "rfinv_common.cuh"

#ifdef RFINV_PHASE_TIMING
// Debug build only (tools/phase_timing.py): per-phase cycle totals of forward_kernel, summed over CTAs.
__device__ unsigned long long g_phase[16];
#define PHASE_MARK(i)                                                            \
  do {                                                                           \
    if (tid == 0) {                                                              \
      const long long now__ = clock64();                                         \
      atomicAdd(&g_phase[i], (unsigned long long)(now__ - t_phase__));           \
      t_phase__ = now__;                                                         \
    }                                                                            \
  } while (0)
#define PHASE_INIT() long long t_phase__ = clock64(); if (tid == 0) atomicAdd(&g_phase[15], 1ULL)
extern "C" int rfinv_debug_get_phases(unsigned long long* out) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_phase, sizeof(g_phase));
  unsigned long long zero[16] = {0};
  cudaMemcpyToSymbol(g_phase, zero, sizeof(zero));
  return e == cudaSuccess ? 0 : 2;
}
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_INIT() do { } while (0)
#endif

namespace {

// Per (model, ray, solid layer) constants of the real propagator B_l, written by prep_kernel.
struct LayerConst {
  double thx, the;            // domg*xi*h, domg*eta*h : phase advance per frequency bin
  double cbx, sbx, cbe, sbe;  // rotation by B*thx, B*the (B = threads per CTA = frequency stride of a thread)
  double g, bp;               // 2 beta^2 p^2, 1 - 2 beta^2 p^2
  double c12a, c12b, c21a, c21b, c13a, c13b, c24a, c24b, c31a, c31b, c42a, c42b;
  double c14, c41;
};
constexpr int LC_DOUBLES = sizeof(LayerConst) / sizeof(double);  // 22

struct HalfSpace {
  double e11, e12, e13, e14, e21, e22, e23, e24;  // E^-1 rows 3,4 with the 1/w factors removed
};

// Per (model, ray) constants written by prep_kernel.
struct RayConst {
  HalfSpace hs;
  double thw, cbw, sbw, rw;   // water layer: phase per bin, stride rotation, rho_w/xi_w
  double tp;                  // direct-arrival delay (src/forward.f90:474-491)
  double2 edge[4];            // fr, fv at the DC pseudo-frequency and at Nyquist
  int k, valid;
};
constexpr int RC_DOUBLES = sizeof(RayConst) / sizeof(double);

__device__ __forceinline__ void rot(double& c, double& s, double cb, double sb) {
  double c2 = c * cb - s * sb;
  double s2 = s * cb + c * sb;
  c = c2;
  s = s2;
}

// y <- B y for the two propagated vectors; (c1,s1) = cos/sin(w xi h), (c2,s2) = cos/sin(w eta h)
__device__ __forceinline__ void layer_step(const LayerConst& L, double c1, double s1, double c2, double s2, double* ya,
                                           double* yb) {
  const double d = c1 - c2;
  const double b11 = fma(L.g, c1, L.bp * c2);
  const double b22 = fma(L.bp, c1, L.g * c2);
  const double b12 = fma(L.c12a, s1, L.c12b * s2);
  const double b21 = fma(L.c21a, s1, L.c21b * s2);
  const double b13 = fma(L.c13a, s1, L.c13b * s2);
  const double b24 = fma(L.c24a, s1, L.c24b * s2);
  const double b31 = fma(L.c31a, s1, L.c31b * s2);
  const double b42 = fma(L.c42a, s1, L.c42b * s2);
  const double b14 = L.c14 * d;
  const double b41 = L.c41 * d;
  {
    const double y1 = ya[0], y2 = ya[1], y3 = ya[2], y4 = ya[3];
    ya[0] = fma(b14, y4, fma(b13, y3, fma(b12, y2, b11 * y1)));
    ya[1] = fma(b24, y4, fma(-b14, y3, fma(b22, y2, b21 * y1)));
    ya[2] = fma(-b21, y4, fma(b11, y3, fma(-b41, y2, b31 * y1)));
    ya[3] = fma(b22, y4, fma(-b12, y3, fma(b42, y2, b41 * y1)));
  }
  {
    const double y1 = yb[0], y2 = yb[1], y3 = yb[2], y4 = yb[3];
    yb[0] = fma(b14, y4, fma(b13, y3, fma(b12, y2, b11 * y1)));
    yb[1] = fma(b24, y4, fma(-b14, y3, fma(b22, y2, b21 * y1)));
    yb[2] = fma(-b21, y4, fma(b11, y3, fma(-b41, y2, b31 * y1)));
    yb[3] = fma(b22, y4, fma(-b12, y3, fma(b42, y2, b41 * y1)));
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Boundary conditions (src/forward.f90:267-287) in the scaled real basis, then the sign conventions of
// calc_rf (src/forward.f90:145-146): fr = conj(ur), fv = -conj(uz).
__device__ __forceinline__ void surface_response(const HalfSpace& H, const double* ya, const double* yb, double cw,
                                                 int ipha, double2& fr, double2& fv) {
  const double2 A3 = make_double2(fma(H.e11, ya[0], H.e14 * ya[3]), fma(-H.e12, ya[1], H.e13 * ya[2]));
  const double2 A4 = make_double2(fma(H.e21, ya[0], -H.e24 * ya[3]), fma(H.e22, ya[1], H.e23 * ya[2]));
  const double2 B3 = make_double2(fma(H.e11, yb[0], H.e14 * yb[3]), fma(-H.e12, yb[1], H.e13 * yb[2]));
  const double2 B4 = make_double2(fma(H.e21, yb[0], -H.e24 * yb[3]), fma(H.e22, yb[1], H.e23 * yb[2]));
  const double2 p = cmul(A3, B4), q = cmul(B3, A4);
  const double2 dl = make_double2(p.x - q.x, p.y - q.y);
  const double rn = 1.0 / (dl.x * dl.x + dl.y * dl.y);
  const double2 inv = make_double2(dl.x * rn, -dl.y * rn);
  double2 ur, uz;
  if (ipha >= 0) {
    ur = cmul(B4, inv);
    const double2 w = cmul(A4, inv);  // uz = -i * A4/Delta * cos_w
    uz = make_double2(w.y * cw, -w.x * cw);
  } else {
    const double2 w0 = cmul(B3, inv);
    ur = make_double2(-w0.x, -w0.y);
    const double2 w = cmul(A3, inv);  // uz = +i * A3/Delta * cos_w
    uz = make_double2(-w.y * cw, w.x * cw);
  }
  fr = make_double2(ur.x, -ur.y);
  fv = make_double2(-uz.x, uz.y);
}

// ------------------------------------------------------------------------------------------------
// prep_kernel: one thread per (model, ray).  format_model (src/model.f90:175-290), the per-layer
// constants of the real propagator for this ray, half-space / water-layer constants, the direct-arrival
// delay, and the two bins that do not fit the regular frequency grid of forward_kernel: the DC
// pseudo-frequency omega = 1.0e-5 (src/forward.f90:246-248) and the Nyquist bin.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prep_kernel(const DevConfig cfg, const ModelBatch mb, double* __restrict__ lc_out,
                                                   double* __restrict__ rc_out, uint8_t* __restrict__ is_valid,
                                                   int n_items, int ntr_eff, int nthr_fwd) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  if (mb.n_active_dev && item >= *mb.n_active_dev * ntr_eff) return;
  const int ci = item / ntr_eff, t0 = item - ci * ntr_eff;
  const int c = mb.active ? mb.active[ci] : ci;
  const int km = cfg.k_max, C = mb.C;
  int k = mb.k[c];
  k = k < 1 ? 1 : (k > km - 1 ? km - 1 : k);
  const double p = cfg.rayp[t0];
  const int ipha = cfg.ipha[t0];
  double z[RFINV_MAX_K], dp[RFINV_MAX_K], ds[RFINV_MAX_K];
  for (int i = 0; i < k; ++i) {
    z[i] = mb.z[(size_t)i * C + c]; dp[i] = mb.dvp[(size_t)i * C + c]; ds[i] = mb.dvs[(size_t)i * C + c];
  }
  for (int i = 1; i < k; ++i) {  // src/sort.f90:34-68 (any correct sort; keys are distinct)
    const double a = z[i], b = dp[i], d = ds[i];
    int m = i - 1;
    while (m >= 0 && z[m] > a) { z[m + 1] = z[m]; dp[m + 1] = dp[m]; ds[m + 1] = ds[m]; --m; }
    z[m + 1] = a; dp[m + 1] = b; ds[m + 1] = d;
  }
  RayConst R;
  // water layer (src/model.f90:201-207, src/forward.f90:424-442) and start vectors of the two edge bins
  double ya0[4] = {1.0, 0.0, 0.0, 0.0}, yb0[4], ya1[4] = {1.0, 0.0, 0.0, 0.0}, yb1[4], cw0 = 1.0, cw1 = 1.0;
  const double nyq = (double)(cfg.nfft / 2);
  if (cfg.sdep > 0.0) {
    const double aw = 1.5, rhow = 1.0, hw = cfg.sdep;
    const double xiw = sqrt(1.0 / (aw * aw) - p * p);
    R.thw = cfg.domg * xiw * hw;
    sincos((double)nthr_fwd * R.thw, &R.sbw, &R.cbw);
    R.rw = rhow / xiw;
    double sw;
    sincos((double)1.0e-5f * xiw * hw, &sw, &cw0);
    yb0[0] = 0.0; yb0[1] = cw0; yb0[2] = 0.0; yb0[3] = -R.rw * sw;
    sincos(nyq * R.thw, &sw, &cw1);
    yb1[0] = 0.0; yb1[1] = cw1; yb1[2] = 0.0; yb1[3] = -R.rw * sw;
  } else {
    R.thw = 0.0; R.cbw = 1.0; R.sbw = 0.0; R.rw = 0.0;
    yb0[0] = 0.0; yb0[1] = 1.0; yb0[2] = 0.0; yb0[3] = 0.0;
    yb1[0] = 0.0; yb1[1] = 1.0; yb1[2] = 0.0; yb1[3] = 0.0;
  }
  bool valid = true;
  double tp = 0.0;
  LayerConst* lc = reinterpret_cast<LayerConst*>(lc_out) + (size_t)item * km;
  const double p2 = __dmul_rn(p, p);
  for (int l = 0; l <= k; ++l) {
    double zc, h, dvs_l, dvp_l;
    if (l == 0) { zc = __dmul_rn(0.5, __dadd_rn(cfg.sdep, z[0])); h = __dsub_rn(z[0], cfg.sdep); dvs_l = ds[0]; dvp_l = dp[0]; }
    else if (l < k) { zc = __dmul_rn(0.5, __dadd_rn(z[l], z[l - 1])); h = __dsub_rn(z[l], z[l - 1]); dvs_l = ds[l]; dvp_l = dp[l]; }
    else { zc = __dmul_rn(0.5, __dadd_rn(cfg.z_max, z[k - 1])); h = 999.0;
           dvs_l = mb.dvs[(size_t)(km - 1) * C + c]; dvp_l = mb.dvp[(size_t)(km - 1) * C + c]; }
    double a, b;
    bool ok = layer_velocity(cfg, zc, dvs_l, dvp_l, a, b);
    if (l == 0) ok = ok && !(h < __dmul_rn(0.125, a));     // src/model.f90:229
    else if (l < k) ok = ok && !(h < cfg.h_min);           // src/model.f90:257
    valid = valid && ok;
    const double rho = vp_to_rho(a);
    const double beta2 = __dmul_rn(b, b);
    const double bp = 1.0 - 2.0 * beta2 * p2;
    const double eta = sqrt(__dsub_rn(__ddiv_rn(1.0, beta2), p2));              // src/forward.f90:395
    const double xi = sqrt(__dsub_rn(__ddiv_rn(1.0, __dmul_rn(a, a)), p2));    // src/forward.f90:396
    if (l < k) {
      if (cfg.deconv_mode == 0) tp = __dadd_rn(tp, __dmul_rn(h, ipha == 1 ? xi : eta));  // src/forward.f90:489-491
      LayerConst L;
      L.thx = cfg.domg * xi * h;
      L.the = cfg.domg * eta * h;
      sincos((double)nthr_fwd * L.thx, &L.sbx, &L.cbx);
      sincos((double)nthr_fwd * L.the, &L.sbe, &L.cbe);
      L.g = 2.0 * beta2 * p2;
      L.bp = bp;
      L.c12a = -p * bp / xi;            L.c12b = 2.0 * p * beta2 * eta;
      L.c21a = 2.0 * p * beta2 * xi;    L.c21b = -p * bp / eta;
      L.c13a = p2 / (xi * rho);         L.c13b = eta / rho;
      L.c24a = xi / rho;                L.c24b = p2 / (eta * rho);
      L.c31a = -4.0 * rho * beta2 * beta2 * p2 * xi;   L.c31b = -rho * bp * bp / eta;
      L.c42a = -rho * bp * bp / xi;                    L.c42b = -4.0 * rho * beta2 * beta2 * p2 * eta;
      L.c14 = p / rho;
      L.c41 = 2.0 * beta2 * rho * p * bp;
      lc[l] = L;
      double c1, s1, c2, s2;
      sincos((double)1.0e-5f * xi * h, &s1, &c1);   // (omega*xi)*z with omega = 1.0e-5 (single precision literal)
      sincos((double)1.0e-5f * eta * h, &s2, &c2);
      layer_step(L, c1, s1, c2, s2, ya0, yb0);
      sincos(nyq * L.thx, &s1, &c1);
      sincos(nyq * L.the, &s2, &c2);
      layer_step(L, c1, s1, c2, s2, ya1, yb1);
    } else {  // half space: rows 3,4 of E^-1 (src/forward.f90:350-380) without their 1/omega factors
      R.hs.e11 = beta2 * p / a;
      R.hs.e12 = bp / (2.0 * a * xi);
      R.hs.e13 = p / (2.0 * rho * a * xi);
      R.hs.e14 = 1.0 / (2.0 * rho * a);
      R.hs.e21 = bp / (2.0 * b * eta);
      R.hs.e22 = b * p;
      R.hs.e23 = 1.0 / (2.0 * rho * b);
      R.hs.e24 = p / (2.0 * rho * b * eta);
    }
  }
  surface_response(R.hs, ya0, yb0, cw0, ipha, R.edge[0], R.edge[1]);
  surface_response(R.hs, ya1, yb1, cw1, ipha, R.edge[2], R.edge[3]);
  R.tp = tp;
  R.k = k;
  R.valid = valid;
  reinterpret_cast<RayConst*>(rc_out)[item] = R;
  if (is_valid && t0 == 0) is_valid[c] = (uint8_t)valid;
}

// exp(+2 pi i m / n) for m < 3n/4 from the quarter-wave table twq[r] = exp(+2 pi i r / n), r < n/4
__device__ __forceinline__ double2 twiddle(const double2* twq, int m, int qmask, int qshift) {
  const double2 w = twq[m & qmask];
  const int q = m >> qshift;
  return q == 0 ? w : (q == 1 ? make_double2(-w.y, w.x) : make_double2(-w.x, -w.y));
}

// Stockham inverse FFT (sign +, unnormalised) of n complex points in shared memory, ping-pong x <-> y.
// Returns the buffer holding the result.
__device__ double2* fft_inverse(double2* x, double2* y, int n, int log2n, const double2* twq, int tid, int nthr) {
  int Ns = 1;
  if (log2n & 1) {  // one radix-2 pass (no twiddles at Ns = 1)
    const int half = n >> 1;
    for (int j = tid; j < half; j += nthr) {
      const double2 a = x[j], b = x[j + half];
      y[2 * j] = make_double2(a.x + b.x, a.y + b.y);
      y[2 * j + 1] = make_double2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
    double2* t = x; x = y; y = t;
    Ns = 2;
  }
  const int quarter = n >> 2, qmask = quarter - 1, qshift = log2n - 2;
  while (Ns < n) {
    const int tstep = n / (4 * Ns);
    for (int j = tid; j < quarter; j += nthr) {
      const int k = j & (Ns - 1);
      double2 v0 = x[j], v1 = x[j + quarter], v2 = x[j + 2 * quarter], v3 = x[j + 3 * quarter];
      if (Ns > 1) {
        v1 = cmul(v1, twiddle(twq, k * tstep, qmask, qshift));
        v2 = cmul(v2, twiddle(twq, 2 * k * tstep, qmask, qshift));
        v3 = cmul(v3, twiddle(twq, 3 * k * tstep, qmask, qshift));
      }
      const double2 a0 = make_double2(v0.x + v2.x, v0.y + v2.y), a1 = make_double2(v0.x - v2.x, v0.y - v2.y);
      const double2 a2 = make_double2(v1.x + v3.x, v1.y + v3.y);
      const double2 a3 = make_double2(-(v1.y - v3.y), v1.x - v3.x);  // i*(v1 - v3)
      const int j0 = ((j - k) << 2) + k;
      y[j0] = make_double2(a0.x + a2.x, a0.y + a2.y);
      y[j0 + Ns] = make_double2(a1.x + a3.x, a1.y + a3.y);
      y[j0 + 2 * Ns] = make_double2(a0.x - a2.x, a0.y - a2.y);
      y[j0 + 3 * Ns] = make_double2(a1.x - a3.x, a1.y - a3.y);
    }
    __syncthreads();
    double2* t = x; x = y; y = t;
    Ns <<= 2;
  }
  return x;
}

__device__ __forceinline__ double block_max(double v, double* scratch, int tid, int nthr) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((tid & 31) == 0) scratch[tid >> 5] = v;
  __syncthreads();
  const int nw = (nthr + 31) >> 5;
  double r = scratch[0];
  for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// forward_kernel: one CTA per (model, ray); thread `tid` owns frequency bins j = tid + m*B, m < J
// (B = blockDim.x, B*J = nfft/2).  Bins 0 (DC) and nfft/2 (Nyquist) come from prep_kernel.
// Shared memory: two FFT buffers of nfft complex doubles; the spectra alias the second one unless rays
// are common to all traces (then they must survive the per-trace FFTs); layer constants behind them.
// ------------------------------------------------------------------------------------------------
template <int J, int BMAX, int MINB>
__global__ void __launch_bounds__(BMAX, MINB) forward_kernel(const DevConfig cfg, const ModelBatch mb, const EvalOutputs out,
                                                             const double* __restrict__ lc_in,
                                                             const double* __restrict__ rc_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n = cfg.nfft, nh = cfg.nh, km = cfg.k_max, C = mb.C;
  const int ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  const int item = blockIdx.x;
  if (mb.n_active_dev && item >= *mb.n_active_dev * ntr_eff) return;   // grid is sized for the upper bound
  const int ci = item / ntr_eff, t0 = item - ci * ntr_eff;
  const int c = mb.active ? mb.active[ci] : ci;

  // layout: [buf1 | buf0 ... trig tables may extend past buf0 | (spectra if rays are common) | layer consts | ray consts]
  const int n_hi = nthr >> 4;                         // table split: tid = 16*hi + lo
  const int tab_per_layer = 2 * (16 + n_hi);          // double2 entries per layer: (xi | eta) x (lo | hi)
  const size_t tab_entries = (size_t)km * tab_per_layer;
  const size_t region0 = tab_entries > (size_t)n ? tab_entries : (size_t)n;   // buf0 region also hosts the tables
  double2* s_buf1 = reinterpret_cast<double2*>(smem_raw);
  double2* s_buf0 = s_buf1 + n + 4;
  double2* s_tab = s_buf0;                            // dead before buf0 is first written (Z build)
  double2* s_fr = cfg.ray_common ? s_buf0 + region0 : s_buf1;   // [nh] unfiltered radial spectrum (or deconvolved RF spectrum)
  double2* s_fv = s_fr + (nh + 1);                              // [nh] unfiltered vertical spectrum   (2*(nh+1) = n + 4)
  double* s_tail = reinterpret_cast<double*>(s_buf0 + region0 + (cfg.ray_common ? 2 * (nh + 1) : 0));
  LayerConst* s_lc = reinterpret_cast<LayerConst*>(s_tail);
  RayConst* s_rc = reinterpret_cast<RayConst*>(s_lc + km);
  double* s_red = reinterpret_cast<double*>(s_rc + 1);     // [32]
  double2* s_twq = reinterpret_cast<double2*>(s_red + 32); // [n/4] quarter-wave twiddles

  PHASE_INIT();
  // ---- stage the constants of this (model, ray): one round trip to L2/HBM ----
  int k = mb.k[c];
  k = k < 1 ? 1 : (k > km - 1 ? km - 1 : k);   // same clamp as prep_kernel
  {
    const double* src = rc_in + (size_t)item * RC_DOUBLES;
    double* dst = reinterpret_cast<double*>(s_rc);
    for (int i = tid; i < RC_DOUBLES; i += nthr) dst[i] = src[i];
    const double* src2 = lc_in + (size_t)item * km * LC_DOUBLES;
    double* dst2 = reinterpret_cast<double*>(s_lc);
    for (int i = tid; i < k * LC_DOUBLES; i += nthr) dst2[i] = src2[i];
    for (int i = tid; i < (n >> 2); i += nthr) s_twq[i] = cfg.tw[i];
  }
  __syncthreads();
  PHASE_MARK(0);
  // ---- two-level rotation tables: cos/sin(tid*theta) = rot(lo[tid & 15], hi[tid >> 4]) ----
  // per layer: [xi lo 0..15 | xi hi 0..n_hi-1 | eta lo 0..15 | eta hi 0..n_hi-1], entries (cos, sin).
  // One thread per (layer, angle, level): one sincos, then a chain of rotations (<= 15 steps, error ~1e-15).
  for (int task = tid; task < 4 * k; task += nthr) {
    const int l = task >> 2, which = (task >> 1) & 1, level = task & 1;
    const double th = (which ? s_lc[l].the : s_lc[l].thx) * (level ? 16.0 : 1.0);
    const int cnt = level ? n_hi : 16;
    double2* dst = s_tab + l * tab_per_layer + which * (16 + n_hi) + level * 16;
    double sb, cb;
    sincos(th, &sb, &cb);
    double cc = 1.0, ss = 0.0;
    dst[0] = make_double2(1.0, 0.0);
    for (int i = 1; i < cnt; ++i) {
      rot(cc, ss, cb, sb);
      dst[i] = make_double2(cc, ss);
    }
  }
  __syncthreads();
  PHASE_MARK(1);
  const int ipha = cfg.ipha[t0];

  // ---- propagator product over the solid layers, top down ----
  double ya[J][4], yb[J][4], cwv[J];
  {
    const double thw = s_rc->thw, cbw = s_rc->cbw, sbw = s_rc->sbw, rw = s_rc->rw;
    double cw, sw;
    sincos((double)tid * thw, &sw, &cw);
#pragma unroll
    for (int m = 0; m < J; ++m) {
      ya[m][0] = 1.0; ya[m][1] = 0.0; ya[m][2] = 0.0; ya[m][3] = 0.0;
      yb[m][0] = 0.0; yb[m][1] = cw; yb[m][2] = 0.0; yb[m][3] = -rw * sw;
      cwv[m] = cw;
      rot(cw, sw, cbw, sbw);
    }
  }
  const int t_lo = tid & 15, t_hi = 16 + (tid >> 4);
  for (int l = 0; l < k; ++l) {
    const LayerConst& L = s_lc[l];
    const double2* tab = s_tab + l * tab_per_layer;
    double c1, s1, c2, s2;
    {
      const double2 a = tab[t_lo], b = tab[t_hi], c = tab[16 + n_hi + t_lo], d = tab[16 + n_hi + t_hi];
      c1 = a.x; s1 = a.y; rot(c1, s1, b.x, b.y);
      c2 = c.x; s2 = c.y; rot(c2, s2, d.x, d.y);
    }
#pragma unroll
    for (int m = 0; m < J; ++m) {
      layer_step(L, c1, s1, c2, s2, ya[m], yb[m]);
      if (m + 1 < J) {
        rot(c1, s1, L.cbx, L.sbx);
        rot(c2, s2, L.cbe, L.sbe);
      }
    }
  }
  PHASE_MARK(2);
  {
    const HalfSpace H = s_rc->hs;
#pragma unroll
    for (int m = 0; m < J; ++m) {
      double2 fr, fv;
      surface_response(H, ya[m], yb[m], cwv[m], ipha, fr, fv);
      s_fr[tid + m * nthr] = fr;
      s_fv[tid + m * nthr] = fv;
    }
    if (tid == 0) {  // the two bins off the regular grid
      s_fr[0] = s_rc->edge[0]; s_fv[0] = s_rc->edge[1];
      s_fr[nh - 1] = s_rc->edge[2]; s_fv[nh - 1] = s_rc->edge[3];
    }
  }
  __syncthreads();
  PHASE_MARK(3);

  // ---- water-level deconvolution (src/forward.f90:148-153, 447-470): overwrites s_fr with rff ----
  if (cfg.deconv_mode == 1) {
    const double2* xs = ipha == 1 ? s_fv : s_fr;  // denominator spectrum
    const double2* ys = ipha == 1 ? s_fr : s_fv;
    double mx = -INFINITY;
    for (int j = tid; j < nh; j += nthr) mx = fmax(mx, xs[j].x * xs[j].x + xs[j].y * xs[j].y);
    mx = block_max(mx, s_red, tid, nthr);
    const double wlvl = 0.001 * mx;
    double2 keep[J + 1];
#pragma unroll
    for (int q = 0; q < J + 1; ++q) {
      const int j = tid + q * nthr;
      if (j < nh) {
        const double2 x = xs[j], y = ys[j];
        const double amp = x.x * x.x + x.y * x.y;
        const double d = fmax(amp, wlvl);
        keep[q] = make_double2((y.x * x.x + y.y * x.y) / d, (y.y * x.x - y.x * x.y) / d);
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < J + 1; ++q) {
      const int j = tid + q * nthr;
      if (j < nh) s_fr[j] = keep[q];
    }
    __syncthreads();
  }

  // ---- per trace: filter -> inverse FFT -> shift / normalise -> outputs ----
  const double tp = s_rc->tp;
  const int S = cfg.nsmp, Sp = cfg.nsmp_pad;
  const int t_begin = cfg.ray_common ? 0 : t0, t_end = cfg.ray_common ? cfg.ntrc : t0 + 1;
  for (int t = t_begin; t < t_end; ++t) {
    const double* __restrict__ flt = cfg.flt + (size_t)t * nh;
    const double2* src_r = (cfg.deconv_mode == 1 || ipha == 1) ? s_fr : s_fv;  // rff
    // packed spectrum Z = X_r + i X_v with Hermitian extension (c2r ignores Im of DC and Nyquist)
    for (int j = tid; j < nh; j += nthr) {
      const double f = flt[j];
      const double2 xr = make_double2(src_r[j].x * f, src_r[j].y * f);
      double2 xv = make_double2(0.0, 0.0);
      if (cfg.deconv_mode == 0) xv = make_double2(s_fv[j].x * f, s_fv[j].y * f);
      if (j == 0 || j == nh - 1) {
        s_buf0[j] = make_double2(xr.x, xv.x);
      } else {
        s_buf0[j] = make_double2(xr.x - xv.y, xr.y + xv.x);
        s_buf0[n - j] = make_double2(xr.x + xv.y, xv.x - xr.y);
      }
    }
    __syncthreads();
    PHASE_MARK(4);
    const double2* res = fft_inverse(s_buf0, s_buf1, n, cfg.log2n, s_twq, tid, nthr);
    PHASE_MARK(5);
    double fac = 1.0;
    if (cfg.deconv_mode == 0) {  // src/forward.f90:197-203
      double mx = -INFINITY;
      for (int i = tid; i < n; i += nthr) mx = fmax(mx, res[i].y);
      fac = block_max(mx, s_red, tid, nthr);
    }
    int npre;
    if (ipha == 1) npre = f_nint((-cfg.t_start - tp) / cfg.delta);   // src/forward.f90:177
    else npre = f_nint((-cfg.t_start + tp) / cfg.delta);             // src/forward.f90:186
    const int nout = out.rft_full ? n : S;
    double* mis = out.misfit + ((size_t)t * C + c) * Sp;
    double* smp_base = out.rft_smp;
    if (out.slot && ((out.slot[c] ^ out.slot_invert) & 1)) smp_base = out.rft_smp_alt;
    double* smp = smp_base ? smp_base + ((size_t)t * C + c) * S : nullptr;
    double* full = out.rft_full ? out.rft_full + ((size_t)c * cfg.ntrc + t) * n : nullptr;
    const double* __restrict__ obs = cfg.obs + (size_t)t * S;
    const double scale = cfg.deconv_mode == 0 ? 1.0 / fac : 1.0;
    const int nmask = n - 1;
    for (int i = tid; i < nout; i += nthr) {
      double v;
      if (ipha == 1) v = res[(i - npre) & nmask].x;       // src/forward.f90:178-184 (n is a power of two)
      else v = -res[(npre - i - 1) & nmask].x;            // src/forward.f90:187-193
      v *= scale;
      if (i < S) {
        mis[i] = v - obs[i];
        if (smp) smp[i] = v;
      }
      if (full) full[i] = v;
    }
    __syncthreads();
    PHASE_MARK(6);
  }
}

__global__ void format_model_kernel(const DevConfig cfg, const ModelBatch mb, int* nlay_out, double* alpha,
                                    double* beta, double* rho, double* h, uint8_t* is_valid) {
  // One thread per model; used by rfinv_format_model_batch (host diagnostics / parity tests).
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= mb.C) return;
  const int km = cfg.k_max, C = mb.C, k = mb.k[c], stride = km + 1;
  double z[RFINV_MAX_K], dp[RFINV_MAX_K], ds[RFINV_MAX_K];
  for (int i = 0; i < k; ++i) {
    z[i] = mb.z[(size_t)i * C + c]; dp[i] = mb.dvp[(size_t)i * C + c]; ds[i] = mb.dvs[(size_t)i * C + c];
  }
  for (int i = 1; i < k; ++i) {
    const double a = z[i], b = dp[i], d = ds[i];
    int m = i - 1;
    while (m >= 0 && z[m] > a) { z[m + 1] = z[m]; dp[m + 1] = dp[m]; ds[m + 1] = ds[m]; --m; }
    z[m + 1] = a; dp[m + 1] = b; ds[m + 1] = d;
  }
  int i = 0;
  bool valid = true;
  double* A = alpha + (size_t)c * stride; double* Bt = beta + (size_t)c * stride;
  double* R = rho + (size_t)c * stride; double* H = h + (size_t)c * stride;
  if (cfg.sdep > 0.0) { A[i] = 1.5; Bt[i] = -999.0; R[i] = 1.0; H[i] = cfg.sdep; ++i; }
  for (int l = 0; l <= k; ++l) {
    double zc, hh, dvs_l, dvp_l;
    if (l == 0) { zc = __dmul_rn(0.5, __dadd_rn(cfg.sdep, z[0])); hh = __dsub_rn(z[0], cfg.sdep); dvs_l = ds[0]; dvp_l = dp[0]; }
    else if (l < k) { zc = __dmul_rn(0.5, __dadd_rn(z[l], z[l - 1])); hh = __dsub_rn(z[l], z[l - 1]); dvs_l = ds[l]; dvp_l = dp[l]; }
    else { zc = __dmul_rn(0.5, __dadd_rn(cfg.z_max, z[k - 1])); hh = 999.0;
           dvs_l = mb.dvs[(size_t)(km - 1) * C + c]; dvp_l = mb.dvp[(size_t)(km - 1) * C + c]; }
    double a, b;
    bool ok = layer_velocity(cfg, zc, dvs_l, dvp_l, a, b);
    if (l == 0) ok = ok && !(hh < __dmul_rn(0.125, a));
    else if (l < k) ok = ok && !(hh < cfg.h_min);
    valid = valid && ok;
    if (i < stride) { A[i] = a; Bt[i] = b; R[i] = vp_to_rho(a); H[i] = hh; }
    ++i;
  }
  nlay_out[c] = i;
  is_valid[c] = (uint8_t)valid;
}

size_t forward_smem_bytes(const DevConfig& cfg, int nthr) {
  const size_t n = cfg.nfft, nh = cfg.nh, km = cfg.k_max;
  const size_t tab_entries = km * 2 * (16 + (nthr >> 4));
  const size_t region0 = tab_entries > n ? tab_entries : n;
  const size_t spectra = cfg.ray_common ? 2 * (nh + 1) : 0;   // aliased onto buf1 otherwise
  return sizeof(double2) * (n + 4 + region0 + spectra + n / 4) + sizeof(LayerConst) * km + sizeof(RayConst) + sizeof(double) * 32;
}

template <int J, int BMAX, int MINB>
int launch_forward_t(const DevConfig& cfg, const ModelBatch& mb, const EvalOutputs& out, const double* lc,
                     const double* rc, int nthr, cudaStream_t stream) {
  const size_t smem = forward_smem_bytes(cfg, nthr);
  RFINV_CUDA_CHECK(cudaFuncSetAttribute(forward_kernel<J, BMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_models = mb.active ? mb.n_active : mb.C;
  const int ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  const long long grid = (long long)n_models * ntr_eff;
  forward_kernel<J, BMAX, MINB><<<(unsigned)grid, nthr, smem, stream>>>(cfg, mb, out, lc, rc);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}

}  // namespace

int rfinv_forward_bins_per_thread(int nfft) {
  if (nfft <= 64) return 1;
  if (nfft <= 256) return 2;
  if (nfft <= 1024) return 4;
  return 8;
}

size_t rfinv_forward_scratch_doubles(const DevConfig& cfg, long long n_models) {
  const long long ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  return (size_t)(n_models * ntr_eff) * ((size_t)cfg.k_max * LC_DOUBLES + RC_DOUBLES);
}

// scratch: rfinv_forward_scratch_doubles(cfg, n_models) doubles of device memory
int rfinv_launch_forward(const DevConfig& cfg, const ModelBatch& mb, const EvalOutputs& out, double* scratch,
                         cudaStream_t stream) {
  const int J = rfinv_forward_bins_per_thread(cfg.nfft);
  const int nthr = (cfg.nfft / 2) / J;
  const int n_models = mb.active ? mb.n_active : mb.C;
  const int ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  const long long n_items = (long long)n_models * ntr_eff;
  if (n_items == 0) return RFINV_OK;
  double* lc = scratch;
  double* rc = scratch + (size_t)n_items * cfg.k_max * LC_DOUBLES;
  prep_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, stream>>>(cfg, mb, lc, rc, out.is_valid, (int)n_items, ntr_eff, nthr);
  RFINV_CUDA_CHECK(cudaGetLastError());
  switch (J) {
    case 1: return launch_forward_t<1, 32, 8>(cfg, mb, out, lc, rc, nthr, stream);
    case 2: return launch_forward_t<2, 64, 6>(cfg, mb, out, lc, rc, nthr, stream);
    case 4: {
      static const int minb = getenv("RFINV_FWD_MINB") ? atoi(getenv("RFINV_FWD_MINB")) : 4;   // tuning knob (CTAs per SM)
      if (minb <= 3) return launch_forward_t<4, 128, 3>(cfg, mb, out, lc, rc, nthr, stream);
      return launch_forward_t<4, 128, 4>(cfg, mb, out, lc, rc, nthr, stream);
    }
    default: return launch_forward_t<8, 256, 1>(cfg, mb, out, lc, rc, nthr, stream);
  }
}

int rfinv_launch_format_model(const DevConfig& cfg, const ModelBatch& mb, int* nlay, double* alpha, double* beta,
                              double* rho, double* h, uint8_t* is_valid, cudaStream_t stream) {
  if (mb.C == 0) return RFINV_OK;
  format_model_kernel<<<(mb.C + 127) / 128, 128, 0, stream>>>(cfg, mb, nlay, alpha, beta, rho, h, is_valid);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}
