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
#include "rfinv_common.cuh"

namespace {

struct LayerConst {
  double thx, the;            // domg*xi*h, domg*eta*h : phase advance per frequency bin
  double cbx, sbx, cbe, sbe;  // rotation by B*thx, B*the (B = blockDim.x = frequency stride of a thread)
  double g, bp;               // 2 beta^2 p^2, 1 - 2 beta^2 p^2
  double c12a, c12b, c21a, c21b, c13a, c13b, c24a, c24b, c31a, c31b, c42a, c42b;
  double c14, c41;
  double dcx, dsx, dce, dse;  // cos/sin at the DC pseudo-frequency 1.0e-5 (src/forward.f90:246-248)
};

struct HalfSpace {
  double e11, e12, e13, e14, e21, e22, e23, e24;  // E^-1 rows 3,4 with the 1/w factors removed
};

struct Water {
  double thw, cbw, sbw, rw, dcw, dsw;  // phase per bin, stride rotation, rho_w/xi_w, DC cos/sin
  int present;
};

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
  const double nrm = dl.x * dl.x + dl.y * dl.y;
  const double2 inv = make_double2(dl.x / nrm, -dl.y / nrm);
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

// In-place-pair Stockham inverse FFT (sign +, unnormalised) of n complex points in shared memory.
// Returns the buffer holding the result.
__device__ double2* fft_inverse(double2* x, double2* y, int n, int log2n, const double2* __restrict__ tw, int tid,
                                int nthr) {
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
  const int quarter = n >> 2;
  while (Ns < n) {
    const int tstep = n / (4 * Ns);
    for (int j = tid; j < quarter; j += nthr) {
      const int k = j & (Ns - 1);
      double2 v0 = x[j], v1 = x[j + quarter], v2 = x[j + 2 * quarter], v3 = x[j + 3 * quarter];
      if (Ns > 1) {
        v1 = cmul(v1, __ldg(&tw[k * tstep]));
        v2 = cmul(v2, __ldg(&tw[2 * k * tstep]));
        v3 = cmul(v3, __ldg(&tw[3 * k * tstep]));
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

template <int J, int BMAX>
__global__ void __launch_bounds__(BMAX) forward_kernel(const DevConfig cfg, const ModelBatch mb, const EvalOutputs out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n = cfg.nfft, nh = cfg.nh, km = cfg.k_max, C = mb.C;
  const int ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  const int item = blockIdx.x;
  const int ci = item / ntr_eff, t0 = item - ci * ntr_eff;
  const int c = mb.active ? mb.active[ci] : ci;

  // ---- shared memory carve-up ----
  double2* s_buf0 = reinterpret_cast<double2*>(smem_raw);
  double2* s_buf1 = s_buf0 + n;
  double2* s_fr = s_buf1 + n;          // [nh] unfiltered radial spectrum (or deconvolved RF spectrum)
  double2* s_fv = s_fr + (nh + 1);     // [nh] unfiltered vertical spectrum
  LayerConst* s_lc = reinterpret_cast<LayerConst*>(s_fv + (nh + 1));
  double* s_z = reinterpret_cast<double*>(s_lc + km);
  double* s_dvp = s_z + km;
  double* s_dvs = s_dvp + km;
  double* s_alpha = s_dvs + km;
  double* s_beta = s_alpha + (km + 1);
  double* s_rho = s_beta + (km + 1);
  double* s_h = s_rho + (km + 1);
  double* s_xi = s_h + (km + 1);
  double* s_eta = s_xi + (km + 1);
  double* s_red = s_eta + (km + 1);    // [32]
  __shared__ HalfSpace s_hs;
  __shared__ Water s_w;
  __shared__ int s_valid;
  __shared__ double s_tp;

  const int k = mb.k[c];
  const double p = cfg.rayp[t0];
  const int ipha = cfg.ipha[t0];

  // ---- format_model (src/model.f90:175-290): rank sort of the k interfaces ----
  if (tid == 0) s_valid = 1;
  double* tz = reinterpret_cast<double*>(s_buf0);  // scratch: unsorted copies
  double* tp_ = tz + km;
  double* ts_ = tp_ + km;
  for (int i = tid; i < k; i += nthr) {
    tz[i] = mb.z[(size_t)i * C + c];
    tp_[i] = mb.dvp[(size_t)i * C + c];
    ts_[i] = mb.dvs[(size_t)i * C + c];
  }
  __syncthreads();
  for (int i = tid; i < k; i += nthr) {
    const double zi = tz[i];
    int r = 0;
    for (int j = 0; j < k; ++j) r += (tz[j] < zi) || (tz[j] == zi && j < i);
    s_z[r] = zi;
    s_dvp[r] = tp_[i];
    s_dvs[r] = ts_[i];
  }
  __syncthreads();
  for (int l = tid; l <= k; l += nthr) {
    double zc, h, dvs_l, dvp_l;
    if (l == 0) {
      zc = __dmul_rn(0.5, __dadd_rn(cfg.sdep, s_z[0]));
      h = __dsub_rn(s_z[0], cfg.sdep);
      dvs_l = s_dvs[0]; dvp_l = s_dvp[0];
    } else if (l < k) {
      zc = __dmul_rn(0.5, __dadd_rn(s_z[l], s_z[l - 1]));
      h = __dsub_rn(s_z[l], s_z[l - 1]);
      dvs_l = s_dvs[l]; dvp_l = s_dvp[l];
    } else {
      zc = __dmul_rn(0.5, __dadd_rn(cfg.z_max, s_z[k - 1]));
      h = 999.0;
      dvs_l = mb.dvs[(size_t)(km - 1) * C + c];
      dvp_l = mb.dvp[(size_t)(km - 1) * C + c];
    }
    double a, b;
    bool ok = layer_velocity(cfg, zc, dvs_l, dvp_l, a, b);
    if (l == 0) ok = ok && !(h < __dmul_rn(0.125, a));     // src/model.f90:229
    else if (l < k) ok = ok && !(h < cfg.h_min);           // src/model.f90:257
    s_alpha[l] = a; s_beta[l] = b; s_rho[l] = vp_to_rho(a); s_h[l] = h;
    if (!ok) s_valid = 0;
    // ---- per-layer propagator constants for this ray ----
    const double beta2 = b * b, p2 = p * p;
    const double bp = 1.0 - 2.0 * beta2 * p2;
    const double eta = sqrt(1.0 / beta2 - p2);
    const double xi = sqrt(1.0 / (a * a) - p2);
    const double rho = s_rho[l];
    s_xi[l] = xi; s_eta[l] = eta;
    if (l < k) {
      LayerConst L;
      L.thx = cfg.domg * xi * h;
      L.the = cfg.domg * eta * h;
      sincos((double)nthr * L.thx, &L.sbx, &L.cbx);
      sincos((double)nthr * L.the, &L.sbe, &L.cbe);
      sincos((double)1.0e-5f * xi * h, &L.dsx, &L.dcx);
      sincos((double)1.0e-5f * eta * h, &L.dse, &L.dce);
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
      s_lc[l] = L;
    } else {  // half space: rows 3,4 of E^-1 (src/forward.f90:350-380) without their 1/omega factors
      HalfSpace H;
      H.e11 = beta2 * p / a;
      H.e12 = bp / (2.0 * a * xi);
      H.e13 = p / (2.0 * rho * a * xi);
      H.e14 = 1.0 / (2.0 * rho * a);
      H.e21 = bp / (2.0 * b * eta);
      H.e22 = b * p;
      H.e23 = 1.0 / (2.0 * rho * b);
      H.e24 = p / (2.0 * rho * b * eta);
      s_hs = H;
    }
  }
  if (tid == 0) {  // water layer (src/model.f90:201-207, src/forward.f90:424-442)
    Water W;
    W.present = cfg.sdep > 0.0;
    if (W.present) {
      const double aw = 1.5, rhow = 1.0, hw = cfg.sdep;
      const double xiw = sqrt(1.0 / (aw * aw) - p * p);
      W.thw = cfg.domg * xiw * hw;
      sincos((double)nthr * W.thw, &W.sbw, &W.cbw);
      sincos((double)1.0e-5f * xiw * hw, &W.dsw, &W.dcw);
      W.rw = rhow / xiw;
    } else {
      W.thw = 0.0; W.cbw = 1.0; W.sbw = 0.0; W.rw = 0.0; W.dcw = 1.0; W.dsw = 0.0;
    }
    s_w = W;
  }
  __syncthreads();
  if (tid == 0) {  // direct_arrival (src/forward.f90:474-491), summed in layer order
    double t = 0.0;
    if (cfg.deconv_mode == 0) {
      const double* sl = ipha == 1 ? s_xi : s_eta;
      for (int l = 0; l < k; ++l) t = t + s_h[l] * sl[l];
    }
    s_tp = t;
    if (out.is_valid && t0 == 0) out.is_valid[c] = (uint8_t)s_valid;
  }

  // ---- propagator product: thread handles bins j = tid + m*nthr (m < J); thread 0 also the Nyquist ----
  double ya[J][4], yb[J][4], cwv[J];
  double yae[4] = {1.0, 0.0, 0.0, 0.0}, ybe[4], cwe;
  {
    const Water W = s_w;
    double cw, sw;
    sincos((double)tid * W.thw, &sw, &cw);
#pragma unroll
    for (int m = 0; m < J; ++m) {
      double ucw = cw, usw = sw;
      if (m == 0 && tid == 0) { ucw = W.dcw; usw = W.dsw; }
      ya[m][0] = 1.0; ya[m][1] = 0.0; ya[m][2] = 0.0; ya[m][3] = 0.0;
      yb[m][0] = 0.0; yb[m][1] = ucw; yb[m][2] = 0.0; yb[m][3] = -W.rw * usw;
      cwv[m] = ucw;
      rot(cw, sw, W.cbw, W.sbw);
    }
    ybe[0] = 0.0; ybe[1] = cw; ybe[2] = 0.0; ybe[3] = -W.rw * sw;
    cwe = cw;
  }
  for (int l = 0; l < k; ++l) {
    const LayerConst L = s_lc[l];
    double c1, s1, c2, s2;
    sincos((double)tid * L.thx, &s1, &c1);
    sincos((double)tid * L.the, &s2, &c2);
#pragma unroll
    for (int m = 0; m < J; ++m) {
      double uc1 = c1, us1 = s1, uc2 = c2, us2 = s2;
      if (m == 0 && tid == 0) { uc1 = L.dcx; us1 = L.dsx; uc2 = L.dce; us2 = L.dse; }
      layer_step(L, uc1, us1, uc2, us2, ya[m], yb[m]);
      rot(c1, s1, L.cbx, L.sbx);
      rot(c2, s2, L.cbe, L.sbe);
    }
    if (tid == 0) layer_step(L, c1, s1, c2, s2, yae, ybe);
  }
  {
    const HalfSpace H = s_hs;
#pragma unroll
    for (int m = 0; m < J; ++m) {
      double2 fr, fv;
      surface_response(H, ya[m], yb[m], cwv[m], ipha, fr, fv);
      s_fr[tid + m * nthr] = fr;
      s_fv[tid + m * nthr] = fv;
    }
    if (tid == 0) {
      double2 fr, fv;
      surface_response(H, yae, ybe, cwe, ipha, fr, fv);
      s_fr[nh - 1] = fr;
      s_fv[nh - 1] = fv;
    }
  }
  __syncthreads();

  // ---- water-level deconvolution (src/forward.f90:148-153, 447-470): overwrites s_fr with rff ----
  if (cfg.deconv_mode == 1) {
    const double2* xs = ipha == 1 ? s_fv : s_fr;  // denominator spectrum
    const double2* ys = ipha == 1 ? s_fr : s_fv;
    double mx = -INFINITY;
    for (int j = tid; j < nh; j += nthr) mx = fmax(mx, xs[j].x * xs[j].x + xs[j].y * xs[j].y);
    mx = block_max(mx, s_red, tid, nthr);
    const double wlvl = 0.001 * mx;
    double2 keep[J + 1];
    int cnt = 0;
    for (int j = tid; j < nh; j += nthr) {
      const double2 x = xs[j], y = ys[j];
      const double amp = x.x * x.x + x.y * x.y;
      const double d = fmax(amp, wlvl);
      keep[cnt++] = make_double2((y.x * x.x + y.y * x.y) / d, (y.y * x.x - y.x * x.y) / d);
    }
    __syncthreads();
    cnt = 0;
    for (int j = tid; j < nh; j += nthr) s_fr[j] = keep[cnt++];
    __syncthreads();
  }

  // ---- per trace: filter -> inverse FFT -> shift / normalise -> outputs ----
  const double tp = s_tp;
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
    const double2* res = fft_inverse(s_buf0, s_buf1, n, cfg.log2n, cfg.tw, tid, nthr);
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
    double* smp = out.rft_smp ? out.rft_smp + ((size_t)t * C + c) * S : nullptr;
    double* full = out.rft_full ? out.rft_full + ((size_t)c * cfg.ntrc + t) * n : nullptr;
    const double* __restrict__ obs = cfg.obs + (size_t)t * S;
    for (int i = tid; i < nout; i += nthr) {
      int src;
      double v;
      if (ipha == 1) {
        src = (i - npre) % n; if (src < 0) src += n;
        v = res[src].x;
      } else {
        src = (npre - i - 1) % n; if (src < 0) src += n;
        v = -res[src].x;
      }
      if (cfg.deconv_mode == 0) v = v / fac;
      if (i < S) {
        mis[i] = v - obs[i];
        if (smp) smp[i] = v;
      }
      if (full) full[i] = v;
    }
    __syncthreads();
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

size_t forward_smem_bytes(const DevConfig& cfg) {
  const size_t n = cfg.nfft, nh = cfg.nh, km = cfg.k_max;
  return sizeof(double2) * (2 * n + 2 * (nh + 1)) + sizeof(LayerConst) * km + sizeof(double) * (3 * km + 6 * (km + 1) + 32);
}

template <int J, int BMAX>
int launch_forward_t(const DevConfig& cfg, const ModelBatch& mb, const EvalOutputs& out, int nthr, cudaStream_t stream) {
  const size_t smem = forward_smem_bytes(cfg);
  static size_t configured = 0;
  if (smem > configured) {
    RFINV_CUDA_CHECK(cudaFuncSetAttribute(forward_kernel<J, BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int n_models = mb.active ? mb.n_active : mb.C;
  const int ntr_eff = cfg.ray_common ? 1 : cfg.ntrc;
  const long long grid = (long long)n_models * ntr_eff;
  if (grid == 0) return RFINV_OK;
  forward_kernel<J, BMAX><<<(unsigned)grid, nthr, smem, stream>>>(cfg, mb, out);
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

int rfinv_launch_forward(const DevConfig& cfg, const ModelBatch& mb, const EvalOutputs& out, cudaStream_t stream) {
  const int J = rfinv_forward_bins_per_thread(cfg.nfft);
  const int nthr = (cfg.nfft / 2) / J;
  switch (J) {
    case 1: return launch_forward_t<1, 32>(cfg, mb, out, nthr, stream);
    case 2: return launch_forward_t<2, 64>(cfg, mb, out, nthr, stream);
    case 4: return launch_forward_t<4, 128>(cfg, mb, out, nthr, stream);
    default: return launch_forward_t<8, 256>(cfg, mb, out, nthr, stream);
  }
}

int rfinv_launch_format_model(const DevConfig& cfg, const ModelBatch& mb, int* nlay, double* alpha, double* beta,
                              double* rho, double* h, uint8_t* is_valid, cudaStream_t stream) {
  if (mb.C == 0) return RFINV_OK;
  format_model_kernel<<<(mb.C + 127) / 128, 128, 0, stream>>>(cfg, mb, nlay, alpha, beta, rho, h, is_valid);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}
