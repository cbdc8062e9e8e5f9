// Layer-loop microbenchmark: FP64 pipe utilisation of the wave-coordinate propagation loop of forward_kernel in
// isolation (synthetic tables in shared memory), as a function of CTAs per SM and of the loop structure.
#include <cstdio>
#include <cuda_runtime.h>
struct Wave { double a1, a2, b1, b2; };
struct LayerConst { double thx, the, cbx, sbx, cbe, sbe, t12, t21, u11, u12, u21, u22; };
__device__ __forceinline__ void rot(double& c, double& s, double cb, double sb) {
  double c2 = c * cb - s * sb; double s2 = s * cb + c * sb; c = c2; s = s2; }
__device__ __forceinline__ void wave_rotate(Wave& w, double c1, double s1, double c2, double s2) {
  const double a1 = fma(-s1, w.a2, c1 * w.a1), a2 = fma(s1, w.a1, c1 * w.a2);
  const double b1 = fma(-s2, w.b2, c2 * w.b1), b2 = fma(s2, w.b1, c2 * w.b2);
  w.a1 = a1; w.a2 = a2; w.b1 = b1; w.b2 = b2; }
__device__ __forceinline__ void wave_interface(Wave& w, double t12, double t21, double u11, double u12, double u21, double u22) {
  const double a1 = fma(t12, w.b1, w.a1), b1 = fma(t21, w.a1, w.b1);
  const double a2 = fma(u12, w.b2, u11 * w.a2), b2 = fma(u22, w.b2, u21 * w.a2);
  w.a1 = a1; w.a2 = a2; w.b1 = b1; w.b2 = b2; }

constexpr int J = 4, KM = 30, NHI = 8, TPL = 2 * (16 + NHI);
// MODE 0: as in forward_kernel; MODE 1: next layer's base trig prefetched one layer ahead; MODE 2: as 0 plus a CTA barrier per item
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(double* out, int items, int kl, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_tab = reinterpret_cast<double2*>(smem_raw);
  LayerConst* s_lc = reinterpret_cast<LayerConst*>(s_tab + KM * TPL);
  const int tid = threadIdx.x;
  for (int i = tid; i < KM * TPL; i += blockDim.x) { double sn, cs; sincos(1e-3 * i, &sn, &cs); s_tab[i] = make_double2(cs, sn); }
  for (int i = tid; i < KM; i += blockDim.x) {
    LayerConst L; L.thx = 0.01; L.the = 0.02; sincos(0.3 + i, &L.sbx, &L.cbx); sincos(0.7 + i, &L.sbe, &L.cbe);
    L.t12 = 0.01; L.t21 = -0.02; L.u11 = 1.0; L.u12 = 0.03; L.u21 = -0.01; L.u22 = 1.0; s_lc[i] = L; }
  __syncthreads();
  const int t_lo = tid & 15, t_hi = 16 + (tid >> 4);
  long long t0 = clock64();
  double acc = 0.0;
  for (int it = 0; it < items; ++it) {
    Wave wa[J], wb[J];
#pragma unroll
    for (int m = 0; m < J; ++m) { wa[m].a1 = 1.0 + it; wa[m].a2 = 0.0; wa[m].b1 = 0.5; wa[m].b2 = 0.0; wb[m].a1 = 0.0; wb[m].a2 = 1.0; wb[m].b1 = 0.0; wb[m].b2 = 0.3 * m; }
    double c1, s1, c2, s2;
    if (MODE == 1) {
      const double2 a = s_tab[t_lo], b = s_tab[t_hi], cc = s_tab[16 + NHI + t_lo], d = s_tab[16 + NHI + t_hi];
      c1 = a.x; s1 = a.y; rot(c1, s1, b.x, b.y); c2 = cc.x; s2 = cc.y; rot(c2, s2, d.x, d.y);
    }
    for (int l = 0; l < kl; ++l) {
      const LayerConst& L = s_lc[l];
      double n1 = 0, ns1 = 0, n2 = 0, ns2 = 0;
      if (MODE == 1) {
        const double2* tab = s_tab + (l + 1) * TPL;   // table of the next layer (one spare layer in the buffer)
        const double2 a = tab[t_lo], b = tab[t_hi], cc = tab[16 + NHI + t_lo], d = tab[16 + NHI + t_hi];
        n1 = a.x; ns1 = a.y; rot(n1, ns1, b.x, b.y); n2 = cc.x; ns2 = cc.y; rot(n2, ns2, d.x, d.y);
      } else {
        const double2* tab = s_tab + l * TPL;
        const double2 a = tab[t_lo], b = tab[t_hi], cc = tab[16 + NHI + t_lo], d = tab[16 + NHI + t_hi];
        c1 = a.x; s1 = a.y; rot(c1, s1, b.x, b.y); c2 = cc.x; s2 = cc.y; rot(c2, s2, d.x, d.y);
      }
      const double t12 = L.t12, t21 = L.t21, u11 = L.u11, u12 = L.u12, u21 = L.u21, u22 = L.u22;
#pragma unroll
      for (int m = 0; m < J; ++m) {
        wave_rotate(wa[m], c1, s1, c2, s2);
        wave_rotate(wb[m], c1, s1, c2, s2);
        wave_interface(wa[m], t12, t21, u11, u12, u21, u22);
        wave_interface(wb[m], t12, t21, u11, u12, u21, u22);
        if (m + 1 < J) { rot(c1, s1, L.cbx, L.sbx); rot(c2, s2, L.cbe, L.sbe); }
      }
      if (MODE == 1) { c1 = n1; s1 = ns1; c2 = n2; s2 = ns2; }
    }
#pragma unroll
    for (int m = 0; m < J; ++m) acc += wa[m].a1 + wa[m].a2 + wa[m].b1 + wa[m].b2 + wb[m].a1 + wb[m].a2 + wb[m].b1 + wb[m].b2;
    if (MODE == 2) __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int ctas_per_sm, double* out, long long* dcyc) {
  const int items = 40, kl = 12;
  const size_t smem = sizeof(double2) * (KM + 1) * TPL + sizeof(LayerConst) * KM;
  const size_t pad = ctas_per_sm == 4 ? smem : (ctas_per_sm == 2 ? 100000 : (ctas_per_sm == 3 ? 70000 : 200000));
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 128, pad);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * ctas_per_sm, 128, pad>>>(out, items, kl, dcyc);
  cudaEventRecord(e0);
  k<MODE><<<148 * ctas_per_sm, 128, pad>>>(out, items, kl, dcyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
  const double fp64_per_warp = (double)items * kl * (8 + J * 28 + (J - 1) * 8);   // warp-instructions per warp
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("%-34s CTAs/SM %d (occupancy %d): block 0 %8lld cycles, grid %8.0f cycles, FP64 pipe utilisation %5.1f%% (%s)\n", name, ctas_per_sm, occ, c,
         cycles, 100.0 * 2.0 * fp64_per_warp * ctas_per_sm / cycles, cudaGetErrorString(e));
}
int main() {
  double* out; long long* dcyc;
  cudaMalloc(&out, 8 * 148 * 4 * 128); cudaMalloc(&dcyc, 8);
  for (int c : {1, 2, 3, 4}) {
    run<0>("layer loop as in forward_kernel", c, out, dcyc);
    run<1>("next layer's trig prefetched", c, out, dcyc);
    run<2>("as first + CTA barrier per item", c, out, dcyc);
  }
  return 0;
}
