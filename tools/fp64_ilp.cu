// DFMA dependent-chain latency and throughput vs ILP x warps per SMSP on B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp, double* out, long long* dcyc) {
  const int iters = 2000, threads = 32 * 4 * warps_per_smsp;  // one CTA per SM
  k<ILP><<<148, threads>>>(out, iters, 1.0000001, 1e-9, dcyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (iters * 4.0 * ILP);   // cycles per DFMA warp-instruction per warp
  printf("ILP %2d warps/SMSP %2d : %6.2f cycles per DFMA per warp, SMSP issue interval %5.2f cycles, pipe util %5.1f%%\n", ILP,
         warps_per_smsp, per, per / warps_per_smsp, 100.0 * 2.0 * warps_per_smsp / per);
}
int main() {
  double* out; long long* dcyc;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dcyc, 8);
  for (int w : {1, 2, 4, 8}) { run<1>(w, out, dcyc); run<2>(w, out, dcyc); run<4>(w, out, dcyc); run<8>(w, out, dcyc); run<16>(w, out, dcyc); }
  return 0;
}
