// FP64 roofline microbenchmark for B200: DFMA (vector pipe), DMMA (mma.sync f64) and both together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Writes one JSON object to stdout; bench.py / DESIGN.md quote `dfma_tflops` as the FP64 peak.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x; c[u][1] = u; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) dmma884(c[u][0], c[u][1], a, b);
  }
  double s = 0;
  for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters, double av, double bv) {
  double c[4][4], a[8], b[4];
  for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) c[u][j] = threadIdx.x + u + j;
  for (int j = 0; j < 8; ++j) a[j] = av + j;
  for (int j = 0; j < 4; ++j) b[j] = bv + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) dmma16816(c[u], a, b);
  }
  double s = 0;
  for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) s += c[u][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// even warps: DFMA, odd warps: DMMA (m16n8k16).  fma_iters / mma_iters scale each side.
__global__ void mixed_kernel(double* out, int fma_iters, int mma_iters, double av, double bv) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if ((warp & 1) == 0) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < fma_iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        x0 = fma(x0, av, bv); x1 = fma(x1, av, bv); x2 = fma(x2, av, bv); x3 = fma(x3, av, bv);
        x4 = fma(x4, av, bv); x5 = fma(x5, av, bv); x6 = fma(x6, av, bv); x7 = fma(x7, av, bv);
      }
    }
    s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  } else {
    double c[4][4], a[8], b[4];
    for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) c[u][j] = threadIdx.x + u + j;
    for (int j = 0; j < 8; ++j) a[j] = av + j;
    for (int j = 0; j < 4; ++j) b[j] = bv + j;
    for (int i = 0; i < mma_iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma16816(c[u], a, b);
    }
    for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) s += c[u][j];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount, threads = 256, blocks = sms * 8;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  const int iters = 4096;
  float t_fma = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
  double fma_flop = 2.0 * 64.0 * iters * (double)blocks * threads;
  float t_884 = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
  double m884_flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
  float t_168 = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 10);
  double m168_flop = 2.0 * 16 * 8 * 16 * 4.0 * (iters / 4) * (double)blocks * (threads / 32);
  // mixed: same per-warp work as the isolated runs, half the warps each
  float t_mix = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, iters, iters / 4, 1.0000001, 1e-9); }, 10);
  double mix_flop = 0.5 * fma_flop + 0.5 * m168_flop;
  CK(cudaDeviceSynchronize());
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
         "\"dmma_m16n8k16_tflops\": %.3f, \"mixed_half_half_tflops\": %.3f, \"t_fma_ms\": %.3f, \"t_m16n8k16_ms\": %.3f, "
         "\"t_mixed_ms\": %.3f}\n",
         prop.name, sms, prop.clockRate, fma_flop / t_fma * 1e-9, m884_flop / t_884 * 1e-9, m168_flop / t_168 * 1e-9,
         mix_flop / t_mix * 1e-9, t_fma, t_168, t_mix);
  return 0;
}
