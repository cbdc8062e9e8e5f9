// FP64 roofline denominators, measured live: MEASURED_PEAKS.json (driver-written) carries HBM and bf16 peaks
// only, and this path is bound by the FP64 pipe.  Same loops as tools/fp64_peak.cu.
#include "rfinv_common.cuh"

namespace {
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
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
__global__ void dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x; c[u][1] = u; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int32_t rfinv_measure_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops) {
  if (!dfma_tflops || !dmma_tflops) { rfinv_set_error("NULL argument"); return RFINV_ERR_ARG; }
  RFINV_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  RFINV_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 2048;
  double* out = nullptr;
  RFINV_CUDA_CHECK(cudaMalloc((void**)&out, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  RFINV_CUDA_CHECK(cudaEventCreate(&e0));
  RFINV_CUDA_CHECK(cudaEventCreate(&e1));
  double best[2] = {0.0, 0.0};
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      RFINV_CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = which == 0 ? 2.0 * 64.0 * iters * (double)blocks * threads
                                     : 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32);
      const double tf = flop / (ms * 1e-3) * 1e-12;
      if (rep > 0 && tf > best[which]) best[which] = tf;
    }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *dfma_tflops = best[0];
  *dmma_tflops = best[1];
  return RFINV_OK;
}
