// DFMA/DMUL issue rate on B200 as a function of how many DISTINCT 64-bit register operands an instruction reads
// (operand-reuse cache vs register-file ports).  ILP 8, 4 warps per SMSP, one CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ILP = 8;
// MODE 0: x = fma(x, a, b)        one distinct operand per instruction (a, b shared by all: reuse cache)
// MODE 1: x = fma(x, y_i, b)      two distinct
// MODE 2: x = fma(x, y_i, z_i)    three distinct
// MODE 3: x = x * y_i             DMUL, two distinct
// MODE 4: plane rotation pattern of the layer loop: (p,q) <- (c p - s q, s p + c q), c,s shared by 4 pairs
template <int MODE>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < ILP; i += 2) {
          const double p = x[i], q = x[i + 1];
          x[i] = fma(-b, q, a * p);
          x[i + 1] = fma(b, p, a * q);
        }
      } else {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
          if (MODE == 0) x[i] = fma(x[i], a, b);
          if (MODE == 1) x[i] = fma(x[i], y[i], b);
          if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);
          if (MODE == 3) x[i] = x[i] * y[i];
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_smsp, double* out, long long* dcyc) {
  const int iters = 2000, threads = 32 * 4 * warps_per_smsp;
  k<MODE><<<148, threads>>>(out, iters, 1.0000001, 1e-9, dcyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
  const double n_inst = iters * 4.0 * ILP * (MODE == 4 ? 2.0 : 1.0);   // rotation: 2 instructions per element
  const double per = (double)c / n_inst;
  printf("%-44s warps/SMSP %d : SMSP issue interval %5.2f cycles per FP64 instruction, pipe util %5.1f%%\n", name, warps_per_smsp,
         per / warps_per_smsp, 100.0 * 2.0 * warps_per_smsp / per);
}
int main() {
  double* out; long long* dcyc;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dcyc, 8);
  for (int w : {1, 4}) {
    run<0>("DFMA x=fma(x,a,b)      1 distinct operand", w, out, dcyc);
    run<1>("DFMA x=fma(x,y_i,b)    2 distinct operands", w, out, dcyc);
    run<2>("DFMA x=fma(x,y_i,z_i)  3 distinct operands", w, out, dcyc);
    run<3>("DMUL x=x*y_i           2 distinct operands", w, out, dcyc);
    run<4>("rotation (DMUL+DFMA, c,s shared)", w, out, dcyc);
  }
  return 0;
}
