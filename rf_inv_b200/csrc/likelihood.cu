// Correlated-noise likelihood for a batch of models (src/likelihood.f90:85-98):
//   phi_t(c) = m^T R_t^-1 m,   logL(c) = sum_t  -0.5 phi_t / sig_t^2 - nsmp log(sig_t)
// The reference evaluates phi with a dense matmul + dot product per chain, streaming R^-1 (S x S) for
// every evaluation.  Here the misfits of all chains form a matrix M (chains x S) and phi is the row-wise
// diagonal of M R^-1 M^T: one fp64 GEMM against the shared, symmetric R^-1 with the row-dot fused into
// the epilogue.  R^-1 is symmetric, so only tiles on or below the diagonal are visited (off-diagonal
// tiles count twice); the summation order is fixed (no atomics), results are run-to-run deterministic.
#include "rfinv_common.cuh"

namespace {

constexpr int TM = 64;   // chains per CTA
constexpr int TN = 64;   // columns of R^-1 per tile
constexpr int TK = 16;   // k-slab
constexpr int QF_THREADS = 256;

// CTA = 64 chains of one trace; loops over column tiles jt and k tiles kt <= jt.
__global__ void __launch_bounds__(QF_THREADS) quadform_kernel(const DevConfig cfg, int C, const double* __restrict__ misfit,
                                                              double* __restrict__ phi, const int* __restrict__ active,
                                                              int n_active) {
  __shared__ double sA[TK][TM + 4];   // M tile, k-major
  __shared__ double sB[TK][TN + 4];   // R^-1 tile
  __shared__ int s_rows[TM];
  const int t = blockIdx.y;
  const int Sp = cfg.nsmp_pad;
  const int tid = threadIdx.x;
  const int n_rows = active ? n_active : C;
  const int row0 = blockIdx.x * TM;
  if (tid < TM) {
    const int r = row0 + tid;
    s_rows[tid] = r < n_rows ? (active ? active[r] : r) : -1;
  }
  __syncthreads();
  const double* __restrict__ Mt = misfit + (size_t)t * C * Sp;
  const double* __restrict__ Rt = cfg.r_inv + (size_t)t * Sp * Sp;
  const int tx = tid & 15, ty = tid >> 4;   // thread owns rows ty*4..+3, cols tx*4..+3 of the 64x64 tile
  double phi_acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int ntile = Sp / TN;
  // load indices: A tile 64 rows x 16 k: thread loads 4 consecutive k of one row; B tile 16 k x 64 cols
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int b_k = tid >> 4, b_col = (tid & 15) * 4;
  const int a_src = s_rows[a_row];
  for (int jt = 0; jt < ntile; ++jt) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int k_end = (jt + 1) * TN;
    for (int k0 = 0; k0 < k_end; k0 += TK) {
      if (k0 == jt * TN && jt > 0) {  // entering the diagonal tile: everything so far counts twice
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] *= 2.0;
      }
      double av[4] = {0.0, 0.0, 0.0, 0.0};
      if (a_src >= 0) {
        const double2* src = reinterpret_cast<const double2*>(Mt + (size_t)a_src * Sp + k0 + a_k);
        const double2 v0 = src[0], v1 = src[1];
        av[0] = v0.x; av[1] = v0.y; av[2] = v1.x; av[3] = v1.y;
      }
      const double2* bsrc = reinterpret_cast<const double2*>(Rt + (size_t)(k0 + b_k) * Sp + jt * TN + b_col);
      const double2 b0 = bsrc[0], b1 = bsrc[1];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) sA[a_k + q][a_row] = av[q];
      sB[b_k][b_col] = b0.x; sB[b_k][b_col + 1] = b0.y; sB[b_k][b_col + 2] = b1.x; sB[b_k][b_col + 3] = b1.y;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
    // epilogue of this column tile: phi += sum_j acc[i][j] * M[row_i][jt*TN + col_j]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int src = s_rows[ty * 4 + i];
      if (src >= 0) {
        const double2* mp = reinterpret_cast<const double2*>(Mt + (size_t)src * Sp + jt * TN + tx * 4);
        const double2 m0 = mp[0], m1 = mp[1];
        phi_acc[i] += acc[i][0] * m0.x + acc[i][1] * m0.y + acc[i][2] * m1.x + acc[i][3] * m1.y;
      }
    }
  }
  // reduce over the 16 threads (tx) that share rows: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v = phi_acc[i];
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int src = s_rows[ty * 4 + i];
    if (tx == 0 && src >= 0) phi[(size_t)t * C + src] = v;
  }
}

__global__ void loglik_kernel(const DevConfig cfg, int C, const double* __restrict__ phi, const double* __restrict__ sig,
                              double* __restrict__ logl) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double ll = 0.0;
  for (int t = 0; t < cfg.ntrc; ++t) {  // src/likelihood.f90:94-96, same operation order
    const double s = sig[(size_t)t * C + c];
    const double ph = phi[(size_t)t * C + c];
    ll = __dsub_rn(__dsub_rn(ll, __ddiv_rn(__dmul_rn(0.5, ph), __dmul_rn(s, s))), __dmul_rn((double)cfg.nsmp, log(s)));
  }
  logl[c] = ll;
}

}  // namespace

int rfinv_launch_quadform(const DevConfig& cfg, int C, const double* misfit, double* phi, const int* active,
                          int n_active, cudaStream_t stream) {
  const int n_rows = active ? n_active : C;
  if (n_rows == 0) return RFINV_OK;
  dim3 grid((n_rows + TM - 1) / TM, cfg.ntrc);
  quadform_kernel<<<grid, QF_THREADS, 0, stream>>>(cfg, C, misfit, phi, active, n_active);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}

int rfinv_launch_loglik(const DevConfig& cfg, int C, const double* phi, const double* sig, double* logl,
                        cudaStream_t stream) {
  if (C == 0) return RFINV_OK;
  loglik_kernel<<<(C + 255) / 256, 256, 0, stream>>>(cfg, C, phi, sig, logl);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}
