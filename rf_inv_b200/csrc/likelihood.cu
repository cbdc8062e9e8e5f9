// Correlated-noise likelihood for a batch of models (src/likelihood.f90:85-98):
//   phi_t(c) = m^T R_t^-1 m,   logL(c) = sum_t  -0.5 phi_t / sig_t^2 - nsmp log(sig_t)
// The reference evaluates phi with a dense matmul + dot product per chain, streaming R^-1 (S x S) for
// every evaluation.  Here the misfits of all chains form a matrix M (chains x S) and phi is the row-wise
// diagonal of M R^-1 M^T: one fp64 GEMM against the shared, symmetric R^-1 with the row-dot fused into
// the epilogue.  R^-1 is symmetric, so only tiles on or below the diagonal are visited (off-diagonal
// tiles count twice); the summation order is fixed (no atomics), results are run-to-run deterministic.
// Where R^-1 = W W^T with few columns (the truncated pseudo-inverse the reference builds is positive semi-definite of
// rank ~0.4 S at a = 4), phi = |W^T m|^2: the same GEMM against W^T (2 S r flop) with a sum of squares as epilogue.
#include "rfinv_common.cuh"

namespace {

constexpr int TM = 64;    // chains per CTA
constexpr int TN = 64;    // columns of R^-1 per tile
constexpr int TK = 16;    // k-slab per pipeline stage
constexpr int LDS_STRIDE = TK + 4;   // doubles; (row*20 + k) mod 16 distinct over an 8x4 fragment: conflict-free LDS.64
constexpr int QF_THREADS = 128;      // 4 warps, each owns 32 rows x 4 interleaved 8-column fragments (columns 8 (2j + wn) ..)
constexpr int QF_CHUNK = 384;        // (trace, row block) pairs per scheduling chunk: 384 x 64 rows x 4 KB = 100 MB of misfits at S = 512

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Work item = (64 chains, one trace, one column tile jt of the symmetric R^-1).  It accumulates
//   Y(64 x 64) = 2 * sum_{kt<jt} M[:,kt] R[kt,jt]  +  M[:,jt] R[jt,jt]          (DMMA, fp64 accumulate)
// and the row-dot  sum_cols Y .* M[:,jt]  -> partial[jt][t][chain].  Both MMA operands are read as "row-major,
// k contiguous": M[chain][k] and, by symmetry, R[col][k].  The cost of an item grows with jt, so persistent CTAs
// take items from an atomic counter, largest first.  The CTA that delivers the last of the ntile partials of a
// (64 chains, trace) block sums them in the fixed order jt = 0..ntile-1 into phi: no floating-point atomics, results
// are bit-identical from run to run whatever the schedule.
__global__ void __launch_bounds__(QF_THREADS, 4) quadform_kernel(const DevConfig cfg, int C, const double* __restrict__ misfit,
                                                              double* __restrict__ phi, double* __restrict__ partial,
                                                              int* __restrict__ arrivals, int* __restrict__ work,
                                                              const int* __restrict__ active, int n_active,
                                                              const int* __restrict__ n_active_dev, const double* __restrict__ sig,
                                                              double* __restrict__ logl) {
  __shared__ __align__(16) double sA[2][TM * LDS_STRIDE];
  __shared__ __align__(16) double sB[2][TN * LDS_STRIDE];
  __shared__ int s_rows[TM];
  __shared__ double s_part[2][TM];
  __shared__ int s_next, s_last;
  const int Sp = cfg.nsmp_pad, T = cfg.ntrc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;          // warp sub-tile: rows wm*32 .., column fragments 2j + wn
  const int n_rb_grid = ((active ? n_active : C) + TM - 1) / TM;   // row blocks the item index space is built on
  const int ntile = cfg.qf_tiles_max;               // column tiles of the trace with the most work items
  const int n_items = ntile * T * n_rb_grid;
  const int fr = lane >> 2, fk = lane & 3;          // fragment coordinates
  if (tid == 0) s_next = (int)gridDim.x + atomicAdd(work, 1);   // (the counter was left at zero by the previous launch)
  pdl_wait();                                       // forward_kernel has delivered the misfit rows
  const int n_rows = active ? (n_active_dev ? *n_active_dev : n_active) : C;
  int item = blockIdx.x;
  while (item < n_items) {
    // item order: (trace, row block) pairs in chunks of QF_CHUNK (their misfit rows, 64 x Sp doubles each, stay in L2
    // while the chunk's column tiles are worked through), inside a chunk the largest column tiles first
    const int nblk = T * n_rb_grid;
    const int chunk = item / (ntile * QF_CHUNK), local = item - chunk * (ntile * QF_CHUNK);
    const int cb = min(QF_CHUNK, nblk - chunk * QF_CHUNK);   // blocks in this chunk (the last one may be short)
    // most expensive column tiles first (dense form: the cost grows with jt; factor forms: full tiles before partial ones),
    // so the cheap ones fill the tail of the launch
    const int jt = cfg.qf_order[local / cb];
    const int blk = chunk * QF_CHUNK + local % cb;
    const int t = blk / n_rb_grid, rb = blk - t * n_rb_grid;
    const int row0 = rb * TM;
    const int rank = cfg.qf_rank[t];                // > 0: factor form for this trace
    const int ntile_t = cfg.qf_tiles[t];
    __syncthreads();                                // previous item is done with s_rows / s_part; s_next is visible
    const int next = s_next;
    int next_raw = 0;
    if (tid == 0) next_raw = atomicAdd(work, 1);    // consumed at the end of the item
    if (row0 < n_rows && jt < ntile_t) {
      if (tid < TM) {
        int r = row0 + tid;
        r = r < n_rows ? r : n_rows - 1;            // clamp: duplicates are computed but never written
        s_rows[tid] = active ? active[r] : r;
      }
      __syncthreads();
      const double* __restrict__ Mt = misfit + (size_t)t * C * Sp;
      const double* __restrict__ Rt = rank > 0 ? cfg.w_fac + (size_t)t * cfg.qf_wrows * Sp : cfg.r_inv + (size_t)t * Sp * Sp;
      // 8-column fragments of this warp that hold columns of W (all 4 in the dense form); k range of the item: the whole
      // misfit row, or in the split form the half that holds s (tiles of the symmetric block) or a (antisymmetric block)
      int jw = 4, koff = 0, nk = (jt + 1) * (TN / TK);   // dense: k-slabs up to and including the diagonal tile
      if (rank > 0) {
        int cols_tile = rank - jt * TN;
        nk = Sp / TK;
        if (cfg.qf_split[t]) {
          const int rs = cfg.qf_rank_s[t], ts = (rs + TN - 1) / TN;
          cols_tile = jt < ts ? rs - jt * TN : (rank - rs) - (jt - ts) * TN;
          koff = jt < ts ? 0 : Sp >> 1;
          nk = (Sp >> 1) / TK;
        }
        const int nfrag = cols_tile >= TN ? 8 : (cols_tile <= 0 ? 0 : (cols_tile + 7) >> 3);   // 8-column fragments in use
        jw = (nfrag + 1 - wn) >> 1;               // fragments 2j + wn of this warp: a partial tile loads both warp columns evenly
      }
      // global -> smem copy assignment: 64 rows x 16 doubles = 512 x 16 B per operand, 4 per thread each
      const double* a_src[4];
      int cp_off[4], cp_row[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int id = tid + q * QF_THREADS;        // 0..511
        const int row = id >> 3, seg = id & 7;      // 8 segments of 2 doubles per row
        cp_row[q] = row;
        cp_off[q] = row * LDS_STRIDE + seg * 2;
        a_src[q] = Mt + (size_t)s_rows[row] * Sp + seg * 2 + koff;
      }
      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      const double* b_base = Rt + (size_t)(jt * TN) * Sp + koff;
      // prologue: stage 0
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        cp_async16(&sA[0][cp_off[q]], a_src[q]);
        cp_async16(&sB[0][cp_off[q]], b_base + (size_t)cp_row[q] * Sp + (cp_off[q] - cp_row[q] * LDS_STRIDE));
      }
      cp_async_commit();
      for (int ks = 0; ks < nk; ++ks) {
        const int cur = ks & 1;
        // ONE barrier per k-slab: it publishes the slab that has just landed and, at the same time, tells every thread that
        // the other buffer (read during the previous slab) is free -- so the next slab is requested right behind it and has
        // the whole of this slab's MMAs to arrive
        cp_async_wait<0>();
        __syncthreads();
        if (ks + 1 < nk) {
          const int k0 = (ks + 1) * TK;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            cp_async16(&sA[cur ^ 1][cp_off[q]], a_src[q] + k0);
            cp_async16(&sB[cur ^ 1][cp_off[q]], b_base + (size_t)cp_row[q] * Sp + k0 + (cp_off[q] - cp_row[q] * LDS_STRIDE));
          }
          cp_async_commit();
        }
        if (rank == 0 && ks == jt * (TN / TK) && jt > 0) {   // entering the diagonal tile: what came before counts twice
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] *= 2.0; acc[i][j][1] *= 2.0; }
        }
        const double* A = &sA[cur][(wm * 32 + fr) * LDS_STRIDE + fk];
        const double* B = &sB[cur][(wn * 8 + fr) * LDS_STRIDE + fk];   // fragment j of this warp = columns 8 (2j + wn) .. + 7
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
          double a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = A[i * 8 * LDS_STRIDE + kk];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = B[j * 16 * LDS_STRIDE + kk];
          if (jw == 4) {                            // the hot path stays one straight run of 16 DMMA
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < jw) {
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
              }
          }
        }
      }
      // row-dot with M[:, jt]: lane holds Y[row = 8i + fr][col = 8j + 2*fk + {0,1}] of its warp sub-tile
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double* mrow = Mt + (size_t)s_rows[wm * 32 + 8 * i + fr] * Sp + jt * TN + wn * 8 + 2 * fk;
        double v = 0.0;
        if (rank > 0) {                             // |W^T m|^2: columns beyond the rank hold exact zeros
#pragma unroll
          for (int j = 0; j < 4; ++j) { v = fma(acc[i][j][0], acc[i][j][0], v); v = fma(acc[i][j][1], acc[i][j][1], v); }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 m = *reinterpret_cast<const double2*>(mrow + 16 * j);
            v = fma(acc[i][j][0], m.x, v);
            v = fma(acc[i][j][1], m.y, v);
          }
        }
        // reduce over the 4 lanes sharing a row; the two warps (wn) covering the 64 columns meet in shared memory
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (fk == 0) s_part[wn][wm * 32 + 8 * i + fr] = v;
      }
      __syncthreads();
      double* part_t = partial + (size_t)t * C;                      // [jt][T][C], indexed by chain
      const size_t jstride = (size_t)T * C;
      if (tid < TM && row0 + tid < n_rows) part_t[(size_t)jt * jstride + s_rows[tid]] = s_part[0][tid] + s_part[1][tid];
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(&arrivals[t * n_rb_grid + rb], 1) == ntile_t - 1);
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (tid < TM && row0 + tid < n_rows) {
          const int c = s_rows[tid];
          double v = 0.0;
          for (int j = 0; j < ntile_t; ++j) v += __ldcg(part_t + (size_t)j * jstride + c);   // fixed order
          phi[(size_t)t * C + c] = v;
        }
        if (tid == 0) arrivals[t * n_rb_grid + rb] = 0;              // ready for the next launch
        // logL of the block's chains (src/likelihood.f90:94-96, same operation order) by the CTA that completes the last
        // trace of the block: one kernel less per evaluation
        if (logl) {
          __threadfence();
          __syncthreads();
          if (tid == 0) s_last = (atomicAdd(&arrivals[T * n_rb_grid + rb], 1) == T - 1);
          __syncthreads();
          if (s_last) {
            __threadfence();
            if (tid < TM && row0 + tid < n_rows) {
              const int c = s_rows[tid];
              double ll = 0.0;
              for (int tt = 0; tt < T; ++tt) {
                const double sg = sig[(size_t)tt * C + c];
                const double ph = __ldcg(phi + (size_t)tt * C + c);
                ll = __dsub_rn(__dsub_rn(ll, __ddiv_rn(__dmul_rn(0.5, ph), __dmul_rn(sg, sg))), __dmul_rn((double)cfg.nsmp, log(sg)));
              }
              logl[c] = ll;
            }
            if (tid == 0) arrivals[T * n_rb_grid + rb] = 0;
          }
        }
      }
    } else {
      __syncthreads();   // skipped item (no such tile / row block): everyone has read s_next before it is rewritten
    }
    if (tid == 0) s_next = (int)gridDim.x + next_raw;
    item = next;
  }
  // the last CTA to leave resets the work counter for the next launch (no memset between the kernels of an evaluation)
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(work + 1, 1) == (int)gridDim.x - 1) { work[0] = 0; work[1] = 0; }
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

// column tiles per (64 chains, trace): nsmp_pad/64 dense, up to 2 ceil(nsmp_pad/128) in the split form
size_t rfinv_quadform_partial_doubles(const DevConfig& cfg, int C) { return (size_t)(cfg.nsmp_pad / TN + 1) * cfg.ntrc * C; }
size_t rfinv_quadform_counter_ints(const DevConfig& cfg, int C) { return (size_t)((C + TM - 1) / TM) * (cfg.ntrc + 1) + 4; }

// partial: rfinv_quadform_partial_doubles(cfg, C) doubles; counters: rfinv_quadform_counter_ints(cfg, capacity) ints, zeroed
// once at allocation (the arrival counters and the work counter reset themselves at the end of every launch)
int rfinv_launch_quadform(const DevConfig& cfg, int C, const double* misfit, double* phi, double* partial, int* counters,
                          const int* active, int n_active, const int* n_active_dev, cudaStream_t stream, const double* sig,
                          double* logl) {
  const int n_rows = active ? n_active : C;
  if (n_rows == 0) return RFINV_OK;
  const size_t ntile = cfg.qf_tiles_max;
  int* work = counters;
  int* arrivals = counters + 4;
  // resident CTAs of the current device, cached per device id (handles may live on different devices; concurrent first
  // calls compute and store the same value)
  static int resident_of_device[64] = {0};
  int dev = 0;
  RFINV_CUDA_CHECK(cudaGetDevice(&dev));
  long long resident = (dev >= 0 && dev < 64) ? resident_of_device[dev] : 0;
  if (resident == 0) {
    int per_sm = 0, n_sm = 0;
    RFINV_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    RFINV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quadform_kernel, QF_THREADS, 0));
    resident = (long long)n_sm * (per_sm < 1 ? 1 : per_sm);
    if (dev >= 0 && dev < 64) resident_of_device[dev] = (int)resident;
  }
  const long long items = (long long)ntile * cfg.ntrc * ((n_rows + TM - 1) / TM);
  RFINV_CUDA_CHECK(rfinv_launch_pdl(2, quadform_kernel, dim3((unsigned)(items < resident ? items : resident)), dim3(QF_THREADS), 0, stream, cfg, C,
                                    misfit, phi, partial, arrivals, work, active, n_active, n_active_dev, sig, logl));
  return RFINV_OK;
}

int rfinv_launch_loglik(const DevConfig& cfg, int C, const double* phi, const double* sig, double* logl,
                        cudaStream_t stream) {
  if (C == 0) return RFINV_OK;
  loglik_kernel<<<(C + 255) / 256, 256, 0, stream>>>(cfg, C, phi, sig, logl);
  RFINV_CUDA_CHECK(cudaGetLastError());
  return RFINV_OK;
}
