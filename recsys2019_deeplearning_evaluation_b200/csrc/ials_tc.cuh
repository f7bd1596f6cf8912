// K4 on the tensor cores (sm_100a): the per-row Gram  Y_p^T diag(c - 1) Y_p  as one 3xTF32 tcgen05 tile per row,
// followed by the fp64 Cholesky of ials.cu and iterative refinement against the exact operator.
//
// STATUS: opt-in (B200REC_IALS_TC=1), n_factors <= 128; written after the round's GPU time was spent, never executed.
// tests/test_ials.py runs unchanged with the variable set -- that is the switch-over check.
//
// Why this is accurate enough (tools/ials_slicing_study.py, tests/test_ials_study.py): with the tensor core's truncating
// fp32 accumulation a 3xTF32 Gram is good to ~2e-5 on a 5 650-entry profile; the per-row systems of the reference's factors
// have cond ~ 1e2..1e4 whenever the other side has many more rows than factors, and ONE refinement step against the
// exact fp64 operator  r = b - (Y^T Y x + Y_p^T((c-1) .* (Y_p x)) + reg x)  brings the solution to ~4e-9, a second to 1e-13.
// The host only takes this path when n_other >= 4 n_factors; a row whose refined residual is still large raises `redo`
// and the host recomputes the half epoch with the fp64 kernel.
//
// One CTA (256 threads) per row, rows strided over the grid.  Per 32 profile entries: all threads gather the factor rows
// (fp64 -> fp32), split them into hi/lo TF32 parts and store them as K-major UMMA tiles -- operand A = Y_p^T (M = factor,
// K = entry), operand B = (diag(c-1) Y_p)^T -- two stages; thread 0 issues the 12 MMAs of the step (hi*hi + hi*lo + lo*hi)
// into a 128 x 128 fp32 accumulator in TMEM and commits them to the stage's mbarrier (the pipeline of gemm_tc.cuh, with a
// chunk counter that runs across rows so that the barrier phases stay consistent).  The accumulator is read back with
// tcgen05.ld, widened to fp64 and completed with Y^T Y + reg I in the packed shared-memory matrix of ials.cu.
#pragma once
#include "gemm_tc.cuh"

namespace b200 {
namespace ials {

constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 2;
constexpr int TC_KC = tc::BK;  // profile entries per step

__device__ __forceinline__ int tc_pidx(int r, int c) { return r * (r + 1) / 2 + c; }

// L z = v, then L^T x = z, in place on v (one warp, packed lower-triangular L in shared memory)
__device__ __forceinline__ void tc_chol_solve(const double* Bp, double* v, int f, int lane) {
  for (int r = 0; r < f; ++r) {
    double part = 0.0;
    for (int c = lane; c < r; c += 32) part += Bp[tc_pidx(r, c)] * v[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) v[r] = (v[r] - part) / Bp[tc_pidx(r, r)];
    __syncwarp();
  }
  for (int r = f - 1; r >= 0; --r) {
    double part = 0.0;
    for (int c = r + 1 + lane; c < f; c += 32) part += Bp[tc_pidx(c, r)] * v[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) v[r] = (v[r] - part) / Bp[tc_pidx(r, r)];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) ials_rows_tc_kernel(const int* __restrict__ rows, int n_solve, const int* __restrict__ ptr,
                                                                     const int* __restrict__ idx, const float* __restrict__ conf,
                                                                     const double* __restrict__ Y, const double* __restrict__ YtY, int f,
                                                                     double reg, double* X, int* info, int* redo, int n_refine) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem_base;
  __shared__ double s_norm[3];  // |r|^2 before the last / the first correction, |b|^2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* tiles = smem;                                             // TC_STAGES x {A_hi, A_lo, B_hi, B_lo}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * 4 * tc::TILE_BYTES);
  double* Bp = reinterpret_cast<double*>(smem + TC_STAGES * 4 * tc::TILE_BYTES + 64);  // packed lower system -> L
  double* b0 = Bp + f * (f + 1) / 2;   // right-hand side
  double* xs = b0 + f;                 // solution
  double* rs = xs + f;                 // residual / correction
  double* part = rs + f;               // [8 warps][f] partial sums of the refinement
  float* wk = reinterpret_cast<float*>(part + 8 * f);  // [TC_KC] c - 1 of the step's entries
  int* rk = reinterpret_cast<int*>(wk + TC_KC);        // [TC_KC] their factor-row indices (-1 = past the profile)
  const uint32_t bar0 = tc::smem_u32(bars);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem_base)), "r"(tc::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < TC_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem_base;

  unsigned gk = 0;  // steps issued by this CTA so far (all rows): stage = gk & 1, its barrier phase = (gk >> 1) & 1
  for (int srow = blockIdx.x; srow < n_solve; srow += gridDim.x) {
    const int row = rows[srow];
    const int p0 = ptr[row], p1 = ptr[row + 1];
    const int nk = (p1 - p0 + TC_KC - 1) / TC_KC;
    double my_rhs = 0.0;  // thread tid < f owns b0[tid]
    for (int kb = 0; kb < nk; ++kb, ++gk) {
      const int s = (int)(gk & 1u);
      unsigned char* a_hi = tiles + (s * 4 + 0) * tc::TILE_BYTES;
      unsigned char* a_lo = tiles + (s * 4 + 1) * tc::TILE_BYTES;
      unsigned char* b_hi = tiles + (s * 4 + 2) * tc::TILE_BYTES;
      unsigned char* b_lo = tiles + (s * 4 + 3) * tc::TILE_BYTES;
      __syncthreads();  // wk / rk of the previous step are consumed
      if (tid < TC_KC) {
        const int q = p0 + kb * TC_KC + tid;
        rk[tid] = q < p1 ? idx[q] : -1;
        wk[tid] = q < p1 ? conf[q] - 1.f : 0.f;
      }
      if (gk >= (unsigned)TC_STAGES) tc::mbar_wait(bar0 + 8u * s, (uint32_t)(((gk >> 1) - 1u) & 1u));  // the MMAs that read this stage are done
      __syncthreads();
      // tile(m, k) = Y[rk[k]][m] (operand A) and (c_k - 1) * Y[rk[k]][m] (operand B); zero past f and past the profile
#pragma unroll
      for (int e = 0; e < (tc::BM * TC_KC) / TC_THREADS; ++e) {
        const int id = tid + e * TC_THREADS;
        const int m = id & (tc::BM - 1), k = id >> 7;  // consecutive threads read consecutive factors of one row
        const int r = rk[k];
        const double y = (r >= 0 && m < f) ? Y[(size_t)r * f + m] : 0.0;
        float h, l;
        tc::split_tf32((float)y, h, l);
        const uint32_t off = tc::tile_offset(m, k);
        *reinterpret_cast<float*>(a_hi + off) = h;
        *reinterpret_cast<float*>(a_lo + off) = l;
        tc::split_tf32((float)(y * (double)wk[k]), h, l);
        *reinterpret_cast<float*>(b_hi + off) = h;
        *reinterpret_cast<float*>(b_lo + off) = l;
      }
      if (tid < f) {  // Y_p^T c (IALSRecommender.py:201), exact in fp64
        for (int k = 0; k < TC_KC; ++k) {
          const int r = rk[k];
          if (r >= 0) my_rhs += ((double)wk[k] + 1.0) * Y[(size_t)r * f + tid];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo), bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
#pragma unroll
        for (int ks = 0; ks < TC_KC / 8; ++ks) {
          const uint32_t o = ks * 256u;
          tc::umma_tf32(tmem, tc::make_smem_desc(ah + o), tc::make_smem_desc(bh + o), (kb > 0 || ks > 0) ? 1u : 0u);
          tc::umma_tf32(tmem, tc::make_smem_desc(ah + o), tc::make_smem_desc(bl + o), 1u);
          tc::umma_tf32(tmem, tc::make_smem_desc(al + o), tc::make_smem_desc(bh + o), 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8u * s) : "memory");
      }
    }
    if (nk > 0) {
      const unsigned last = gk - 1u;
      tc::mbar_wait(bar0 + 8u * (last & 1u), (uint32_t)((last >> 1) & 1u));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- accumulator -> packed fp64 system: B = A + Y^T Y + reg I (lower triangle); warp w owns TMEM lanes 32 (w % 4) .. + 31
    {
      const int r = (warp & 3) * 32 + lane;
      const int col0 = (warp >> 2) * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        if (nk > 0) {
          const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col0 + half * 32);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int c = col0 + half * 32 + q;
          if (r < f && c <= r) Bp[tc_pidx(r, c)] = (double)__uint_as_float(v[q]) + YtY[(size_t)r * f + c] + (r == c ? reg : 0.0);
        }
      }
    }
    if (tid < f) { b0[tid] = my_rhs; xs[tid] = my_rhs; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // every warp has read the accumulator: the next row may overwrite it

    // ---- in-place Cholesky (right-looking), as ials.cu
    for (int j = 0; j < f; ++j) {
      if (tid == 0) {
        const double d = Bp[tc_pidx(j, j)];
        if (!(d > 0.0)) atomicExch(info, row + 1);
        Bp[tc_pidx(j, j)] = sqrt(fmax(d, 1e-300));
      }
      __syncthreads();
      const double djj = Bp[tc_pidx(j, j)];
      for (int r = j + 1 + tid; r < f; r += TC_THREADS) Bp[tc_pidx(r, j)] /= djj;
      __syncthreads();
      for (int r = j + 1 + (tid >> 4); r < f; r += 16) {
        const double lrj = Bp[tc_pidx(r, j)];
        for (int c = j + 1 + (tid & 15); c <= r; c += 16) Bp[tc_pidx(r, c)] -= lrj * Bp[tc_pidx(c, j)];
      }
      __syncthreads();
    }
    if (warp == 0) tc_chol_solve(Bp, xs, f, lane);
    __syncthreads();

    // ---- iterative refinement against the exact operator, matrix-free in fp64
    for (int it = 0; it < n_refine; ++it) {
      // part[w][m] = sum over this warp's profile entries of (c_k - 1) (y_k . x) y_k[m]
      double acc[4] = {0.0, 0.0, 0.0, 0.0};  // f <= 128: lane owns m = lane, lane + 32, lane + 64, lane + 96
      for (int q = p0 + warp; q < p1; q += 8) {
        const double* y = Y + (size_t)idx[q] * f;
        double yv[4], dot = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int m = lane + 32 * t;
          yv[t] = m < f ? y[m] : 0.0;
          dot += m < f ? yv[t] * xs[m] : 0.0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        const double wd = ((double)conf[q] - 1.0) * dot;
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] += wd * yv[t];
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < f) part[warp * f + lane + 32 * t] = acc[t];
      __syncthreads();
      if (tid < f) {
        double ax = reg * xs[tid];
        for (int w = 0; w < 8; ++w) ax += part[w * f + tid];
        const double* yrow = YtY + (size_t)tid * f;
        for (int n = 0; n < f; ++n) ax += yrow[n] * xs[n];
        rs[tid] = b0[tid] - ax;
      }
      __syncthreads();
      if (warp == 0) {
        if (it == 0 || it == n_refine - 1) {  // |r|^2 before the first and before the last correction: their ratio is the contraction
          double r2 = 0.0;
          for (int m = lane; m < f; m += 32) r2 += rs[m] * rs[m];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, off);
          double b2 = 0.0;
          for (int m = lane; m < f; m += 32) b2 += b0[m] * b0[m];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) b2 += __shfl_xor_sync(0xffffffffu, b2, off);
          if (lane == 0) {
            if (it == 0) { s_norm[1] = r2; s_norm[2] = b2; }
            if (it == n_refine - 1) s_norm[0] = r2;
          }
          __syncwarp();
        }
        tc_chol_solve(Bp, rs, f, lane);
      }
      __syncthreads();
      if (tid < f) xs[tid] += rs[tid];
      __syncthreads();
    }
    // the refinement contracts the error by rho ~ cond x (Gram error) per step and the unrefined error is ~rho too, so after two
    // steps the error is ~rho^3: rho <= 0.03 keeps it below 3e-5.  The residual must therefore have shrunk at least 30-fold
    // between the first and the last correction, otherwise the approximate factor is too far from the exact operator for this
    // row (ill-conditioned system) and the fp64 path has to redo the half epoch
    // (or be at the rounding floor of the exact operator already, as for tiny systems whose 3xTF32 Gram is exact)
    if (n_refine >= 2 && tid == 0 && !(s_norm[0] <= 1e-3 * s_norm[1] || s_norm[0] <= 1e-22 * s_norm[2])) atomicExch(redo, 1);
    if (tid < f) X[(size_t)row * f + tid] = xs[tid];
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tc::TMEM_COLS) : "memory");
}

inline size_t tc_smem_bytes(int f) {
  return (size_t)TC_STAGES * 4 * tc::TILE_BYTES + 64 + sizeof(double) * ((size_t)f * (f + 1) / 2 + 3 * (size_t)f + 8 * (size_t)f) +
         (sizeof(float) + sizeof(int)) * TC_KC;
}

}  // namespace ials
}  // namespace b200
