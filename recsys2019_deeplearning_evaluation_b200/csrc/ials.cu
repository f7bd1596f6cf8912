// K4: implicit-ALS half epochs (per-row normal equations), sm_100a.
//
// Replaces MatrixFactorization/IALSRecommender.py:137-201: for every warm user (then every warm item)
//     A = Y_p^T diag(c - 1) Y_p,   B = Y^T Y + A + reg * I,   x = B^{-1} Y_p^T c
// with Y the other side's factors, p the row's interaction profile and c its confidences (:170-201).
//
// One CTA per row.  The f x f matrix A is accumulated in REGISTERS: the 256 threads form a 16 x 16 grid, thread
// (ty, tx) owns the lower-triangular entries (r, c) with r = ty + 16 i, c = tx + 16 j, j <= i, and consumes the
// profile's factor rows from a shared-memory tile that all threads fill with coalesced loads.  B is then completed in
// shared memory in packed lower-triangular form, factorised by an in-place Cholesky and used for the two triangular
// solves (the reference forms the explicit inverse, :201; the solution is the same).  Y^T Y is accumulated by the
// same register tiling over row chunks with fp64 atomics.
// Everything is fp64 like the reference (:204-210): with the reference's all-positive initial factors the systems
// have condition numbers ~1e6, beyond what an fp32 factorisation resolves to the 1e-4 parity bar.
// ROUND-1 STATUS: CUDA-core fp64; the tensor-core (tcgen05) Gram accumulation is the round-2 item.
// Limit: n_factors <= 208 (packed fp64 matrix must fit the 227 KB of shared memory).
#include <algorithm>

#include "common.cuh"

namespace b200 {
namespace ials {

constexpr int THREADS = 256;
constexpr int TROWS = 8;  // factor rows of the profile staged per tile

__device__ __forceinline__ int pidx(int r, int c) { return r * (r + 1) / 2 + c; }  // packed lower, c <= r

// acc[i][j] (j <= i) += w * y[r_i] * y[c_j] for the rows in the tile
template <int FB>
__device__ __forceinline__ void rank_update(double (&acc)[FB][FB], const double* tile, const double* wt, int nrows, int f,
                                            int ty, int tx) {
  for (int t = 0; t < nrows; ++t) {
    const double* y = tile + t * f;
    const double w = wt[t];
    double yr[FB], yc[FB];
#pragma unroll
    for (int i = 0; i < FB; ++i) {
      const int r = ty + 16 * i, c = tx + 16 * i;
      yr[i] = r < f ? y[r] * w : 0.0;
      yc[i] = c < f ? y[c] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < FB; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) acc[i][j] += yr[i] * yc[j];
  }
}

// YtY[r, c] (full symmetric f x f, fp64) += sum over rows n in this CTA's chunk of Y[n, r] * Y[n, c]
template <int FB>
__global__ void __launch_bounds__(THREADS) gram_kernel(const double* __restrict__ Y, int n_rows, int f, double* YtY) {
  extern __shared__ double sm[];
  double* tile = sm;               // TROWS x f
  double* wt = sm + TROWS * f;     // TROWS
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  double acc[FB][FB];
#pragma unroll
  for (int i = 0; i < FB; ++i)
#pragma unroll
    for (int j = 0; j < FB; ++j) acc[i][j] = 0.0;
  const int per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n_rows, lo + per);
  for (int n0 = lo; n0 < hi; n0 += TROWS) {
    const int nr = min(TROWS, hi - n0);
    __syncthreads();
    for (int e = tid; e < nr * f; e += THREADS) tile[e] = Y[(size_t)n0 * f + e];
    if (tid < nr) wt[tid] = 1.0;
    __syncthreads();
    rank_update<FB>(acc, tile, wt, nr, f, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < FB; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      if (r < f && c < f && c <= r && acc[i][j] != 0.0) {
        atomicAdd(YtY + (size_t)r * f + c, acc[i][j]);
        if (c != r) atomicAdd(YtY + (size_t)c * f + r, acc[i][j]);
      }
    }
}

template <int FB>
__global__ void __launch_bounds__(THREADS) ials_rows_kernel(const int* __restrict__ rows, int n_solve, const int* __restrict__ ptr,
                                                            const int* __restrict__ idx, const float* __restrict__ conf,
                                                            const double* __restrict__ Y, const double* __restrict__ YtY, int f,
                                                            double reg, double* X, int* info) {
  extern __shared__ double sm[];
  const int np = f * (f + 1) / 2;
  double* Bp = sm;                     // packed lower B -> L
  double* tile = Bp + np;              // TROWS x f
  double* wt = tile + TROWS * f;       // TROWS: c - 1
  double* rhs = wt + TROWS;            // f
  double* cw = rhs + f;                // TROWS: c
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  for (int s = blockIdx.x; s < n_solve; s += gridDim.x) {
    const int row = rows[s];
    const int p0 = ptr[row], p1 = ptr[row + 1];
    double acc[FB][FB];
#pragma unroll
    for (int i = 0; i < FB; ++i)
#pragma unroll
      for (int j = 0; j < FB; ++j) acc[i][j] = 0.0;
    double my_rhs = 0.0;  // thread tid < f owns rhs[tid]
    for (int k0 = p0; k0 < p1; k0 += TROWS) {
      const int nr = min(TROWS, p1 - k0);
      __syncthreads();
      for (int e = tid; e < nr * f; e += THREADS) {
        const int t = e / f, q = e % f;
        tile[e] = Y[(size_t)idx[k0 + t] * f + q];
      }
      if (tid < nr) { const double c = (double)conf[k0 + tid]; wt[tid] = c - 1.0; cw[tid] = c; }
      __syncthreads();
      rank_update<FB>(acc, tile, wt, nr, f, ty, tx);
      if (tid < f)
        for (int t = 0; t < nr; ++t) my_rhs += cw[t] * tile[t * f + tid];  // Y_p^T c (:201)
    }
    __syncthreads();
    // B = YtY + A + reg I (:199), packed lower
#pragma unroll
    for (int i = 0; i < FB; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        if (r < f && c <= r) Bp[pidx(r, c)] = acc[i][j] + YtY[(size_t)r * f + c] + (r == c ? reg : 0.0);
      }
    if (tid < f) rhs[tid] = my_rhs;
    __syncthreads();
    // in-place Cholesky (right-looking)
    for (int j = 0; j < f; ++j) {
      if (tid == 0) {
        const double d = Bp[pidx(j, j)];
        if (!(d > 0.0)) atomicExch(info, row + 1);
        Bp[pidx(j, j)] = sqrt(fmax(d, 1e-300));
      }
      __syncthreads();
      const double djj = Bp[pidx(j, j)];
      for (int r = j + 1 + tid; r < f; r += THREADS) Bp[pidx(r, j)] /= djj;
      __syncthreads();
      const int rem = f - j - 1;
      // rows r > j, columns j < c <= r: thread-strided over rows and columns (16 x 16 grid)
      for (int r = j + 1 + ty; r < f; r += 16) {
        const double lrj = Bp[pidx(r, j)];
        for (int c = j + 1 + tx; c <= r; c += 16) Bp[pidx(r, c)] -= lrj * Bp[pidx(c, j)];
      }
      (void)rem;
      __syncthreads();
    }
    // forward solve L z = rhs, then backward L^T x = z (one warp; f is small)
    if (tid < 32) {
      for (int r = 0; r < f; ++r) {
        double part = 0.0;
        for (int c = tid; c < r; c += 32) part += Bp[pidx(r, c)] * rhs[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (tid == 0) rhs[r] = (rhs[r] - part) / Bp[pidx(r, r)];
        __syncwarp();
      }
      for (int r = f - 1; r >= 0; --r) {
        double part = 0.0;
        for (int c = r + 1 + tid; c < f; c += 32) part += Bp[pidx(c, r)] * rhs[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (tid == 0) rhs[r] = (rhs[r] - part) / Bp[pidx(r, r)];
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < f) X[(size_t)row * f + tid] = rhs[tid];
    __syncthreads();
  }
}

template <int FB>
void run(cudaStream_t st, const int* rows, int n_solve, const int* ptr, const int* idx, const float* conf, const double* Y,
         int n_other, int f, double reg, double* X, double* YtY, int* info) {
  const size_t smem_g = sizeof(double) * ((size_t)TROWS * f + TROWS);
  B200_CUDA(cudaMemsetAsync(YtY, 0, sizeof(double) * (size_t)f * f, st));
  const int grid_g = std::max(1, std::min(sm_count() * 4, (n_other + 63) / 64));
  gram_kernel<FB><<<grid_g, THREADS, smem_g, st>>>(Y, n_other, f, YtY);
  count_launch();
  const size_t smem = sizeof(double) * ((size_t)f * (f + 1) / 2 + (size_t)TROWS * f + 2 * TROWS + f);
  B200_CUDA(cudaFuncSetAttribute(ials_rows_kernel<FB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ials_rows_kernel<FB>, THREADS, smem));
  const int grid = std::max(1, std::min(n_solve, sm_count() * std::max(per_sm, 1)));
  ials_rows_kernel<FB><<<grid, THREADS, smem, st>>>(rows, n_solve, ptr, idx, conf, Y, YtY, f, reg, X, info);
  B200_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ials
}  // namespace b200

using namespace b200;

extern "C" {

int b200_ials_half_epoch_device(const int32_t* d_rows, int n_solve, const int32_t* d_ptr, const int32_t* d_idx, const float* d_conf,
                                const double* d_Y, int n_other, int n_factors, double reg, double* d_X, double* d_YtY_work,
                                void* stream) {
  return guarded([&] {
    B200_REQUIRE(d_rows && d_ptr && d_idx && d_conf && d_Y && d_X && d_YtY_work, "b200_ials_half_epoch: NULL argument");
    B200_REQUIRE(n_factors >= 1 && n_factors <= 208, "b200_ials_half_epoch: n_factors must be in [1, 208] (got %d)", n_factors);
    B200_REQUIRE(n_solve >= 0 && n_other > 0, "b200_ials_half_epoch: bad shape");
    if (n_solve == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf<int> info(1);
    B200_CUDA(cudaMemsetAsync(info.get(), 0, sizeof(int), st));
    const int f = n_factors;
    if (f <= 32) ials::run<2>(st, d_rows, n_solve, d_ptr, d_idx, d_conf, d_Y, n_other, f, reg, d_X, d_YtY_work, info.get());
    else if (f <= 64) ials::run<4>(st, d_rows, n_solve, d_ptr, d_idx, d_conf, d_Y, n_other, f, reg, d_X, d_YtY_work, info.get());
    else if (f <= 128) ials::run<8>(st, d_rows, n_solve, d_ptr, d_idx, d_conf, d_Y, n_other, f, reg, d_X, d_YtY_work, info.get());
    else ials::run<13>(st, d_rows, n_solve, d_ptr, d_idx, d_conf, d_Y, n_other, f, reg, d_X, d_YtY_work, info.get());
    int h_info = 0;
    B200_CUDA(cudaMemcpyAsync(&h_info, info.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_REQUIRE(h_info == 0, "b200_ials_half_epoch: normal equations of row %d are not positive definite", h_info - 1);
  });
}

}  // extern "C"
