// K1: sparse column-column similarity with exact top-K selection, sm_100a.
//
// Replaces Base/Similarity/Cython/Compute_Similarity_Cython.pyx:327-408 (computeItemSimilarities, the
// Gustavson row-gather into an n_columns accumulator) and :467-568 (normalise, top-K, emit).
//
// Design (DESIGN.md "K1"): one persistent CTA per SM pulls target columns from an atomic counter.  For a
// target column i the CTA stages i's CSC entries (user, x_ui * w_u) in shared memory, then -- one window
// of the neighbour axis at a time -- every warp streams whole CSR row segments of those users with
// coalesced vector loads and scatter-adds x_ui*x_uj into a shared-memory accumulator that covers the
// window (fp32 CAS-add, or a native integer ATOMS add when every stored value is 1).  Rows are sorted, so
// the segment of a row that falls in window w is a contiguous range whose bounds are precomputed
// (`split`).  The window is then scanned: a dot-product histogram bootstraps a lower bound of the K-th
// similarity, an upper bound of each cell's similarity (using the extreme column norms) prunes cells that
// cannot qualify before their norm is gathered, survivors go to a candidate buffer of 64-bit keys
// (similarity bits << 32 | ~index => ties resolve to the ascending index), and an 8-bit MSB radix select
// keeps the K best.  Bytes per gathered entry: 8 (index + value) or 4 (binary path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace sim {

typedef unsigned long long u64;

constexpr int THREADS = 1024;
constexpr int NWARPS = THREADS / 32;
constexpr int COLCHUNK = 1024;  // users of the target column staged per chunk
constexpr int HBINS = 4096;     // bootstrap histogram bins (aliases the candidate buffer)
constexpr int UB = 4;           // row segments in flight per warp

enum Formula { F_PROD = 0, F_NONORM = 1, F_JACCARD = 2, F_DICE = 3, F_TVERSKY = 4 };

struct KParams {
  int n_cols, K, n_win, win, cap;
  int formula;
  float se;          // shrink + 1e-6
  float shrink_div;  // shrink if != 0 else 1
  float ta, tb;
  const int* __restrict__ csr_ptr;
  const int2* __restrict__ csr_ent;
  const int* __restrict__ csr_idx;
  const int* __restrict__ split;
  const int* __restrict__ csc_ptr;
  const int2* __restrict__ csc_ent;
  const int* __restrict__ csc_idx;
  const float* __restrict__ A;
  const float* __restrict__ B;
  float B_min, B_max;
  int col_begin, n_range;
  const int* __restrict__ order;  // processing order of local columns (descending work), or nullptr
  int* counter;
  int* out_idx;
  float* out_val;
  int* out_cnt;
  int signed_data;
};

__device__ __forceinline__ float sim_value(const KParams& p, float d, float a, float b) {
  switch (p.formula) {
    case F_PROD: return d / (a * b + p.se);
    case F_NONORM: return d / p.shrink_div;
    case F_JACCARD: return d / (a + b - d + p.se);
    case F_DICE: return d / (a + b + p.se);
    default: return d / (d + (a - d) * p.ta + (b - d) * p.tb + p.se);
  }
}

template <bool NEG>
__device__ __forceinline__ unsigned key32_of(float v) {
  return NEG ? ~__float_as_uint(v) : __float_as_uint(v);
}

struct Shared {
  int col;
  int nbuf;
  int overflow;
  int npos, nneg;
  int digit, need, bincnt;
  int b0;
  int cnt;
  int warp_tot[NWARPS];
  int dig[256];
};

// exclusive suffix sum over the block: returns sum of v over all threads with a larger thread index
__device__ __forceinline__ int block_suffix_excl(int v, int* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int t = __shfl_down_sync(0xffffffffu, incl, off);
    if (lane + off < 32) incl += t;
  }
  if (lane == 0) warp_tot[warp] = incl;
  __syncthreads();
  int above = 0;
  for (int w = warp + 1; w < NWARPS; ++w) above += warp_tot[w];
  __syncthreads();
  return above + incl - v;
}

// Keeps the `K` largest of buf[0..n) (n > K, distinct keys) in buf[0..K) and returns a threshold t such that
// exactly those keys are >= t.  MSB-first radix select, 8 bits per pass, early exit when a bin is taken whole.
// If lo<hi, entries whose index field (0xFFFFFFFF - low32) lies in [lo,hi) are dropped after the selection
// (used by the overflow path); *n_out receives the surviving count.
__device__ u64 block_select(u64* buf, int n, int K, Shared* sh, int drop_lo, int drop_hi, int* n_out) {
  const int tid = threadIdx.x;
  u64 prefix = 0, mask = 0;
  int need = K;
  for (int shift = 56; shift >= 0; shift -= 8) {
    if (tid < 256) sh->dig[tid] = 0;
    __syncthreads();
    for (int q = tid; q < n; q += THREADS) {
      u64 k = buf[q];
      if ((k & mask) == prefix) atomicAdd(&sh->dig[(int)((k >> shift) & 255ull)], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int c[8], local = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) { c[b] = sh->dig[tid * 8 + b]; local += c[b]; }
      int incl = local;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_down_sync(0xffffffffu, incl, off);
        if (tid + off < 32) incl += t;
      }
      int cum = incl - local;  // keys in higher bins
#pragma unroll
      for (int b = 7; b >= 0; --b) {
        if (cum < need && cum + c[b] >= need) { sh->digit = tid * 8 + b; sh->need = need - cum; sh->bincnt = c[b]; }
        cum += c[b];
      }
    }
    __syncthreads();
    prefix |= ((u64)sh->digit) << shift;
    mask |= 255ull << shift;
    need = sh->need;
    const int bincnt = sh->bincnt;
    __syncthreads();
    if (bincnt == need) break;
  }
  // compaction through registers (cap <= 8 * THREADS)
  u64 keep[8];
  if (tid == 0) sh->cnt = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int i = q * THREADS + tid;
    u64 k = (i < n) ? buf[i] : 0ull;
    bool ok = (i < n) && (k >= prefix);
    if (ok && drop_lo < drop_hi) {
      int idx = (int)(0xFFFFFFFFu - (unsigned)k);
      if (idx >= drop_lo && idx < drop_hi) ok = false;
    }
    keep[q] = ok ? k : 0ull;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (keep[q]) buf[atomicAdd(&sh->cnt, 1)] = keep[q];
  }
  __syncthreads();
  *n_out = sh->cnt;
  __syncthreads();
  return prefix;
}

template <bool BINARY, bool NEG>
__device__ void process_column(const KParams& p, int col, int target, int out_base, unsigned char* smem_raw,
                               Shared* sh, int* n_emitted) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* accf = reinterpret_cast<float*>(smem_raw);
  int* acci = reinterpret_cast<int*>(smem_raw);
  u64* buf = reinterpret_cast<u64*>(smem_raw + (size_t)p.win * 4);
  int* hist = reinterpret_cast<int*>(buf);
  float* sm_x = reinterpret_cast<float*>(buf + p.cap);
  int* sm_s = reinterpret_cast<int*>(sm_x + COLCHUNK);
  int* sm_e = sm_s + COLCHUNK;

  const int cs = p.csc_ptr[col], ce = p.csc_ptr[col + 1];
  const float Ai = p.A[col];
  u64 thr = 0;
  int nbuf = 0;
  if (tid == 0) { sh->nbuf = 0; if (!NEG) { sh->npos = 0; sh->nneg = 0; } }
  __syncthreads();

  for (int w = 0; w < p.n_win; ++w) {
    const int win_lo = w * p.win;
    const int win_n = min(p.win, p.n_cols - win_lo);
    const int win4 = (win_n + 3) >> 2;

    // ---------------- accumulate: acc[j - win_lo] += x_ui * x_uj over users u of column `col`
    for (int k0 = cs; k0 < ce; k0 += COLCHUNK) {
      const int n = min(COLCHUNK, ce - k0);
      __syncthreads();
      for (int t = tid; t < n; t += THREADS) {
        int u;
        float x;
        if (BINARY) {
          u = p.csc_idx[k0 + t];
          x = 1.0f;
        } else {
          int2 e = p.csc_ent[k0 + t];
          u = e.x;
          x = __int_as_float(e.y);
        }
        int s, e_;
        if (p.n_win == 1) {
          s = p.csr_ptr[u];
          e_ = p.csr_ptr[u + 1];
        } else {
          const int* sp = p.split + (size_t)u * (p.n_win + 1) + w;
          s = sp[0];
          e_ = sp[1];
        }
        sm_x[t] = x;
        sm_s[t] = s;
        sm_e[t] = e_;
      }
      __syncthreads();
      for (int t0 = warp; t0 < n; t0 += NWARPS * UB) {
        int s[UB], len[UB];
        float x[UB];
        int maxlen = 0;
#pragma unroll
        for (int k = 0; k < UB; ++k) {
          const int t = t0 + k * NWARPS;
          if (t < n) {
            s[k] = sm_s[t];
            len[k] = sm_e[t] - s[k];
            x[k] = sm_x[t];
          } else {
            s[k] = 0; len[k] = 0; x[k] = 0.f;
          }
          maxlen = max(maxlen, len[k]);
        }
        for (int r0 = 0; r0 < maxlen; r0 += 32) {
          const int r = r0 + lane;
          if (BINARY) {
            int j[UB];
#pragma unroll
            for (int k = 0; k < UB; ++k) j[k] = (r < len[k]) ? __ldg(p.csr_idx + s[k] + r) : -1;
#pragma unroll
            for (int k = 0; k < UB; ++k)
              if (j[k] >= 0 && j[k] != col) atomicAdd(&acci[j[k] - win_lo], 1);
          } else {
            int2 e[UB];
#pragma unroll
            for (int k = 0; k < UB; ++k) e[k] = (r < len[k]) ? __ldg(p.csr_ent + s[k] + r) : make_int2(-1, 0);
#pragma unroll
            for (int k = 0; k < UB; ++k)
              if (e[k].x >= 0 && e[k].x != col) atomicAdd(&accf[e[k].x - win_lo], x[k] * __int_as_float(e[k].y));
          }
        }
      }
    }
    __syncthreads();

    // ---------------- bootstrap a lower bound of the target-th best similarity from the dot histogram
    nbuf = sh->nbuf;
    if (!NEG && thr == 0 && nbuf == 0 && win_n > 2 * target) {
      for (int i = tid; i < HBINS; i += THREADS) hist[i] = 0;
      if (tid == 0) sh->b0 = -1;
      __syncthreads();
      for (int i4 = tid; i4 < win4; i4 += THREADS) {
        float d[4];
        if (BINARY) {
          int4 v = reinterpret_cast<const int4*>(acci)[i4];
          d[0] = (float)v.x; d[1] = (float)v.y; d[2] = (float)v.z; d[3] = (float)v.w;
        } else {
          float4 v = reinterpret_cast<const float4*>(accf)[i4];
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (d[c] > 0.f) atomicAdd(&hist[min(__float_as_uint(d[c]) >> 19, (unsigned)(HBINS - 1))], 1);
      }
      __syncthreads();
      int4 h = reinterpret_cast<const int4*>(hist)[tid];
      const int local = h.x + h.y + h.z + h.w;
      int cum = block_suffix_excl(local, sh->warp_tot);
      const int hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int b = 3; b >= 0; --b) {
        if (cum < target && cum + hh[b] >= target) sh->b0 = tid * 4 + b;
        cum += hh[b];
      }
      __syncthreads();
      const int b0 = sh->b0;
      if (b0 > 0) {
        const float d0 = __uint_as_float(((unsigned)b0) << 19);
        const float lb = fminf(sim_value(p, d0, Ai, p.B_min), sim_value(p, d0, Ai, p.B_max));
        if (lb > 0.f) thr = ((u64)__float_as_uint(lb)) << 32;
      }
      __syncthreads();  // hist (aliasing buf) fully consumed before candidates are pushed
    }

    // ---------------- scan: prune by upper bound, evaluate, push candidates; retry after an overflow
    bool first = true;
    while (true) {
      if (tid == 0) sh->overflow = 0;
      __syncthreads();
      const unsigned thr32 = (unsigned)(thr >> 32);
      int cpos = 0, cneg = 0;
      for (int i4 = tid; i4 < win4; i4 += THREADS) {
        float d[4];
        if (BINARY) {
          int4 v = reinterpret_cast<const int4*>(acci)[i4];
          d[0] = (float)v.x; d[1] = (float)v.y; d[2] = (float)v.z; d[3] = (float)v.w;
        } else {
          float4 v = reinterpret_cast<const float4*>(accf)[i4];
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float dd = d[c];
          if (dd == 0.f) continue;
          if (!NEG && first) { cpos += dd > 0.f; cneg += dd < 0.f; }
          if (NEG ? (dd < 0.f) : (dd > 0.f)) {
            const unsigned ub = max(key32_of<NEG>(sim_value(p, dd, Ai, p.B_min)),
                                    key32_of<NEG>(sim_value(p, dd, Ai, p.B_max)));
            if (ub >= thr32) {
              const int j = win_lo + i4 * 4 + c;
              const float sv = sim_value(p, dd, Ai, __ldg(p.B + j));
              const u64 key = (((u64)key32_of<NEG>(sv)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)j);
              if (key >= thr && (NEG ? (sv < 0.f) : (sv > 0.f))) {
                const int pos = atomicAdd(&sh->nbuf, 1);
                if (pos < p.cap) buf[pos] = key; else sh->overflow = 1;
              }
            }
          }
        }
      }
      if (!NEG && first && p.signed_data) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          cpos += __shfl_xor_sync(0xffffffffu, cpos, off);
          cneg += __shfl_xor_sync(0xffffffffu, cneg, off);
        }
        if (lane == 0) { if (cpos) atomicAdd(&sh->npos, cpos); if (cneg) atomicAdd(&sh->nneg, cneg); }
      }
      first = false;
      __syncthreads();
      if (!sh->overflow) break;
      // overflow: keep the best `target` of the full buffer, drop this window's entries (the rescan re-pushes
      // the ones that still qualify), tighten the threshold, rescan.
      int kept;
      thr = max(thr, block_select(buf, p.cap, target, sh, win_lo, win_lo + win_n, &kept));
      if (tid == 0) sh->nbuf = kept;
      __syncthreads();
    }
    nbuf = sh->nbuf;
    if (nbuf > target && (nbuf > p.cap / 2 || w == p.n_win - 1)) {
      int kept;
      thr = max(thr, block_select(buf, nbuf, target, sh, 0, 0, &kept));
      if (tid == 0) sh->nbuf = kept;
      nbuf = kept;
    }
    // ---------------- clear the window
    {
      int4 z = make_int4(0, 0, 0, 0);
      for (int i4 = tid; i4 < win4; i4 += THREADS) reinterpret_cast<int4*>(acci)[i4] = z;
    }
    __syncthreads();
  }

  // ---------------- emit
  for (int t = tid; t < nbuf; t += THREADS) {
    const u64 k = buf[t];
    const unsigned hi = (unsigned)(k >> 32);
    p.out_idx[(size_t)out_base + t] = (int)(0xFFFFFFFFu - (unsigned)k);
    p.out_val[(size_t)out_base + t] = __uint_as_float(NEG ? ~hi : hi);
  }
  *n_emitted = nbuf;
  __syncthreads();
}

template <bool BINARY>
__global__ void __launch_bounds__(THREADS, 1) sim_topk_kernel(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Shared sh;
  const int tid = threadIdx.x;
  {
    int4 z = make_int4(0, 0, 0, 0);
    for (int i4 = tid; i4 < (p.win >> 2); i4 += THREADS) reinterpret_cast<int4*>(smem_raw)[i4] = z;
  }
  __syncthreads();
  while (true) {
    if (tid == 0) sh.col = atomicAdd(p.counter, 1);
    __syncthreads();
    const int c = sh.col;
    if (c >= p.n_range) break;
    const int lc = p.order ? p.order[c] : c;
    const int col = p.col_begin + lc;
    const int out_base_row = lc;
    int n_out = 0;
    process_column<BINARY, false>(p, col, p.K, out_base_row * p.K, smem_raw, &sh, &n_out);
    if (p.signed_data && n_out < p.K) {
      // zeros outrank negatives (Compute_Similarity_Python.py:335-345): negatives are only emitted when the
      // positives plus the implicit zeros (every column without a non-zero similarity, the diagonal
      // included) do not fill K slots.
      const int npos = sh.npos, nneg = sh.nneg;
      const int nzero = p.n_cols - npos - nneg;
      const int m = p.K - n_out - nzero;
      __syncthreads();
      if (m > 0 && nneg > 0) {
        int n_neg_out = 0;
        process_column<BINARY, true>(p, col, m, out_base_row * p.K + n_out, smem_raw, &sh, &n_neg_out);
        n_out += n_neg_out;
      }
    }
    for (int t = n_out + tid; t < p.K; t += THREADS) {
      p.out_idx[(size_t)out_base_row * p.K + t] = -1;
      p.out_val[(size_t)out_base_row * p.K + t] = 0.f;
    }
    if (tid == 0) p.out_cnt[out_base_row] = n_out;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// preprocessing kernels (constructor work of pyx:147-209, on the device)
// ------------------------------------------------------------------------------------------------------
__global__ void flags_kernel(const float* __restrict__ data, long long nnz, int* flags) {
  int f = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    float v = data[i];
    if (v != 1.0f) f |= 1;
    if (v < 0.f) f |= 2;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

__global__ void fill_ones_kernel(float* data, long long nnz) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x)
    data[i] = 1.0f;
}

// adjusted cosine: subtract the mean of the stored entries of each row (pyx:277-312); one warp per row
__global__ void row_center_kernel(const int* __restrict__ ptr, float* data, int n_rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int s = ptr[warp], e = ptr[warp + 1];
  if (e <= s) return;
  double sum = 0.0;
  for (int q = s + lane; q < e; q += 32) sum += (double)data[q];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const double mean = sum / (double)(e - s);
  for (int q = s + lane; q < e; q += 32) data[q] = (float)((double)data[q] - mean);
}

__global__ void col_accum_kernel(const int* __restrict__ idx, const float* __restrict__ data, long long nnz,
                                 double* colsum, double* colsq, int* colcnt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    const int j = idx[i];
    const double v = (double)data[i];
    if (colsum) atomicAdd(colsum + j, v);
    if (colsq) atomicAdd(colsq + j, v * v);
    if (colcnt) atomicAdd(colcnt + j, 1);
  }
}

// pearson: subtract the per-column mean of stored entries (pyx:236-273)
__global__ void col_center_kernel(const int* __restrict__ idx, float* data, long long nnz,
                                  const double* __restrict__ colsum, const int* __restrict__ colcnt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    const int j = idx[i];
    const int c = colcnt[j];
    if (c > 0) data[i] = (float)((double)data[i] - colsum[j] / (double)c);
  }
}

__global__ void norms_kernel(const double* __restrict__ colsq, int n_cols, int mode, float alpha, float* A, float* B) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_cols) return;
  const double sq = colsq[j];
  if (mode == 0) {  // set kinds: plain sum of squares (pyx:170-174)
    A[j] = (float)sq; B[j] = (float)sq;
  } else if (mode == 1) {  // cosine family
    const float n = (float)sqrt(sq);
    A[j] = n; B[j] = n;
  } else {  // asymmetric (pyx:176-180)
    const double n = sqrt(sq) + 1e-6;
    A[j] = (float)pow(n, 2.0 * (double)alpha);
    B[j] = (float)pow(n, 2.0 * (1.0 - (double)alpha));
  }
}

__global__ void rowid_iota_kernel(const int* __restrict__ ptr, int n_rows, int* rowid, int* iota) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int s = ptr[warp], e = ptr[warp + 1];
  for (int q = s + lane; q < e; q += 32) { rowid[q] = warp; iota[q] = q; }
}

__global__ void build_csr_ent_kernel(const int* __restrict__ idx, const float* __restrict__ data, long long nnz, int2* ent) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x)
    ent[i] = make_int2(idx[i], __float_as_int(data[i]));
}

// CSC side: entry q of the column-sorted permutation -> (user, x_ui * w_u)
__global__ void build_csc_kernel(const int* __restrict__ perm, const int* __restrict__ rowid,
                                 const float* __restrict__ data, const float* __restrict__ row_w, long long nnz,
                                 int2* ent, int* idx_only) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nnz; q += (long long)gridDim.x * blockDim.x) {
    const int pp = perm[q];
    const int u = rowid[pp];
    if (idx_only) {
      idx_only[q] = u;
    } else {
      float x = data[pp];
      if (row_w) x *= row_w[u];
      ent[q] = make_int2(u, __float_as_int(x));
    }
  }
}

// split[u*(n_win+1)+w] = first position of row u whose column index >= w*win  (rows are sorted)
__global__ void split_kernel(const int* __restrict__ ptr, const int* __restrict__ idx, int n_rows, int n_win,
                             int win, int* split) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)n_rows * (n_win + 1);
  if (g >= total) return;
  const int u = (int)(g / (n_win + 1)), w = (int)(g % (n_win + 1));
  int lo = ptr[u], hi = ptr[u + 1];
  if (w == 0) { split[g] = lo; return; }
  if (w == n_win) { split[g] = hi; return; }
  const int bound = w * win;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (idx[mid] < bound) lo = mid + 1; else hi = mid;
  }
  split[g] = lo;
}

// work[c] = sum over users u of column c of len_u  (the gathered-entry count of SURVEY 8(d))
__global__ void col_work_kernel(const int* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                                const int2* __restrict__ csc_ent, const int* __restrict__ csr_ptr, int n_cols,
                                unsigned long long* work) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_cols) return;
  unsigned long long s = 0;
  for (int q = csc_ptr[warp] + lane; q < csc_ptr[warp + 1]; q += 32) {
    const int u = csc_idx ? csc_idx[q] : csc_ent[q].x;
    s += (unsigned long long)(csr_ptr[u + 1] - csr_ptr[u]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) work[warp] = s;
}

}  // namespace sim
}  // namespace b200

using namespace b200;
using namespace b200::sim;

struct b200_sim_s {
  int n_rows = 0, n_cols = 0;
  long long nnz = 0;
  int kind = 0, K = 0, normalize = 0;
  float shrink = 0.f, asym_alpha = 0.5f, ta = 1.f, tb = 1.f;
  int formula = F_PROD;
  bool binary = false, signed_data = false;
  int n_win = 1, win = 0, cap = 2048;
  size_t smem_bytes = 0;
  int n_sm = 0;
  DevBuf<int> csr_ptr, csr_idx, csc_ptr, csc_idx, split;
  DevBuf<int2> csr_ent, csc_ent;
  DevBuf<float> A, B;
  DevBuf<unsigned long long> work;
  std::vector<unsigned long long> h_work;
  float B_min = 0.f, B_max = 0.f;
  DevBuf<int> counter, order;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
};

namespace {

constexpr int GRID1D = 148 * 8;

void build(b200_sim_s* h, const int32_t* h_indptr, const int32_t* h_indices, const float* h_data,
           const float* h_row_weights, cudaStream_t st) {
  const int n_rows = h->n_rows, n_cols = h->n_cols;
  const long long nnz = h->nnz;
  h->n_sm = sm_count();
  B200_CUDA(cudaEventCreate(&h->ev0));
  B200_CUDA(cudaEventCreate(&h->ev1));

  h->csr_ptr.alloc((size_t)n_rows + 1);
  h->csr_idx.alloc((size_t)std::max<long long>(nnz, 1));
  DevBuf<float> data((size_t)std::max<long long>(nnz, 1));
  B200_CUDA(cudaMemcpyAsync(h->csr_ptr.get(), h_indptr, sizeof(int) * ((size_t)n_rows + 1), cudaMemcpyHostToDevice, st));
  if (nnz) {
    B200_CUDA(cudaMemcpyAsync(h->csr_idx.get(), h_indices, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(data.get(), h_data, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  }
  DevBuf<float> row_w;
  if (h_row_weights) {
    row_w.alloc((size_t)n_rows);
    B200_CUDA(cudaMemcpyAsync(row_w.get(), h_row_weights, sizeof(float) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
  }

  // ---- per-kind data transform (pyx:160-165)
  const bool set_kind = h->kind == B200_SIM_JACCARD || h->kind == B200_SIM_DICE || h->kind == B200_SIM_TVERSKY;
  DevBuf<double> colsum((size_t)n_cols), colsq((size_t)n_cols);
  DevBuf<int> colcnt((size_t)n_cols);
  B200_CUDA(cudaMemsetAsync(colsum.get(), 0, sizeof(double) * (size_t)n_cols, st));
  B200_CUDA(cudaMemsetAsync(colsq.get(), 0, sizeof(double) * (size_t)n_cols, st));
  B200_CUDA(cudaMemsetAsync(colcnt.get(), 0, sizeof(int) * (size_t)n_cols, st));
  if (nnz) {
    if (set_kind) {
      fill_ones_kernel<<<GRID1D, 256, 0, st>>>(data.get(), nnz); count_launch();
    } else if (h->kind == B200_SIM_ADJUSTED) {
      row_center_kernel<<<div_up((long long)n_rows * 32, 256), 256, 0, st>>>(h->csr_ptr.get(), data.get(), n_rows); count_launch();
    } else if (h->kind == B200_SIM_PEARSON) {
      col_accum_kernel<<<GRID1D, 256, 0, st>>>(h->csr_idx.get(), data.get(), nnz, colsum.get(), nullptr, colcnt.get()); count_launch();
      col_center_kernel<<<GRID1D, 256, 0, st>>>(h->csr_idx.get(), data.get(), nnz, colsum.get(), colcnt.get()); count_launch();
      B200_CUDA(cudaMemsetAsync(colcnt.get(), 0, sizeof(int) * (size_t)n_cols, st));
    }
  }
  // ---- flags: binary path (every stored value == 1, no row weights), signed data
  DevBuf<int> flags(1);
  B200_CUDA(cudaMemsetAsync(flags.get(), 0, sizeof(int), st));
  if (nnz) { flags_kernel<<<GRID1D, 256, 0, st>>>(data.get(), nnz, flags.get()); count_launch(); }
  int hflags = 0;
  B200_CUDA(cudaMemcpyAsync(&hflags, flags.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
  // ---- column sums of squares (before the row weights, pyx:169-194) and counts
  if (nnz) { col_accum_kernel<<<GRID1D, 256, 0, st>>>(h->csr_idx.get(), data.get(), nnz, nullptr, colsq.get(), colcnt.get()); count_launch(); }
  h->A.alloc((size_t)n_cols);
  h->B.alloc((size_t)n_cols);
  const int norm_mode = set_kind ? 0 : (h->kind == B200_SIM_ASYMMETRIC ? 2 : 1);
  norms_kernel<<<div_up(n_cols, 256), 256, 0, st>>>(colsq.get(), n_cols, norm_mode, h->asym_alpha, h->A.get(), h->B.get()); count_launch();

  // ---- CSC: exclusive scan of the column counts, stable sort of (column, position) pairs
  h->csc_ptr.alloc((size_t)n_cols + 1);
  B200_CUDA(cudaMemsetAsync(h->csc_ptr.get(), 0, sizeof(int) * ((size_t)n_cols + 1), st));
  {
    size_t tmp_bytes = 0;
    B200_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, colcnt.get(), h->csc_ptr.get() + 1, n_cols, st));
    DevBuf<unsigned char> tmp(tmp_bytes + 16);
    B200_CUDA(cub::DeviceScan::InclusiveSum(tmp.get(), tmp_bytes, colcnt.get(), h->csc_ptr.get() + 1, n_cols, st));
    count_launch(2);
    B200_CUDA(cudaStreamSynchronize(st));
  }
  B200_CUDA(cudaStreamSynchronize(st));
  h->signed_data = (hflags & 2) != 0;
  h->binary = ((hflags & 1) == 0) && !h_row_weights;

  DevBuf<int> rowid((size_t)std::max<long long>(nnz, 1)), iota((size_t)std::max<long long>(nnz, 1));
  DevBuf<int> keys_out((size_t)std::max<long long>(nnz, 1)), perm((size_t)std::max<long long>(nnz, 1));
  if (nnz) {
    rowid_iota_kernel<<<div_up((long long)n_rows * 32, 256), 256, 0, st>>>(h->csr_ptr.get(), n_rows, rowid.get(), iota.get()); count_launch();
    int end_bit = 1;
    while ((1ll << end_bit) < (long long)n_cols) ++end_bit;
    size_t tmp_bytes = 0;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->csr_idx.get(), keys_out.get(), iota.get(), perm.get(),
                                              (int)nnz, 0, end_bit, st));
    DevBuf<unsigned char> tmp(tmp_bytes + 16);
    B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, h->csr_idx.get(), keys_out.get(), iota.get(), perm.get(),
                                              (int)nnz, 0, end_bit, st));
    count_launch(4);
    if (h->binary) {
      h->csc_idx.alloc((size_t)nnz);
      build_csc_kernel<<<GRID1D, 256, 0, st>>>(perm.get(), rowid.get(), data.get(), nullptr, nnz, nullptr, h->csc_idx.get());
    } else {
      h->csc_ent.alloc((size_t)nnz);
      h->csr_ent.alloc((size_t)nnz);
      build_csc_kernel<<<GRID1D, 256, 0, st>>>(perm.get(), rowid.get(), data.get(), row_w.get(), nnz, h->csc_ent.get(), nullptr);
      build_csr_ent_kernel<<<GRID1D, 256, 0, st>>>(h->csr_idx.get(), data.get(), nnz, h->csr_ent.get()); count_launch();
    }
    count_launch();
    B200_CUDA(cudaStreamSynchronize(st));
  }

  // ---- formula and norm extremes (over columns that hold entries)
  if (set_kind) {
    h->formula = h->kind == B200_SIM_JACCARD ? F_JACCARD : (h->kind == B200_SIM_DICE ? F_DICE : F_TVERSKY);
  } else {
    h->formula = h->normalize ? F_PROD : F_NONORM;
  }
  {
    std::vector<float> hB((size_t)n_cols);
    std::vector<int> hcnt((size_t)n_cols);
    B200_CUDA(cudaMemcpy(hB.data(), h->B.get(), sizeof(float) * (size_t)n_cols, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(hcnt.data(), colcnt.get(), sizeof(int) * (size_t)n_cols, cudaMemcpyDeviceToHost));
    float mn = 0.f, mx = 0.f;
    bool any = false;
    for (int j = 0; j < n_cols; ++j) {
      if (hcnt[j] == 0) continue;
      if (!any) { mn = mx = hB[j]; any = true; }
      mn = std::min(mn, hB[j]);
      mx = std::max(mx, hB[j]);
    }
    h->B_min = mn;
    h->B_max = mx;
  }

  // ---- window geometry: the accumulator covers `win` neighbour columns; n_win passes per target column
  int dev = 0, max_smem = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int cap = 2048;
  while (cap < 4 * h->K) cap <<= 1;
  B200_REQUIRE(cap <= 8 * THREADS, "topK=%d too large for the top-K kernel (max %d); use the dense path", h->K, 2 * THREADS);
  h->cap = cap;
  const size_t fixed = (size_t)cap * 8 + (size_t)COLCHUNK * 12 + sizeof(Shared) + 1024;
  const long long max_cells = ((long long)max_smem - (long long)fixed) / 4;
  B200_REQUIRE(max_cells >= 4096, "not enough shared memory (%d bytes) for the similarity kernel", max_smem);
  int n_win = (int)((n_cols + max_cells - 1) / max_cells);
  if (n_win < 1) n_win = 1;
  int win = (n_cols + n_win - 1) / n_win;
  win = (win + 3) & ~3;
  if (win < 4) win = 4;
  h->n_win = n_win;
  h->win = win;
  h->smem_bytes = (size_t)win * 4 + (size_t)cap * 8 + (size_t)COLCHUNK * 12;
  B200_CUDA(cudaFuncSetAttribute(sim_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  B200_CUDA(cudaFuncSetAttribute(sim_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  if (n_win > 1) {
    h->split.alloc((size_t)n_rows * (n_win + 1));
    const long long total = (long long)n_rows * (n_win + 1);
    split_kernel<<<div_up(total, 256), 256, 0, st>>>(h->csr_ptr.get(), h->csr_idx.get(), n_rows, n_win, win, h->split.get()); count_launch();
  }
  // ---- per-column work (for LPT ordering and the bytes model)
  h->work.alloc((size_t)n_cols);
  col_work_kernel<<<div_up((long long)n_cols * 32, 256), 256, 0, st>>>(h->csc_ptr.get(), h->binary ? h->csc_idx.get() : nullptr,
                                                                      h->binary ? nullptr : h->csc_ent.get(), h->csr_ptr.get(), n_cols, h->work.get());
  count_launch();
  h->h_work.resize((size_t)n_cols);
  B200_CUDA(cudaMemcpyAsync(h->h_work.data(), h->work.get(), sizeof(unsigned long long) * (size_t)n_cols, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (!h->binary) h->csr_idx.release();  // the AoS copy carries the indices
  h->counter.alloc(1);
  h->order.alloc((size_t)n_cols);
}

}  // namespace

extern "C" {

int b200_sim_create(b200_sim_t* out, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* h_indptr,
                    const int32_t* h_indices, const float* h_data, int kind, int topK, float shrink, int normalize,
                    float asymmetric_alpha, float tversky_alpha, float tversky_beta, const float* h_row_weights,
                    void* stream) {
  if (out) *out = nullptr;
  b200_sim_s* h = nullptr;
  int rc = guarded([&] {
    B200_REQUIRE(out != nullptr, "b200_sim_create: out is NULL");
    B200_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0, "b200_sim_create: bad shape %lld x %lld nnz %lld",
                 (long long)n_rows, (long long)n_cols, (long long)nnz);
    B200_REQUIRE(n_rows < (1ll << 31) - 1 && n_cols < (1ll << 31) - 1 && nnz < (1ll << 31) - 1,
                 "b200_sim_create: int32 index range exceeded");
    B200_REQUIRE(kind >= B200_SIM_COSINE && kind <= B200_SIM_TVERSKY, "b200_sim_create: unknown similarity kind %d", kind);
    B200_REQUIRE(topK >= 1, "b200_sim_create: topK must be >= 1 (dense output goes through b200_sim_compute_dense)");
    B200_REQUIRE(h_indptr && (nnz == 0 || (h_indices && h_data)), "b200_sim_create: NULL input array");
    h = new b200_sim_s();
    h->n_rows = (int)n_rows;
    h->n_cols = (int)n_cols;
    h->nnz = nnz;
    h->kind = kind;
    h->K = (int)std::min<int64_t>(topK, n_cols);
    const bool set_kind = kind == B200_SIM_JACCARD || kind == B200_SIM_DICE || kind == B200_SIM_TVERSKY;
    h->normalize = set_kind ? 0 : (normalize != 0);
    h->shrink = shrink;
    h->asym_alpha = asymmetric_alpha;
    h->ta = tversky_alpha;
    h->tb = tversky_beta;
    build(h, h_indptr, h_indices, h_data, h_row_weights, (cudaStream_t)stream);
    *out = h;
  });
  if (rc != B200_OK && h) delete h;
  return rc;
}

int b200_sim_destroy(b200_sim_t h) {
  if (!h) return B200_OK;
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return B200_OK;
}

int b200_sim_info(b200_sim_t h, int* K, int* n_windows, int* window_cells, int* binary_path, int* signed_data) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_sim_info: NULL handle");
    if (K) *K = h->K;
    if (n_windows) *n_windows = h->n_win;
    if (window_cells) *window_cells = h->win;
    if (binary_path) *binary_path = h->binary ? 1 : 0;
    if (signed_data) *signed_data = h->signed_data ? 1 : 0;
  });
}

int b200_sim_compute_device(b200_sim_t h, int start_col, int end_col, int32_t* d_idx, float* d_val, int32_t* d_cnt,
                            void* stream) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_sim_compute: NULL handle");
    B200_REQUIRE(0 <= start_col && start_col <= end_col && end_col <= h->n_cols, "b200_sim_compute: bad column range [%d,%d)",
                 start_col, end_col);
    const int n_range = end_col - start_col;
    if (n_range == 0) return;
    B200_REQUIRE(d_idx && d_val && d_cnt, "b200_sim_compute: NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    // longest-processing-time-first order of the local columns
    std::vector<int> order((size_t)n_range);
    for (int i = 0; i < n_range; ++i) order[(size_t)i] = i;
    const unsigned long long* w = h->h_work.data() + start_col;
    std::stable_sort(order.begin(), order.end(), [w](int a, int b) { return w[a] > w[b]; });
    B200_CUDA(cudaMemcpyAsync(h->order.get(), order.data(), sizeof(int) * (size_t)n_range, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemsetAsync(h->counter.get(), 0, sizeof(int), st));
    KParams p;
    p.n_cols = h->n_cols; p.K = h->K; p.n_win = h->n_win; p.win = h->win; p.cap = h->cap;
    p.formula = h->formula;
    p.se = h->shrink + 1e-6f;
    p.shrink_div = h->shrink != 0.f ? h->shrink : 1.f;
    p.ta = h->ta; p.tb = h->tb;
    p.csr_ptr = h->csr_ptr.get(); p.csr_ent = h->csr_ent.get(); p.csr_idx = h->csr_idx.get();
    p.split = h->split.get();
    p.csc_ptr = h->csc_ptr.get(); p.csc_ent = h->csc_ent.get(); p.csc_idx = h->csc_idx.get();
    p.A = h->A.get(); p.B = h->B.get(); p.B_min = h->B_min; p.B_max = h->B_max;
    p.col_begin = start_col; p.n_range = n_range;
    p.order = h->order.get();
    p.counter = h->counter.get();
    p.out_idx = d_idx; p.out_val = d_val; p.out_cnt = d_cnt;
    p.signed_data = h->signed_data ? 1 : 0;
    const int grid = std::min(n_range, h->n_sm);
    B200_CUDA(cudaStreamSynchronize(st));  // `order` is a host vector about to go out of scope
    B200_CUDA(cudaEventRecord(h->ev0, st));
    if (h->binary)
      sim_topk_kernel<true><<<grid, THREADS, h->smem_bytes, st>>>(p);
    else
      sim_topk_kernel<false><<<grid, THREADS, h->smem_bytes, st>>>(p);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(h->ev1, st));
    h->timed = true;
    count_launch();
  });
}

int b200_sim_compute(b200_sim_t h, int start_col, int end_col, int32_t* h_idx, float* h_val, int32_t* h_cnt) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_sim_compute: NULL handle");
    B200_REQUIRE(0 <= start_col && start_col <= end_col && end_col <= h->n_cols, "b200_sim_compute: bad column range [%d,%d)",
                 start_col, end_col);
    const size_t n_range = (size_t)(end_col - start_col);
    if (n_range == 0) return;
    DevBuf<int> d_idx(n_range * h->K), d_cnt(n_range);
    DevBuf<float> d_val(n_range * h->K);
    int rc = b200_sim_compute_device(h, start_col, end_col, d_idx.get(), d_val.get(), d_cnt.get(), nullptr);
    if (rc != B200_OK) throw CudaFail{rc};
    B200_CUDA(cudaMemcpy(h_idx, d_idx.get(), sizeof(int) * n_range * h->K, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(h_val, d_val.get(), sizeof(float) * n_range * h->K, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(h_cnt, d_cnt.get(), sizeof(int) * n_range, cudaMemcpyDeviceToHost));
  });
}

int b200_sim_last_kernel_ms(b200_sim_t h, float* ms) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr && ms != nullptr, "b200_sim_last_kernel_ms: NULL argument");
    B200_REQUIRE(h->timed, "b200_sim_last_kernel_ms: no kernel launched yet");
    B200_CUDA(cudaEventSynchronize(h->ev1));
    B200_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  });
}

int b200_sim_work(b200_sim_t h, int start_col, int end_col, int64_t* gathered_entries) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr && gathered_entries != nullptr, "b200_sim_work: NULL argument");
    B200_REQUIRE(0 <= start_col && start_col <= end_col && end_col <= h->n_cols, "b200_sim_work: bad column range");
    unsigned long long s = 0;
    for (int c = start_col; c < end_col; ++c) s += h->h_work[(size_t)c];
    *gathered_entries = (int64_t)s;
  });
}

}  // extern "C"
