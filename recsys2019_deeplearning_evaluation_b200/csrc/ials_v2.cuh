// K4 on the tensor cores, second generation (sm_100a): implicit-ALS half epochs for n_factors <= 256.
//
// Replaces MatrixFactorization/IALSRecommender.py:137-201 like ials.cu (per warm row:  A = Y_p^T diag(c - 1) Y_p,
// B = Y^T Y + A + reg I,  x = B^-1 Y_p^T c), with the dense contraction on tcgen05 as BASELINE.json's north_star asks.
//
// Why a second generation.  ials.cu accumulates the Gram in fp64 registers and factors a packed fp64 matrix that no longer
// fits shared memory above 208 factors: one C4 epoch at 256 factors takes 11.7 s.  The first tensor-core kernel (one
// 128 x 128 accumulator, fp64 Cholesky, round 1) only covered 128 factors and was 3x slower than the fp64 kernel it was
// meant to replace -- scalar staging and an unblocked factorisation (profiles/r02_hotpath3_c4_timings.txt).  Here:
//   * Gram on the tensor core, 3xTF32 (fp32-accurate).  The profile's factor rows are scaled by sqrt(c - 1) while they are
//     staged, so ONE operand serves both sides of  A = (sqrt(c-1) Y_p)^T (sqrt(c-1) Y_p): per 32 profile entries a
//     256 x 32 hi tile and a lo tile (K-major UMMA layout, 128-bit shared stores, two stages).  Only the lower triangle is
//     formed: accumulator D0 = rows 0..127 x columns 0..127 (M = 128, N = 128), D1 = rows 128..255 x columns 0..255
//     (M = 128, N = 256); 384 of the 512 TMEM columns.  One elected thread issues the 24 MMAs of a step and commits them to
//     the stage's mbarrier; the staging of the next step overlaps them.
//   * B in fp32, lower triangle, rows padded to 16 bytes (133 KB at 256 factors: it takes over the staging buffers), then a
//     right-looking blocked Cholesky with 32-wide panels: the diagonal block in the registers of one warp (shuffles), the
//     panel by one thread per row, the trailing update by warps that own 16 x 32 blocks of it so that their 128-bit
//     operand loads are broadcasts; the inverses of the diagonal blocks are kept for the solves.
//   * The solution is refined against the EXACT operator in fp64,  r = b - (Y^T Y x + Y_p^T((c-1) .* (Y_p x)) + reg x),
//     matrix-free, twice: only the fp32 factor is approximate, so each step contracts the error by ~cond * 1e-7.  A row whose
//     residual does not contract (ill-conditioned system) raises `redo`; the host then repeats the half epoch with ials.cu.
// Y^T Y comes from gram_kernel (ials.cu, fp64).  One CTA (512 threads) per SM, rows strided over the grid.
#pragma once
#include "gemm_tc.cuh"

namespace b200 {
namespace ials2 {

constexpr int T = 512;                   // 16 warps: the phases between the barriers are latency-bound, more warps hide more of it
constexpr int NW = T / 32;
constexpr int KC = 32;                   // profile entries per step
constexpr int STAGES = 2;
constexpr int TILE = 256 * KC * 4;       // one 256 x 32 fp32 tile: 32 KB
constexpr int NB = 32;                   // Cholesky panel width
constexpr int NBP = NB + 1;              // row stride of the stored diagonal-block inverses (no shared bank across rows)
constexpr uint32_t TMEM_COLS = 512;
// kind::tf32, D = fp32, A and B K-major, M = 128, N = 128 / 256 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
constexpr uint32_t IDESC_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_N256 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// lower triangle, row r starts at roff(r) floats: rows are padded to multiples of 4 floats so that every 4-aligned column
// offset of every row is 16-byte aligned
__device__ __host__ __forceinline__ int roff(int r) { const int q = r >> 2, m = r & 3; return 4 * (q + 1) * (2 * q + m); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
}

struct Smem {
  float* L;        // padded lower triangle (aliases the staging tiles)
  unsigned char* tiles;
  uint64_t* bars;
  double *b0, *xs, *rs, *part;  // rhs, solution, residual / correction, [NW][f] partial sums
  float* dinv;     // [f / 32][32][33] inverses of the diagonal blocks of L (lower triangular, row-major, padded rows)
  float* sw;       // [KC] sqrt(c - 1) of the step's entries
  double* cw;      // [KC] c
  int* rk;         // [KC] factor-row indices (-1 past the profile)
  float* rdiag;    // [NB] reciprocals of the current diagonal block's diagonal
};

// ---- L z = v (forward) or L^T z = v (backward) in place on v (fp64), blocked by NB with the stored diagonal-block inverses
__device__ __forceinline__ void solve_forward(const Smem& S, double* v, int f) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < f; k0 += NB) {
    const int nb = min(NB, f - k0);
    if (warp == 0) {  // z_blk = inv(L11) v_blk
      const float* inv = S.dinv + (k0 / NB) * NB * NBP;
      double z = 0.0;
      if (lane < nb)
        for (int c = 0; c <= lane; ++c) z += (double)inv[lane * NBP + c] * v[k0 + c];
      __syncwarp();
      if (lane < nb) v[k0 + lane] = z;
    }
    __syncthreads();
    const int r = k0 + nb + tid;  // rows below the block: v[r] -= L[r, blk] z_blk
    if (r < f) {
      const float* Lr = S.L + roff(r) + k0;
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s += (double)Lr[c] * v[k0 + c];
      v[r] -= s;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void solve_backward(const Smem& S, double* v, int f) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (f + NB - 1) / NB;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int k0 = kb * NB, nb = min(NB, f - k0);
    if (warp == 0) {  // x_blk = inv(L11)^T v_blk
      const float* inv = S.dinv + kb * NB * NBP;
      double z = 0.0;
      if (lane < nb)
        for (int c = lane; c < nb; ++c) z += (double)inv[c * NBP + lane] * v[k0 + c];
      __syncwarp();
      if (lane < nb) v[k0 + lane] = z;
    }
    __syncthreads();
    if (tid < k0) {  // columns left of the block: v[c] -= sum_r L[k0 + r, c] x[k0 + r]
      double s = 0.0;
      for (int r = 0; r < nb; ++r) s += (double)S.L[roff(k0 + r) + tid] * v[k0 + r];
      v[tid] -= s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(T, 1) ials_rows_v2_kernel(const int* __restrict__ rows, int n_solve, const int* __restrict__ ptr,
                                                            const int* __restrict__ idx, const float* __restrict__ conf,
                                                            const double* __restrict__ Y, const double* __restrict__ YtY, int f,
                                                            double reg, double* X, int* info, int* redo, int n_refine) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem_base;
  __shared__ double s_norm[3];  // |r|^2 before the last / the first correction, |b|^2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem S;
  {
    const size_t lbytes = (size_t)roff(f) * 4;
    size_t o = std::max<size_t>(lbytes, (size_t)STAGES * 2 * TILE);
    o = (o + 127) & ~(size_t)127;
    S.L = reinterpret_cast<float*>(smem);
    S.tiles = smem;
    S.bars = reinterpret_cast<uint64_t*>(smem + o); o += 64;
    S.b0 = reinterpret_cast<double*>(smem + o); o += sizeof(double) * f;
    S.xs = reinterpret_cast<double*>(smem + o); o += sizeof(double) * f;
    S.rs = reinterpret_cast<double*>(smem + o); o += sizeof(double) * f;
    S.part = reinterpret_cast<double*>(smem + o); o += sizeof(double) * NW * f;
    S.cw = reinterpret_cast<double*>(smem + o); o += sizeof(double) * KC;
    S.dinv = reinterpret_cast<float*>(smem + o); o += sizeof(float) * ((f + NB - 1) / NB) * NB * NBP;
    S.sw = reinterpret_cast<float*>(smem + o); o += sizeof(float) * KC;
    S.rk = reinterpret_cast<int*>(smem + o); o += sizeof(int) * KC;
    S.rdiag = reinterpret_cast<float*>(smem + o);
  }
  const uint32_t bar0 = tc::smem_u32(S.bars);
  const bool two = f > 128;  // rows 128.. exist: the second accumulator is in use

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s) : "memory");
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
    const int nk = (p1 - p0 + KC - 1) / KC;

    // ================= Gram on the tensor core
    double my_rhs = 0.0;  // threads m and m + 256 each hold half of b0[m] = (Y_p^T c)[m]  (IALSRecommender.py:201)
    for (int kb = 0; kb < nk; ++kb, ++gk) {
      const int s = (int)(gk & 1u);
      unsigned char* t_hi = S.tiles + (size_t)(s * 2 + 0) * TILE;
      unsigned char* t_lo = S.tiles + (size_t)(s * 2 + 1) * TILE;
      __syncthreads();  // rk / sw / cw of the previous step are consumed
      if (tid < KC) {
        const int q = p0 + kb * KC + tid;
        const float c = q < p1 ? conf[q] : 1.f;
        S.rk[tid] = q < p1 ? idx[q] : -1;
        S.sw[tid] = sqrtf(fmaxf(c - 1.f, 0.f));
        S.cw[tid] = q < p1 ? (double)c : 0.0;
      }
      if (gk >= (unsigned)STAGES) tc::mbar_wait(bar0 + 8u * s, (uint32_t)(((gk >> 1) - 1u) & 1u));  // the MMAs that read this stage are done
      __syncthreads();
      // tile(m, k) = sqrt(c_k - 1) * Y[rk[k]][m]; thread = (factor m, half of the step's entries), four entries per 128-bit
      // store; zero past f / the profile
      {
        const int m = tid & 255, kh = tid >> 8;
#pragma unroll
        for (int kq = kh * (KC / 8); kq < (kh + 1) * (KC / 8); ++kq) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = kq * 4 + e;
            const int r = S.rk[k];
            const double y = (r >= 0 && m < f) ? Y[(size_t)r * f + m] : 0.0;
            my_rhs += S.cw[k] * y;
            tc::split_tf32((float)y * S.sw[k], h[e], l[e]);
          }
          const uint32_t off = tc::tile_offset(m, kq * 4);
          *reinterpret_cast<float4*>(t_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(t_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t th = tc::smem_u32(t_hi), tl = tc::smem_u32(t_lo);
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint32_t o = ks * 256u;
          const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
          // D0: rows 0..127 x columns 0..127
          umma(tmem, tc::make_smem_desc(th + o), tc::make_smem_desc(th + o), IDESC_N128, acc);
          umma(tmem, tc::make_smem_desc(th + o), tc::make_smem_desc(tl + o), IDESC_N128, 1u);
          umma(tmem, tc::make_smem_desc(tl + o), tc::make_smem_desc(th + o), IDESC_N128, 1u);
          if (two) {  // D1: rows 128..255 (operand A starts 128 rows = 16 KB into the tile) x columns 0..255
            umma(tmem + 128u, tc::make_smem_desc(th + 16384u + o), tc::make_smem_desc(th + o), IDESC_N256, acc);
            umma(tmem + 128u, tc::make_smem_desc(th + 16384u + o), tc::make_smem_desc(tl + o), IDESC_N256, 1u);
            umma(tmem + 128u, tc::make_smem_desc(tl + 16384u + o), tc::make_smem_desc(th + o), IDESC_N256, 1u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8u * s) : "memory");
      }
    }
    if (nk > 0) {
      // every outstanding step: the tiles are about to be overwritten by L
      if (nk > 1) { const unsigned prev = gk - 2u; tc::mbar_wait(bar0 + 8u * (prev & 1u), (uint32_t)((prev >> 1) & 1u)); }
      const unsigned last = gk - 1u;
      tc::mbar_wait(bar0 + 8u * (last & 1u), (uint32_t)((last >> 1) & 1u));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // ================= accumulators -> B = A + Y^T Y + reg I, fp32, lower triangle.  Warp w owns TMEM lanes 32 (w % 4) .. + 31
    // (a row of each accumulator per thread) and half of the columns.
    {
      const int g = warp & 3, hcol = warp >> 2;  // hcol in [0, NW / 4)
      for (int acc_i = 0; acc_i < (two ? 2 : 1); ++acc_i) {
        const int r = acc_i * 128 + g * 32 + lane;             // row of B
        const int ncol = acc_i == 0 ? 128 : 256;               // columns this accumulator holds
        const int cbeg = hcol * (ncol / (NW / 4)), cend = cbeg + ncol / (NW / 4);
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          if (c0 > acc_i * 128 + g * 32 + 31) continue;         // above the diagonal for every row of this warp (uniform)
          uint32_t v[32];
          if (nk > 0) {
            tmem_ld32(tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc_i * 128 + c0), v);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0u;
          }
          if (r < f) {
            float* Lr = S.L + roff(r);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int c = c0 + q;
              if (c <= r) Lr[c] = (float)((double)__uint_as_float(v[q]) + YtY[(size_t)c * f + r] + (r == c ? reg : 0.0));  // Y^T Y is symmetric: coalesced over r
            }
          }
        }
      }
    }
    if (tid >= 256 && tid - 256 < f) S.rs[tid - 256] = my_rhs;  // the upper half's share of the right-hand side
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // every warp has read the accumulators: the next row may overwrite them
    if (tid < f) { const double b = my_rhs + S.rs[tid]; S.b0[tid] = b; S.xs[tid] = b; }
    __syncthreads();

    // ================= blocked Cholesky, fp32, in place
    for (int k0 = 0; k0 < f; k0 += NB) {
      const int nb = min(NB, f - k0);
      // ---- diagonal block in the registers of warp 0: lane i holds row i; its inverse for the solves
      if (warp == 0) {
        float a[NB];
        {
          // row k0 + lane, columns k0 .. k0 + 31 (16-byte aligned); entries past the diagonal belong to the next row or to
          // padding and are never used (see below), lanes >= nb read a valid address of the last row
          const float* src = S.L + roff(k0 + min(lane, nb - 1)) + k0;
#pragma unroll
          for (int c4 = 0; c4 < NB / 4; ++c4) {
            const float4 t4 = (c4 * 4 <= lane) ? *reinterpret_cast<const float4*>(src + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            a[c4 * 4 + 0] = t4.x; a[c4 * 4 + 1] = t4.y; a[c4 * 4 + 2] = t4.z; a[c4 * 4 + 3] = t4.w;
          }
        }
        // Lane i holds row i; only the lower triangle (c <= i) of a[] is ever read -- by this lane, or through the shuffles that
        // fetch lane c's a[j] with c > j -- so nothing below is predicated on the lane: the entries above the diagonal hold
        // garbage that feeds only garbage.  One reciprocal square root per step instead of a square root and a division (the
        // factor only has to be fp32-good: the refinement is against the exact operator).
        bool bad = false;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (j < nb) {
            const float ajj = __shfl_sync(0xffffffffu, a[j], j);
            bad = bad || !(ajj > 0.f);
            const float rinv = rsqrtf(fmaxf(ajj, 1e-30f));
            if (lane == 0) S.rdiag[j] = rinv;  // 1 / L11[j][j] for the forward substitutions below
            a[j] = (lane == j) ? ajj * rinv : a[j] * rinv;
#pragma unroll
            for (int c = j + 1; c < NB; ++c) a[c] = fmaf(-a[j], __shfl_sync(0xffffffffu, a[j], c), a[c]);
          }
        }
        if (bad && lane == 0) atomicExch(info, row + 1);
        if (lane < nb) {
#pragma unroll
          for (int c = 0; c < NB; ++c)
            if (c <= lane) S.L[roff(k0 + lane) + k0 + c] = a[c];
        }
      }
      __syncthreads();
      const int R0 = k0 + nb;
      // ---- one forward substitution against L11 for 32 + n_below right-hand sides: the rows of the panel below the block
      // (row r of L21 = A21 L11^-T, threads 32..) and, in warp 0, the unit vectors -- column `lane` of inv(L11), kept for the
      // solves.  The inverse costs the critical path nothing: it runs beside the panel.
      {
        const bool is_inv = tid < 32;
        const int r = R0 + tid - 32;
        if (is_inv || r < f) {
          float* Lr = is_inv ? nullptr : S.L + roff(r) + k0;
          float x[NB];
          if (is_inv) {
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = (c == lane) ? 1.f : 0.f;
          } else {
#pragma unroll
            for (int c4 = 0; c4 < NB / 4; ++c4) {
              const float4 t4 = *reinterpret_cast<const float4*>(Lr + c4 * 4);
              x[c4 * 4 + 0] = t4.x; x[c4 * 4 + 1] = t4.y; x[c4 * 4 + 2] = t4.z; x[c4 * 4 + 3] = t4.w;
            }
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            if (j < nb) {
              const float* Lj = S.L + roff(k0 + j) + k0;  // same address in every thread: broadcast
              float s = x[j];
#pragma unroll
              for (int c = 0; c < NB; ++c)
                if (c < j) s -= x[c] * Lj[c];
              x[j] = s * S.rdiag[j];
            }
          }
          if (is_inv) {
            float* inv = S.dinv + (k0 / NB) * NB * NBP;
#pragma unroll
            for (int i = 0; i < NB; ++i) inv[i * NBP + lane] = (i >= lane && i < nb && lane < nb) ? x[i] : 0.f;  // inv[i][j] = (L11^-1 e_j)[i]
          } else {
#pragma unroll
            for (int c4 = 0; c4 < NB / 4; ++c4)
              *reinterpret_cast<float4*>(Lr + c4 * 4) = make_float4(x[c4 * 4 + 0], x[c4 * 4 + 1], x[c4 * 4 + 2], x[c4 * 4 + 3]);
          }
        }
      }
      __syncthreads();
      if (R0 >= f) break;
      // ---- trailing update A22 -= L21 L21^T (lower part).  A warp owns a 16-row x 32-column block: lane -> 4 x 4 tile
      // (rows 4 (lane >> 3).., columns 4 (lane & 7)..): the 128-bit loads of a row / column panel are shared by 8 / 4 lanes
      {
        const int nt = f - R0;                    // trailing order
        const int nsr = (nt + 15) / 16;           // 16-row strips; strip si meets (si >> 1) + 1 column strips of 32
        // super-tiles in row-major order of the lower triangle, dealt round-robin to the warps: strips 2a and 2a + 1 start at
        // a (a + 1) and (a + 1)^2
        const int n_super = (nsr & 1) ? ((nsr + 1) / 2) * ((nsr + 1) / 2) : (nsr / 2) * (nsr / 2 + 1);
        for (int n = warp; n < n_super; n += NW) {
          {
            int a = (int)sqrtf((float)n);
            while (a * a > n) --a;
            while ((a + 1) * (a + 1) <= n) ++a;
            int si, sj;
            if (n >= a * (a + 1)) { si = 2 * a; sj = n - a * (a + 1); } else { si = 2 * a - 1; sj = n - a * a; }
            const int r0 = R0 + si * 16 + (lane >> 3) * 4, c0 = R0 + sj * 32 + (lane & 7) * 4;
            if (c0 > r0 + 3 || r0 >= f) continue;
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk) {
              float4 av[4], bv[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const int r = min(r0 + x, f - 1), c = min(c0 + x, f - 1);
                av[x] = *reinterpret_cast<const float4*>(S.L + roff(r) + k0 + kk * 4);
                bv[x] = *reinterpret_cast<const float4*>(S.L + roff(c) + k0 + kk * 4);
              }
#pragma unroll
              for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                  acc[x][y] += av[x].x * bv[y].x + av[x].y * bv[y].y + av[x].z * bv[y].z + av[x].w * bv[y].w;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) {
                const int r = r0 + x, c = c0 + y;
                if (r < f && c <= r) S.L[roff(r) + c] -= acc[x][y];
              }
          }
        }
      }
      __syncthreads();
    }

    // ================= solve and refine
    solve_forward(S, S.xs, f);
    solve_backward(S, S.xs, f);
    for (int it = 0; it < n_refine; ++it) {
      // part[w][m] = sum over this warp's profile entries of (c_k - 1) (y_k . x) y_k[m]; lane owns m = lane + 32 t
      double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
      for (int q = p0 + warp; q < p1; q += NW) {  // two entries' loads in flight
        const double* y = Y + (size_t)idx[q] * f;
        double yv[8], dot = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int m = lane + 32 * t;
          yv[t] = m < f ? y[m] : 0.0;
          dot += m < f ? yv[t] * S.xs[m] : 0.0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        const double wd = ((double)conf[q] - 1.0) * dot;
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += wd * yv[t];
      }
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (lane + 32 * t < f) S.part[warp * f + lane + 32 * t] = acc[t];
      __syncthreads();
      if (tid < f) {
        double ax = reg * S.xs[tid];
        for (int w = 0; w < NW; ++w) ax += S.part[w * f + tid];
#pragma unroll 8
        for (int n = 0; n < f; ++n) ax += YtY[(size_t)n * f + tid] * S.xs[n];  // symmetric: coalesced over tid; eight loads in flight
        S.rs[tid] = S.b0[tid] - ax;
      }
      __syncthreads();
      if (warp == 0 && (it == 0 || it == n_refine - 1)) {  // |r|^2 before the first and before the last correction
        double r2 = 0.0, b2 = 0.0;
        for (int m = lane; m < f; m += 32) { r2 += S.rs[m] * S.rs[m]; b2 += S.b0[m] * S.b0[m]; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { r2 += __shfl_xor_sync(0xffffffffu, r2, off); b2 += __shfl_xor_sync(0xffffffffu, b2, off); }
        if (lane == 0) {
          if (it == 0) { s_norm[1] = r2; s_norm[2] = b2; }
          if (it == n_refine - 1) s_norm[0] = r2;
        }
      }
      solve_forward(S, S.rs, f);
      solve_backward(S, S.rs, f);
      if (tid < f) S.xs[tid] += S.rs[tid];
      __syncthreads();
    }
    // the residual must have shrunk at least 1000-fold between the first and the last correction (or sit at the rounding floor
    // of the exact operator already); otherwise the fp32 factor is too far from the exact operator for this row
    if (n_refine >= 2 && tid == 0 && !(s_norm[0] <= 1e-3 * s_norm[1] || s_norm[0] <= 1e-22 * s_norm[2])) atomicExch(redo, 1);
    if (tid < f) X[(size_t)row * f + tid] = S.xs[tid];
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

inline size_t smem_bytes(int f) {
  size_t o = std::max<size_t>((size_t)roff(f) * 4, (size_t)STAGES * 2 * TILE);
  o = (o + 127) & ~(size_t)127;
  o += 64 + sizeof(double) * (3 * (size_t)f + NW * (size_t)f + KC) + sizeof(float) * ((size_t)((f + NB - 1) / NB) * NB * NBP + KC) + sizeof(int) * KC + sizeof(float) * NB;
  return o + 64;
}

}  // namespace ials2
}  // namespace b200
