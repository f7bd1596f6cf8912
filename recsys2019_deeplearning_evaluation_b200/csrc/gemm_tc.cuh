// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, sm_100a): C = alpha * op(A) * op(B) + beta * C.
//
// Each operand element x is split into two TF32 numbers, x = hi + lo (hi = rna_tf32(x), lo = rna_tf32(x - hi)), and every
// 128 x 128 x 8 step issues three tcgen05.mma.kind::tf32 instructions, hi*hi + hi*lo + lo*hi, into one fp32 accumulator
// tile in tensor memory ("3xTF32": the dropped lo*lo term is ~2^-22 relative, so the result is fp32-accurate, which the
// 1e-4 parity bar of the EASE_R inverse needs; a single TF32 pass is not).
//
// One CTA (256 threads) per 128 x 128 output tile.  K is consumed in chunks of 32: all threads load the two operand
// chunks from global memory, split them and store hi / lo tiles in shared memory in the canonical no-swizzle K-major
// UMMA layout (8-row x 16-byte core matrices; LBO = 128 B between K-adjacent cores, SBO = 1024 B between 8-row groups),
// two stages deep.  One elected thread issues the 12 MMAs of a chunk and commits them to the stage's mbarrier, which
// gates the reuse of that stage's shared memory; loads of chunk k+1 overlap the MMAs of chunk k.  The accumulator
// (128 lanes x 128 columns of TMEM) is read back with tcgen05.ld (32x32b.x32) by all 8 warps for the alpha/beta epilogue.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;   // BK in fp32/tf32 elements (128 bytes per row)
constexpr int TILE_BYTES = BM * BK * 4;      // 16 KB
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * 4 * TILE_BYTES + 64;  // {A_hi, A_lo, B_hi, B_lo} per stage + barriers
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);                     // start address, 16-byte units
  d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 16;                // leading byte offset: next core matrix along K
  d |= (uint64_t)((((BK / 4) * 128u) >> 4) & 0x3FFFu) << 32;   // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                                      // descriptor version 1 (sm_100)
  return d;                                                    // base offset 0, layout type 0 = no swizzle
}

// kind::tf32, D = fp32, A and B K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t tile_offset(int row, int k) {
  return (uint32_t)((((row >> 3) * (BK / 4) + (k >> 2)) << 7) + ((row & 7) << 4) + ((k & 3) << 2));
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}

// tile(row, k) = src[(row0 + row) * ld + k0 + k]   (k contiguous in memory)
__device__ __forceinline__ void load_tile_kcontig(const float* __restrict__ src, long long ld, int row0, int k0,
                                                  unsigned char* hi_tile, unsigned char* lo_tile, int tid) {
#pragma unroll
  for (int e = 0; e < (BM * BK / 4) / THREADS; ++e) {
    const int idx = tid + e * THREADS;
    const int r7 = idx & 7, k4 = (idx >> 3) & 7, m8 = idx >> 6;
    const int row = m8 * 8 + r7;
    const float4 v = *reinterpret_cast<const float4*>(src + (long long)(row0 + row) * ld + k0 + k4 * 4);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    const uint32_t off = tile_offset(row, k4 * 4);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

// tile(row, k) = src[(k0 + k) * ld + row0 + row]   (row contiguous in memory)
__device__ __forceinline__ void load_tile_rowcontig(const float* __restrict__ src, long long ld, int row0, int k0,
                                                    unsigned char* hi_tile, unsigned char* lo_tile, int tid) {
#pragma unroll
  for (int e = 0; e < (BM * BK / 4) / THREADS; ++e) {
    const int idx = tid + e * THREADS;
    const int m4 = idx & 31, k = idx >> 5;  // 32 float4 along the rows, 32 values of k
    const float4 v = *reinterpret_cast<const float4*>(src + (long long)(k0 + k) * ld + row0 + m4 * 4);
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float h, l;
      split_tf32(x[c], h, l);
      const uint32_t off = tile_offset(m4 * 4 + c, k);
      *reinterpret_cast<float*>(hi_tile + off) = h;
      *reinterpret_cast<float*>(lo_tile + off) = l;
    }
  }
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}

// TA: op(A)(m, k) = A[k * lda + m]; TB: op(B)(k, n) = B[n * ldb + k].  TRI: op(A) = L^T, op(B) = L with L lower
// triangular, so only k >= max(m0, n0) contributes.  M, N multiples of 128, K a multiple of 32.  blockIdx.z = batch.
template <bool TA, bool TB, bool TRI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                                                             long long strideA, const float* __restrict__ B, int ldb, long long strideB,
                                                             float beta, float* C, int ldc, long long strideC) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem_base;
  A += (long long)blockIdx.z * strideA;
  B += (long long)blockIdx.z * strideB;
  C += (long long)blockIdx.z * strideC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * 4 * TILE_BYTES);
  const uint32_t bar0 = smem_u32(bars);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS) : "memory");
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

  const int k_begin = TRI ? (max(m0, n0) / BK) * BK : 0;
  const int nk = (K - k_begin) / BK;
  for (int kb = 0; kb < nk; ++kb) {
    const int s = kb & 1;
    const int k0 = k_begin + kb * BK;
    unsigned char* a_hi = tiles + (s * 4 + 0) * TILE_BYTES;
    unsigned char* a_lo = tiles + (s * 4 + 1) * TILE_BYTES;
    unsigned char* b_hi = tiles + (s * 4 + 2) * TILE_BYTES;
    unsigned char* b_lo = tiles + (s * 4 + 3) * TILE_BYTES;
    if (kb >= STAGES) mbar_wait(bar0 + 8u * s, (uint32_t)(((kb >> 1) - 1) & 1));  // MMAs that read this stage are done
    if (TA) load_tile_rowcontig(A, lda, m0, k0, a_hi, a_lo, tid); else load_tile_kcontig(A, lda, m0, k0, a_hi, a_lo, tid);
    if (TB) load_tile_kcontig(B, ldb, n0, k0, b_hi, b_lo, tid); else load_tile_rowcontig(B, ldb, n0, k0, b_hi, b_lo, tid);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {  // one MMA consumes K = 8 tf32 = two core matrices = 256 bytes
        const uint32_t o = ks * 256u;
        umma_tf32(tmem, make_smem_desc(ah + o), make_smem_desc(bh + o), (kb > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(tmem, make_smem_desc(ah + o), make_smem_desc(bl + o), 1u);
        umma_tf32(tmem, make_smem_desc(al + o), make_smem_desc(bh + o), 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8u * s) : "memory");
    }
  }
  if (nk > 0) {
    const int last = nk - 1;
    mbar_wait(bar0 + 8u * (last & 1), (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (its row of the tile), columns 64*(w/4) .. +63
  const int row = (warp & 3) * 32 + lane;
  const int col0 = (warp >> 2) * 64;
  float* crow = C + (long long)(m0 + row) * ldc + n0 + col0;
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
      for (int c = 0; c < 32; ++c) v[c] = 0u;
    }
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float4 o;
      o.x = alpha * __uint_as_float(v[c4 * 4 + 0]);
      o.y = alpha * __uint_as_float(v[c4 * 4 + 1]);
      o.z = alpha * __uint_as_float(v[c4 * 4 + 2]);
      o.w = alpha * __uint_as_float(v[c4 * 4 + 3]);
      float4* dst = reinterpret_cast<float4*>(crow + half * 32 + c4 * 4);
      if (beta != 0.f) {
        const float4 old = *dst;
        o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
      }
      *dst = o;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace tc
}  // namespace b200
