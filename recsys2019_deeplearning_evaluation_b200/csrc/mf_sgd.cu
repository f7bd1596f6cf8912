// K2: matrix-factorisation SGD epochs (BPR-MF and FunkSVD), sm_100a.
//
// Replaces MatrixFactorization/Cython/MatrixFactorization_Cython_Epoch.pyx:
//   epochIteration_Cython_BPR_SGD :583-678, epochIteration_Cython_FUNK_SVD_SGD :289-390,
//   _apply_minibatch_updates_to_latent_factors :773-832, adaptive_gradient :838-876,
//   sampleBPR_Cython :943-987, sampleMSE_Cython :881-938.
//
// Two execution modes (DESIGN.md "K2"):
//   * mini-batch (the reference's semantics, exactly): parameters are frozen inside a batch, every sample's
//     gradient is accumulated into per-row accumulators, then every touched row takes one step with the mean
//     gradient (divided by batch_size, pyx:805,828) through the adaptive rule.  One persistent cooperative
//     kernel runs the whole epoch: phase 1 (one warp per sample: three coalesced row gathers, warp-shuffle dot,
//     vector RED.ADD into the accumulators), grid sync, phase 2 (one warp per touched row), grid sync.
//     The sample stream is either replayed on the host with glibc's rand() (bit-compatible with the reference's
//     libc calls) or drawn on the device with Philox4x32-10.
//   * hogwild: no batch barrier -- every warp applies its sample's update immediately (the reference's
//     batch_size=1 recursion run concurrently; races between warps are the usual Hogwild races).
// Roofline: HBM; bytes per BPR sample = 6 * f * 4 (three rows read, three accumulator rows RMW).
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace mf {

enum Algo { MF_BPR = 0, FUNK_SVD = 1 };
enum SgdMode { SGD = 0, ADAGRAD = 1, RMSPROP = 2, ADAM = 3 };

struct Params {
  int n_users, n_items, f, batch_size, algorithm, use_bias, sgd_mode, hogwild;
  float lr, user_reg, item_reg, bias_reg, positive_reg, negative_reg;
  float gamma, beta1, beta2;
  double b1_pow, b2_pow;  // adam powers at the start of the epoch
  float *U, *V, *bu, *bi, *mu;
  // batch gradient sums are fp64 like pyx:305-354: a double atomic sum is order-independent to ~1e-16 relative, so the
  // mini-batch mode is run-to-run deterministic at the fp32 precision of the parameters (fp32 RED.ADD was not: the
  // summation order of a row's samples changed the rounded sum, and Adam's m/(sqrt(v)+eps) amplified it)
  double *accU, *accV, *accbu, *accbi, *accmu;
  float *cU, *cV, *cbu, *cbi, *cmu;                                    // adagrad / rmsprop cache
  float *m1U, *m2U, *m1V, *m2V, *m1bu, *m2bu, *m1bi, *m2bi, *m1mu, *m2mu;  // adam
  int *flagI, *flagU, *listI, *listU, *cnt;  // cnt[4]: items/users counters, double-buffered by batch parity
  const int* su; const int* si; const int* sj; const float* sr;  // sample stream of the epoch
  long long n_batches;
  double* pow_out;  // [2] adam powers after the epoch
};

struct AdaptCtx {
  int mode;
  float gamma, beta1, beta2, inv1, inv2;  // inv = 1 / (1 - beta^t)
};

// pyx:838-876 on one element; c / m1 / m2 point at this element's state
__device__ __forceinline__ float adapt(const AdaptCtx& a, float g, float* c, float* m1, float* m2) {
  if (a.mode == ADAGRAD) {
    const float cc = *c + g * g;
    *c = cc;
    return g / (sqrtf(cc) + 1e-8f);
  } else if (a.mode == RMSPROP) {
    const float cc = *c * a.gamma + (1.f - a.gamma) * g * g;
    *c = cc;
    return g / (sqrtf(cc) + 1e-8f);
  } else if (a.mode == ADAM) {
    const float mm1 = *m1 * a.beta1 + (1.f - a.beta1) * g;
    const float mm2 = *m2 * a.beta2 + (1.f - a.beta2) * g * g;
    *m1 = mm1;
    *m2 = mm2;
    return (mm1 * a.inv1) / (sqrtf(mm2 * a.inv2) + 1e-8f);
  }
  return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
#if __CUDA_ARCH__ >= 900
  atomicAdd(reinterpret_cast<float4*>(addr), v);
#else
  atomicAdd(addr, v.x); atomicAdd(addr + 1, v.y); atomicAdd(addr + 2, v.z); atomicAdd(addr + 3, v.w);
#endif
}

__device__ __forceinline__ void touch(int* flag, int* list, int* counter, int row) {
  if (atomicExch(flag + row, 1) == 0) list[atomicAdd(counter, 1)] = row;
}

// ---- phase 1: gradients of one sample, accumulated (mini-batch mode)
template <bool VEC4>
__device__ __forceinline__ void bpr_accumulate(const Params& p, int u, int i, int j, int lane) {
  const int f = p.f;
  const float* Uu = p.U + (size_t)u * f;
  const float* Vi = p.V + (size_t)i * f;
  const float* Vj = p.V + (size_t)j * f;
  float x = 0.f;
  if (VEC4) {
    for (int q = lane * 4; q < f; q += 128) {
      const float4 a = *reinterpret_cast<const float4*>(Uu + q), b = *reinterpret_cast<const float4*>(Vi + q),
                   c = *reinterpret_cast<const float4*>(Vj + q);
      x += a.x * (b.x - c.x) + a.y * (b.y - c.y) + a.z * (b.z - c.z) + a.w * (b.w - c.w);
    }
  } else {
    for (int q = lane; q < f; q += 32) x += Uu[q] * (Vi[q] - Vj[q]);
  }
  x = warp_sum(x);
  const float sig = 1.f / (1.f + expf(x));  // pyx:622
  double* aU = p.accU + (size_t)u * f;
  double* aI = p.accV + (size_t)i * f;
  double* aJ = p.accV + (size_t)j * f;
  const double sg = (double)sig, rp = (double)p.positive_reg, rn = (double)p.negative_reg, ru = (double)p.user_reg;
  for (int q = lane; q < f; q += 32) {  // consecutive lanes -> consecutive 8-byte RED.ADD.F64 (coalesced)
    const double a = (double)Uu[q], b = (double)Vi[q], c = (double)Vj[q];
    atomicAdd(aI + q, sg * a - rp * b);        // pyx:633
    atomicAdd(aJ + q, -sg * a - rn * c);       // pyx:634
    atomicAdd(aU + q, sg * (b - c) - ru * a);  // pyx:635
  }
}

template <bool VEC4>
__device__ __forceinline__ void mse_accumulate(const Params& p, int u, int i, float r, int lane) {
  const int f = p.f;
  const float* Uu = p.U + (size_t)u * f;
  const float* Vi = p.V + (size_t)i * f;
  float x = 0.f;
  if (VEC4) {
    for (int q = lane * 4; q < f; q += 128) {
      const float4 a = *reinterpret_cast<const float4*>(Uu + q), b = *reinterpret_cast<const float4*>(Vi + q);
      x += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
  } else {
    for (int q = lane; q < f; q += 32) x += Uu[q] * Vi[q];
  }
  x = warp_sum(x);
  if (p.use_bias) x += p.mu[0] + p.bu[u] + p.bi[i];  // pyx:313-316
  const float err = r - x;
  const double er = (double)err;
  if (p.use_bias && lane == 0) {  // pyx:332-339
    const double rb = (double)p.bias_reg;
    atomicAdd(p.accmu, er - rb * (double)p.mu[0]);
    atomicAdd(p.accbi + i, er - rb * (double)p.bi[i]);
    atomicAdd(p.accbu + u, er - rb * (double)p.bu[u]);
  }
  double* aU = p.accU + (size_t)u * f;
  double* aI = p.accV + (size_t)i * f;
  const double rp = (double)p.positive_reg, ru = (double)p.user_reg;
  for (int q = lane; q < f; q += 32) {
    const double a = (double)Uu[q], b = (double)Vi[q];
    atomicAdd(aI + q, er * a - rp * b);  // item regulariser is positive_reg, not item_reg (pyx:349)
    atomicAdd(aU + q, er * b - ru * a);
  }
}

// ---- phase 2: one touched row takes its step (pyx:792-832)
__device__ __forceinline__ void apply_row(const Params& p, const AdaptCtx& ad, float* P, double* acc, float* c, float* m1,
                                          float* m2, size_t row, int lane, double inv_bs) {
  const int f = p.f;
  const size_t o = row * (size_t)f;
  for (int q = lane; q < f; q += 32) {
    float g = (float)(acc[o + q] * inv_bs);
    g = adapt(ad, g, c ? c + o + q : nullptr, m1 ? m1 + o + q : nullptr, m2 ? m2 + o + q : nullptr);
    P[o + q] += p.lr * g;
    acc[o + q] = 0.0;
  }
}

__device__ __forceinline__ void apply_scalar(const Params& p, const AdaptCtx& ad, float* P, double* acc, float* c, float* m1,
                                             float* m2, size_t k, double inv_bs) {
  float g = (float)(acc[k] * inv_bs);
  g = adapt(ad, g, c ? c + k : nullptr, m1 ? m1 + k : nullptr, m2 ? m2 + k : nullptr);
  P[k] += p.lr * g;
  acc[k] = 0.0;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) mf_epoch_kernel(const Params p) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double inv_bs = 1.0 / (double)p.batch_size;
  double b1p = p.b1_pow, b2p = p.b2_pow;
  AdaptCtx ad;
  ad.mode = p.sgd_mode; ad.gamma = p.gamma; ad.beta1 = p.beta1; ad.beta2 = p.beta2;
  for (long long b = 0; b < p.n_batches; ++b) {
    int* cnt = p.cnt + 2 * (int)(b & 1);
    // ---------------- phase 1
    for (long long s = warp; s < p.batch_size; s += n_warps) {
      const long long g = b * p.batch_size + s;
      const int u = p.su[g], i = p.si[g];
      if (p.algorithm == MF_BPR) {
        const int j = p.sj[g];
        if (lane == 0) { touch(p.flagI, p.listI, cnt, i); touch(p.flagI, p.listI, cnt, j); touch(p.flagU, p.listU, cnt + 1, u); }
        bpr_accumulate<VEC4>(p, u, i, j, lane);
      } else {
        if (lane == 0) { touch(p.flagI, p.listI, cnt, i); touch(p.flagU, p.listU, cnt + 1, u); }
        mse_accumulate<VEC4>(p, u, i, p.sr[g], lane);
      }
    }
    grid.sync();
    // ---------------- phase 2
    ad.inv1 = (float)(1.0 / (1.0 - b1p));
    ad.inv2 = (float)(1.0 / (1.0 - b2p));
    const int nI = cnt[0], nU = cnt[1];
    if (p.use_bias && warp == 0 && lane == 0) apply_scalar(p, ad, p.mu, p.accmu, p.cmu, p.m1mu, p.m2mu, 0, inv_bs);
    for (long long t = warp; t < nI + nU; t += n_warps) {
      if (t < nI) {
        const int k = p.listI[t];
        if (p.use_bias && lane == 0) apply_scalar(p, ad, p.bi, p.accbi, p.cbi, p.m1bi, p.m2bi, k, inv_bs);
        apply_row(p, ad, p.V, p.accV, p.cV, p.m1V, p.m2V, k, lane, inv_bs);
        if (lane == 0) p.flagI[k] = 0;
      } else {
        const int k = p.listU[t - nI];
        if (p.use_bias && lane == 0) apply_scalar(p, ad, p.bu, p.accbu, p.cbu, p.m1bu, p.m2bu, k, inv_bs);
        apply_row(p, ad, p.U, p.accU, p.cU, p.m1U, p.m2U, k, lane, inv_bs);
        if (lane == 0) p.flagU[k] = 0;
      }
    }
    if (warp == 0 && lane == 0) { int* nxt = p.cnt + 2 * (int)((b + 1) & 1); nxt[0] = 0; nxt[1] = 0; }
    if (p.sgd_mode == ADAM) { b1p *= (double)p.beta1; b2p *= (double)p.beta2; }  // once per batch, pyx:649-652
    grid.sync();
  }
  if (warp == 0 && lane == 0) { p.pow_out[0] = b1p; p.pow_out[1] = b2p; p.cnt[0] = p.cnt[1] = p.cnt[2] = p.cnt[3] = 0; }
}

// ---- hogwild: every warp applies its samples' updates at once (batch_size = 1 recursion, concurrent)
__global__ void __launch_bounds__(256) mf_hogwild_kernel(const Params p, long long n_samples) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int f = p.f;
  AdaptCtx ad;
  ad.mode = p.sgd_mode; ad.gamma = p.gamma; ad.beta1 = p.beta1; ad.beta2 = p.beta2;
  for (long long g = warp; g < n_samples; g += n_warps) {
    if (p.sgd_mode == ADAM) {  // the reference advances the powers once per (size-1) batch
      ad.inv1 = (float)(1.0 / (1.0 - p.b1_pow * pow((double)p.beta1, (double)g)));
      ad.inv2 = (float)(1.0 / (1.0 - p.b2_pow * pow((double)p.beta2, (double)g)));
    }
    const int u = p.su[g], i = p.si[g];
    float* Uu = p.U + (size_t)u * f;
    float* Vi = p.V + (size_t)i * f;
    if (p.algorithm == MF_BPR && p.sgd_mode == SGD && (f & 3) == 0) {
      // plain-SGD BPR, rows as float4: one 16-byte load and store per lane and row for f = 128
      const int j = p.sj[g];
      float* Vj = p.V + (size_t)j * f;
      float x = 0.f;
      for (int q = lane * 4; q < f; q += 128) {
        const float4 a = *reinterpret_cast<const float4*>(Uu + q), b = *reinterpret_cast<const float4*>(Vi + q),
                     c = *reinterpret_cast<const float4*>(Vj + q);
        x += a.x * (b.x - c.x) + a.y * (b.y - c.y) + a.z * (b.z - c.z) + a.w * (b.w - c.w);
      }
      x = warp_sum(x);
      const float sig = 1.f / (1.f + expf(x));
      const float lr = p.lr;
      for (int q = lane * 4; q < f; q += 128) {
        const float4 a = *reinterpret_cast<const float4*>(Uu + q), b = *reinterpret_cast<const float4*>(Vi + q),
                     c = *reinterpret_cast<const float4*>(Vj + q);
        // deltas go through RED.ADD so that concurrent samples sharing a row add up instead of overwriting each other
        red_add4(Vi + q, make_float4(lr * (sig * a.x - p.positive_reg * b.x), lr * (sig * a.y - p.positive_reg * b.y),
                                     lr * (sig * a.z - p.positive_reg * b.z), lr * (sig * a.w - p.positive_reg * b.w)));
        red_add4(Vj + q, make_float4(lr * (-sig * a.x - p.negative_reg * c.x), lr * (-sig * a.y - p.negative_reg * c.y),
                                     lr * (-sig * a.z - p.negative_reg * c.z), lr * (-sig * a.w - p.negative_reg * c.w)));
        red_add4(Uu + q, make_float4(lr * (sig * (b.x - c.x) - p.user_reg * a.x), lr * (sig * (b.y - c.y) - p.user_reg * a.y),
                                     lr * (sig * (b.z - c.z) - p.user_reg * a.z), lr * (sig * (b.w - c.w) - p.user_reg * a.w)));
      }
    } else if (p.algorithm == MF_BPR) {
      const int j = p.sj[g];
      float* Vj = p.V + (size_t)j * f;
      float x = 0.f;
      for (int q = lane; q < f; q += 32) x += Uu[q] * (Vi[q] - Vj[q]);
      x = warp_sum(x);
      const float sig = 1.f / (1.f + expf(x));
      for (int q = lane; q < f; q += 32) {
        const float a = Uu[q], b = Vi[q], c = Vj[q];
        const size_t oi = (size_t)i * f + q, oj = (size_t)j * f + q, ou = (size_t)u * f + q;
        // items first, then the user, as pyx:792-832 orders the apply
        atomicAdd(Vi + q, p.lr * adapt(ad, sig * a - p.positive_reg * b, p.cV ? p.cV + oi : nullptr, p.m1V ? p.m1V + oi : nullptr, p.m2V ? p.m2V + oi : nullptr));
        atomicAdd(Vj + q, p.lr * adapt(ad, -sig * a - p.negative_reg * c, p.cV ? p.cV + oj : nullptr, p.m1V ? p.m1V + oj : nullptr, p.m2V ? p.m2V + oj : nullptr));
        atomicAdd(Uu + q, p.lr * adapt(ad, sig * (b - c) - p.user_reg * a, p.cU ? p.cU + ou : nullptr, p.m1U ? p.m1U + ou : nullptr, p.m2U ? p.m2U + ou : nullptr));
      }
    } else {
      float x = 0.f;
      for (int q = lane; q < f; q += 32) x += Uu[q] * Vi[q];
      x = warp_sum(x);
      if (p.use_bias) x += p.mu[0] + p.bu[u] + p.bi[i];
      const float err = p.sr[g] - x;
      if (p.use_bias && lane == 0) {
        p.mu[0] += p.lr * adapt(ad, err - p.bias_reg * p.mu[0], p.cmu, p.m1mu, p.m2mu);
        p.bi[i] += p.lr * adapt(ad, err - p.bias_reg * p.bi[i], p.cbi ? p.cbi + i : nullptr, p.m1bi ? p.m1bi + i : nullptr, p.m2bi ? p.m2bi + i : nullptr);
        p.bu[u] += p.lr * adapt(ad, err - p.bias_reg * p.bu[u], p.cbu ? p.cbu + u : nullptr, p.m1bu ? p.m1bu + u : nullptr, p.m2bu ? p.m2bu + u : nullptr);
      }
      for (int q = lane; q < f; q += 32) {
        const float a = Uu[q], b = Vi[q];
        const size_t oi = (size_t)i * f + q, ou = (size_t)u * f + q;
        atomicAdd(Vi + q, p.lr * adapt(ad, err * a - p.positive_reg * b, p.cV ? p.cV + oi : nullptr, p.m1V ? p.m1V + oi : nullptr, p.m2V ? p.m2V + oi : nullptr));
        atomicAdd(Uu + q, p.lr * adapt(ad, err * b - p.user_reg * a, p.cU ? p.cU + ou : nullptr, p.m1U ? p.m1U + ou : nullptr, p.m2U ? p.m2U + ou : nullptr));
      }
    }
  }
}

// ---- device sampler: Philox4x32-10, counter = (sample index, draw block), key = (seed, epoch)
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}
__device__ __forceinline__ uint4 philox(unsigned long long idx, unsigned blk, unsigned seed, unsigned epoch) {
  unsigned c0 = (unsigned)idx, c1 = (unsigned)(idx >> 32), c2 = blk, c3 = 0x9E3779B9u;
  unsigned k0 = seed, k1 = epoch;
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  return make_uint4(c0, c1, c2, c3);
}

struct Draws {
  unsigned long long idx; unsigned seed, epoch, blk; uint4 cur; int pos;
  __device__ Draws(unsigned long long i, unsigned s, unsigned e) : idx(i), seed(s), epoch(e), blk(0), pos(4) {}
  __device__ unsigned next() {
    if (pos == 4) { cur = philox(idx, blk++, seed, epoch); pos = 0; }
    const unsigned v = pos == 0 ? cur.x : (pos == 1 ? cur.y : (pos == 2 ? cur.z : cur.w));
    ++pos;
    return v;
  }
};

// same acceptance rules as sampleBPR_Cython / sampleMSE_Cython (users with 0 < profile < n_items; negative item
// not in the sorted profile, binary search instead of the linear scan), different random stream
__global__ void mf_sample_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
                                 int user_lo, int n_users, int n_items, int algorithm, float quota, long long n_samples, unsigned seed,
                                 unsigned epoch, int* su, int* si, int* sj, float* sr) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_samples) return;
  Draws d((unsigned long long)g, seed, epoch);
  int u, s, n;
  do {
    u = user_lo + (int)(d.next() % (unsigned)n_users);  // n_users = size of this rank's user shard
    s = indptr[u];
    n = indptr[u + 1] - s;
  } while (n == 0 || n == n_items);
  bool positive = true;
  if (algorithm == FUNK_SVD && quota != 0.f) positive = (float)(d.next() >> 8) * (1.f / 16777216.f) <= quota;
  int item;
  float r = 0.f;
  if (algorithm == MF_BPR || positive) {
    const int k = (int)(d.next() % (unsigned)n);
    item = indices[s + k];
    if (algorithm == FUNK_SVD) r = data[s + k];
  }
  if (algorithm == MF_BPR || !positive) {
    int neg;
    while (true) {
      neg = (int)(d.next() % (unsigned)n_items);
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (indices[s + mid] < neg) lo = mid + 1; else hi = mid; }
      if (lo == n || indices[s + lo] != neg) break;
    }
    if (algorithm == MF_BPR) sj[g] = neg; else { item = neg; r = 0.f; }
  }
  su[g] = u;
  si[g] = item;
  if (algorithm == FUNK_SVD) sr[g] = r;
}

// host replay of glibc srand()/rand() (TYPE_3 additive feedback, r[i] = r[i-31] + r[i-3], 310 discarded, >> 1)
struct GlibcRand {
  int32_t r[31];
  int f = 3, b = 0;
  void seed(unsigned s) {
    int32_t word = s == 0 ? 1 : (int32_t)s;
    r[0] = word;
    for (int i = 1; i < 31; ++i) {
      const long hi = word / 127773, lo = word % 127773;
      long w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      word = (int32_t)w;
      r[i] = word;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; ++i) next_raw();
  }
  uint32_t next_raw() {
    const uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
    r[f] = (int32_t)v;
    f = (f + 1) % 31;
    b = (b + 1) % 31;
    return v;
  }
  int next() { return (int)(next_raw() >> 1); }
};

}  // namespace mf
}  // namespace b200

using namespace b200;
using namespace b200::mf;

struct b200_mf_s {
  Params p{};
  int sampler = 0;  // 0 glibc replay on the host, 1 Philox on the device
  unsigned seed = 1;
  unsigned epoch = 0;
  long long nnz = 0;
  float quota = 0.5f;
  GlibcRand rng;
  std::vector<int> h_indptr, h_indices;
  std::vector<float> h_data;
  DevBuf<int> d_indptr, d_indices;
  DevBuf<float> d_data;
  std::vector<DevBuf<float>> fbufs;  // owns every float device array referenced by p
  std::vector<DevBuf<double>> dbufs;  // the fp64 gradient accumulators of the mini-batch mode
  double* dalloc(size_t n) {
    dbufs.emplace_back(std::max<size_t>(n, 1));
    double* d = dbufs.back().get();
    B200_CUDA(cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(double)));
    return d;
  }
  DevBuf<int> flagI, flagU, listI, listU, cnt, su, si, sj;
  DevBuf<float> sr;
  DevBuf<double> pow_out;
  std::vector<int> hs_u, hs_i, hs_j;
  std::vector<float> hs_r;
  long long samples_last = 0, cap_samples = 0, epoch_samples_override = 0;
  int shard_lo = 0, shard_hi = 0;  // device sampler draws users from [shard_lo, shard_hi) when set (multi-GPU user sharding)
  int grid = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  float* falloc(size_t n, const double* init) {
    fbufs.emplace_back(std::max<size_t>(n, 1));
    float* d = fbufs.back().get();
    if (init) {
      std::vector<float> tmp(n);
      for (size_t i = 0; i < n; ++i) tmp[i] = (float)init[i];
      B200_CUDA(cudaMemcpy(d, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    } else {
      B200_CUDA(cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(float)));
    }
    return d;
  }
};

namespace {

long long epoch_batches(const b200_mf_s* h) {
  // pyx:586 (BPR: n_users / batch_size + 1) and pyx:292 (FunkSVD: nnz / batch_size + 1)
  return (h->p.algorithm == MF_BPR ? (long long)h->p.n_users : h->nnz) / h->p.batch_size + 1;
}

void host_samples(b200_mf_s* h, long long n) {
  // sampleBPR_Cython pyx:943-987 / sampleMSE_Cython pyx:881-938, draw for draw
  h->hs_u.resize((size_t)n); h->hs_i.resize((size_t)n);
  if (h->p.algorithm == MF_BPR) h->hs_j.resize((size_t)n); else h->hs_r.resize((size_t)n);
  const int* indptr = h->h_indptr.data();
  const int* indices = h->h_indices.data();
  const int nU = h->p.n_users, nI = h->p.n_items;
  for (long long g = 0; g < n; ++g) {
    long u = 0, start = 0, len = 0;
    while (len == 0 || len == nI) {
      u = h->rng.next() % nU;
      start = indptr[u];
      len = indptr[u + 1] - start;
    }
    bool positive = true;
    if (h->p.algorithm == FUNK_SVD && h->quota != 0.0f) positive = h->rng.next() <= (double)h->quota * 2147483647.0;
    long item = -1;
    float r = 0.f;
    if (h->p.algorithm == MF_BPR || positive) {
      const long k = h->rng.next() % len;
      item = indices[start + k];
      if (h->p.algorithm == FUNK_SVD) r = h->h_data[(size_t)(start + k)];
    }
    if (h->p.algorithm == MF_BPR || !positive) {
      long neg;
      for (;;) {
        neg = h->rng.next() % nI;
        const int* lo = std::lower_bound(indices + start, indices + start + len, (int)neg);
        if (lo == indices + start + len || *lo != neg) break;
      }
      if (h->p.algorithm == MF_BPR) h->hs_j[(size_t)g] = (int)neg; else { item = neg; r = 0.f; }
    }
    h->hs_u[(size_t)g] = (int)u;
    h->hs_i[(size_t)g] = (int)item;
    if (h->p.algorithm == FUNK_SVD) h->hs_r[(size_t)g] = r;
  }
}

}  // namespace

extern "C" {

int b200_mf_create(b200_mf_t* out, int64_t n_users, int64_t n_items, int64_t nnz, const int32_t* h_indptr,
                   const int32_t* h_indices, const float* h_data, int n_factors, int algorithm, int batch_size,
                   float negative_interactions_quota, float learning_rate, int use_bias, float user_reg, float item_reg,
                   float bias_reg, float positive_reg, float negative_reg, int sgd_mode, float gamma, float beta_1,
                   float beta_2, const double* h_user_factors, const double* h_item_factors, int has_seed,
                   uint32_t random_seed, int sampler, int hogwild) {
  if (out) *out = nullptr;
  b200_mf_s* h = nullptr;
  int rc = guarded([&] {
    B200_REQUIRE(out && h_indptr && h_user_factors && h_item_factors, "b200_mf_create: NULL argument");
    B200_REQUIRE(n_users > 0 && n_items > 0 && nnz >= 0 && nnz < (1ll << 31) - 1, "b200_mf_create: bad shape");
    B200_REQUIRE(n_factors >= 1 && batch_size >= 1, "b200_mf_create: n_factors and batch_size must be >= 1");
    B200_REQUIRE(algorithm == MF_BPR || algorithm == FUNK_SVD, "b200_mf_create: unknown algorithm %d", algorithm);
    B200_REQUIRE(sgd_mode >= SGD && sgd_mode <= ADAM, "b200_mf_create: unknown sgd_mode %d", sgd_mode);
    h = new b200_mf_s();
    Params& p = h->p;
    p.n_users = (int)n_users; p.n_items = (int)n_items; p.f = n_factors; p.batch_size = batch_size;
    p.algorithm = algorithm; p.use_bias = use_bias != 0; p.sgd_mode = sgd_mode; p.hogwild = hogwild != 0;
    p.lr = learning_rate; p.user_reg = user_reg; p.item_reg = item_reg; p.bias_reg = bias_reg;
    p.positive_reg = positive_reg; p.negative_reg = negative_reg;
    p.gamma = gamma; p.beta1 = beta_1; p.beta2 = beta_2;
    p.b1_pow = beta_1; p.b2_pow = beta_2;  // pyx:220-221
    h->nnz = nnz;
    h->quota = negative_interactions_quota;
    h->sampler = sampler;
    h->seed = has_seed ? random_seed : 1u;
    h->rng.seed(h->seed);
    h->h_indptr.assign(h_indptr, h_indptr + n_users + 1);
    h->h_indices.assign(h_indices, h_indices + nnz);
    if (algorithm == FUNK_SVD) h->h_data.assign(h_data, h_data + nnz);
    h->d_indptr.alloc((size_t)n_users + 1);
    h->d_indices.alloc((size_t)std::max<int64_t>(nnz, 1));
    h->d_data.alloc((size_t)std::max<int64_t>(nnz, 1));
    B200_CUDA(cudaMemcpy(h->d_indptr.get(), h_indptr, sizeof(int) * ((size_t)n_users + 1), cudaMemcpyHostToDevice));
    if (nnz) {
      B200_CUDA(cudaMemcpy(h->d_indices.get(), h_indices, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
      B200_CUDA(cudaMemcpy(h->d_data.get(), h_data, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    const size_t nUf = (size_t)n_users * n_factors, nIf = (size_t)n_items * n_factors;
    h->fbufs.reserve(40);
    p.U = h->falloc(nUf, h_user_factors);
    p.V = h->falloc(nIf, h_item_factors);
    h->dbufs.reserve(8);
    if (!p.hogwild) { p.accU = h->dalloc(nUf); p.accV = h->dalloc(nIf); }
    if (p.use_bias) {
      p.bu = h->falloc((size_t)n_users, nullptr); p.bi = h->falloc((size_t)n_items, nullptr); p.mu = h->falloc(1, nullptr);
      p.accbu = h->dalloc((size_t)n_users); p.accbi = h->dalloc((size_t)n_items); p.accmu = h->dalloc(1);
    }
    if (sgd_mode == ADAGRAD || sgd_mode == RMSPROP) {
      p.cU = h->falloc(nUf, nullptr); p.cV = h->falloc(nIf, nullptr);
      if (p.use_bias) { p.cbu = h->falloc((size_t)n_users, nullptr); p.cbi = h->falloc((size_t)n_items, nullptr); p.cmu = h->falloc(1, nullptr); }
    } else if (sgd_mode == ADAM) {
      p.m1U = h->falloc(nUf, nullptr); p.m2U = h->falloc(nUf, nullptr); p.m1V = h->falloc(nIf, nullptr); p.m2V = h->falloc(nIf, nullptr);
      if (p.use_bias) {
        p.m1bu = h->falloc((size_t)n_users, nullptr); p.m2bu = h->falloc((size_t)n_users, nullptr);
        p.m1bi = h->falloc((size_t)n_items, nullptr); p.m2bi = h->falloc((size_t)n_items, nullptr);
        p.m1mu = h->falloc(1, nullptr); p.m2mu = h->falloc(1, nullptr);
      }
    }
    h->flagI.alloc((size_t)n_items); h->flagU.alloc((size_t)n_users);
    h->listI.alloc((size_t)2 * batch_size); h->listU.alloc((size_t)batch_size); h->cnt.alloc(4);
    B200_CUDA(cudaMemset(h->flagI.get(), 0, sizeof(int) * (size_t)n_items));
    B200_CUDA(cudaMemset(h->flagU.get(), 0, sizeof(int) * (size_t)n_users));
    B200_CUDA(cudaMemset(h->cnt.get(), 0, sizeof(int) * 4));
    p.flagI = h->flagI.get(); p.flagU = h->flagU.get(); p.listI = h->listI.get(); p.listU = h->listU.get(); p.cnt = h->cnt.get();
    h->pow_out.alloc(2);
    p.pow_out = h->pow_out.get();
    h->cap_samples = epoch_batches(h) * batch_size;
    h->su.alloc((size_t)h->cap_samples); h->si.alloc((size_t)h->cap_samples);
    if (algorithm == MF_BPR) h->sj.alloc((size_t)h->cap_samples); else h->sr.alloc((size_t)h->cap_samples);
    p.su = h->su.get(); p.si = h->si.get(); p.sj = h->sj.get(); p.sr = h->sr.get();
    // cooperative grid: every block resident
    int per_sm = 0;
    const bool vec4 = (n_factors % 4) == 0;
    if (vec4) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mf_epoch_kernel<true>, 256, 0));
    else B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mf_epoch_kernel<false>, 256, 0));
    B200_REQUIRE(per_sm >= 1, "b200_mf_create: epoch kernel does not fit on an SM");
    h->grid = sm_count() * std::min(per_sm, 8);
    B200_CUDA(cudaEventCreate(&h->ev0));
    B200_CUDA(cudaEventCreate(&h->ev1));
    *out = h;
  });
  if (rc != B200_OK && h) delete h;
  return rc;
}

int b200_mf_destroy(b200_mf_t h) {
  if (!h) return B200_OK;
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return B200_OK;
}

int b200_mf_epoch(b200_mf_t h, void* stream) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_mf_epoch: NULL handle");
    cudaStream_t st = (cudaStream_t)stream;
    Params& p = h->p;
    p.n_batches = epoch_batches(h);
    if (h->epoch_samples_override > 0) p.n_batches = std::max<long long>(1, h->epoch_samples_override / p.batch_size);
    const long long n = p.n_batches * p.batch_size;
    if (h->sampler == 0) {
      host_samples(h, n);
      B200_CUDA(cudaMemcpyAsync(h->su.get(), h->hs_u.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(h->si.get(), h->hs_i.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
      if (p.algorithm == MF_BPR) B200_CUDA(cudaMemcpyAsync(h->sj.get(), h->hs_j.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
      else B200_CUDA(cudaMemcpyAsync(h->sr.get(), h->hs_r.data(), sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    B200_CUDA(cudaEventRecord(h->ev0, st));
    if (h->sampler != 0) {
      mf_sample_kernel<<<div_up(n, 256), 256, 0, st>>>(h->d_indptr.get(), h->d_indices.get(), h->d_data.get(), h->shard_lo,
                                                      (h->shard_hi > h->shard_lo ? h->shard_hi - h->shard_lo : p.n_users), p.n_items,
                                                      p.algorithm, h->quota, n, h->seed, h->epoch, h->su.get(), h->si.get(),
                                                      h->sj.get(), h->sr.get());
      count_launch();
    }
    if (p.hogwild) {
      mf_hogwild_kernel<<<sm_count() * 8, 256, 0, st>>>(p, n);
      B200_CUDA(cudaGetLastError());
      if (p.sgd_mode == ADAM) {  // powers advance once per size-1 batch
        p.b1_pow *= pow((double)p.beta1, (double)n);
        p.b2_pow *= pow((double)p.beta2, (double)n);
      }
    } else {
      void* args[] = {(void*)&p};
      if (p.f % 4 == 0) B200_CUDA(cudaLaunchCooperativeKernel((void*)mf_epoch_kernel<true>, dim3(h->grid), dim3(256), args, 0, st));
      else B200_CUDA(cudaLaunchCooperativeKernel((void*)mf_epoch_kernel<false>, dim3(h->grid), dim3(256), args, 0, st));
    }
    count_launch();
    B200_CUDA(cudaEventRecord(h->ev1, st));
    h->timed = true;
    if (!p.hogwild && p.sgd_mode == ADAM) {
      double pw[2];
      B200_CUDA(cudaMemcpyAsync(pw, h->pow_out.get(), sizeof(pw), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      p.b1_pow = pw[0];
      p.b2_pow = pw[1];
    } else if (h->sampler == 0) {
      B200_CUDA(cudaStreamSynchronize(st));  // the host sample vectors are reused by the next epoch
    }
    h->samples_last = n;
    h->epoch += 1;
  });
}

int b200_mf_set_user_shard(b200_mf_t h, int user_lo, int user_hi, int64_t samples_per_epoch, uint32_t stream_id) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_mf_set_user_shard: NULL handle");
    B200_REQUIRE(h->sampler != 0, "b200_mf_set_user_shard: only the device (Philox) sampler can be sharded");
    B200_REQUIRE(0 <= user_lo && user_lo < user_hi && user_hi <= h->p.n_users, "b200_mf_set_user_shard: bad range [%d,%d)", user_lo, user_hi);
    B200_REQUIRE(samples_per_epoch >= 0 && samples_per_epoch <= h->cap_samples, "b200_mf_set_user_shard: samples_per_epoch out of range");
    h->shard_lo = user_lo;
    h->shard_hi = user_hi;
    h->epoch_samples_override = samples_per_epoch;
    h->seed += 0x9E3779B9u * stream_id;  // decorrelates the ranks' Philox streams
  });
}

int b200_mf_samples_last_epoch(b200_mf_t h, int64_t* n) {
  return guarded([&] {
    B200_REQUIRE(h && n, "b200_mf_samples_last_epoch: NULL argument");
    *n = h->samples_last;
  });
}

int b200_mf_get_samples(b200_mf_t h, int32_t* u, int32_t* i, int32_t* j, float* r) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_mf_get_samples: NULL handle");
    B200_CUDA(cudaDeviceSynchronize());
    const size_t n = (size_t)h->samples_last;
    if (u) B200_CUDA(cudaMemcpy(u, h->su.get(), sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (i) B200_CUDA(cudaMemcpy(i, h->si.get(), sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (j && h->p.algorithm == MF_BPR) B200_CUDA(cudaMemcpy(j, h->sj.get(), sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (r && h->p.algorithm == FUNK_SVD) B200_CUDA(cudaMemcpy(r, h->sr.get(), sizeof(float) * n, cudaMemcpyDeviceToHost));
  });
}

int b200_mf_get_factors(b200_mf_t h, double* user_factors, double* item_factors, double* user_bias, double* item_bias,
                        double* global_bias) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_mf_get_factors: NULL handle");
    B200_CUDA(cudaDeviceSynchronize());
    auto fetch = [&](const float* d, double* out, size_t n) {
      if (!out || !d) return;
      std::vector<float> tmp(n);
      B200_CUDA(cudaMemcpy(tmp.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
      for (size_t k = 0; k < n; ++k) out[k] = (double)tmp[k];
    };
    fetch(h->p.U, user_factors, (size_t)h->p.n_users * h->p.f);
    fetch(h->p.V, item_factors, (size_t)h->p.n_items * h->p.f);
    fetch(h->p.bu, user_bias, (size_t)h->p.n_users);
    fetch(h->p.bi, item_bias, (size_t)h->p.n_items);
    fetch(h->p.mu, global_bias, 1);
  });
}

int b200_mf_device_factors(b200_mf_t h, float** d_user_factors, float** d_item_factors) {
  return guarded([&] {
    B200_REQUIRE(h != nullptr, "b200_mf_device_factors: NULL handle");
    if (d_user_factors) *d_user_factors = h->p.U;
    if (d_item_factors) *d_item_factors = h->p.V;
  });
}

int b200_mf_last_epoch_ms(b200_mf_t h, float* ms) {
  return guarded([&] {
    B200_REQUIRE(h && ms && h->timed, "b200_mf_last_epoch_ms: no epoch run yet");
    B200_CUDA(cudaEventSynchronize(h->ev1));
    B200_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  });
}

}  // extern "C"
