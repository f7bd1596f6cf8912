// K1-B: sparse-candidate variant of the binary similarity kernel (included by sim_topk.cu inside namespace b200::sim).
//
// STATUS: opt-in (B200REC_K1B=1 at handle creation); written after the round's GPU time was spent, never executed.
// tests/test_similarity_gpu.py runs unchanged with the variable set -- that is the switch-over check.
//
// Why: the window kernel spends 32 K of its 74.5 K cycles per C5 column sweeping every accumulator
// cell twice (bootstrap + scan), although ~90 % of a column's gathered entries touch their cell for the only time and a
// count-1 cell can reach the top-K only if its index lies below a bound that follows from the current floor (the
// neighbour axis is sorted by ascending norm term, and every formula decreases with it).  So, per target column:
//   bm1 / bm2   two bitmaps over the neighbour axis: touched at least once / at least twice (atomicOr returns the old word);
//   table       open-addressing (key = neighbour, value = hits after the first) for the cells hit again;
//   bootstrap, scan, evaluate run over the table slots (16 K) instead of 200 K cells; the table is cleared while it is read;
//   count-1 cells (bm1 & ~bm2) are enumerated only over the leading coarse tiles whose BEST count-1 similarity still
//   reaches the floor -- none at C5 densities;
//   one window whatever n_items (n_win = 1 row layout `csr_idx1` / `split1`), both bitmaps cleared in two vector sweeps.
// A table that fills up sets p.fail: the host then recomputes the whole range with the window kernel (eligibility is
// estimated per handle from the per-column work, so this is the rare case).
// Exactness is the window kernel's: 64-bit keys (similarity bits, ~original index), block_select for the K best.

constexpr int CB_LOG2 = 12;  // coarse norm tile: 4096 neighbours

template <int F>
__global__ void __launch_bounds__(THREADS, 1) sim_k1b_kernel(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Shared sh;
  __shared__ int s_ntab, s_fail, s_tstop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = 1 << p.tbits;
  const int tab_limit = (T >> 1) + (T >> 3);  // 62.5 % load
  unsigned* bm1 = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* bm2 = bm1 + p.bm_words;
  int* keys = reinterpret_cast<int*>(bm2 + p.bm_words);
  int* cnts = keys + T;
  u64* buf = reinterpret_cast<u64*>(cnts + T);
  int* hist = reinterpret_cast<int*>(buf);
  int* stage = reinterpret_cast<int*>(buf + p.cap_alloc);
  float* cbs = reinterpret_cast<float*>(stage + STAGE_INTS);  // [ncb + 1] norm term at the coarse tile boundaries
  float* tile_f = cbs + p.ncb + 1;                            // [ncb] per-column scratch: lower-bound scales, then dot thresholds
  for (int i = tid; i < 2 * p.bm_words; i += THREADS) bm1[i] = 0u;
  for (int i = tid; i < T; i += THREADS) { keys[i] = -1; cnts[i] = 0; }
  for (int i = tid; i <= p.ncb; i += THREADS) cbs[i] = p.cb[i];
  __syncthreads();

  const int lpu = 1 << p.lpu_log2, upw = 32 >> p.lpu_log2;
  const int sub = lane & (lpu - 1), uslot = lane >> p.lpu_log2;
  const int colchunk = STAGE_INTS / 2;
  const int target = p.K;

  while (true) {
    if (tid == 0) sh.col = atomicAdd(p.counter, 1);
    __syncthreads();
    const int c = sh.col;
    if (c >= p.n_range) break;
    const int lc = p.order ? p.order[c] : c;
    const int col = p.old2new[p.col_begin + lc];
    const size_t out_base = (size_t)lc * p.K;
    const int cs = p.csc_ptr[col], ce = p.csc_ptr[col + 1];
    const float Ai = p.A[col];
    long long prof_t = p.prof ? clock64() : 0;
    if (tid == 0) { s_ntab = 0; s_fail = 0; s_tstop = p.ncb; sh.nbuf = 0; }

    // ---------------- stage + accumulate, one chunk of the column's users at a time
    for (int k0 = cs; k0 < ce; k0 += colchunk) {
      const int n = min(colchunk, ce - k0);
      __syncthreads();
      for (int t = tid; t < n; t += THREADS) {
        const int u = p.csc_idx[k0 + t];
        stage[2 * t] = p.split1[2 * (size_t)u];
        stage[2 * t + 1] = p.split1[2 * (size_t)u + 1];
      }
      __syncthreads();
      PROF_MARK(0);
      for (int t0 = warp * upw; t0 < n; t0 += NWARPS * upw * UB) {
        int s[UB], e[UB];
        int mych = 0;
#pragma unroll
        for (int k = 0; k < UB; ++k) {
          const int t = t0 + k * NWARPS * upw + uslot;
          s[k] = t < n ? stage[2 * t] : 0;
          e[k] = t < n ? stage[2 * t + 1] : 0;
          mych = max(mych, (e[k] - s[k]) >> 2);  // rows are padded to whole 16-byte chunks
        }
        const int maxch = __reduce_max_sync(0xffffffffu, mych);
        for (int c0 = 0; c0 < maxch; c0 += lpu) {
          const int ch = c0 + sub;
          int4 v[UB];
#pragma unroll
          for (int k = 0; k < UB; ++k) {
            const int g = s[k] + ch * 4;
            v[k] = make_int4(-1, -1, -1, -1);
            if (g < e[k]) v[k] = __ldg(reinterpret_cast<const int4*>(p.csr_idx1 + g));
          }
#pragma unroll
          for (int k = 0; k < UB; ++k) {
            const int jj[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = jj[q];
              if (j < 0 || j >= p.n_cols || j == col) continue;  // -1: no chunk; n_cols..: row padding; col: the diagonal (pyx:396)
              const unsigned bit = 1u << (j & 31);
              const unsigned old = atomicOr(&bm1[j >> 5], bit);
              if (old & bit) {  // hit again: count it in the table
                atomicOr(&bm2[j >> 5], bit);
                unsigned h = ((unsigned)j * 2654435761u) >> (32 - p.tbits);
                for (int probe = 0;; ++probe) {
                  const int k2 = atomicCAS(&keys[h], -1, j);
                  if (k2 == -1) {
                    if (atomicAdd(&s_ntab, 1) >= tab_limit) s_fail = 1;
                    atomicAdd(&cnts[h], 1);
                    break;
                  }
                  if (k2 == j) { atomicAdd(&cnts[h], 1); break; }
                  h = (h + 1) & (unsigned)(T - 1);
                  if (probe > 256) { s_fail = 1; break; }
                }
              }
            }
          }
        }
      }
    }
    __syncthreads();
    PROF_MARK(1);
    const int ntab = s_ntab;
    if (s_fail) {
      // the table ran out of room: this launch's results are discarded by the host; leave clean state behind
      if (tid == 0) { atomicExch(p.fail, 1); p.out_cnt[lc] = 0; }
      for (int i = tid; i < 2 * p.bm_words; i += THREADS) bm1[i] = 0u;
      for (int i = tid; i < T; i += THREADS) { keys[i] = -1; cnts[i] = 0; }
      __syncthreads();
      continue;
    }

    // ---------------- bootstrap over the table: floor of the target-th best similarity from per-slot lower bounds
    u64 thr = 0;
    if (ntab > 2 * target) {
      for (int i = tid; i < HBINS; i += THREADS) hist[i] = 0;
      if (tid == 0) sh.b0 = -1;
      for (int t = tid; t < p.ncb; t += THREADS) tile_f[t] = lower_bound_scale<F>(p, Ai, cbs[t], cbs[t + 1]);
      __syncthreads();
      for (int sidx = tid; sidx < T; sidx += THREADS) {
        const int k2 = keys[sidx];
        if (k2 >= 0) {
          const float lb = (float)(cnts[sidx] + 1) * tile_f[k2 >> CB_LOG2];
          if (lb > 0.f) atomicAdd(&hist[lb_bin(lb)], 1);
        }
      }
      __syncthreads();
      int hh[HBINS / THREADS], local = 0;
#pragma unroll
      for (int b = 0; b < HBINS / THREADS; ++b) { hh[b] = hist[tid * (HBINS / THREADS) + b]; local += hh[b]; }
      int cum = block_suffix_excl(local, sh.warp_tot);
#pragma unroll
      for (int b = HBINS / THREADS - 1; b >= 0; --b) {
        if (cum < target && cum + hh[b] >= target) sh.b0 = tid * (HBINS / THREADS) + b;
        cum += hh[b];
      }
      __syncthreads();
      const int b0 = sh.b0;
      if (b0 > 0) thr = ((u64)lb_bin_floor_bits(b0)) << 32;
      __syncthreads();  // hist (aliasing buf) fully consumed before candidates are pushed
    }
    PROF_MARK(2);

    // ---------------- table pass: evaluate the slots that can reach the floor, clear every slot on the way
    for (int t = tid; t < p.ncb; t += THREADS)
      tile_f[t] = thr ? dot_threshold<F>(p, __uint_as_float((unsigned)(thr >> 32)), Ai, cbs[t], cbs[t + 1]) : 0.f;
    int ub = 0;  // block-uniform upper bound of sh.nbuf (a step pushes at most THREADS keys)
    __syncthreads();
    for (int s0 = 0; s0 < T; s0 += THREADS) {
      if (ub + THREADS > p.cap) {
        ub = sh.nbuf;
        __syncthreads();
        if (ub + THREADS > p.cap) {
          int kept;
          const u64 t2 = block_select(buf, ub, target, &sh, stage, &kept);
          if (tid == 0) sh.nbuf = kept;
          ub = kept;
          if (t2 > thr) {
            thr = t2;
            __syncthreads();
            for (int t = tid; t < p.ncb; t += THREADS)
              tile_f[t] = dot_threshold<F>(p, __uint_as_float((unsigned)(thr >> 32)), Ai, cbs[t], cbs[t + 1]);
          }
          __syncthreads();
        }
      }
      ub += THREADS;
      const int sidx = s0 + tid;
      const int k2 = sidx < T ? keys[sidx] : -1;  // T < THREADS only under the debug hook
      if (k2 >= 0) {
        const float d = (float)(cnts[sidx] + 1);
        keys[sidx] = -1;
        cnts[sidx] = 0;
        if (d >= tile_f[k2 >> CB_LOG2]) {
          const int2 bn = __ldg(p.BN + k2);
          const float sv = sim_value<F>(p, d, Ai, __int_as_float(bn.x));
          const u64 key = (((u64)__float_as_uint(sv)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bn.y);
          if (sv > 0.f && key >= thr) buf[atomicAdd(&sh.nbuf, 1)] = key;
        }
      }
      __syncthreads();
    }
    {  // an exact floor (the target-th best key so far) decides how much of the count-1 set matters
      int kept;
      const u64 t2 = block_select(buf, sh.nbuf, target, &sh, stage, &kept);
      if (tid == 0) sh.nbuf = kept;
      thr = max(thr, t2);
      __syncthreads();
    }
    PROF_MARK(3);

    // ---------------- count-1 cells: only the leading coarse tiles whose best count-1 similarity reaches the floor
    {
      const float tsim = __uint_as_float((unsigned)(thr >> 32));
      for (int t = tid; t < p.ncb; t += THREADS) {
        // cbs[t] is the smallest norm term of tile t and every formula decreases with it: the tile's best case
        const float best = sim_value<F>(p, 1.f, Ai, cbs[t]);
        if (thr != 0 && !(best >= tsim)) atomicMin(&s_tstop, t);
      }
      __syncthreads();
      const int j_end = min(p.n_cols, s_tstop << CB_LOG2);
      const int w_end = (j_end + 31) >> 5;
      ub = sh.nbuf;
      __syncthreads();
      for (int w0 = 0; w0 < w_end; w0 += THREADS) {
        const int w = w0 + tid;
        unsigned m = w < w_end ? (bm1[w] & ~bm2[w]) : 0u;
        while (__syncthreads_or(m != 0u)) {
          if (ub + THREADS > p.cap) {
            ub = sh.nbuf;
            __syncthreads();
            if (ub + THREADS > p.cap) {
              int kept;
              thr = max(thr, block_select(buf, ub, target, &sh, stage, &kept));
              if (tid == 0) sh.nbuf = kept;
              ub = kept;
              __syncthreads();
            }
          }
          ub += THREADS;
          if (m) {
            const int j = w * 32 + __ffs(m) - 1;
            m &= m - 1;
            if (j < p.n_cols) {
              const int2 bn = __ldg(p.BN + j);
              const float sv = sim_value<F>(p, 1.f, Ai, __int_as_float(bn.x));
              const u64 key = (((u64)__float_as_uint(sv)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bn.y);
              if (sv > 0.f && key >= thr) buf[atomicAdd(&sh.nbuf, 1)] = key;
            }
          }
        }
      }
      __syncthreads();
    }
    PROF_MARK(4);

    // ---------------- the K best, emit, clear the bitmaps
    int nbuf;
    {
      int kept;
      thr = max(thr, block_select(buf, sh.nbuf, target, &sh, stage, &kept));
      nbuf = kept;
    }
    PROF_MARK(5);
    for (int t = tid; t < nbuf; t += THREADS) {
      const u64 k64 = buf[t];
      p.out_idx[out_base + t] = (int)(0xFFFFFFFFu - (unsigned)k64);
      p.out_val[out_base + t] = __uint_as_float((unsigned)(k64 >> 32));
    }
    for (int t = nbuf + tid; t < p.K; t += THREADS) {
      p.out_idx[out_base + t] = -1;
      p.out_val[out_base + t] = 0.f;
    }
    if (tid == 0) p.out_cnt[lc] = nbuf;
    for (int i = tid; i < (2 * p.bm_words) >> 2; i += THREADS) reinterpret_cast<int4*>(bm1)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    PROF_MARK(6);
  }
}

// cb[t] = norm term at neighbour min(t << CB_LOG2, n_cols - 1), t = 0 .. ncb
__global__ void coarse_bounds_kernel(const int2* __restrict__ BN, int n_cols, int ncb, float* cb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ncb) return;
  cb[t] = __int_as_float(BN[min(t << CB_LOG2, n_cols - 1)].x);
}
