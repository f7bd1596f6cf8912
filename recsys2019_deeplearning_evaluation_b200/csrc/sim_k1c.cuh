// K1-C: the binary similarity kernel for large catalogues with sparse co-occurrence counts (included by sim_topk.cu inside
// namespace b200::sim).  Replaces Compute_Similarity_Cython.pyx:327-408 (gather / accumulate) and :467-568 (normalise,
// top-K, emit) for every-stored-value-is-1 data, like the window kernel, but with a different on-chip representation.
//
// Why.  At 200 K columns the window kernel (sim_topk_kernel) needs two passes over a 16-bit counter window per target
// column and sweeps every one of the 200 K cells twice (bootstrap + scan): ncu attributes 43 % of its time to those
// shared-memory sweeps and most of the rest to the latency of the row gather (LDG -> ATOMS with two loads in flight per
// lane).  But a target column gathers ~50 K entries over 200 K neighbours: 88 % of the touched neighbours are hit once,
// and a neighbour hit c times can only reach the top-K if its norm term lies below a bound that follows from the
// current K-th best similarity -- the neighbour axis is numbered by ascending norm term, so that set is a PREFIX.
//
// Representation ("thermometer" bitmaps).  bm1 / bm2 / bm3 hold one bit per neighbour: hit at least once / twice /
// three times; the atomicOr that sets a bit returns whether it was already set and escalates to the next level; hits
// beyond the third go to a small open-addressing table (key = neighbour, value = extra hits).  3 x n_cols / 8 bytes of
// shared memory (75 KB at 200 K columns) instead of 2 x 200 KB of counters: ONE pass per column, and the levels are
// enumerated with bit scans instead of cell sweeps:
//   level >= 3  every bit of bm3 (count = 3 + table lookup) is evaluated exactly -> keys -> the K best give a floor thr;
//   level 2     bits of bm2 & ~bm3, only in the leading norm tiles whose best count-2 similarity still reaches thr;
//   level 1     bits of bm1 & ~bm2, likewise (no tile qualifies at C5 densities).
// Exactness is the window kernel's: 64-bit keys (similarity bits << 32 | ~original index), select_group for the K best.
//
// Data movement (north_star: "TMA staging of column blocks into shared memory").  Warp 31 is a producer that never
// touches the bitmaps: for each target column it pulls the column's (row start, row chunks) list with ONE bulk copy
// (csc_seg: the CSC side stores where each user's padded row lives, so there is no dependent pointer chase), then
// issues one `cp.async.bulk` (SASS UBLKCP, the TMA engine) per <= 512-byte piece of every row into a ring of shared-memory
// slots, each completing on its own mbarrier.  The 31 consumer warps wait on the slots in order, read them with
// conflict-free LDS.128 and do the atomics.  The ring (64 KB and more) is far deeper than the bandwidth-delay product
// of one SM (~18 KB), keeps filling with the NEXT column's rows while the consumers run the selection of the current
// one, and costs no registers.
//
// Columns whose deep table overflows (dense co-occurrence: popular items) are appended to a redo list and recomputed by
// the window kernel; the host routes columns whose expected count per neighbour is high there directly.

constexpr int C_CONS_WARPS = NWARPS - 1;
constexpr int C_CONS = C_CONS_WARPS * 32;   // consumer threads (warps 0 .. 30); warp 31 is the producer
constexpr int C_PIECE_CHUNKS = 32;          // 16-byte chunks per ring slot: one LDS.128 per consumer lane
constexpr int C_PIECE_BYTES = C_PIECE_CHUNKS * 16;
constexpr int C_STAGE = 1024;               // (row start, row chunks) pairs staged per bulk copy
constexpr int C_TILE_LOG2 = 10;             // norm tile of the level bounds: 1024 neighbours
constexpr int C_AHEAD = 2;                  // columns the producer may run ahead of the consumers

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void c_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "C_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra C_DONE;\n"
      "bra C_WAIT;\n"
      "C_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void c_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void c_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void c_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 2, %0;" ::"n"(C_CONS) : "memory"); }
__device__ __forceinline__ bool cons_or(bool v) {
  unsigned r;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %1, 0;\n"
      "barrier.red.or.pred p, 2, %2, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r)
      : "r"((unsigned)v), "n"(C_CONS)
      : "memory");
  return r != 0;
}
// consumer-wide select: keeps the K best keys of buf[0..n) compacted at the front
__device__ __forceinline__ u64 cons_select(u64* buf, int n, int K, Shared* sh, int* hist, int* n_out) {
  cons_bar();
  if (threadIdx.x < SELT) select_group(buf, n, K, sh, hist);
  cons_bar();
  *n_out = sh->cnt;
  return sh->sel_thr;
}

struct K1CShared {
  int4 colinfo[4];        // (new column index, local column index, csc begin, csc end) by column sequence number & 3
  volatile int cons_seq;  // columns the consumers have finished
  int ntab, fail, tstop;
};

template <int F>
__global__ void __launch_bounds__(THREADS, 1) sim_k1c_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Shared sh;
  __shared__ K1CShared cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.ring_slots;             // power of two
  const int T = 1 << p.t4bits;
  // layout: [ring S*512][meta S*16][stage 2*(C_STAGE+2)*8][bm1 bm2 bm3][tkeys T*4][tcnts T*4][buf cap*8][hist SBINS*4]
  //         [tbs (ntile+1)*4][mbarriers (2S+2)*8]
  unsigned char* sp = smem_raw;
  int4* ring = reinterpret_cast<int4*>(sp); sp += (size_t)S * C_PIECE_BYTES;
  int4* meta = reinterpret_cast<int4*>(sp); sp += (size_t)S * 16;
  int2* stagebuf = reinterpret_cast<int2*>(sp); sp += (size_t)2 * (C_STAGE + 2) * 8;
  unsigned* bm1 = reinterpret_cast<unsigned*>(sp); sp += (size_t)3 * p.bm_words * 4;
  unsigned* bm2 = bm1 + p.bm_words;
  unsigned* bm3 = bm2 + p.bm_words;
  int* tkeys = reinterpret_cast<int*>(sp); sp += (size_t)T * 4;
  int* tcnts = reinterpret_cast<int*>(sp); sp += (size_t)T * 4;
  u64* buf = reinterpret_cast<u64*>(sp); sp += (size_t)p.cap_alloc * 8;
  int* hist = reinterpret_cast<int*>(sp); sp += (size_t)SBINS * 4;
  float* tbs = reinterpret_cast<float*>(sp); sp += (size_t)(p.ntile + 1) * 4;
  sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sp) + 7) & ~(uintptr_t)7);
  const uint32_t bar_full = c_smem_u32(sp);       // full[s]  = bar_full + 8 s
  const uint32_t bar_empty = bar_full + 8u * S;   // empty[s] = bar_empty + 8 s
  const uint32_t bar_stage = bar_empty + 8u * S;  // stage[b] = bar_stage + 8 b

  for (int i = tid; i < 3 * p.bm_words; i += THREADS) bm1[i] = 0u;
  for (int i = tid; i < T; i += THREADS) { tkeys[i] = -1; tcnts[i] = 0; }
  for (int i = tid; i <= p.ntile; i += THREADS) tbs[i] = p.tbnd[i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { c_mbar_init(bar_full + 8u * s, 1); c_mbar_init(bar_empty + 8u * s, 1); }
    c_mbar_init(bar_stage, 1);
    c_mbar_init(bar_stage + 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cs.cons_seq = 0;
  }
  __syncthreads();

  if (warp == NWARPS - 1) {
    // =========================================================== producer warp
    uint32_t q = 0;  // running piece index (identical in every lane)
    int seq = 0;
    uint32_t stage_phase[2] = {0u, 0u};
    const int4* __restrict__ wl = p.worklist;
    int t_cur = 0;
    if (lane == 0) t_cur = atomicAdd(p.counter, 1);
    t_cur = __shfl_sync(0xffffffffu, t_cur, 0);
    int4 w_cur = make_int4(-1, 0, 0, 0);
    if (t_cur < p.n_range) w_cur = __ldg(wl + t_cur);
    int sb = 0;
    auto issue_stage = [&](int buf_id, int k0, int n) {  // entries [k0, k0 + n) of csc_seg, widened to 16-byte alignment
      const int a0 = k0 & ~1;
      const int nb = ((k0 + n + 1) & ~1) - a0;
      if (lane == 0) {
        const uint32_t bar = bar_stage + 8u * buf_id;
        c_mbar_expect_tx(bar, (uint32_t)nb * 8u);
        c_bulk_g2s(c_smem_u32(stagebuf + (size_t)buf_id * (C_STAGE + 2)), p.csc_seg + a0, (uint32_t)nb * 8u, bar);
      }
    };
    if (t_cur < p.n_range) issue_stage(0, w_cur.z, min(C_STAGE, w_cur.w - w_cur.z));
    while (t_cur < p.n_range) {
      // the next column's work item and first stage copy are in flight while this column's pieces are issued
      int t_next = 0;
      if (lane == 0) t_next = atomicAdd(p.counter, 1);
      t_next = __shfl_sync(0xffffffffu, t_next, 0);
      int4 w_next = make_int4(-1, 0, 0, 0);
      if (t_next < p.n_range) w_next = __ldg(wl + t_next);
      while (seq - cs.cons_seq > C_AHEAD) __nanosleep(64);
      if (lane == 0) cs.colinfo[seq & 3] = w_cur;
      __syncwarp();
      const int col = w_cur.x, c_lo = w_cur.z, c_hi = w_cur.w;
      bool next_issued = false;
      for (int k0 = c_lo; k0 < c_hi; k0 += C_STAGE) {
        const int n = min(C_STAGE, c_hi - k0);
        if (k0 > c_lo) {  // long column: later chunks reuse this buffer (every lane has finished reading it)
          __syncwarp();
          issue_stage(sb, k0, n);
        }
        if (!next_issued && k0 + C_STAGE >= c_hi) {  // last chunk of this column: the other buffer is free for the next one
          if (t_next < p.n_range) issue_stage(sb ^ 1, w_next.z, min(C_STAGE, w_next.w - w_next.z));
          next_issued = true;
        }
        c_mbar_wait(bar_stage + 8u * sb, stage_phase[sb]);
        stage_phase[sb] ^= 1u;
        const int2* st = stagebuf + (size_t)sb * (C_STAGE + 2) + (k0 & 1);
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          int2 seg = make_int2(0, 0);
          if (i < n) seg = st[i];
          const int np = (seg.y + C_PIECE_CHUNKS - 1) / C_PIECE_CHUNKS;
          int incl = np;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          const uint32_t q0 = q + (uint32_t)(incl - np);
          for (int k = 0; k < np; ++k) {
            const uint32_t qq = q0 + (uint32_t)k;
            const uint32_t slot = qq & (uint32_t)(S - 1);
            const uint32_t round = qq / (uint32_t)S;
            c_mbar_wait(bar_empty + 8u * slot, (round & 1u) ^ 1u);
            const int nch = min(C_PIECE_CHUNKS, seg.y - C_PIECE_CHUNKS * k);
            meta[slot] = make_int4(nch, seq, col, 0);
            const uint32_t full = bar_full + 8u * slot;
            c_mbar_expect_tx(full, (uint32_t)nch * 16u);
            c_bulk_g2s(c_smem_u32(ring + (size_t)slot * C_PIECE_CHUNKS),
                       p.csr_idx1 + ((size_t)seg.x + (size_t)C_PIECE_CHUNKS * k) * 4, (uint32_t)nch * 16u, full);
          }
          q += (uint32_t)total;
        }
      }
      sb ^= 1;
      ++seq;
      t_cur = t_next;
      w_cur = w_next;
    }
    // end marker: one sentinel piece per consumer warp (no data; the arrive alone completes the phase)
    if (lane < C_CONS_WARPS) {
      const uint32_t qq = q + (uint32_t)lane;
      const uint32_t slot = qq & (uint32_t)(S - 1);
      const uint32_t round = qq / (uint32_t)S;
      c_mbar_wait(bar_empty + 8u * slot, (round & 1u) ^ 1u);
      meta[slot] = make_int4(-1, seq, -1, 0);
      c_mbar_arrive(bar_full + 8u * slot);
    }
    return;
  }

  // ============================================================= consumer warps
  const int target = p.K;
  uint32_t q = (uint32_t)warp;
  long long prof_t = p.prof ? clock64() : 0;
  for (int cur = 0;; ++cur) {
    // ---------------- accumulate: every piece of column `cur`
    bool end = false;
    for (;;) {
      const uint32_t slot = q & (uint32_t)(S - 1);
      c_mbar_wait(bar_full + 8u * slot, (q / (uint32_t)S) & 1u);
      const int4 m = meta[slot];
      if (m.y != cur) break;  // a later column's piece: stays in the ring until this column's selection is done
      if (m.x < 0) { end = true; break; }
      const int col = m.z;
      if (lane < m.x) {
        const int4 v = ring[(size_t)slot * C_PIECE_CHUNKS + lane];
        const int jj[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = jj[c];
          if ((unsigned)j >= (unsigned)p.n_cols || j == col) continue;  // row padding; the diagonal (pyx:396)
          const unsigned bit = 1u << (j & 31);
          const int w = j >> 5;
          if (!(atomicOr(&bm1[w], bit) & bit)) continue;
          if (!(atomicOr(&bm2[w], bit) & bit)) continue;
          if (!(atomicOr(&bm3[w], bit) & bit)) continue;
          // fourth and later hits: the deep table
          if (cs.fail) continue;
          unsigned h = ((unsigned)j * 2654435761u) >> (32 - p.t4bits);
          for (int probe = 0;; ++probe) {
            const int k2 = atomicCAS(&tkeys[h], -1, j);
            if (k2 == -1) {
              if (atomicAdd(&cs.ntab, 1) >= (T >> 1) + (T >> 3)) cs.fail = 1;  // 62.5 % load
              atomicAdd(&tcnts[h], 1);
              break;
            }
            if (k2 == j) { atomicAdd(&tcnts[h], 1); break; }
            h = (h + 1) & (unsigned)(T - 1);
            if (probe > 128) { cs.fail = 1; break; }
          }
        }
      }
      __syncwarp();
      if (lane == 0) c_mbar_arrive(bar_empty + 8u * slot);
      q += (uint32_t)C_CONS_WARPS;
    }
    if (end) break;  // every consumer warp meets its own sentinel at the same `cur`
    cons_bar();
    PROF_MARK(1);
    const int4 info = cs.colinfo[cur & 3];
    const int col = info.x, lc = info.y;
    const size_t out_base = (size_t)lc * p.K;
    const float Ai = p.A[col];
    const int W = p.bm_words;
    const int ntab = cs.ntab;
    if (cs.fail) {
      // the deep table ran out of room: the window kernel redoes this column (host side); leave clean state behind
      cons_bar();
      if (tid == 0) { p.redo[atomicAdd(p.fail, 1)] = lc; p.out_cnt[lc] = 0; cs.ntab = 0; cs.fail = 0; sh.nbuf = 0; }
      for (int i = tid; i < (3 * W) >> 2; i += C_CONS) reinterpret_cast<int4*>(bm1)[i] = make_int4(0, 0, 0, 0);
      for (int i = tid; i < T; i += C_CONS) { tkeys[i] = -1; tcnts[i] = 0; }
      cons_bar();
      if (tid == 0) cs.cons_seq = cur + 1;
      continue;
    }
    if (tid == 0) { sh.nbuf = 0; cs.tstop = p.ntile; }
    cons_bar();

    u64 thr = 0;
    int ub = 0;  // block-uniform upper bound of sh.nbuf (a round pushes at most C_CONS keys)
    // ---------------- levels 3+, 2, 1
#pragma unroll 1
    for (int level = 3; level >= 1; --level) {
      int w_end = W;
      if (level < 3) {
        // exactly-`level` cells can reach thr only in the leading norm tiles: tbs[t] is the smallest norm term of tile t and
        // every formula decreases with it
        if (thr != 0) {
          const float tsim = __uint_as_float((unsigned)(thr >> 32));
          for (int t = tid; t < p.ntile; t += C_CONS) {
            const float best = sim_value<F>(p, (float)level, Ai, tbs[t]);
            if (!(best >= tsim)) atomicMin(&cs.tstop, t);
          }
          cons_bar();
          w_end = min(W, cs.tstop << (C_TILE_LOG2 - 5));
          cons_bar();
        }
        if (w_end == 0) continue;
      }
      const unsigned* hi = level == 3 ? bm3 : (level == 2 ? bm2 : bm1);
      const unsigned* lo = level == 3 ? nullptr : (level == 2 ? bm3 : bm2);
      for (int w0 = 0; w0 < w_end; w0 += C_CONS) {
        const int w = w0 + tid;
        unsigned m = 0u;
        if (w < w_end) m = lo ? (hi[w] & ~lo[w]) : hi[w];
        while (cons_or(m != 0u)) {
          if (ub + C_CONS > p.cap) {
            ub = sh.nbuf;
            cons_bar();
            if (ub + C_CONS > p.cap) {
              int kept;
              thr = max(thr, cons_select(buf, ub, target, &sh, hist, &kept));
              if (tid == 0) sh.nbuf = kept;
              ub = kept;
              cons_bar();
            }
          }
          ub += C_CONS;
          if (m) {
            const int j = w * 32 + __ffs(m) - 1;
            m &= m - 1;
            if (j < p.n_cols) {
              float d = (float)level;
              if (level == 3 && ntab > 0) {
                unsigned h = ((unsigned)j * 2654435761u) >> (32 - p.t4bits);
                for (;;) {
                  const int k2 = tkeys[h];
                  if (k2 == j) { d += (float)tcnts[h]; break; }
                  if (k2 == -1) break;
                  h = (h + 1) & (unsigned)(T - 1);
                }
              }
              const int2 bn = __ldg(p.BN + j);
              const float sv = sim_value<F>(p, d, Ai, __int_as_float(bn.x));
              const u64 key = (((u64)__float_as_uint(sv)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bn.y);
              if (sv > 0.f && key >= thr) buf[atomicAdd(&sh.nbuf, 1)] = key;
            }
          }
        }
      }
      // an exact floor from everything seen so far decides how much of the next level matters
      {
        cons_bar();
        const int nb = sh.nbuf;
        int kept;
        const u64 t2 = cons_select(buf, nb, target, &sh, hist, &kept);
        if (tid == 0) sh.nbuf = kept;
        ub = kept;
        thr = max(thr, t2);
        cons_bar();
      }
      if (level == 3) PROF_MARK(2); else if (level == 2) PROF_MARK(3); else PROF_MARK(4);
    }

    // ---------------- emit, clear
    const int nbuf = sh.nbuf;
    for (int t = tid; t < nbuf; t += C_CONS) {
      const u64 k64 = buf[t];
      p.out_idx[out_base + t] = (int)(0xFFFFFFFFu - (unsigned)k64);
      p.out_val[out_base + t] = __uint_as_float((unsigned)(k64 >> 32));
    }
    for (int t = nbuf + tid; t < p.K; t += C_CONS) {
      p.out_idx[out_base + t] = -1;
      p.out_val[out_base + t] = 0.f;
    }
    if (tid == 0) { p.out_cnt[lc] = nbuf; cs.ntab = 0; }
    for (int i = tid; i < (3 * W) >> 2; i += C_CONS) reinterpret_cast<int4*>(bm1)[i] = make_int4(0, 0, 0, 0);
    if (ntab > 0)
      for (int i = tid; i < T; i += C_CONS) { tkeys[i] = -1; tcnts[i] = 0; }
    cons_bar();
    if (tid == 0) cs.cons_seq = cur + 1;
    PROF_MARK(6);
  }
}

// tb[t] = norm term at neighbour min(t << C_TILE_LOG2, n_cols - 1), t = 0 .. ntile
__global__ void k1c_tile_bounds_kernel(const int2* __restrict__ BN, int n_cols, int ntile, float* tb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntile) return;
  tb[t] = __int_as_float(BN[min(t << C_TILE_LOG2, n_cols - 1)].x);
}

// csc_seg[q] = where the padded single-window row of CSC entry q's user lives: (start, length) in 16-byte chunks
__global__ void k1c_csc_seg_kernel(const int* __restrict__ csc_idx, const int* __restrict__ split1, long long nnz, int2* seg) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nnz; q += (long long)gridDim.x * blockDim.x) {
    const int u = csc_idx[q];
    const int s = split1[2 * (size_t)u], e = split1[2 * (size_t)u + 1];
    seg[q] = make_int2(s >> 2, (e - s) >> 2);
  }
}
