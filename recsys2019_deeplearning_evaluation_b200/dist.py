"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

K1 shards the item axis: the URM is replicated, rank r computes the top-K rows of a contiguous
column range chosen so that the ranges carry equal gathered-entry work, and one all-gather of the fixed-shape
[cols_per_rank_max, K] idx/val slabs (+ counts) gives every rank the full table.  The reference has no
distributed path; its hook is compute_similarity(start_col, end_col) (Compute_Similarity_Cython.pyx:413,450-454).
"""
import numpy as np


def balanced_ranges(work, world_size):
    """Contiguous ranges [lo_r, hi_r) covering [0, len(work)) with (nearly) equal sums of `work`.
    Returns an int64 array of world_size + 1 boundaries.  Pure host logic (tested on CPU)."""
    work = np.asarray(work, dtype=np.float64)
    n = len(work)
    bounds = np.zeros(world_size + 1, np.int64)
    bounds[-1] = n
    if n == 0 or world_size == 1:
        return bounds
    csum = np.cumsum(work + 1e-9)  # strictly increasing so that empty columns still get spread
    total = csum[-1]
    for r in range(1, world_size):
        bounds[r] = int(np.searchsorted(csum, total * r / world_size, side="left")) + 1
    bounds[1:-1] = np.clip(bounds[1:-1], 0, n)
    bounds = np.maximum.accumulate(bounds)
    return bounds


def allgather_topk_tables(idx, val, cnt, bounds, group=None):
    """idx/val: [n_local, K] torch tensors of this rank's range, cnt: [n_local].  Every rank contributes a slab
    padded to the largest range; returns the full [n_cols, K] idx/val and [n_cols] cnt on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = (bounds[1:] - bounds[:-1]).astype(np.int64)
    m = int(sizes.max())
    K = idx.shape[1]
    n_local = int(sizes[rank])

    def pad(t, fill):
        out = torch.full((m,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        if n_local:
            out[:n_local] = t[:n_local]
        return out

    send_idx, send_val, send_cnt = pad(idx, -1), pad(val, 0), pad(cnt, 0)
    g_idx = torch.empty((world * m, K), dtype=idx.dtype, device=idx.device)
    g_val = torch.empty((world * m, K), dtype=val.dtype, device=val.device)
    g_cnt = torch.empty((world * m,), dtype=cnt.dtype, device=cnt.device)
    dist.all_gather_into_tensor(g_idx, send_idx, group=group)
    dist.all_gather_into_tensor(g_val, send_val, group=group)
    dist.all_gather_into_tensor(g_cnt, send_cnt, group=group)
    if all(int(s) == m for s in sizes):
        return g_idx, g_val, g_cnt
    keep = torch.cat([torch.arange(r * m, r * m + int(sizes[r]), device=idx.device) for r in range(world)])
    return g_idx[keep], g_val[keep], g_cnt[keep]


def compute_similarity_sharded(sim, group=None, assemble=True):
    """Item-sharded compute_similarity: every rank holds the same `sim` (Compute_Similarity_Cython built from the
    replicated URM).  Returns the scipy CSR W (on every rank) or, with assemble=False, the gathered device table."""
    import torch.distributed as dist
    from .similarity import topk_table_to_csr
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = balanced_ranges(sim.column_work(), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    tab = sim.compute_topk_device(lo, hi)
    g_idx, g_val, g_cnt = allgather_topk_tables(tab.idx, tab.val, tab.cnt, bounds, group)
    if not assemble:
        return g_idx, g_val, g_cnt
    return topk_table_to_csr(sim.n_columns, sim.K, g_idx.contiguous(), g_val.contiguous(), g_cnt.contiguous())


# ---------------------------------------------------------------------------------------------------------------------
# K2 data parallelism: user-sharded Hogwild BPR-MF with a replicated item-factor table.
# Rank r draws its samples from users [lo_r, hi_r) only, so user rows are private.  The item factors are replicated; what
# the ranks exchange is each rank's own MOVEMENT of the table since its last snapshot:
#     snapshot k:  own_k = V - B,  B = V          (fused pass, csrc/mf_sgd.cu mf_delta_snapshot_kernel)
#     all-reduce:  sum_k = sum_r own_k            (NCCL, asynchronous, on a side stream)
#     apply k:     V += sum_k - own_k,  B += ...  (fused pass with RED.ADD: the trainer keeps writing V meanwhile)
# The apply of exchange k runs when exchange k + 1 starts, i.e. one epoch later: the collective is hidden behind the next
# epoch's training kernel and every update of every rank reaches every replica exactly once (bounded staleness: one
# epoch -- the Hogwild reading of "all updates are applied"; flush() drains the pipeline).

def sync_replicated_delta(V, V_prev, group=None):
    """Blocking form: V <- V_prev + all_reduce_sum(V - V_prev); V_prev <- V.  CPU tensors (gloo) and CUDA (NCCL)."""
    import torch.distributed as dist
    delta = V - V_prev
    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
    V.copy_(V_prev + delta)
    V_prev.copy_(V)
    return V


class ReplicatedDeltaExchange:
    """Overlapped exchange of one replicated fp32 table (see the protocol above).  `exchange()` is called after a
    training epoch has been enqueued on the current stream; it never blocks the host or the training stream."""

    def __init__(self, V, group=None):
        import torch
        self.V, self.group = V, group
        self.B = V.clone()
        self.own = torch.zeros_like(V)
        self.sum = torch.zeros_like(V)
        self.work = None
        self.cuda = V.is_cuda
        self.comm = torch.cuda.Stream(device=V.device) if self.cuda else None
        assert V.numel() % 4 == 0 or not self.cuda, "the fused exchange kernels move float4 elements"

    def _snapshot(self):
        if self.cuda:
            import ctypes
            import torch
            from . import _lib
            _lib.check(_lib.load().b200_mf_delta_snapshot_device(self.V.data_ptr(), self.B.data_ptr(), self.own.data_ptr(), self.sum.data_ptr(),
                                                                 self.V.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:  # gloo tests of the protocol (host logic only)
            self.own.copy_(self.V - self.B)
            self.sum.copy_(self.own)
            self.B.copy_(self.V)

    def _apply(self):
        if self.cuda:
            import ctypes
            import torch
            from . import _lib
            _lib.check(_lib.load().b200_mf_delta_apply_device(self.V.data_ptr(), self.B.data_ptr(), self.sum.data_ptr(), self.own.data_ptr(),
                                                              self.V.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            t = self.sum - self.own
            self.V.add_(t)
            self.B.add_(t)

    def _step(self):
        import torch.distributed as dist
        if self.work is not None:
            self.work.wait()  # CUDA: the current (side) stream waits for the collective; CPU: blocks
            self._apply()
        self._snapshot()
        self.work = dist.all_reduce(self.sum, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def exchange(self):
        if not self.cuda:
            self._step()
            return
        import torch
        ev = torch.cuda.Event()
        ev.record()  # the epoch just enqueued on the training stream
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ev)
            self._step()

    def flush(self):
        """Drains the pipeline: afterwards V holds every rank's movement up to the last exchange()."""
        if self.work is None:
            return
        if not self.cuda:
            self.work.wait()
            self._apply()
            self.work = None
            return
        import torch
        with torch.cuda.stream(self.comm):
            self.work.wait()
            self._apply()
        self.work = None
        torch.cuda.current_stream().wait_stream(self.comm)


class ShardedBPR:
    """One rank of a user-sharded Hogwild BPR-MF run (MatrixFactorization_Cython_Epoch semantics per sample).
    scaling="strong": the reference's epoch ((n_users / bs + 1) * bs samples) is split over the ranks;
    scaling="weak": every rank draws a full epoch's worth of samples from its own user shard per step."""

    def __init__(self, URM, group=None, scaling="strong", overlap=True, **mf_kwargs):
        import torch.distributed as dist
        from .mf_epoch import MatrixFactorization_Cython_Epoch
        if mf_kwargs.get("random_seed") is None:
            # every rank draws its initial factors from its own numpy RNG: without a common seed the replicas of the item
            # table start different and only deltas are exchanged, so they would never agree
            raise ValueError("ShardedBPR: random_seed is required (the ranks must start from identical factors)")
        self.group, self.overlap = group, overlap
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n_users = URM.shape[0]
        bounds = balanced_ranges(np.diff(URM.indptr), self.world)  # equal interaction mass per rank
        self.lo, self.hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        mf_kwargs.setdefault("algorithm_name", "MF_BPR")
        mf_kwargs.update(sampler="philox", hogwild=True)
        self.epoch_obj = MatrixFactorization_Cython_Epoch(URM, **mf_kwargs)  # same seed -> identical initial factors on every rank
        bs = self.epoch_obj.batch_size
        total = (n_users // bs + 1) * bs  # the reference's epoch length
        self.samples_per_rank = total if scaling == "weak" else max(bs, (total // self.world // bs) * bs)
        self.epoch_obj.set_user_shard(self.lo, self.hi, self.samples_per_rank, stream_id=self.rank)
        self.U, self.V = self.epoch_obj.device_factors()
        self.U0 = self.U.clone()
        self.xchg = ReplicatedDeltaExchange(self.V.view(-1), group)

    def epoch(self):
        self.epoch_obj.epochIteration_Cython()
        self.xchg.exchange()
        if not self.overlap:
            self.xchg.flush()
        return self.samples_per_rank * self.world

    def flush(self):
        self.xchg.flush()

    def gather_user_factors(self):
        """Every rank ends with the full user-factor table (rows outside a rank's shard never moved there)."""
        import torch.distributed as dist
        delta = self.U - self.U0
        dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=self.group)
        return self.U0 + delta


# ---------------------------------------------------------------------------------------------------------------------
# K4 row parallelism: implicit ALS with the rows of each half epoch sharded over the ranks.
# Every row solve is independent given the other side's factors, so rank r solves a contiguous slice of the warm rows
# (balanced by profile length), the other side's factor table is replicated, and one all-reduce per half epoch of a
# buffer that holds every rank's freshly solved rows (zeros elsewhere: the sum adds nothing to a solved row)
# gives every rank the full updated table.  Y^T Y is recomputed from the replicated table on every rank.

def sync_owned_rows(X, rows, lo, hi, group=None):
    """In place: X[rows[k]] for k in [lo, hi) are this rank's new values; afterwards X[rows] holds every rank's.
    `rows` is an integer tensor (the warm rows, same on every rank).  Works with gloo (CPU tensors) and NCCL."""
    import torch
    import torch.distributed as dist
    Z = torch.zeros_like(X)
    mine = rows[lo:hi].long()
    if hi > lo:
        Z[mine] = X[mine]
    dist.all_reduce(Z, op=dist.ReduceOp.SUM, group=group)
    X[rows.long()] = Z[rows.long()]
    return X


def make_sharded_ials(group=None):
    """Returns a subclass of recommenders.IALSRecommender whose _run_epoch shards the row solves over the ranks of
    `group`.  Every rank must seed numpy identically before fit() (the initial item factors come from np.random,
    MatrixFactorization/IALSRecommender.py:204-210) and ends every epoch with the full factor tables."""
    import torch.distributed as dist
    from .recommenders import IALSRecommender

    class ShardedIALSRecommender(IALSRecommender):
        RECOMMENDER_NAME = "IALSRecommender"

        def _slice(self, rows, csr):
            key = rows.data_ptr()
            cache = self.__dict__.setdefault("_shard_cache", {})
            if key not in cache:
                ptr = csr[0].cpu().numpy()
                r = rows.cpu().numpy()
                b = balanced_ranges((ptr[r + 1] - ptr[r]).astype(np.float64) + self.num_factors / 8.0, dist.get_world_size(group))
                rank = dist.get_rank(group)
                cache[key] = (int(b[rank]), int(b[rank + 1]))
            return cache[key]

        def _half_sharded(self, rows, csr, Y, X):
            lo, hi = self._slice(rows, csr)
            if hi > lo:
                self._half(rows[lo:hi].contiguous(), csr, Y, X)
            sync_owned_rows(X, rows, lo, hi, group)

        def _run_epoch(self, num_epoch):
            self._half_sharded(self._d_warm_users, self._d_C, self._d_V, self._d_U)
            self._half_sharded(self._d_warm_items, self._d_Ct, self._d_U, self._d_V)

    return ShardedIALSRecommender
