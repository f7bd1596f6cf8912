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


class SymmetricTopKTable:
    """The full [n_columns, K] idx / val + [n_columns] cnt table of an item-sharded similarity run, allocated once per rank as
    symmetric memory (torch.distributed._symmetric_memory: every rank's copy is mapped into every other rank's address space
    over NVLink / NVSwitch).  `fill(sim, lo, hi)` runs the similarity kernel on this rank's columns with EVERY rank's copy as
    an output table: the CTA that finishes a column stores its row into all the tables, so the kernel itself is the all-gather
    (b200_sim_compute_peers_device) and the only collective left is the barrier that says every rank's kernel has finished."""

    MAX_WORLD = 8  # output tables one kernel launch can address (include/b200rec.h)

    def __init__(self, n_columns, K, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > self.MAX_WORLD:
            raise ValueError("SymmetricTopKTable: at most %d ranks" % self.MAX_WORLD)
        self.n, self.K = int(n_columns), int(K)
        self.off_idx, self.off_val, self.off_cnt = 0, self.n * self.K, 2 * self.n * self.K
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((2 * self.n * self.K + self.n,), dtype=torch.int32, device=dev)
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.tables = [ptrs[self.rank]] + [ptrs[r] for r in range(self.world) if r != self.rank]  # the local copy first
        self.idx = self.buf[:self.off_val].view(self.n, self.K)
        self.val = self.buf[self.off_val:self.off_cnt].view(torch.float32).view(self.n, self.K)
        self.cnt = self.buf[self.off_cnt:]

    def fill(self, sim, lo, hi):
        """Every rank calls this with its own column range; returns once the launches are enqueued on the current stream
        (the closing barrier is stream-ordered too): afterwards idx / val / cnt hold every rank's rows."""
        import ctypes
        import torch
        from . import _lib
        self.hdl.barrier(channel=0)  # nobody is still reading the previous fill of its copy
        arr = (ctypes.c_void_p * len(self.tables))(*self.tables)
        _lib.check(_lib.load().b200_sim_compute_peers_device(sim._h, int(lo), int(hi), len(self.tables), arr, self.off_idx, self.off_val,
                                                             self.off_cnt, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.hdl.barrier(channel=1)  # every rank's kernel (and with it its stores into this copy) has finished
        return self.idx, self.val, self.cnt


def compute_similarity_sharded(sim, group=None, assemble=True, table=None):
    """Item-sharded compute_similarity: every rank holds the same `sim` (Compute_Similarity_Cython built from the
    replicated URM).  Returns the scipy CSR W (on every rank) or, with assemble=False, the full device table.
    With a SymmetricTopKTable the kernel writes every rank's copy directly (no collective); without one the ranks' slabs are
    exchanged with NCCL all-gathers."""
    import torch.distributed as dist
    from .similarity import topk_table_to_csr
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = balanced_ranges(sim.column_work(), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if table is not None:
        g_idx, g_val, g_cnt = table.fill(sim, lo, hi)
    else:
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


# ---------------------------------------------------------------------------------------------------------------------
# K5: EASE_R with the Gram sharded over the users: G = X^T X = sum over user shards of X_r^T X_r.  Every
# rank runs the dense mode of the similarity kernel on its own rows (equal interaction mass), one all-reduce of the n_items^2
# fp32 matrix (1.25 GB at C4) gives every rank the full Gram, and every rank inverts it (the dense inverse does not shard at
# this size: replicas).

def make_sharded_ease(group=None):
    """Returns a subclass of recommenders.EASE_R_Recommender whose Gram is accumulated over the ranks of `group`."""
    import torch.distributed as dist
    from .recommenders import EASE_R_Recommender

    class ShardedEASE_R_Recommender(EASE_R_Recommender):
        RECOMMENDER_NAME = "EASE_R_Recommender"

        def _gram_device(self, rows=None):
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            b = balanced_ranges(np.diff(self.URM_train.indptr), world)
            G = EASE_R_Recommender._gram_device(self, rows=(int(b[rank]), int(b[rank + 1])))
            dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
            return G

    return ShardedEASE_R_Recommender


# ---------------------------------------------------------------------------------------------------------------------
# K3 model parallelism: SLIM-BPR with S sharded by columns.  At 200 K items the dense S is 160 GB: no
# single GPU holds it (the reference trains such catalogues in its tree-sparse mode, SLIM_BPR_Cython_Epoch.pyx:509-1031).
# Rank g owns S[:, cols_g]; every rank walks the SAME sample stream (counter-based Philox, common seed); per batch each rank
# sums its own cells into a partial x_uij per sample, ONE all-reduce of the [batch] vector adds the partials, and every
# rank updates the cells it owns (csrc/slim_bpr.cu slim_shard_*_kernel).  batch_size = 1 is the reference's recursion.

class ShardedSLIM_BPR:
    """One rank of a column-sharded SLIM-BPR run.  Mirrors the epoch API of SLIM_BPR_Cython_Epoch (epochIteration_Cython,
    get_S); the full, non-symmetric S is trained (the triangular storage of symmetric=True does not shard by columns)."""

    _MODE = {"sgd": 0, "adagrad": 1, "rmsprop": 2, "adam": 3}

    def __init__(self, URM_mask, group=None, batch_size=8192, learning_rate=0.01, li_reg=0.0, lj_reg=0.0, topK=150, symmetric=False,
                 random_seed=None, sgd_mode="adam", gamma=0.995, beta_1=0.9, beta_2=0.999, col_range=None, world_rank=None):
        import ctypes
        import scipy.sparse as sps
        import torch
        from . import _lib
        if symmetric:
            raise NotImplementedError("ShardedSLIM_BPR trains the full S; symmetric=True (triangular storage) is single-GPU only")
        if random_seed is None:
            raise ValueError("ShardedSLIM_BPR: random_seed is required (every rank must draw the same sample stream)")
        if sgd_mode not in self._MODE:
            raise ValueError("SLIM_BPR_Cython_Epoch: sgd_mode '{}' not recognized".format(sgd_mode))
        self._lib = _lib.load()
        self.group = group
        if world_rank is None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:  # explicit placement (tests run several shards in one process and add the partial sums themselves)
            self.world, self.rank = world_rank
        X = sps.csr_matrix(URM_mask, dtype=np.float32)
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        self.n_users, self.n_items = X.shape
        self.topK = min(int(topK), self.n_items)
        self.batch_size = int(batch_size)
        if col_range is None:
            b = np.linspace(0, self.n_items, self.world + 1).astype(np.int64)
            col_range = (int(b[self.rank]), int(b[self.rank + 1]))
        self.lo, self.hi = col_range
        self._h = ctypes.c_void_p()
        indptr = np.ascontiguousarray(X.indptr, np.int32)
        indices = np.ascontiguousarray(X.indices, np.int32)
        _lib.check(self._lib.b200_slim_create_sharded(
            ctypes.byref(self._h), self.n_users, self.n_items, X.nnz, _lib.ptr(indptr), _lib.ptr(indices), float(learning_rate),
            float(li_reg), float(lj_reg), self._MODE[sgd_mode], float(gamma), float(beta_1), float(beta_2),
            int(random_seed) & 0xFFFFFFFF, self.lo, self.hi))
        self._x = torch.empty(max(1, min(self.batch_size, self.n_users)), dtype=torch.float32, device="cuda")

    def _stream(self):
        import ctypes
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def partial(self, first, n):
        from . import _lib
        _lib.check(self._lib.b200_slim_shard_partial_device(self._h, int(first), int(n), self._x.data_ptr(), self._stream()))
        return self._x[:n]

    def apply(self, first, n, x_sum):
        from . import _lib
        _lib.check(self._lib.b200_slim_shard_apply_device(self._h, int(first), int(n), x_sum.data_ptr(), self._stream()))

    def epochIteration_Cython(self):
        """n_users samples (pyx:231) in batches; returns the number of samples."""
        import torch.distributed as dist
        for first in range(0, self.n_users, self.batch_size):
            n = min(self.batch_size, self.n_users - first)
            x = self.partial(first, n)
            if self.world > 1:
                dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
            self.apply(first, n, x)
        return self.n_users

    def slab(self):
        """This rank's [n_items, hi - lo] columns of S as a torch CUDA tensor aliasing the trainer's memory."""
        import ctypes
        import torch
        from . import _lib
        ptr = ctypes.c_void_p()
        _lib.check(self._lib.b200_slim_shard_device(self._h, ctypes.byref(ptr), None, None))

        class _Arr:
            def __init__(s, p, shape):
                s.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (p, False), "version": 3, "strides": None}

        return torch.as_tensor(_Arr(ptr.value, (self.n_items, self.hi - self.lo)), device=torch.device("cuda", torch.cuda.current_device()))

    def local_row_topk(self):
        """[n_items, K] idx (global column) / val candidates of every row from this rank's columns (K largest non-zero)."""
        import torch
        from . import _lib
        K, w = self.topK, self.hi - self.lo
        dev = torch.device("cuda", torch.cuda.current_device())
        idx = torch.empty((self.n_items, K), dtype=torch.int32, device=dev)
        val = torch.empty((self.n_items, K), dtype=torch.float32, device=dev)
        cnt = torch.empty((self.n_items,), dtype=torch.int32, device=dev)
        _lib.check(self._lib.b200_dense_topk_rect_device(self.slab().data_ptr(), self.n_items, w, w, 1, self.lo, K, 0, idx.data_ptr(),
                                                         val.data_ptr(), cnt.data_ptr(), self._stream()))
        return idx, val

    @staticmethod
    def merge_row_topk(cands, n_items, K):
        """cands: list of (idx, val) [n_items, K] tables of the shards -> scipy CSR with the K largest non-zero cells per row
        (the dense branch of get_S: similarityMatrixTopK(S.T).T, pyx:371,386)."""
        import scipy.sparse as sps
        import torch
        from . import _lib
        from .similarity import topk_table_to_csr
        idx = torch.cat([c[0] for c in cands], dim=1).contiguous()
        val = torch.cat([c[1] for c in cands], dim=1).contiguous()
        m = idx.shape[1]
        ptr = (torch.arange(n_items + 1, dtype=torch.int64, device=idx.device) * m).to(torch.int32)
        oi = torch.empty((n_items, K), dtype=torch.int32, device=idx.device)
        ov = torch.empty((n_items, K), dtype=torch.float32, device=idx.device)
        oc = torch.empty((n_items,), dtype=torch.int32, device=idx.device)
        import ctypes
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().b200_sparse_topk_device(n_items, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), K, 0, oi.data_ptr(),
                                                       ov.data_ptr(), oc.data_ptr(), st))
        T = topk_table_to_csr(n_items, K, oi, ov, oc)  # T[column, row]: the transpose
        return sps.csc_matrix((T.data, T.indices, T.indptr), shape=(n_items, n_items)).tocsr()

    def get_S(self):
        """Row top-K of the full S on every rank: local candidates, one all-gather of the [n_items, K] tables, merge."""
        import torch
        import torch.distributed as dist
        idx, val = self.local_row_topk()
        if self.world == 1:
            return self.merge_row_topk([(idx, val)], self.n_items, self.topK)
        gi = [torch.empty_like(idx) for _ in range(self.world)]
        gv = [torch.empty_like(val) for _ in range(self.world)]
        dist.all_gather(gi, idx, group=self.group)
        dist.all_gather(gv, val, group=self.group)
        return self.merge_row_topk(list(zip(gi, gv)), self.n_items, self.topK)

    def _dealloc(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_slim_destroy(self._h)
            import ctypes
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self._dealloc()
        except Exception:
            pass
