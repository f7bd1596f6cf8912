"""CPU: host-side logic -- synthetic generator, column partition, the multi-rank gather (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from recsys2019_deeplearning_evaluation_b200.dist import balanced_ranges
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm, CONFIGS


def test_synth_deterministic_sorted_binary():
    A = synth_urm(300, 100, 0.05, seed=3)
    B = synth_urm(300, 100, 0.05, seed=3)
    assert (A.indices == B.indices).all() and (A.indptr == B.indptr).all()
    assert A.dtype == np.float32 and (A.data == 1).all()
    for r in range(300):
        row = A.indices[A.indptr[r]:A.indptr[r + 1]]
        assert (np.diff(row) > 0).all()
    Z = synth_urm(2000, 300, 0.02, seed=1, values="ratings", popularity=1.0)
    cnt = np.bincount(Z.indices, minlength=300)
    assert cnt[:10].sum() > 5 * cnt[-10:].sum()
    assert set(CONFIGS) == {"C1", "C2", "C3", "C4", "C5"}


def test_balanced_ranges_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        w = rng.integers(0, 1000, 997)
        b = balanced_ranges(w, world)
        assert b[0] == 0 and b[-1] == 997 and (np.diff(b) >= 0).all()
        sums = [w[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(sums) - min(sums) <= 2 * w.max()
    assert balanced_ranges(np.zeros(10), 4)[-1] == 10
    assert list(balanced_ranges([5], 3))[-1] == 1
    assert list(balanced_ranges([], 2)) == [0, 0, 0]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gather_worker(rank, world, port, sizes, K, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recsys2019_deeplearning_evaluation_b200.dist import allgather_topk_tables
    bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(sizes[rank])
    base = int(bounds[rank])
    idx = (torch.arange(n * K, dtype=torch.int32).reshape(n, K) + base * K)
    val = idx.to(torch.float32) * 0.5
    cnt = torch.full((n,), rank + 1, dtype=torch.int32)
    g_idx, g_val, g_cnt = allgather_topk_tables(idx, val, cnt, bounds)
    q.put((rank, g_idx.numpy().copy(), g_val.numpy().copy(), g_cnt.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [[5, 5], [7, 3], [0, 4]])
def test_allgather_topk_tables_gloo_world2(sizes):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    K = 4
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, sizes, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = sum(sizes)
    want_idx = np.arange(n * K, dtype=np.int32).reshape(n, K)
    want_cnt = np.concatenate([np.full(sizes[0], 1), np.full(sizes[1], 2)]).astype(np.int32)
    for rank, gi, gv, gc in res:
        assert gi.shape == (n, K) and (gi == want_idx).all()
        assert np.allclose(gv, want_idx * 0.5) and (gc == want_cnt).all()


def _delta_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recsys2019_deeplearning_evaluation_b200.dist import sync_replicated_delta
    V_prev = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    V = V_prev.clone()
    V[rank] += 1.0 + rank          # each rank moved a different row ...
    V[3] += 0.5                    # ... and both moved row 3
    sync_replicated_delta(V, V_prev)
    q.put((rank, V.numpy().copy(), V_prev.numpy().copy()))
    dist.destroy_process_group()


def test_sync_replicated_delta_gloo_world2():
    """K2 data parallelism: V <- V_prev + sum_r (V_r - V_prev) on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_delta_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(12, dtype=np.float32).reshape(4, 3)
    want[0] += 1.0; want[1] += 2.0; want[3] += 1.0
    for rank, V, Vp in res:
        assert np.allclose(V, want) and np.allclose(Vp, want)


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recsys2019_deeplearning_evaluation_b200.dist import ReplicatedDeltaExchange
    V = torch.zeros(8, dtype=torch.float32)
    x = ReplicatedDeltaExchange(V)
    seen = []
    for epoch in range(3):
        V[rank] += 1.0            # a private cell per rank ...
        V[4] += 0.25 * (rank + 1)  # ... and a shared one
        x.exchange()
        seen.append(V.numpy().copy())
    x.flush()
    q.put((rank, seen, V.numpy().copy()))
    dist.destroy_process_group()


def test_overlapped_delta_exchange_gloo_world2():
    """dist.ReplicatedDeltaExchange: the other rank's movement of epoch k becomes visible when exchange k + 1 starts (one
    epoch of staleness), and after flush() both replicas hold every movement exactly once."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros(8, np.float32)
    want[0] = want[1] = 3.0
    want[4] = 3 * 0.25 * (1 + 2)
    for rank, seen, V in res:
        assert np.allclose(V, want), (rank, V)
        other = 1 - rank
        # after exchange k (0-based) the replica holds its own k + 1 movements and the other rank's first k
        for k, s_ in enumerate(seen):
            assert np.isclose(s_[rank], k + 1) and np.isclose(s_[other], k), (rank, k, s_)


def _owned_rows_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recsys2019_deeplearning_evaluation_b200.dist import sync_owned_rows
    X = torch.zeros((6, 2), dtype=torch.float64)
    X[5] = -7.0                                   # a cold row: never listed, must survive
    rows = torch.tensor([0, 2, 3, 4], dtype=torch.int32)
    lo, hi = (0, 3) if rank == 0 else (3, 4)      # rank 0 solved rows 0, 2, 3; rank 1 solved row 4
    X[rows[lo:hi].long()] = 10.0 * (rank + 1) + rows[lo:hi].double()[:, None]
    sync_owned_rows(X, rows, lo, hi)
    q.put((rank, X.numpy().copy()))
    dist.destroy_process_group()


def test_sync_owned_rows_gloo_world2():
    """K4 row parallelism: after the exchange every rank holds every rank's solved rows; unlisted rows are untouched."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owned_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros((6, 2))
    want[0], want[2], want[3], want[4], want[5] = 10.0, 12.0, 13.0, 24.0, -7.0
    for rank, X in res:
        assert np.array_equal(X, want)


def test_check_matrix_formats_and_dtypes():
    """Base/Recommender_utils.py:13-52 semantics of the host helper."""
    import scipy.sparse as sps
    from recsys2019_deeplearning_evaluation_b200.recommender_utils import check_matrix
    D = np.array([[1.0, 0.0, 2.5], [0.0, 0.0, 3.0]])
    C = sps.csr_matrix(D)
    for fmt, cls in (("csc", sps.csc_matrix), ("csr", sps.csr_matrix), ("coo", sps.coo_matrix), ("lil", sps.lil_matrix)):
        out = check_matrix(C, fmt)
        assert isinstance(out, cls) and out.dtype == np.float32 and np.array_equal(out.toarray(), D.astype(np.float32))
        out = check_matrix(D, fmt, dtype=np.float64)  # ndarray in: sparse out, explicit zeros dropped
        assert isinstance(out, cls) and out.dtype == np.float64 and out.nnz == 3
    same = check_matrix(C, "csr", dtype=np.float64)
    assert isinstance(same, sps.csr_matrix) and same.dtype == np.float64
    assert isinstance(check_matrix(C, "npy"), np.ndarray) and check_matrix(C, "npy").dtype == np.float32
    assert np.array_equal(check_matrix(D, "npy"), D)


def test_argument_rules_of_the_next_row_trainers_are_checked_before_any_device_work():
    """The mirrors refuse what the reference refuses (same exception types) before they touch CUDA, so these run on a CPU box:
    ASY_SVD with a batch (pyx:399) or off the reference's sample stream, the tree-sparse SLIM mode in the throughput mode,
    SLIM ElasticNet with l1_ratio outside [0, 1] (SLIMElasticNetRecommender.py:43)."""
    from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch
    from recsys2019_deeplearning_evaluation_b200.recommenders import SLIMElasticNetRecommender
    from recsys2019_deeplearning_evaluation_b200.slim_bpr_epoch import SLIM_BPR_Cython_Epoch
    X = synth_urm(40, 15, 0.2, values="ratings")
    with pytest.raises(AssertionError, match="Batch size other than 1"):
        MatrixFactorization_Cython_Epoch(X, n_factors=4, algorithm_name="ASY_SVD", batch_size=2, random_seed=1)
    with pytest.raises(ValueError, match="ASY_SVD"):
        MatrixFactorization_Cython_Epoch(X, n_factors=4, algorithm_name="ASY_SVD", batch_size=1, sampler="philox", random_seed=1)
    with pytest.raises(ValueError, match="ASY_SVD"):
        MatrixFactorization_Cython_Epoch(X, n_factors=4, algorithm_name="ASY_SVD", batch_size=1, hogwild=True, random_seed=1)
    with pytest.raises(ValueError, match="algorithm_name"):
        MatrixFactorization_Cython_Epoch(X, n_factors=4, algorithm_name="SVD++")
    with pytest.raises(ValueError, match="train_with_sparse_weights"):
        SLIM_BPR_Cython_Epoch(X, train_with_sparse_weights=True, hogwild=True, sampler="philox", random_seed=1)
    with pytest.raises(AssertionError, match="l1_ratio must be between 0 and 1"):
        SLIMElasticNetRecommender(X, verbose=False).fit(l1_ratio=1.5)
