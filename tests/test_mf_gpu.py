"""-m gpu: BPR-MF / FunkSVD epochs through the C ABI against the C oracle (oracle/sgd_oracle.c), which is itself
pinned to the compiled reference (tests/test_oracle_sgd.py).  Learned factors within 1e-4 relative (north_star)."""
import numpy as np
import pytest

from oracle.sgd_oracle import MFOracle
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-6  # north_star tolerance; atol covers factors that sit at ~0


def _cls():
    from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch
    return MatrixFactorization_Cython_Epoch


def _compare(g, o, bias=False):
    for name in ("get_USER_factors", "get_ITEM_factors") + (("get_USER_bias", "get_ITEM_bias", "get_GLOBAL_bias") if bias else ()):
        a, b = getattr(g, name)(), getattr(o, name)()
        assert np.allclose(a, b, rtol=RTOL, atol=ATOL), "%s: max abs diff %.3e" % (name, float(np.abs(a - b).max()))


CASES = [
    ("MF_BPR", dict(sgd_mode="sgd", batch_size=32, n_factors=16, user_reg=1e-3, positive_reg=2e-3, negative_reg=3e-3)),
    ("MF_BPR", dict(sgd_mode="adagrad", batch_size=100, n_factors=128, user_reg=1e-3, positive_reg=1e-3, negative_reg=1e-3)),
    ("MF_BPR", dict(sgd_mode="adam", batch_size=7, n_factors=10)),
    ("MF_BPR", dict(sgd_mode="rmsprop", batch_size=1, n_factors=5)),
    ("MF_BPR", dict(sgd_mode="sgd", batch_size=1000, n_factors=64)),  # batch larger than n_users: one (full) batch
    ("FUNK_SVD", dict(sgd_mode="adam", batch_size=16, n_factors=12, use_bias=True, negative_interactions_quota=0.3,
                      bias_reg=1e-3, user_reg=1e-3, positive_reg=1e-3)),
    ("FUNK_SVD", dict(sgd_mode="sgd", batch_size=50, n_factors=32, use_bias=False, negative_interactions_quota=0.0)),
    ("FUNK_SVD", dict(sgd_mode="adagrad", batch_size=64, n_factors=20, use_bias=True, negative_interactions_quota=0.5)),
]


@pytest.mark.parametrize("dataflow", ["1", "0"])
@pytest.mark.parametrize("algo,kw", CASES)
def test_glibc_stream_parity(algo, kw, dataflow, monkeypatch):
    """Same seed => same replayed libc sample stream => factors agree with the reference semantics.  Both mini-batch
    kernels: the dataflow kernel (default; the bias cases fall back to the cooperative one by design) and the
    cooperative two-barriers-per-batch kernel."""
    monkeypatch.setenv("B200REC_MF_DATAFLOW", dataflow)
    X = synth_urm(300, 120, 0.08, seed=3, values="ratings")
    common = dict(algorithm_name=algo, learning_rate=0.05, random_seed=42, **kw)
    g, o = _cls()(X, **common), MFOracle(X, **common)
    for _ in range(3):
        g.epochIteration_Cython()
        o.epochIteration_Cython()
    _compare(g, o, bias=kw.get("use_bias", False))
    u, i, j = g.get_samples()
    assert len(u) == g.samples_last_epoch() == ((300 if algo == "MF_BPR" else X.nnz) // kw["batch_size"] + 1) * kw["batch_size"]


def test_minibatch_mode_is_run_to_run_deterministic():
    """The batch gradient sums are fp64 atomics (order-independent at fp32 precision): twenty runs of the FunkSVD /
    adam / bias case -- the one an fp32 atomic sum made flaky -- agree with each other to the last fp32 bits (1e-6 relative) and are all within tolerance."""
    algo, kw = CASES[5]
    X = synth_urm(300, 120, 0.08, seed=3, values="ratings")
    common = dict(algorithm_name=algo, learning_rate=0.05, random_seed=42, **kw)
    o = MFOracle(X, **common)
    for _ in range(3):
        o.epochIteration_Cython()
    first = None
    for _ in range(20):
        g = _cls()(X, **common)
        for _ in range(3):
            g.epochIteration_Cython()
        _compare(g, o, bias=True)
        got = (g.get_USER_factors(), g.get_ITEM_factors(), g.get_USER_bias(), g.get_ITEM_bias())
        if first is None:
            first = got
        else:
            for a, b in zip(got, first):
                assert np.allclose(a, b, rtol=1e-6, atol=1e-8), float(np.abs(a - b).max())


@pytest.mark.parametrize("shape", ["sparse", "dense_profiles", "cold_users"])
@pytest.mark.parametrize("algo", ["MF_BPR", "FUNK_SVD"])
def test_glibc_stream_resolved_on_the_device_equals_the_host_replay(algo, shape, monkeypatch):
    """The reference's rand() stream with its rejection loops, resolved in parallel on the device (lengths at every stream
    position + pointer doubling), against the sequential host replay: identical (u, i, j | rating) streams over several
    epochs (the unread tail of one epoch's buffer opens the next), with many negative rejections (dense profiles: the buffer
    has to be extended) and many user rejections (empty profiles)."""
    if shape == "sparse":
        X = synth_urm(3000, 900, 0.01, seed=21, values="ratings")
    elif shape == "dense_profiles":
        X = synth_urm(200, 40, 0.7, seed=22, values="ratings")   # most negative draws are rejected
    else:
        X = synth_urm(500, 200, 0.004, seed=23, values="ratings")  # most users have an empty profile
    kw = dict(algorithm_name=algo, n_factors=8, batch_size=50, learning_rate=0.01, random_seed=77, sgd_mode="sgd",
              negative_interactions_quota=0.35)
    monkeypatch.setenv("B200REC_GLIBC_HOST", "1")
    host = _cls()(X, **kw)
    monkeypatch.setenv("B200REC_GLIBC_HOST", "0")
    dev = _cls()(X, **kw)
    for _ in range(3):
        host.epochIteration_Cython()
        dev.epochIteration_Cython()
        a, b = host.get_samples(), dev.get_samples()
        assert len(a[0]) == len(b[0]) > 0
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    _compare(dev, host)


def test_dataflow_kernel_under_heavy_row_sharing():
    """Few rows, large batches: almost every (row, batch) pair is hit many times, so the accumulate-and-last-arriver path,
    the waits on the previous batch and the per-row arrival counters all run hot (the C5 shape exercises almost only the
    hit-once fast path)."""
    X = synth_urm(40, 12, 0.4, seed=8, values="ratings")
    for algo, kw in (("MF_BPR", dict(sgd_mode="adam", batch_size=64, n_factors=8, user_reg=1e-3, positive_reg=1e-3, negative_reg=1e-3)),
                     ("FUNK_SVD", dict(sgd_mode="adagrad", batch_size=97, n_factors=33, negative_interactions_quota=0.4)),
                     ("MF_BPR", dict(sgd_mode="sgd", batch_size=5, n_factors=128))):
        common = dict(algorithm_name=algo, learning_rate=0.05, random_seed=13, **kw)
        g, o = _cls()(X, **common), MFOracle(X, **common)
        for _ in range(4):
            g.epochIteration_Cython()
            o.epochIteration_Cython()
        _compare(g, o)


@pytest.mark.parametrize("algo", ["MF_BPR", "FUNK_SVD"])
def test_philox_stream_parity_and_sample_validity(algo):
    """Device-drawn samples obey the reference's acceptance rules, and replaying them through the oracle gives
    the same factors."""
    X = synth_urm(400, 150, 0.06, seed=5, values="ratings")
    kw = dict(algorithm_name=algo, n_factors=24, batch_size=48, learning_rate=0.03, random_seed=9, sgd_mode="adagrad",
              user_reg=1e-3, positive_reg=1e-3, negative_reg=1e-3, negative_interactions_quota=0.4)
    g = _cls()(X, sampler="philox", **kw)
    init = (g.get_USER_factors(), g.get_ITEM_factors())
    streams = []
    for _ in range(2):
        g.epochIteration_Cython()
        streams.append(g.get_samples())
    su = np.concatenate([s[0] for s in streams]); si = np.concatenate([s[1] for s in streams]); s3 = np.concatenate([s[2] for s in streams])
    dense = X.toarray()
    lens = np.diff(X.indptr)
    assert (lens[su] > 0).all()
    if algo == "MF_BPR":
        assert (dense[su, si] != 0).all() and (dense[su, s3] == 0).all()
        assert len(np.unique(su)) > 200 and len(np.unique(s3)) > 100
    else:
        pos = s3 != 0
        assert np.allclose(dense[su[pos], si[pos]], s3[pos]) and (dense[su[~pos], si[~pos]] == 0).all()
        assert 0.3 < pos.mean() < 0.5
    assert not np.array_equal(streams[0][0], streams[1][0])
    o = MFOracle(X, init_factors=init, samples=(su, si, s3), **kw)
    for _ in range(2):
        o.epochIteration_Cython()
    _compare(g, o)


def test_hogwild_descends_like_the_sequential_recursion():
    """Hogwild mode (no batch barrier; every sample of an epoch may read stale rows) against the batch_size=1
    recursion on the same Philox stream and initial point: statistical parity -- the BPR
    objective on a fixed probe set must fall by a comparable amount."""
    X = synth_urm(4000, 600, 0.03, seed=7, popularity=0.8)
    kw = dict(algorithm_name="MF_BPR", n_factors=32, batch_size=1, learning_rate=0.05, random_seed=3, sgd_mode="sgd",
              user_reg=1e-4, positive_reg=1e-4, negative_reg=1e-4)
    g = _cls()(X, sampler="philox", hogwild=True, **kw)
    s = _cls()(X, sampler="philox", hogwild=False, **kw)  # sequential batch-1 semantics, same stream
    rng = np.random.default_rng(0)
    us = rng.integers(0, 4000, 30000)
    us = us[np.diff(X.indptr)[us] > 0]
    pos = np.array([X.indices[X.indptr[u] + rng.integers(0, X.indptr[u + 1] - X.indptr[u])] for u in us])
    neg = rng.integers(0, 600, len(us))

    def loss(m):
        U, V = m.get_USER_factors(), m.get_ITEM_factors()
        x = np.einsum("ij,ij->i", U[us], V[pos] - V[neg])
        return float(np.mean(np.log1p(np.exp(-x))))

    l0 = loss(s)
    for _ in range(40):
        g.epochIteration_Cython()
        s.epochIteration_Cython()
    assert np.array_equal(g.get_samples()[0], s.get_samples()[0])
    ls, lh = loss(s), loss(g)
    assert l0 - ls > 0.01, (l0, ls)
    assert 0.7 < (l0 - lh) / (l0 - ls) < 1.3, (l0, ls, lh)


def test_argument_errors():
    X = synth_urm(50, 20, 0.2)
    with pytest.raises(ValueError, match="sgd_mode"):
        _cls()(X, algorithm_name="MF_BPR", sgd_mode="momentum")
    with pytest.raises(ValueError, match="algorithm_name"):
        _cls()(X, algorithm_name="SVD++")
