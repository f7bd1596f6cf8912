"""CPU: pins oracle/sgd_oracle.c -- glibc rand() replay against libc itself, the MF / SLIM trainers against the
compiled reference (oracle/_ref, when present) and against golden vectors generated from it."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ref_loader, sgd_oracle
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sgd_golden.npz")


@pytest.mark.parametrize("seed", [1, 42, 12345, 2 ** 31 + 5])
def test_glibc_rand_replay_matches_libc(seed):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    g = sgd_oracle.GlibcRand(seed & 0xFFFFFFFF)
    assert [libc.rand() for _ in range(5000)] == [g.rand() for _ in range(5000)]


MF_CASES = [
    ("MF_BPR", dict(sgd_mode="sgd", batch_size=32, user_reg=1e-3, positive_reg=2e-3, negative_reg=3e-3)),
    ("MF_BPR", dict(sgd_mode="adagrad", batch_size=32, user_reg=1e-3, positive_reg=2e-3, negative_reg=3e-3)),
    ("MF_BPR", dict(sgd_mode="adam", batch_size=7)),
    ("MF_BPR", dict(sgd_mode="rmsprop", batch_size=1)),
    ("FUNK_SVD", dict(sgd_mode="adam", batch_size=16, use_bias=True, negative_interactions_quota=0.3, bias_reg=1e-3,
                      user_reg=1e-3, positive_reg=1e-3)),
    ("FUNK_SVD", dict(sgd_mode="sgd", batch_size=50, use_bias=False, negative_interactions_quota=0.0)),
]
SLIM_CASES = [(sym, mode) for sym in (True, False) for mode in ("sgd", "adagrad", "adam", "rmsprop")]


def _urm():
    return synth_urm(300, 120, 0.08, seed=3, values="ratings")


def _run_mf(cls, algo, kw):
    m = cls(_urm(), n_factors=16, algorithm_name=algo, learning_rate=0.05, random_seed=42, **kw)
    for _ in range(3):
        m.epochIteration_Cython()
    return m.get_USER_factors(), m.get_ITEM_factors()


def _run_slim_oracle(sym, mode):
    o = sgd_oracle.SLIMOracle(_urm(), learning_rate=0.05, li_reg=1e-3, lj_reg=2e-3, topK=120, symmetric=sym, random_seed=7,
                              sgd_mode=mode)
    for _ in range(3):
        o.epochIteration_Cython()
    S = o.S_full()
    np.fill_diagonal(S, 0)
    return S


@pytest.mark.parametrize("n", range(len(MF_CASES)))
def test_mf_oracle_matches_golden(n):
    z = np.load(GOLD)
    U, V = _run_mf(sgd_oracle.MFOracle, *MF_CASES[n])
    assert np.abs(U - z["mf%d_U" % n]).max() < 1e-12 and np.abs(V - z["mf%d_V" % n]).max() < 1e-12


@pytest.mark.parametrize("n", range(len(SLIM_CASES)))
def test_slim_oracle_matches_golden(n):
    z = np.load(GOLD)
    S = _run_slim_oracle(*SLIM_CASES[n])
    # dense (non-symmetric) get_S goes through a float32 similarityMatrixTopK in the reference (pyx:371,386)
    tol = 1e-12 if SLIM_CASES[n][0] else 1e-6
    assert np.abs(S - z["slim%d_S" % n]).max() < tol


@pytest.mark.skipif(ref_loader.load("MatrixFactorization_Cython_Epoch") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [0, 2, 4])
def test_mf_oracle_matches_compiled_reference_live(n):
    ref = ref_loader.load("MatrixFactorization_Cython_Epoch").MatrixFactorization_Cython_Epoch
    a, b = _run_mf(ref, *MF_CASES[n]), _run_mf(sgd_oracle.MFOracle, *MF_CASES[n])
    assert np.abs(a[0] - b[0]).max() < 1e-12 and np.abs(a[1] - b[1]).max() < 1e-12


def test_external_sample_stream_equals_internal():
    X = _urm()
    kw = dict(n_factors=8, algorithm_name="MF_BPR", batch_size=10, learning_rate=0.05, random_seed=5, sgd_mode="adagrad")
    a = sgd_oracle.MFOracle(X, record=10000, **kw)
    a.epochIteration_Cython()
    u, i, j = a.recorded()
    np.random.seed(5)
    init = (np.random.normal(0, 0.1, (300, 8)), np.random.normal(0, 0.1, (120, 8)))
    b = sgd_oracle.MFOracle(X, init_factors=init, samples=(u, i, j), **kw)
    b.epochIteration_Cython()
    assert np.array_equal(a.get_USER_factors(), b.get_USER_factors())
