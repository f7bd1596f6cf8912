"""-m gpu: similarity checks added after the round's last GPU call (they first run at round end, so they sort last), and
the opt-in check of the sparse-candidate kernel (B200REC_K1B=1)."""
import os

import numpy as np
import pytest

from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from test_similarity_gpu import _check

pytestmark = pytest.mark.gpu


def test_python_named_implementation_is_the_same_device_path():
    """Compute_Similarity(use_implementation='python') / dense inputs (Compute_Similarity.py:71-113, a5) run the same kernel."""
    from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity, Compute_Similarity_Python
    X = synth_urm(300, 120, 0.05, seed=12, values="continuous")
    kw = dict(topK=9, shrink=2, similarity="cosine")
    a = Compute_Similarity(X, use_implementation="cython", **kw)
    b = Compute_Similarity(X, use_implementation="python", **kw)
    c = Compute_Similarity(X.toarray(), **kw)
    assert isinstance(b.compute_similarity_object, Compute_Similarity_Python) and c.dense
    Wa, Wb, Wc = a.compute_similarity(), b.compute_similarity(block_size=50), c.compute_similarity()
    assert abs(Wa - Wb).nnz == 0 and abs(Wa - Wc).nnz == 0
    with pytest.raises(ValueError):
        Compute_Similarity(X, use_implementation="numba", **kw)


@pytest.mark.skipif(os.environ.get("B200REC_K1B") != "1",
                    reason="the sparse-candidate kernel (csrc/sim_k1b.cuh) is opt-in until validated: B200REC_K1B=1")
def test_sparse_candidate_kernel_k1b(monkeypatch):
    """B200REC_K1B=1: the bitmap + table kernel is chosen for uniform binary data, gives the window kernel's answer (and the
    oracle's), and a table that is too small falls back to the window kernel."""
    import ctypes
    from recsys2019_deeplearning_evaluation_b200 import _lib
    L = _lib.load()
    for shape, kw in (((10_000, 5_000, 0.01), dict(topK=200, shrink=100, similarity="cosine")),
                      ((20_000, 230_000, 0.0003), dict(topK=50, shrink=10, similarity="jaccard")),
                      ((3_000, 1_500, 0.03), dict(topK=30, shrink=0, similarity="tversky", tversky_alpha=0.7, tversky_beta=1.3))):
        X = synth_urm(*shape, seed=42, values="binary")
        cols = np.arange(0, X.shape[1], max(1, X.shape[1] // 200))
        W1, sim1, _ = _check(X, cols=cols, **kw)
        en, tb = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(L.b200_sim_debug_k1b(sim1._h, 0, ctypes.byref(en), ctypes.byref(tb)))
        assert en.value == 1 and tb.value >= 12
        _lib.check(L.b200_sim_debug_k1b(sim1._h, 6, None, None))  # 64 slots: every column overflows -> window kernel
        W1b = sim1.compute_similarity()
        monkeypatch.delenv("B200REC_K1B")
        W0, sim0, _ = _check(X, cols=cols, **kw)
        _lib.check(L.b200_sim_debug_k1b(sim0._h, 0, ctypes.byref(en), None))
        assert en.value == 0
        monkeypatch.setenv("B200REC_K1B", "1")
        assert abs(W1 - W0).nnz == 0 and abs(W1b - W0).nnz == 0
    # skewed popularity: not eligible (or overflowing) -> still exact through the window kernel
    Xs = synth_urm(30_000, 2_000, 0.01, seed=13, values="binary", popularity=1.1)
    _check(Xs, cols=np.arange(0, 2000, 13), topK=100, shrink=10, similarity="cosine")


def test_negative_item_sample_evaluator_matches_reference_golden():
    """EvaluatorNegativeItemSample on the device pipeline (one user per step, candidate set through items_to_compute)."""
    import test_evaluation as TE
    from recsys2019_deeplearning_evaluation_b200.evaluation import EvaluatorNegativeItemSample
    train, test, neg, S, kw = TE.eval_negative_case()
    res, _ = EvaluatorNegativeItemSample(test, neg, verbose=False, **kw).evaluateRecommender(TE._stub(train, S))
    TE._assert_close(res, TE._golden_negative(kw["cutoff_list"]), 1e-6, "cuda negative-sample")


def test_similarity_matrix_topk_reference_recipes():
    """Base/Recommender_utils_Test.py:18-50 through recommender_utils.similarityMatrixTopK: k non-zeros per column on a dense
    input, and dense input == sparse input."""
    import scipy.sparse as sps
    from recsys2019_deeplearning_evaluation_b200.recommender_utils import similarityMatrixTopK
    rng = np.random.default_rng(4)
    dense = rng.random((100, 100)).astype(np.float32)
    out = similarityMatrixTopK(dense, k=20)
    assert sps.isspmatrix_csc(out) and (out.toarray() != 0).sum() == 20 * 100
    small = rng.random((20, 20)).astype(np.float32)
    a = similarityMatrixTopK(small, k=5).toarray()
    b = similarityMatrixTopK(sps.csc_matrix(small), k=5).toarray()
    assert np.allclose(a, b)
    kept = np.sort(small, axis=0)[-5:, :]
    assert np.allclose(np.sort(a, axis=0)[-5:, :], kept)
