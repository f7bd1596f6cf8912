"""-m gpu: further similarity / evaluator checks (the Python-named implementation, the negative-sample evaluator)."""
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
    # fp32 shared-memory atomics add a column's products in a run-dependent order: same pattern, values to the last bits
    for Wo in (Wb, Wc):
        assert np.array_equal(Wa.indptr, Wo.indptr) and np.array_equal(Wa.indices, Wo.indices)
        assert np.allclose(Wa.data, Wo.data, rtol=1e-5, atol=0)
    with pytest.raises(ValueError):
        Compute_Similarity(X, use_implementation="numba", **kw)


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
