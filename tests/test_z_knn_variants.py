"""Content-based, hybrid and custom-similarity KNN recommenders (knn.py): the reference's golden W_sparse restated through
the oracles on CPU (pins the orientation / stacking / weighting logic of the mirrors), and the mirrors themselves on the
GPU (-m gpu)."""
import os
import runpy

import numpy as np
import pytest
import scipy.sparse as sps

from golden_util import same_sparse
from oracle import weighting_oracle as wo
from oracle.similarity_oracle import SimilarityOracle, check_topk_against_dense

HERE = os.path.dirname(os.path.abspath(__file__))
_mk = runpy.run_path(os.path.join(HERE, "golden", "make_golden.py"), run_name="cases")
CASES, inputs = _mk["KNN_VARIANT_CASES"], _mk["knn_variant_inputs"]
Z = np.load(os.path.join(HERE, "golden", "knn_variants_golden.npz"))
TIE_FREE = (0, 1, 2)  # continuous content values (or BM25-weighted ones): the reference's index sets are reproducible


def _content(n):
    """The matrix whose COLUMNS the reference compares in case n, restated with the oracles."""
    name, kw = CASES[n]
    URM, ICM, UCM = inputs()
    kw = dict(kw)
    fw = kw.pop("feature_weighting", "none")
    if name == "ItemKNN_CFCBF_Hybrid_Recommender":
        M = sps.hstack([ICM * kw.pop("ICM_weight"), URM.T], format="csr")
    elif name == "UserKNN_CFCBF_Hybrid_Recommender":
        M = sps.hstack([UCM * kw.pop("UCM_weight"), URM], format="csr")
    else:
        M = ICM if name.startswith("Item") else UCM
    if fw != "none":
        M = sps.csr_matrix((wo.okapi_BM_25 if fw == "BM25" else wo.TF_IDF)(M), dtype=np.float32)
    return sps.csr_matrix(M.T, dtype=np.float32), kw


def _golden_W(n):
    k = "k%d" % n
    m = len(Z[k + "_indptr"]) - 1
    return sps.csr_matrix((Z[k + "_data"], Z[k + "_indices"], Z[k + "_indptr"]), shape=(m, m))


@pytest.mark.parametrize("n", range(len(CASES)))
def test_golden_is_the_similarity_of_the_restated_content_matrix(n):
    X, kw = _content(n)
    W = _golden_W(n)
    check_topk_against_dense(W, SimilarityOracle(X, **kw), np.arange(W.shape[0]), rtol=1e-4)


def _fit(n):
    from recsys2019_deeplearning_evaluation_b200 import knn
    name, kw = CASES[n]
    URM, ICM, UCM = inputs()
    r = getattr(knn, name)(URM.copy(), (ICM if name.startswith("Item") else UCM).copy(), verbose=False)
    r.fit(**kw)
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("n", range(len(CASES)))
def test_cuda_matches_reference_golden(n):
    r = _fit(n)
    W = _golden_W(n)
    assert r.W_sparse.shape == W.shape and r.W_sparse.nnz == W.nnz
    X, kw = _content(n)
    check_topk_against_dense(r.W_sparse, SimilarityOracle(X, **kw), np.arange(W.shape[0]), rtol=1e-4)
    if n in TIE_FREE:
        assert same_sparse(r.W_sparse, W, rtol=1e-4, atol=1e-7)
        assert np.allclose(r._compute_item_score(np.arange(25)), Z["k%d_scores" % n], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_custom_similarity_and_argument_checks():
    from recsys2019_deeplearning_evaluation_b200 import knn
    URM, ICM, UCM = inputs()
    Wc = sps.random(150, 150, 0.2, format="csr", random_state=9, dtype=np.float32)
    r = knn.ItemKNNCustomSimilarityRecommender(URM.copy(), verbose=False)
    r.fit(Wc, selectTopK=True, topK=6)
    Wref = sps.csr_matrix((Z["custom_data"], Z["custom_indices"], Z["custom_indptr"]), shape=(150, 150))
    assert same_sparse(r.W_sparse, Wref, rtol=1e-6)
    with pytest.raises(AssertionError):
        r.fit(Wc[:, :100])
    with pytest.raises(AssertionError):
        knn.ItemKNNCBFRecommender(URM, ICM[:100], verbose=False)
    with pytest.raises(ValueError):
        knn.UserKNNCBFRecommender(URM, UCM, verbose=False).fit(feature_weighting="tfidf")
