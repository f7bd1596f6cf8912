"""BM25 / TF-IDF feature weighting (Base/IR_feature_weighting.py): the numpy restatement against the reference's golden
outputs (CPU), and the CUDA path against both plus the KNN recommenders fitted with feature_weighting (-m gpu)."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from golden_util import same_sparse
from oracle import weighting_oracle as wo
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "weighting_golden.npz"))
VALUES = ("continuous", "ratings", "binary")


def _urm(values):
    return synth_urm(400, 150, 0.06, seed=17, values=values)


def _golden(tag, values):
    return sps.csr_matrix((Z["%s_%s_data" % (tag, values)], Z["%s_%s_indices" % (tag, values)], Z["%s_%s_indptr" % (tag, values)]),
                          shape=(150, 400))


@pytest.mark.parametrize("values", VALUES)
@pytest.mark.parametrize("tag", ["bm25", "tfidf"])
def test_oracle_matches_reference_golden(tag, values):
    fn = wo.okapi_BM_25 if tag == "bm25" else wo.TF_IDF
    assert same_sparse(fn(_urm(values).T), _golden(tag, values), rtol=1e-5, atol=1e-9)  # the reference's fp32 sums


@pytest.mark.gpu
@pytest.mark.parametrize("values", VALUES)
@pytest.mark.parametrize("tag", ["bm25", "tfidf"])
def test_cuda_matches_reference_golden(tag, values):
    from recsys2019_deeplearning_evaluation_b200 import weighting
    fn = weighting.okapi_BM_25 if tag == "bm25" else weighting.TF_IDF
    W = fn(_urm(values).T)
    assert sps.isspmatrix_csr(W) and W.shape == (150, 400)
    assert same_sparse(W, _golden(tag, values), rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
def test_cuda_matches_oracle_larger_and_parameters():
    from recsys2019_deeplearning_evaluation_b200 import weighting
    X = synth_urm(20_000, 3_000, 0.01, seed=5, values="ratings", popularity=1.0)
    assert same_sparse(weighting.okapi_BM_25(X.T, K1=2.0, B=0.4), wo.okapi_BM_25(X.T, K1=2.0, B=0.4), rtol=1e-4, atol=1e-7)
    assert same_sparse(weighting.TF_IDF(X.T), wo.TF_IDF(X.T), rtol=1e-4, atol=1e-7)
    with pytest.raises(AssertionError):
        weighting.okapi_BM_25(X.T, B=1.5)
    Xn = X.copy()
    Xn.data[0] = -1.0
    with pytest.raises(AssertionError):
        weighting.TF_IDF(Xn.T)


@pytest.mark.gpu
@pytest.mark.parametrize("side", ["item", "user"])
@pytest.mark.parametrize("fw", ["BM25", "TF-IDF"])
def test_knn_with_feature_weighting_matches_reference(side, fw):
    from recsys2019_deeplearning_evaluation_b200.recommenders import ItemKNNCFRecommender, UserKNNCFRecommender
    cls = ItemKNNCFRecommender if side == "item" else UserKNNCFRecommender
    r = cls(_urm("ratings"), verbose=False)
    r.fit(topK=10, shrink=2, similarity="cosine", feature_weighting=fw)
    k = "knn_%s_%s" % (side, fw.replace("-", ""))
    n = 150 if side == "item" else 400
    Wref = sps.csr_matrix((Z[k + "_data"], Z[k + "_indices"], Z[k + "_indptr"]), shape=(n, n))
    assert same_sparse(r.W_sparse, Wref, rtol=1e-4, atol=1e-7)
    with pytest.raises(ValueError):
        cls(_urm("ratings"), verbose=False).fit(feature_weighting="bm25")
