"""-m gpu: the recommender-level API (fit / _compute_item_score / recommend) against the reference's golden
ItemKNN fit + scores and numpy restatements of BaseRecommender.recommend (Base/BaseRecommender.py:131-222)."""
import numpy as np
import pytest
import scipy.sparse as sps

from golden_util import load_golden, same_sparse
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

pytestmark = pytest.mark.gpu
URMS, CASES, KNN = load_golden()


def _ref_recommend(URM, scores, users, cutoff, remove_seen=True):
    """BaseRecommender.py:164-212 restated with the deterministic tie rule (score desc, item asc)."""
    out = []
    for r, u in enumerate(users):
        s = scores[r].astype(np.float64).copy()
        if remove_seen:
            s[URM.indices[URM.indptr[u]:URM.indptr[u + 1]]] = -np.inf
        order = np.lexsort((np.arange(len(s)), -s))[:cutoff]
        out.append(order[np.isfinite(s[order])].tolist())
    return out


def test_itemknn_fit_and_scores_match_reference_golden():
    from recsys2019_deeplearning_evaluation_b200.recommenders import ItemKNNCFRecommender
    X = URMS["ratings"]
    rec = ItemKNNCFRecommender(X, verbose=False)
    rec.fit(topK=15, shrink=5, similarity="cosine", normalize=True)
    # ratings data has ties at the K boundary: compare through the scores of the tie-free part and structure size
    Wg = KNN["W"]
    assert rec.W_sparse.shape == (150, 150) and abs(rec.W_sparse.nnz - Wg.nnz) <= 0.02 * Wg.nnz
    users = np.arange(40)
    sc = rec._compute_item_score(users)
    ref = X[users].dot(rec.W_sparse).toarray()
    assert sc.shape == (40, 150) and np.allclose(sc, ref, rtol=1e-5, atol=1e-6)
    # scores against the reference's own golden scores: identical wherever the two W agree (tie classes aside)
    agree = np.isclose(sc, KNN["scores"], rtol=1e-4, atol=1e-5).mean()
    assert agree > 0.97
    sc2 = rec._compute_item_score(users, items_to_compute=[1, 5, 9])
    assert np.isneginf(np.delete(sc2, [1, 5, 9], axis=1)).all() and np.allclose(sc2[:, [1, 5, 9]], ref[:, [1, 5, 9]], rtol=1e-5, atol=1e-6)


def test_recommend_matches_restatement():
    from recsys2019_deeplearning_evaluation_b200.recommenders import ItemKNNCFRecommender, UserKNNCFRecommender
    X = synth_urm(1200, 500, 0.03, seed=4, values="continuous")
    for cls in (ItemKNNCFRecommender, UserKNNCFRecommender):
        rec = cls(X, verbose=False)
        rec.fit(topK=30, shrink=2)
        users = np.array([0, 5, 17, 400, 1199])
        lists, scores = rec.recommend(users, cutoff=20, return_scores=True)
        raw = rec._compute_item_score(users)
        assert lists == _ref_recommend(rec.URM_train, raw, users, 20)
        assert np.isneginf(scores[0, X.indices[X.indptr[0]:X.indptr[1]]]).all()
        assert rec.recommend(5, cutoff=7) == _ref_recommend(rec.URM_train, raw[1:2], [5], 7)[0]
        nos = rec.recommend(users, cutoff=10, remove_seen_flag=False)
        assert nos == _ref_recommend(rec.URM_train, raw, users, 10, remove_seen=False)
    if cls is UserKNNCFRecommender:
        ref = rec.W_sparse[users].dot(rec.URM_train).toarray()
        assert np.allclose(raw, ref, rtol=1e-5, atol=1e-6)


def test_mf_recommenders_fit_and_score():
    from recsys2019_deeplearning_evaluation_b200.recommenders import MatrixFactorization_BPR_Cython, MatrixFactorization_FunkSVD_Cython
    X = synth_urm(600, 200, 0.05, seed=6, values="ratings")
    bpr = MatrixFactorization_BPR_Cython(X, verbose=False)
    bpr.fit(epochs=3, batch_size=50, num_factors=24, learning_rate=0.05, sgd_mode="adagrad", random_seed=3)
    assert bpr.USER_factors.shape == (600, 24) and bpr.ITEM_factors.shape == (200, 24) and not bpr.use_bias
    users = np.arange(0, 600, 37)
    sc = bpr._compute_item_score(users)
    assert np.allclose(sc, bpr.USER_factors[users] @ bpr.ITEM_factors.T, rtol=1e-4, atol=1e-6)
    assert bpr.recommend(users, cutoff=15) == _ref_recommend(bpr.URM_train, sc, users, 15)
    fk = MatrixFactorization_FunkSVD_Cython(X, verbose=False)
    fk.fit(epochs=2, batch_size=64, num_factors=10, learning_rate=0.02, sgd_mode="adam", use_bias=True, random_seed=3,
           negative_interactions_quota=0.3)
    sc = fk._compute_item_score(users)
    ref = fk.USER_factors[users] @ fk.ITEM_factors.T + fk.GLOBAL_bias + fk.USER_bias[users][:, None] + fk.ITEM_bias[None, :]
    assert np.allclose(sc, ref, rtol=1e-4, atol=1e-5)


class _CountingEvaluator:
    """Stand-in for Base/Evaluation/Evaluator.py: returns a metric that rises, then falls."""

    def __init__(self, values):
        self.values, self.calls = list(values), 0

    def evaluateRecommender(self, rec):
        v = self.values[min(self.calls, len(self.values) - 1)]
        self.calls += 1
        assert rec._compute_item_score(np.arange(3)).shape[0] == 3
        return {10: {"MAP": v}}, ""


def test_early_stopping_loop_and_slim_wrapper():
    from recsys2019_deeplearning_evaluation_b200.recommenders import SLIM_BPR_Cython, MatrixFactorization_BPR_Cython
    X = synth_urm(500, 120, 0.06, seed=8)
    ev = _CountingEvaluator([0.1, 0.2, 0.15, 0.12, 0.11])
    m = MatrixFactorization_BPR_Cython(X, verbose=False)
    m.fit(epochs=50, batch_size=32, num_factors=8, random_seed=1, validation_every_n=2, stop_on_validation=True,
          validation_metric="MAP", lower_validations_allowed=2, evaluator_object=ev)
    assert ev.calls == 4 and m.epochs_best == 4 and m.best_validation_metric == 0.2  # stopped after two worse validations
    s = SLIM_BPR_Cython(X, verbose=False)
    s.fit(epochs=4, topK=12, learning_rate=0.05, random_seed=2, sgd_mode="adagrad", symmetric=True)
    assert sps.isspmatrix_csr(s.W_sparse) and s.W_sparse.shape == (120, 120)
    assert (np.diff(s.W_sparse.tocsc().indptr) <= 12).all() and s.W_sparse.nnz > 0
    assert len(s.recommend(3, cutoff=5)) == 5
    with pytest.raises(ValueError):
        SLIM_BPR_Cython(X, verbose=False).fit(epochs=1, topK=0)
