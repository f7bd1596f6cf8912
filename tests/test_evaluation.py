"""Hold-out evaluation: the numpy restatement (oracle/evaluator_oracle.py) against the reference's EvaluatorHoldout golden
results (CPU), and the device evaluator (evaluation.EvaluatorHoldout -> csrc/eval.cu) against both (-m gpu)."""
import os
import runpy

import numpy as np
import pytest

from oracle.evaluator_oracle import evaluate_scores
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

HERE = os.path.dirname(os.path.abspath(__file__))
eval_case = runpy.run_path(os.path.join(HERE, "golden", "make_golden.py"), run_name="cases")["eval_case"]
Z = np.load(os.path.join(HERE, "golden", "evaluator_golden.npz"))


def _assert_close(res, ref, rtol, what):
    assert set(res.keys()) == set(ref.keys())
    for c in ref:
        assert list(res[c].keys()) == list(ref[c].keys()) or set(res[c].keys()) == set(ref[c].keys())
        for k, v in ref[c].items():
            assert np.isclose(res[c][k], v, rtol=rtol, atol=1e-9), "%s cutoff %s %s: %r vs %r" % (what, c, k, res[c][k], v)


def _golden(n, cutoffs):
    return {c: {k.split("_", 2)[2]: float(Z[k]) for k in Z.files if k.startswith("e%d_c%d_" % (n, c))} for c in cutoffs}


@pytest.mark.parametrize("n", range(3))
def test_oracle_matches_reference_golden(n):
    train, test, S, kw = eval_case(n)
    _assert_close(evaluate_scores(train, test, S, **kw), _golden(n, kw["cutoff_list"]), 1e-9, "oracle")


def _stub(train, S):
    import torch
    from recsys2019_deeplearning_evaluation_b200.recommenders import BaseRecommender

    class Stub(BaseRecommender):
        RECOMMENDER_NAME = "Stub"

        def __init__(self, URM_train, S):
            super(Stub, self).__init__(URM_train, verbose=False)
            self._S = torch.from_numpy(np.ascontiguousarray(S, np.float32)).cuda()

        def _scores_device(self, d_users, items_to_compute=None):
            return self._S[d_users.long()].contiguous()

    return Stub(train, S)


@pytest.mark.gpu
@pytest.mark.parametrize("n", range(3))
def test_cuda_matches_reference_golden(n):
    from recsys2019_deeplearning_evaluation_b200.evaluation import EvaluatorHoldout, METRIC_NAMES
    train, test, S, kw = eval_case(n)
    ev = EvaluatorHoldout(test, verbose=False, **kw)
    res, text = ev.evaluateRecommender(_stub(train, S), block_size=64)  # several blocks accumulate on the device
    _assert_close(res, _golden(n, kw["cutoff_list"]), 1e-6, "cuda")
    assert list(res[kw["cutoff_list"][0]].keys()) == METRIC_NAMES
    assert text.startswith("CUTOFF: %d - PRECISION: " % kw["cutoff_list"][0])


@pytest.mark.gpu
def test_cuda_evaluates_itemknn_like_the_oracle():
    """A real model end to end: ItemKNN scores, seen items removed, cutoffs up to 100, against the restatement fed with
    the model's own score matrix."""
    from recsys2019_deeplearning_evaluation_b200.evaluation import EvaluatorHoldout
    from recsys2019_deeplearning_evaluation_b200.recommenders import ItemKNNCFRecommender
    train = synth_urm(3000, 1500, 0.02, seed=61, values="ratings", popularity=0.8)
    test = synth_urm(3000, 1500, 0.004, seed=62, values="ratings")
    rec = ItemKNNCFRecommender(train, verbose=False)
    rec.fit(topK=40, shrink=10)
    kw = dict(cutoff_list=[5, 20, 100], min_ratings_per_user=2)
    res, _ = EvaluatorHoldout(test, verbose=False, **kw).evaluateRecommender(rec)
    # ties (the zero scores) resolve to the ascending item id on both sides; the score block is summed with atomics, so
    # near-equal scores may swap between two runs: averaged metrics agree to 1e-4, not bit for bit
    ref = evaluate_scores(train, test, rec._compute_item_score(np.arange(3000)), **kw)
    _assert_close(res, ref, 1e-4, "itemknn")


def test_host_side_preprocessing_matches_the_restatement():
    """CPU: the evaluator's one-off host work -- per-user ideal DCG table and the users_to_evaluate selection -- against
    per-user loops (metrics.py:268, :277-279; Evaluator.py:199-222)."""
    from recsys2019_deeplearning_evaluation_b200.evaluation import EvaluatorHoldout, _ideal_dcg
    train, test, S, kw = eval_case(2)
    cut = [1, 5, 10, 40]
    T = test.copy()
    T.sort_indices()
    table = _ideal_dcg(T, cut)
    for u in range(T.shape[0]):
        rel = np.sort(T.data[T.indptr[u]:T.indptr[u + 1]].astype(np.float64))[::-1]
        for k, c in enumerate(cut):
            r = rel[:c]
            want = np.sum((np.power(2.0, r) - 1) / np.log2(np.arange(len(r), dtype=np.float64) + 2))
            assert np.isclose(table[u, k], want, rtol=1e-12, atol=1e-12)
    ev = EvaluatorHoldout(test, verbose=False, **kw)
    pruned = test.tolil()
    pruned[:, kw["ignore_items"]] = 0
    n_left = np.diff(pruned.tocsr().indptr)
    want_users = [u for u in range(test.shape[0]) if n_left[u] >= 1 and u not in set(kw["ignore_users"])]
    assert ev.users_to_evaluate == want_users
    with pytest.raises(ValueError):
        EvaluatorHoldout([test], [5])
    with pytest.raises(ValueError):
        EvaluatorHoldout(test, [5, 2000])


# ------------------------------------------------------------------ EvaluatorNegativeItemSample (Evaluator.py:466-578)
eval_negative_case = runpy.run_path(os.path.join(HERE, "golden", "make_golden.py"), run_name="cases")["eval_negative_case"]
ZN = np.load(os.path.join(HERE, "golden", "evaluator_negative_golden.npz"))


def _golden_negative(cutoffs):
    return {c: {k.split("_", 1)[1]: float(ZN[k]) for k in ZN.files if k.startswith("c%d_" % c)} for c in cutoffs}


def test_negative_sample_oracle_matches_reference_golden():
    train, test, neg, S, kw = eval_negative_case()
    _assert_close(evaluate_scores(train, test, S, URM_test_negative=neg, **kw), _golden_negative(kw["cutoff_list"]), 1e-9, "oracle")
