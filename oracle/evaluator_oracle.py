"""CPU restatement (numpy, per-user loops) of the reference's hold-out evaluation: Base/Evaluation/Evaluator.py:152-461
and the metric definitions of Base/Evaluation/metrics.py.  TEST INFRASTRUCTURE ONLY.

Input is a dense score matrix (what a recommender's _compute_item_score returns); ranking follows
Base/BaseRecommender.py:164-207 (seen / ignored items to -inf, best first, -inf entries dropped).  Pinned against the
reference's own EvaluatorHoldout through tests/golden/evaluator_golden.npz (tests/test_evaluation.py)."""
import numpy as np
import scipy.sparse as sps


def _dcg(scores):  # metrics.py:277-279
    return np.sum((np.power(2.0, scores) - 1) / np.log2(np.arange(len(scores), dtype=np.float64) + 2))


def evaluate_scores(URM_train, URM_test, S, cutoff_list, min_ratings_per_user=1, exclude_seen=True, ignore_items=None,
                    ignore_users=None, URM_test_negative=None):
    """URM_test_negative given: EvaluatorNegativeItemSample (Evaluator.py:466-578) -- every user is ranked over her test items
    plus her negative items only; otherwise EvaluatorHoldout."""
    URM_train = sps.csr_matrix(URM_train)
    T = sps.csr_matrix(URM_test, dtype=np.float64).copy()
    T.eliminate_zeros()  # metrics.py:375 (in place on the evaluator's own matrix, before any user is scored)
    T.sort_indices()
    n_users, n_items = T.shape
    ignore_items = np.array([] if ignore_items is None else ignore_items, dtype=np.int64)
    ignore_users_arr = np.array([] if ignore_users is None else ignore_users, dtype=np.int64)
    pruned = sps.csc_matrix(T.copy())  # Evaluator.py:137-152, :199
    for i in ignore_items:
        pruned.data[pruned.indptr[i]:pruned.indptr[i + 1]] = 0
    pruned.eliminate_zeros()
    users = np.flatnonzero(np.ediff1d(sps.csr_matrix(pruned).indptr) >= min_ratings_per_user)  # :204-214
    users = np.array(sorted(set(users.tolist()) - set(ignore_users_arr.tolist())), dtype=np.int64)  # :216-222
    rank_items = None
    if URM_test_negative is not None:  # Evaluator.py:497-499
        rank_items = sps.csr_matrix(T.astype(bool)) + sps.csr_matrix(sps.csr_matrix(URM_test_negative).astype(bool))
        rank_items.eliminate_zeros()
    pop = np.ediff1d(sps.csc_matrix(URM_train).indptr).astype(np.float64)
    n_inter = pop.sum()
    pop_norm = pop / pop.max()
    max_cutoff = max(cutoff_list)
    acc = {c: dict(PRECISION=0.0, PRECISION_RECALL_MIN_DEN=0.0, RECALL=0.0, MAP=0.0, MAP_MIN_DEN=0.0, MRR=0.0, NDCG=0.0,
                   HIT_RATE=0.0, ARHR_ALL_HITS=0.0, NOVELTY=0.0, AVERAGE_POPULARITY=0.0, with_recs=0.0) for c in cutoff_list}
    rec_cnt = {c: np.zeros(n_items) for c in cutoff_list}
    hit_cnt = {c: np.zeros(n_items) for c in cutoff_list}
    for u in users:
        s = np.array(S[u], dtype=np.float64)
        if exclude_seen:
            s[URM_train.indices[URM_train.indptr[u]:URM_train.indptr[u + 1]]] = -np.inf  # BaseRecommender.py:166-169
        s[ignore_items] = -np.inf  # :192-193
        if rank_items is not None:  # Evaluator.py:555-563, BaseSimilarityMatrixRecommender.py:84-90: the other items score -inf
            keep = np.zeros(n_items, bool)
            keep[rank_items.indices[rank_items.indptr[u]:rank_items.indptr[u + 1]]] = True
            s[~keep] = -np.inf
        order = np.lexsort((np.arange(n_items), -s))[:max_cutoff]
        rec = order[np.isfinite(s[order])]  # :203-207
        rel_items = T.indices[T.indptr[u]:T.indptr[u + 1]]
        rel_rating = T.data[T.indptr[u]:T.indptr[u + 1]]
        is_rel = np.in1d(rec, rel_items, assume_unique=True)  # Evaluator.py:329
        it2rel = dict(zip(rel_items.tolist(), rel_rating.tolist()))
        for c in cutoff_list:
            a = acc[c]
            r, ir = rec[:c], is_rel[:c]
            L = len(ir)
            if L:
                a["PRECISION"] += ir.sum() / L  # metrics.py:214-222
                a["PRECISION_RECALL_MIN_DEN"] += ir.sum() / min(len(rel_items), L)  # :225-234
                p_at_k = ir * np.cumsum(ir, dtype=np.float64) / (1 + np.arange(L))  # :65-76
                a["MAP"] += p_at_k.sum() / L
                a["MAP_MIN_DEN"] += p_at_k.sum() / min(len(rel_items), L)  # :106-116
                a["AVERAGE_POPULARITY"] += pop_norm[r].sum() / L  # :693-698
                pr = pop[r] / n_inter
                a["NOVELTY"] += np.sum(-np.log2(pr[pr != 0]) / n_items)  # :642-651
                a["with_recs"] += 1
            a["RECALL"] += ir.sum() / len(rel_items)  # :237-243
            ranks = np.arange(1, L + 1)[ir]
            a["MRR"] += 1.0 / ranks[0] if len(ranks) else 0.0  # :146-158
            a["HIT_RATE"] += float(ir.any())
            a["ARHR_ALL_HITS"] += float(ir.dot(1 / np.arange(1, L + 1, 1.0))) if L else 0.0  # :200-210
            rank_dcg = _dcg(np.array([it2rel.get(int(it), 0.0) for it in r], dtype=np.float64))  # :247-273
            ideal = _dcg(np.sort(rel_rating)[::-1][:c])
            a["NDCG"] += rank_dcg / ideal if rank_dcg != 0.0 and ideal != 0.0 else 0.0
            rec_cnt[c][r] += 1
            hit_cnt[c][r[ir]] += 1
    n_eval = len(users)
    keep = np.ones(n_items, dtype=bool)
    keep[ignore_items] = False
    gt_items = np.ediff1d(sps.csc_matrix(T).indptr) > 0
    gt_items[ignore_items] = False
    gt_users = np.ediff1d(T.indptr) > 0
    gt_users[ignore_users_arr] = False
    out = {}
    for c in cutoff_list:
        a = acc[c]
        r = {k: a[k] / n_eval for k in a if k != "with_recs"}
        p_, r_ = r["PRECISION"], r["RECALL"]
        r["F1"] = 2 * p_ * r_ / (p_ + r_) if p_ + r_ != 0 else 0.0
        cnt = rec_cnt[c][keep]
        tot = cnt.sum()
        n = len(cnt)
        r["COVERAGE_ITEM"] = (cnt > 0).sum() / n
        r["COVERAGE_ITEM_HIT"] = (hit_cnt[c][keep] > 0).sum() / n
        r["ITEMS_IN_GT"] = gt_items.sum() / (n_items - len(ignore_items))
        r["COVERAGE_USER"] = a["with_recs"] / (n_users - len(ignore_users_arr))
        r["COVERAGE_USER_HIT"] = a["HIT_RATE"] / (n_users - len(ignore_users_arr))
        r["USERS_IN_GT"] = gt_users.sum() / (n_users - len(ignore_users_arr))
        r["DIVERSITY_GINI"] = 2 * np.sum((n + 1 - np.arange(1, n + 1)) / (n + 1) * np.sort(cnt) / tot)
        r["DIVERSITY_HERFINDAHL"] = 1 - np.sum((cnt / tot) ** 2)
        pr = cnt[cnt > 0] / tot
        r["SHANNON_ENTROPY"] = -np.sum(pr * np.log2(pr))
        couples = n_eval ** 2 - n_eval
        r["DIVERSITY_MEAN_INTER_LIST"] = (couples - (np.sum(rec_cnt[c] ** 2) - n_eval * c) / c) / couples
        out[c] = {k: float(v) for k, v in r.items()}
    return out
