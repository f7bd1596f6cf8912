"""Generates tests/golden/similarity_*.npz by running the UNMODIFIED reference in this container:
the compiled Cython class from oracle/_ref (oracle/build_ref.py) and the reference's own Python wrappers
imported from /root/reference (KNN/ItemKNNCFRecommender.py, Base/Similarity/Compute_Similarity_Python.py).

    python tests/golden/make_golden.py

The fixtures pin (a) the oracle restatement (tests/test_oracle_similarity.py, CPU) and (b) the CUDA path
(tests/test_golden_gpu.py, -m gpu) on the GPU box, where /root/reference does not exist.
Inputs are regenerated from the seeds by recsys2019_deeplearning_evaluation_b200.synth, and also stored.
"""
import os
import sys

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm  # noqa: E402

CASES = []
for values in ("continuous", "ratings", "binary"):
    for kind in ("cosine", "asymmetric", "adjusted", "pearson", "jaccard", "dice", "tversky", "tanimoto"):
        CASES.append(dict(values=values, similarity=kind, topK=12, shrink=3, normalize=True,
                          asymmetric_alpha=0.35, tversky_alpha=0.8, tversky_beta=1.2))
CASES.append(dict(values="continuous", similarity="cosine", topK=7, shrink=0, normalize=False))
CASES.append(dict(values="continuous", similarity="cosine", topK=7, shrink=25, normalize=False))
CASES.append(dict(values="ratings", similarity="cosine", topK=400, shrink=0, normalize=True))  # topK > n_columns


def main():
    cy = ref_loader.load("Compute_Similarity_Cython").Compute_Similarity_Cython
    assert ref_loader.reference_python_available(), "needs /root/reference"
    from Base.Similarity.Compute_Similarity_Python import Compute_Similarity_Python
    from KNN.ItemKNNCFRecommender import ItemKNNCFRecommender
    out = {}
    urms = {}
    for values in ("continuous", "ratings", "binary"):
        X = synth_urm(400, 150, 0.06, seed=17, values=values)
        urms[values] = X
        out["urm_%s_indptr" % values] = X.indptr
        out["urm_%s_indices" % values] = X.indices
        out["urm_%s_data" % values] = X.data
    meta = []
    for n, c in enumerate(CASES):
        c = dict(c)
        values = c.pop("values")
        X = urms[values]
        Wc = sps.csr_matrix(cy(X, **c).compute_similarity())
        Wp = sps.csr_matrix(Compute_Similarity_Python(X, **c).compute_similarity())
        for tag, W in (("cy", Wc), ("py", Wp)):
            W.sort_indices()
            out["case%d_%s_indptr" % (n, tag)] = W.indptr
            out["case%d_%s_indices" % (n, tag)] = W.indices
            out["case%d_%s_data" % (n, tag)] = W.data
        meta.append(repr(dict(values=values, **c)))
    # the wrapper level: ItemKNNCFRecommender.fit -> W_sparse, and scores of a user block
    rec = ItemKNNCFRecommender(urms["ratings"])
    rec.fit(topK=15, shrink=5, similarity="cosine", normalize=True)
    W = sps.csr_matrix(rec.W_sparse)
    W.sort_indices()
    out["knn_W_indptr"], out["knn_W_indices"], out["knn_W_data"] = W.indptr, W.indices, W.data
    out["knn_scores_users0_40"] = rec._compute_item_score(np.arange(40)).astype(np.float32)
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "similarity_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "similarity_golden.npz"), len(CASES), "cases")


if __name__ == "__main__":
    main()


def make_sgd_golden():
    """tests/golden/sgd_golden.npz: factors / S after 3 epochs of the compiled reference trainers."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_sgd import MF_CASES, SLIM_CASES, _urm
    out = {}
    MF = ref_loader.load("MatrixFactorization_Cython_Epoch").MatrixFactorization_Cython_Epoch
    SL = ref_loader.load("SLIM_BPR_Cython_Epoch").SLIM_BPR_Cython_Epoch
    for n, (algo, kw) in enumerate(MF_CASES):
        m = MF(_urm(), n_factors=16, algorithm_name=algo, learning_rate=0.05, random_seed=42, **kw)
        for _ in range(3):
            m.epochIteration_Cython()
        out["mf%d_U" % n], out["mf%d_V" % n] = m.get_USER_factors(), m.get_ITEM_factors()
    for n, (sym, mode) in enumerate(SLIM_CASES):
        r = SL(_urm(), train_with_sparse_weights=False, learning_rate=0.05, li_reg=1e-3, lj_reg=2e-3, topK=120, symmetric=sym,
               random_seed=7, sgd_mode=mode)
        for _ in range(3):
            r.epochIteration_Cython()
        S = r.get_S()
        out["slim%d_S" % n] = np.asarray(S.toarray() if hasattr(S, "toarray") else S, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "sgd_golden.npz"), **out)
    print("wrote sgd_golden.npz")


if __name__ == "__main__":
    make_sgd_golden()


GRAPH_CASES = [
    dict(cls="P3alpha", topK=10, alpha=1.0, normalize_similarity=False),
    dict(cls="P3alpha", topK=10, alpha=0.7, normalize_similarity=True),
    dict(cls="RP3beta", topK=10, alpha=1.0, beta=0.6, normalize_similarity=True),
    dict(cls="RP3beta", topK=12, alpha=0.5, beta=0.3, normalize_similarity=False),
    dict(cls="RP3beta", topK=10, alpha=1.2, beta=0.6, min_rating=2, implicit=True, normalize_similarity=True),
]


def make_graph_golden():
    """tests/golden/graph_golden.npz: W_sparse of the reference's P3alphaRecommender / RP3betaRecommender."""
    ref_loader.ensure_import_path()
    from GraphBased.P3alphaRecommender import P3alphaRecommender
    from GraphBased.RP3betaRecommender import RP3betaRecommender
    out = {}
    for n, c in enumerate(GRAPH_CASES):
        c = dict(c)
        cls = P3alphaRecommender if c.pop("cls") == "P3alpha" else RP3betaRecommender
        X = synth_urm(400, 150, 0.06, seed=17, values="continuous" if n % 2 == 0 else "ratings")
        r = cls(X)
        r.fit(**c)
        W = sps.csr_matrix(r.W_sparse)
        W.sort_indices()
        out["g%d_indptr" % n], out["g%d_indices" % n], out["g%d_data" % n] = W.indptr, W.indices, W.data
    np.savez_compressed(os.path.join(HERE, "graph_golden.npz"), **out)
    print("wrote graph_golden.npz")


if __name__ == "__main__":
    make_graph_golden()


def make_ease_golden():
    """tests/golden/ease_golden.npz: dense B of the reference's EASE_R_Recommender (fp32 LAPACK inverse)."""
    ref_loader.ensure_import_path()
    ref_loader.load("Compute_Similarity_Cython")
    from EASE_R.EASE_R_Recommender import EASE_R_Recommender
    out = {}
    for n, (values, l2) in enumerate((("binary", 50.0), ("ratings", 500.0))):
        X = synth_urm(600, 200, 0.05, seed=23, values=values)
        r = EASE_R_Recommender(X)
        r.fit(topK=None, l2_norm=l2, verbose=False)
        out["ease%d_B" % n] = np.asarray(r.W_sparse, dtype=np.float32)
        out["ease%d_scores" % n] = np.asarray(r._compute_item_score(np.arange(30)), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "ease_golden.npz"), **out)
    print("wrote ease_golden.npz")


if __name__ == "__main__":
    make_ease_golden()


IALS_CASES = [dict(num_factors=24, confidence_scaling="linear", alpha=2.0, reg=1e-3, values="binary"),
              dict(num_factors=40, confidence_scaling="log", alpha=5.0, epsilon=0.5, reg=1e-2, values="ratings")]


def make_ials_golden():
    """tests/golden/ials_golden.npz: initial item factors + factors after 2 epochs of the reference's IALSRecommender."""
    ref_loader.ensure_import_path()
    from MatrixFactorization.IALSRecommender import IALSRecommender
    out = {}
    for n, c in enumerate(IALS_CASES):
        c = dict(c)
        X = synth_urm(500, 180, 0.05, seed=29, values=c.pop("values"))
        X = X[:, :] .tolil(); X[7, :] = 0; X[:, 11] = 0  # a cold user and a cold item
        X = sps.csr_matrix(X.tocsr(), dtype=np.float32); X.eliminate_zeros()
        np.random.seed(100 + n)
        r = IALSRecommender(X)
        r.fit(epochs=2, **c)
        np.random.seed(100 + n)
        out["ials%d_V0" % n] = c["num_factors"] ** -0.5 * np.random.random_sample((180, c["num_factors"]))
        out["ials%d_U" % n], out["ials%d_V" % n] = r.USER_factors, r.ITEM_factors
    np.savez_compressed(os.path.join(HERE, "ials_golden.npz"), **out)
    print("wrote ials_golden.npz")


if __name__ == "__main__":
    make_ials_golden()


EUCLID_CASES = [dict(values=v, topK=12, shrink=s, normalize=nz, normalize_avg_row=av, similarity_from_distance_mode=m)
                for v, s, nz, av, m in (("continuous", 0, False, False, "lin"), ("continuous", 2, True, False, "exp"),
                                        ("continuous", 1, False, True, "log"), ("ratings", 3, True, True, "lin"),
                                        ("ratings", 0, False, False, "log"), ("binary", 1, False, False, "lin"),
                                        ("binary", 0, True, False, "exp"), ("continuous", 0.5, True, False, "lin"))]
EUCLID_CASES.append(dict(values="ratings", topK=400, shrink=1, normalize=False, normalize_avg_row=False,
                         similarity_from_distance_mode="lin"))  # topK > n_columns


def make_euclid_golden():
    """tests/golden/euclid_golden.npz: W of the reference's Compute_Similarity_Euclidean (fp32 arithmetic) on the same
    three URMs as similarity_golden.npz."""
    ref_loader.ensure_import_path()
    from Base.Similarity.Compute_Similarity_Euclidean import Compute_Similarity_Euclidean
    out = {}
    meta = []
    for n, c in enumerate(EUCLID_CASES):
        c = dict(c)
        values = c.pop("values")
        X = synth_urm(400, 150, 0.06, seed=17, values=values)
        W = sps.csr_matrix(Compute_Similarity_Euclidean(X, **c).compute_similarity())
        W.sort_indices()
        out["eu%d_indptr" % n], out["eu%d_indices" % n], out["eu%d_data" % n] = W.indptr, W.indices, W.data
        meta.append(repr(dict(values=values, **c)))
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "euclid_golden.npz"), **out)
    print("wrote euclid_golden.npz", len(EUCLID_CASES), "cases")


if __name__ == "__main__":
    make_euclid_golden()


def make_weighting_golden():
    """tests/golden/weighting_golden.npz: the reference's okapi_BM_25 / TF_IDF on URM.T of the three golden URMs, and
    ItemKNNCFRecommender / UserKNNCFRecommender W_sparse fitted with feature_weighting."""
    ref_loader.ensure_import_path()
    ref_loader.load("Compute_Similarity_Cython")
    from Base.IR_feature_weighting import okapi_BM_25, TF_IDF
    from KNN.ItemKNNCFRecommender import ItemKNNCFRecommender
    from KNN.UserKNNCFRecommender import UserKNNCFRecommender
    out = {}
    for values in ("continuous", "ratings", "binary"):
        X = synth_urm(400, 150, 0.06, seed=17, values=values)
        for tag, fn in (("bm25", okapi_BM_25), ("tfidf", TF_IDF)):
            W = sps.csr_matrix(fn(X.T.astype(np.float32)))
            W.sort_indices()
            out["%s_%s_indptr" % (tag, values)], out["%s_%s_indices" % (tag, values)] = W.indptr, W.indices
            out["%s_%s_data" % (tag, values)] = W.data.astype(np.float64)
    X = synth_urm(400, 150, 0.06, seed=17, values="ratings")
    for tag, cls in (("item", ItemKNNCFRecommender), ("user", UserKNNCFRecommender)):
        for fw in ("BM25", "TF-IDF"):
            r = cls(X.copy())
            r.fit(topK=10, shrink=2, similarity="cosine", feature_weighting=fw)
            W = sps.csr_matrix(r.W_sparse)
            W.sort_indices()
            k = "knn_%s_%s" % (tag, fw.replace("-", ""))
            out[k + "_indptr"], out[k + "_indices"], out[k + "_data"] = W.indptr, W.indices, W.data
    np.savez_compressed(os.path.join(HERE, "weighting_golden.npz"), **out)
    print("wrote weighting_golden.npz")


if __name__ == "__main__":
    make_weighting_golden()


def eval_case(n):
    """Inputs of evaluator case n (shared with tests/test_evaluation.py): train/test URMs, dense tie-free scores, kwargs."""
    rng = np.random.default_rng(300 + n)
    train = synth_urm(300, 120, 0.06, seed=41 + n, values="ratings").tolil()
    train[0:4, 0:112] = 1.0  # users whose unseen items do not fill the longest list
    train = sps.csr_matrix(train.tocsr(), dtype=np.float32)
    test = synth_urm(300, 120, 0.04, seed=51 + n, values="ratings").tolil()
    test[10:20, :] = 0  # users without test interactions
    test = sps.csr_matrix(test.tocsr(), dtype=np.float32)
    test.eliminate_zeros()
    S = rng.random((300, 120)).astype(np.float32)
    kw = [dict(cutoff_list=[5, 10, 20]),
          dict(cutoff_list=[1, 7], min_ratings_per_user=3, exclude_seen=False),
          dict(cutoff_list=[10], ignore_items=[3, 17, 44, 90], ignore_users=[1, 25, 26, 200])][n]
    return train, test, S, kw


def make_evaluator_golden():
    """tests/golden/evaluator_golden.npz: results of the reference's EvaluatorHoldout on a stub recommender that returns a
    fixed dense score matrix."""
    ref_loader.ensure_import_path()
    from Base.BaseRecommender import BaseRecommender
    from Base.Evaluation.Evaluator import EvaluatorHoldout

    class Stub(BaseRecommender):
        RECOMMENDER_NAME = "Stub"

        def __init__(self, URM_train, S):
            super(Stub, self).__init__(URM_train)
            self.S = S

        def _compute_item_score(self, user_id_array, items_to_compute=None):
            return self.S[np.asarray(user_id_array)].astype(np.float32).copy()

    out = {}
    for n in range(3):
        train, test, S, kw = eval_case(n)
        res, _ = EvaluatorHoldout(test, **kw).evaluateRecommender(Stub(train, S))
        for c, d in res.items():
            for k, v in d.items():
                out["e%d_c%d_%s" % (n, c, k)] = np.float64(v)
    np.savez_compressed(os.path.join(HERE, "evaluator_golden.npz"), **out)
    print("wrote evaluator_golden.npz")


if __name__ == "__main__":
    make_evaluator_golden()


def make_dataio_golden():
    """tests/golden/dataio_ref.zip: an archive written by the reference's own DataIO (Base/DataIO.py) holding every
    member type, so that the format-compatible reader is pinned without /root/reference."""
    ref_loader.ensure_import_path()
    from Base.DataIO import DataIO
    import pandas as pd
    rng = np.random.default_rng(5)
    d = {"W_sparse": sps.random(30, 30, 0.1, format="csr", random_state=3, dtype=np.float32),
         "USER_factors": rng.random((7, 4)), "use_bias": False, "topK": np.int64(50), "name": "x",
         "mapper": {3: "a", 9: "b"}, "nested": {"A": rng.random(3), "k": 2},
         "frame": pd.DataFrame({"a": [1, 2], "b": [0.5, 1.5]})}
    DataIO(HERE + "/").save_data("dataio_ref", d)
    print("wrote dataio_ref.zip")


if __name__ == "__main__":
    make_dataio_golden()


def knn_variant_inputs():
    """URM, ICM (items x features), UCM (users x features) of the KNN-variant golden cases (shared with tests/test_z_knn_variants.py)."""
    URM = synth_urm(400, 150, 0.06, seed=17, values="ratings")
    ICM = synth_urm(150, 60, 0.10, seed=71, values="continuous")
    UCM = synth_urm(400, 45, 0.12, seed=72, values="ratings")
    return URM, ICM, UCM


KNN_VARIANT_CASES = [("ItemKNNCBFRecommender", dict(topK=8, shrink=1, similarity="cosine", feature_weighting="none")),
                     ("ItemKNNCBFRecommender", dict(topK=8, shrink=0, similarity="asymmetric", asymmetric_alpha=0.4, feature_weighting="TF-IDF")),
                     ("UserKNNCBFRecommender", dict(topK=9, shrink=2, similarity="cosine", feature_weighting="BM25")),
                     ("ItemKNN_CFCBF_Hybrid_Recommender", dict(ICM_weight=2.5, topK=10, shrink=3, similarity="cosine")),
                     ("UserKNN_CFCBF_Hybrid_Recommender", dict(UCM_weight=0.5, topK=7, shrink=1, similarity="cosine"))]


def make_knn_variant_golden():
    """tests/golden/knn_variants_golden.npz: W_sparse of the reference's CBF / hybrid KNN recommenders and the scores of
    ItemKNNCustomSimilarityRecommender with selectTopK."""
    ref_loader.ensure_import_path()
    ref_loader.load("Compute_Similarity_Cython")
    import importlib
    URM, ICM, UCM = knn_variant_inputs()
    out = {}
    for n, (name, kw) in enumerate(KNN_VARIANT_CASES):
        cls = getattr(importlib.import_module("KNN." + name), name)
        r = cls(URM.copy(), (ICM if name.startswith("Item") else UCM).copy())
        r.fit(**kw)
        W = sps.csr_matrix(r.W_sparse)
        W.sort_indices()
        out["k%d_indptr" % n], out["k%d_indices" % n], out["k%d_data" % n] = W.indptr, W.indices, W.data
        out["k%d_scores" % n] = r._compute_item_score(np.arange(25)).astype(np.float32)
    from KNN.ItemKNNCustomSimilarityRecommender import ItemKNNCustomSimilarityRecommender
    Wc = sps.random(150, 150, 0.2, format="csr", random_state=9, dtype=np.float32)
    r = ItemKNNCustomSimilarityRecommender(URM.copy())
    r.fit(Wc, selectTopK=True, topK=6)
    W = sps.csr_matrix(r.W_sparse)
    W.sort_indices()
    out["custom_indptr"], out["custom_indices"], out["custom_data"] = W.indptr, W.indices, W.data
    np.savez_compressed(os.path.join(HERE, "knn_variants_golden.npz"), **out)
    print("wrote knn_variants_golden.npz")


if __name__ == "__main__":
    make_knn_variant_golden()


def eval_negative_case():
    """Inputs of the negative-item-sample evaluator golden case: eval_case(0) plus 30 sampled negatives per user."""
    train, test, S, _ = eval_case(0)
    rng = np.random.default_rng(77)
    neg = sps.lil_matrix(test.shape, dtype=np.float32)
    seen = (train + test).tocsr()
    for u in range(test.shape[0]):
        cand = np.setdiff1d(np.arange(test.shape[1]), seen.indices[seen.indptr[u]:seen.indptr[u + 1]])
        neg[u, rng.choice(cand, size=min(30, len(cand)), replace=False)] = 1.0
    return train, test, sps.csr_matrix(neg.tocsr(), dtype=np.float32), S, dict(cutoff_list=[1, 5, 10])


def make_evaluator_negative_golden():
    """tests/golden/evaluator_negative_golden.npz: the reference's EvaluatorNegativeItemSample on the stub recommender."""
    ref_loader.ensure_import_path()
    from Base.BaseRecommender import BaseRecommender
    from Base.Evaluation.Evaluator import EvaluatorNegativeItemSample

    class Stub(BaseRecommender):
        RECOMMENDER_NAME = "Stub"

        def __init__(self, URM_train, S):
            super(Stub, self).__init__(URM_train)
            self.S = S

        def _compute_item_score(self, user_id_array, items_to_compute=None):
            s = self.S[np.asarray(user_id_array)].astype(np.float32).copy()
            if items_to_compute is not None:  # BaseSimilarityMatrixRecommender.py:84-90
                out = -np.ones_like(s) * np.inf
                out[:, items_to_compute] = s[:, items_to_compute]
                s = out
            return s

    train, test, neg, S, kw = eval_negative_case()
    res, _ = EvaluatorNegativeItemSample(test, neg, **kw).evaluateRecommender(Stub(train, S))
    out = {}
    for c, d in res.items():
        for k, v in d.items():
            out["c%d_%s" % (c, k)] = np.float64(v)
    np.savez_compressed(os.path.join(HERE, "evaluator_negative_golden.npz"), **out)
    print("wrote evaluator_negative_golden.npz")


if __name__ == "__main__":
    make_evaluator_negative_golden()
