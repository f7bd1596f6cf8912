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
