import ast
import os

import numpy as np
import scipy.sparse as sps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "similarity_golden.npz")


def load_golden():
    z = np.load(GOLDEN, allow_pickle=False)
    urms = {}
    for v in ("continuous", "ratings", "binary"):
        n_items = int(z["urm_%s_indices" % v].max()) + 1
        urms[v] = sps.csr_matrix((z["urm_%s_data" % v], z["urm_%s_indices" % v], z["urm_%s_indptr" % v]),
                                 shape=(len(z["urm_%s_indptr" % v]) - 1, 150), dtype=np.float32)
    cases = []
    for n, m in enumerate(z["meta"]):
        kw = ast.literal_eval(str(m))
        values = kw.pop("values")
        W = {}
        for tag in ("cy", "py"):
            W[tag] = sps.csr_matrix((z["case%d_%s_data" % (n, tag)], z["case%d_%s_indices" % (n, tag)],
                                     z["case%d_%s_indptr" % (n, tag)]), shape=(150, 150), dtype=np.float32)
        cases.append((values, kw, W))
    knn = dict(W=sps.csr_matrix((z["knn_W_data"], z["knn_W_indices"], z["knn_W_indptr"]), shape=(150, 150)),
               scores=z["knn_scores_users0_40"])
    return urms, cases, knn


def tie_free(kw, values):
    """Cases whose top-K boundary has no exact ties: there the reference's index sets are reproducible."""
    return values == "continuous" and kw["similarity"] not in ("jaccard", "dice", "tversky", "tanimoto")


def same_sparse(A, B, rtol=1e-4, atol=1e-7):
    A = sps.csr_matrix(A); B = sps.csr_matrix(B)
    A.sort_indices(); B.sort_indices()
    if A.nnz != B.nnz or (A.indptr != B.indptr).any() or (A.indices != B.indices).any():
        return False
    return bool(np.allclose(A.data, B.data, rtol=rtol, atol=atol))


def load_euclid_golden():
    """[(values, kwargs, W)] of tests/golden/euclid_golden.npz (reference Compute_Similarity_Euclidean outputs)."""
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "euclid_golden.npz"), allow_pickle=False)
    cases = []
    for n, m in enumerate(z["meta"]):
        kw = ast.literal_eval(str(m))
        values = kw.pop("values")
        W = sps.csr_matrix((z["eu%d_data" % n], z["eu%d_indices" % n], z["eu%d_indptr" % n]), shape=(150, 150), dtype=np.float32)
        cases.append((values, kw, W))
    return cases
