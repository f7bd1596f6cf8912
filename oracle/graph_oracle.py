"""fp64 numpy restatement of P3alpha / RP3beta (TEST INFRASTRUCTURE).  Follows GraphBased/P3alphaRecommender.py:47-144
and GraphBased/RP3betaRecommender.py:47-154 step by step (normalize -> power -> Piu*Pui -> degree -> diagonal ->
row top-K -> optional row normalise -> column top-K), with the deterministic tie rule "larger value, then ascending
index" where the reference's argsort order is unspecified.  Pinned against the reference's own classes by
tests/golden/graph_golden.npz (tests/test_oracle_graph.py)."""
import numpy as np
import scipy.sparse as sps


def _row_norm(M):
    M = sps.csr_matrix(M, dtype=np.float64, copy=True)
    s = np.asarray(np.abs(M).sum(axis=1)).ravel()
    s[s == 0] = 1.0
    M.data = M.data / np.repeat(s, np.diff(M.indptr))
    return M


def p3_dense_rows(URM, alpha=1.0, beta=0.0, min_rating=0, implicit=False):
    """Dense (n_items, n_items) fp64 matrix BEFORE any top-K: row i = item i."""
    URM = sps.csr_matrix(URM, dtype=np.float32, copy=True)
    if min_rating > 0:
        URM.data[URM.data < min_rating] = 0
        URM.eliminate_zeros()
        if implicit:
            URM.data = np.ones(URM.data.size, dtype=np.float32)
    Pui = _row_norm(URM)
    Xb = sps.csr_matrix(URM.T, copy=True)
    Xb.data = np.ones(Xb.data.size)
    deg = np.asarray(Xb.sum(axis=1)).ravel()
    Piu = _row_norm(Xb)
    if alpha != 1.0:
        Pui, Piu = Pui.power(alpha), Piu.power(alpha)
    D = np.asarray((Piu @ Pui).todense())
    if beta != 0.0:
        degree = np.zeros(URM.shape[1])
        degree[deg != 0] = np.power(deg[deg != 0], -beta)
        D = D * degree[None, :]
    np.fill_diagonal(D, 0.0)
    return D


def p3_similarity(URM, topK=100, alpha=1.0, beta=0.0, min_rating=0, implicit=False, normalize_similarity=False):
    D = p3_dense_rows(URM, alpha, beta, min_rating, implicit)
    n = D.shape[0]
    k = min(topK, n) if topK is not False else n
    W = np.zeros_like(D)
    for i in range(n):
        order = np.lexsort((np.arange(n), -D[i]))[:k]
        order = order[D[i, order] != 0]
        W[i, order] = D[i, order]
    W = sps.csr_matrix(W)
    if normalize_similarity:
        W = _row_norm(W)
    if topK is not False:
        Wd = W.toarray()
        out = np.zeros_like(Wd)
        for c in range(n):
            col = Wd[:, c]
            nz = np.flatnonzero(col)
            keep = nz[np.lexsort((nz, -col[nz]))][:k]
            out[keep, c] = col[keep]
        W = sps.csr_matrix(out)
    return sps.csr_matrix(W, dtype=np.float32)
