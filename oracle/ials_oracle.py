"""numpy restatement of MatrixFactorization/IALSRecommender.py:137-201 (TEST INFRASTRUCTURE): one epoch = user loop
with VV = V^T V then item loop with UU = U^T U, each row solved through the explicit inverse like :201.  fp64.
Pinned to the reference's own class by tests/golden/ials_golden.npz."""
import numpy as np
import scipy.sparse as sps


def confidence(URM, scaling="linear", alpha=1.0, epsilon=1.0):
    C = sps.csr_matrix(URM, dtype=np.float32, copy=True)
    C.data = (1.0 + alpha * C.data) if scaling == "linear" else (1.0 + alpha * np.log(1.0 + C.data / epsilon))
    return sps.csr_matrix(C, dtype=np.float32)


def update_row(profile, conf, Y, YtY, reg):
    Yi = Y[profile, :]
    A = Yi.T.dot(((conf - 1) * Yi.T).T)
    B = YtY + A + np.diag(reg * np.ones(Y.shape[1]))
    return np.dot(np.linalg.inv(B), Yi.T.dot(conf))


def run_epoch(C, U, V, reg):
    Ct = sps.csc_matrix(C, dtype=np.float32)
    VV = V.T.dot(V)
    for u in np.flatnonzero(np.diff(C.indptr) > 0):
        s, e = C.indptr[u], C.indptr[u + 1]
        U[u] = update_row(C.indices[s:e], C.data[s:e], V, VV, reg)
    UU = U.T.dot(U)
    for i in np.flatnonzero(np.diff(Ct.indptr) > 0):
        s, e = Ct.indptr[i], Ct.indptr[i + 1]
        V[i] = update_row(Ct.indices[s:e], Ct.data[s:e], U, UU, reg)
    return U, V
