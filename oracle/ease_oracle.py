"""numpy restatement of EASE_R/EASE_R_Recommender.py:55-69 (TEST INFRASTRUCTURE).  The Gram matrix is X^T X with the
diagonal replaced by the stored-entry count per column + l2_norm (:62-63 -- popularity, not sum of squares);
P = inv(G); B = P / (-diag P) column-wise; diag(B) = 0.  fp64 by default (the reference runs np.linalg.inv on float32).
Pinned to the reference's own class by tests/golden/ease_golden.npz."""
import numpy as np
import scipy.sparse as sps


def ease_B(URM, l2_norm=1e3, dtype=np.float64):
    X = sps.csr_matrix(URM, dtype=np.float32)
    G = np.asarray((X.T @ X).todense(), dtype=dtype)
    pop = np.diff(X.tocsc().indptr)
    G[np.diag_indices_from(G)] = pop + l2_norm
    P = np.linalg.inv(G)
    B = P / (-np.diag(P))
    B[np.diag_indices_from(B)] = 0.0
    return B
