"""CPU restatement (numpy, fp64) of Base/IR_feature_weighting.py:13-78.  TEST INFRASTRUCTURE ONLY.
Pinned against the reference's own functions through tests/golden/weighting_golden.npz (tests/test_weighting.py)."""
import numpy as np
import scipy.sparse as sps


def okapi_BM_25(dataMatrix, K1=1.2, B=0.75):
    """:13-51 -- items on rows."""
    M = sps.coo_matrix(dataMatrix, dtype=np.float64)
    N = float(M.shape[0])
    idf = np.log(N / (1 + np.bincount(M.col, minlength=M.shape[1])))  # :37
    row_sums = np.ravel(M.sum(axis=1))  # :40
    length_norm = (1.0 - B) + B * row_sums / row_sums.mean()  # :42-43
    den = K1 * length_norm[M.row] + M.data  # :46
    den[den == 0.0] += 1e-9  # :47
    M.data = M.data * (K1 + 1.0) / den * idf[M.col]  # :49
    return M.tocsr()


def TF_IDF(dataMatrix):
    """:56-78 -- items on rows."""
    M = sps.coo_matrix(dataMatrix, dtype=np.float64)
    N = float(M.shape[0])
    idf = np.log(N / (1 + np.bincount(M.col, minlength=M.shape[1])))  # :71
    M.data = np.sqrt(M.data) * idf[M.col]  # :74
    return M.tocsr()
