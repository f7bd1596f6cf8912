"""Import-time dependency shim for the compiled reference modules in oracle/_ref (TEST INFRASTRUCTURE).

The reference's Cython modules do `from Base.Recommender_utils import check_matrix[, similarityMatrixTopK]`
at import (Compute_Similarity_Cython.pyx:41, SLIM_BPR_Cython_Epoch.pyx:34,
MatrixFactorization_Cython_Epoch.pyx:18).  On the GPU box /root/reference does not exist, so these two
helpers are restated here (behaviour of Base/Recommender_utils.py:13-52 and :55-122).  They only do
sparse-format conversion / per-column top-K; none of the timed reference arithmetic lives here.
Used only when /root/reference is absent (oracle/ref_loader.py).
"""
import numpy as np
import scipy.sparse as sps

_FORMATS = {"csc": (sps.csc_matrix, "tocsc"), "csr": (sps.csr_matrix, "tocsr"), "coo": (sps.coo_matrix, "tocoo")}


def check_matrix(X, format="csc", dtype=np.float32):
    if format in _FORMATS:
        cls, conv = _FORMATS[format]
        if isinstance(X, np.ndarray):
            X = sps.csr_matrix(X, dtype=dtype)
            X.eliminate_zeros()
        if not isinstance(X, cls):
            return getattr(X, conv)().astype(dtype)
        return X.astype(dtype)
    if format == "npy":
        return X.toarray().astype(dtype) if sps.issparse(X) else np.array(X)
    raise ValueError("check_matrix shim: unsupported format %r" % (format,))


def similarityMatrixTopK(item_weights, k=100, verbose=False):
    assert item_weights.shape[0] == item_weights.shape[1]
    n = item_weights.shape[1]
    k = min(k, n)
    dense = isinstance(item_weights, np.ndarray)
    if not dense:
        item_weights = check_matrix(item_weights, "csc", np.float32)
    data, rows, indptr = [], [], [0]
    for c in range(n):
        if dense:
            col = item_weights[:, c]
            idx = np.arange(n, dtype=np.int32)
        else:
            s, e = item_weights.indptr[c], item_weights.indptr[c + 1]
            col, idx = item_weights.data[s:e], item_weights.indices[s:e]
        nz = col != 0
        order = np.argsort(col[nz])[-k:] if k > 0 else np.zeros(0, np.int64)
        data.extend(col[nz][order])
        rows.extend(idx[nz][order])
        indptr.append(len(data))
    return sps.csc_matrix((np.asarray(data, np.float32), np.asarray(rows, np.int32), np.asarray(indptr)),
                          shape=(n, n), dtype=np.float32)
