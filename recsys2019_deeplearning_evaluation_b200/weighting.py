"""Host-side mirror of Base/IR_feature_weighting.py (okapi_BM_25 :13-51, TF_IDF :56-78), backed by
b200_feature_weighting_device (csrc/weighting.cu).  Same signatures and assertion texts; like the reference's functions
they take the matrix with the ITEMS ON ROWS (the KNN recommenders pass URM.T, KNN/ItemKNNCFRecommender.py:42-50) and
return a CSR matrix of the same orientation.  The arithmetic runs on the device on the transposed (user-row) layout."""
import ctypes

import numpy as np
import scipy.sparse as sps

from . import _lib

_MODE = {"BM25": 0, "TF-IDF": 1}


def _weight(dataMatrix, mode, K1=1.2, B=0.75):
    import torch
    lib = _lib.load()
    M = sps.csc_matrix(dataMatrix, dtype=np.float32)  # CSC of (items x users) == CSR of the user-row URM
    M.sort_indices()
    n_items, n_users = M.shape
    if M.nnz >= 2 ** 31 - 1:
        raise ValueError("feature weighting: more than 2^31 stored values are not supported")
    dev = torch.device("cuda", torch.cuda.current_device())
    ptr = torch.from_numpy(np.ascontiguousarray(M.indptr, np.int32)).to(dev)
    idx = torch.from_numpy(np.ascontiguousarray(M.indices, np.int32)).to(dev)
    val = torch.from_numpy(np.ascontiguousarray(M.data, np.float32)).to(dev)
    _lib.check(lib.b200_feature_weighting_device(_MODE[mode], n_users, n_items, M.nnz, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                                 float(K1), float(B), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    out = sps.csc_matrix((val.cpu().numpy(), M.indices, M.indptr), shape=M.shape)
    return out.tocsr()


def okapi_BM_25(dataMatrix, K1=1.2, B=0.75):
    """IR_feature_weighting.py:13-51."""
    assert B > 0 and B < 1, "okapi_BM_25: B must be in (0,1)"
    assert K1 > 0, "okapi_BM_25: K1 must be > 0"
    assert np.all(np.isfinite(dataMatrix.data)), \
        "okapi_BM_25: Data matrix contains {} non finite values".format(np.sum(np.logical_not(np.isfinite(dataMatrix.data))))
    return _weight(dataMatrix, "BM25", K1, B)


def TF_IDF(dataMatrix):
    """IR_feature_weighting.py:56-78."""
    assert np.all(np.isfinite(dataMatrix.data)), \
        "TF_IDF: Data matrix contains {} non finite values.".format(np.sum(np.logical_not(np.isfinite(dataMatrix.data))))
    assert np.all(dataMatrix.data >= 0.0), \
        "TF_IDF: Data matrix contains {} negative values, computing the square root is not possible.".format(
            np.sum(dataMatrix.data < 0.0))
    return _weight(dataMatrix, "TF-IDF")
