"""The two helpers of Base/Recommender_utils.py the recommenders lean on, under their reference names:

    check_matrix(X, format='csc', dtype=np.float32)        :13-52   host-side format / dtype normalisation
    similarityMatrixTopK(item_weights, k=100, verbose)     :55-122  per-column top-k of a square matrix -> CSC float32

similarityMatrixTopK runs on the device for both input kinds: sparse matrices through the compressed-line selection
(`b200_sparse_topk_device`, graph.sparse_column_topk), dense ndarrays / CUDA tensors through the dense-line selection
(`b200_dense_topk_device`, slim_bpr_epoch.similarityMatrixTopK).  Ties resolve to the ascending row index."""
import numpy as np
import scipy.sparse as sps

_FORMATS = {"csc": (sps.csc_matrix, "tocsc"), "csr": (sps.csr_matrix, "tocsr"), "coo": (sps.coo_matrix, "tocoo"),
            "dok": (sps.dok_matrix, "todok"), "bsr": (sps.bsr_matrix, "tobsr"), "dia": (sps.dia_matrix, "todia"),
            "lil": (sps.lil_matrix, "tolil")}


def check_matrix(X, format="csc", dtype=np.float32):
    """:13-52.  A matrix already in the requested sparse format is returned with the dtype applied; another sparse format
    is converted; 'npy' densifies a sparse input; an ndarray becomes sparse (explicit zeros dropped) unless 'npy' is asked."""
    if format in _FORMATS:
        cls, conv = _FORMATS[format]
        if isinstance(X, np.ndarray):  # :47-50
            S = sps.csr_matrix(X, dtype=dtype)
            S.eliminate_zeros()
            return check_matrix(S, format=format, dtype=dtype)
        if not isinstance(X, cls):
            return getattr(X, conv)().astype(dtype)
        return X.astype(dtype)
    if format == "npy":  # :42-46
        return X.toarray().astype(dtype) if sps.issparse(X) else np.array(X)
    if isinstance(X, np.ndarray):
        S = sps.csr_matrix(X, dtype=dtype)
        S.eliminate_zeros()
        return S
    return X.astype(dtype)


def similarityMatrixTopK(item_weights, k=100, verbose=False):
    """:55-122."""
    assert item_weights.shape[0] == item_weights.shape[1], "selectTopK: ItemWeights is not a square matrix"
    if sps.issparse(item_weights):
        from .graph import sparse_column_topk
        return sps.csc_matrix(sparse_column_topk(check_matrix(item_weights, "csc", np.float32), k), dtype=np.float32)
    from .slim_bpr_epoch import similarityMatrixTopK as _dense
    return _dense(item_weights, k=k, verbose=verbose)
