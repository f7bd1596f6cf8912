"""Host-side mirrors of the content-based / hybrid / custom-similarity KNN recommenders (SURVEY.md section 2 row 2,
section 8 a7): thin callers of the same similarity kernel as ItemKNNCF / UserKNNCF, on the item-content matrix, the
user-content matrix, or the content matrix stacked with the interactions.

    ItemKNNCBFRecommender               KNN/ItemKNNCBFRecommender.py:17-49          similarity of the columns of ICM^T
    UserKNNCBFRecommender               KNN/UserKNNCBFRecommender.py:17-49          similarity of the columns of UCM^T
    ItemKNN_CFCBF_Hybrid_Recommender    KNN/ItemKNN_CFCBF_Hybrid_Recommender.py     [ICM * w | URM^T]
    UserKNN_CFCBF_Hybrid_Recommender    KNN/UserKNN_CFCBF_Hybrid_Recommender.py     [UCM * w | URM]
    ItemKNNCustomSimilarityRecommender  KNN/ItemKNNCustomSimilarityRecommender.py   a caller-supplied W_sparse (+ column top-K)

Same constructor / fit() arguments, assertion and error texts as the reference; feature weighting runs on the device
(weighting.py).  The CBF bases are Base/BaseCBFRecommender.py:15-66."""
import numpy as np
import scipy.sparse as sps

from .recommenders import BaseItemSimilarityMatrixRecommender, BaseUserSimilarityMatrixRecommender
from .similarity import Compute_Similarity

_FW = ["BM25", "TF-IDF", "none"]


def _weighted(M, feature_weighting, values):
    if feature_weighting not in values:
        raise ValueError("Value for 'feature_weighting' not recognized. Acceptable values are {}, provided was '{}'".format(
            values, feature_weighting))
    if feature_weighting == "none":
        return M
    from .weighting import okapi_BM_25, TF_IDF
    fn = okapi_BM_25 if feature_weighting == "BM25" else TF_IDF
    return sps.csr_matrix(fn(M.astype(np.float32)), dtype=np.float32)  # the content matrix has the items / users on rows already


def _fit_similarity(rec, content, topK, shrink, similarity, normalize, similarity_args):
    rec.topK, rec.shrink = topK, shrink
    sim = Compute_Similarity(sps.csr_matrix(content.T, dtype=np.float32), shrink=shrink, topK=topK, normalize=normalize,
                             similarity=similarity, **similarity_args)
    rec.W_sparse = sps.csr_matrix(sim.compute_similarity(), dtype=np.float32)
    sim.compute_similarity_object._dealloc()


class ItemKNNCBFRecommender(BaseItemSimilarityMatrixRecommender):
    RECOMMENDER_NAME = "ItemKNNCBFRecommender"
    FEATURE_WEIGHTING_VALUES = _FW

    def __init__(self, URM_train, ICM_train, verbose=True):
        super(ItemKNNCBFRecommender, self).__init__(URM_train, verbose=verbose)
        assert self.n_items == ICM_train.shape[0], "{}: URM_train has {} items but ICM_train has {}".format(
            self.RECOMMENDER_NAME, self.n_items, ICM_train.shape[0])  # BaseCBFRecommender.py:24
        self.ICM_train = sps.csr_matrix(ICM_train.copy(), dtype=np.float32)
        self.ICM_train.eliminate_zeros()
        _, self.n_features = self.ICM_train.shape
        self._cold_item_CBF_mask = np.ediff1d(self.ICM_train.indptr) == 0
        if self._cold_item_CBF_mask.any():
            self._print("ICM Detected {} ({:.2f} %) items with no features.".format(
                self._cold_item_CBF_mask.sum(), self._cold_item_CBF_mask.sum() / self.n_items * 100))

    def fit(self, topK=50, shrink=100, similarity="cosine", normalize=True, feature_weighting="none", **similarity_args):
        self.ICM_train = _weighted(self.ICM_train, feature_weighting, self.FEATURE_WEIGHTING_VALUES)
        _fit_similarity(self, self.ICM_train, topK, shrink, similarity, normalize, similarity_args)


class ItemKNN_CFCBF_Hybrid_Recommender(ItemKNNCBFRecommender):
    RECOMMENDER_NAME = "ItemKNN_CFCBF_HybridRecommender"

    def fit(self, ICM_weight=1.0, **fit_args):
        self.ICM_train = sps.hstack([self.ICM_train * ICM_weight, self.URM_train.T], format="csr")  # :22-23
        super(ItemKNN_CFCBF_Hybrid_Recommender, self).fit(**fit_args)

    def _get_cold_item_mask(self):
        return np.logical_and(self._cold_item_CBF_mask, self._cold_item_mask)


class UserKNNCBFRecommender(BaseUserSimilarityMatrixRecommender):
    RECOMMENDER_NAME = "UserKNNCBFRecommender"
    FEATURE_WEIGHTING_VALUES = _FW

    def __init__(self, URM_train, UCM_train, verbose=True):
        super(UserKNNCBFRecommender, self).__init__(URM_train, verbose=verbose)
        assert self.n_users == UCM_train.shape[0], "{}: URM_train has {} users but UCM_train has {}".format(
            self.RECOMMENDER_NAME, self.n_items, UCM_train.shape[0])  # BaseCBFRecommender.py:50 (the text prints n_items)
        self.UCM_train = sps.csr_matrix(UCM_train.copy(), dtype=np.float32)
        self.UCM_train.eliminate_zeros()
        _, self.n_features = self.UCM_train.shape
        self._cold_user_CBF_mask = np.ediff1d(self.UCM_train.indptr) == 0
        if self._cold_user_CBF_mask.any():
            self._print("UCM Detected {} ({:.2f} %) cold users.".format(
                self._cold_user_CBF_mask.sum(), self._cold_user_CBF_mask.sum() / self.n_users * 100))

    def fit(self, topK=50, shrink=100, similarity="cosine", normalize=True, feature_weighting="none", **similarity_args):
        self.UCM_train = _weighted(self.UCM_train, feature_weighting, self.FEATURE_WEIGHTING_VALUES)
        _fit_similarity(self, self.UCM_train, topK, shrink, similarity, normalize, similarity_args)


class UserKNN_CFCBF_Hybrid_Recommender(UserKNNCBFRecommender):
    RECOMMENDER_NAME = "UserKNN_CFCBF_Hybrid_Recommender"

    def fit(self, UCM_weight=1.0, **fit_args):
        self.UCM_train = sps.hstack([self.UCM_train * UCM_weight, self.URM_train], format="csr")  # :21-22
        super(UserKNN_CFCBF_Hybrid_Recommender, self).fit(**fit_args)

    def _get_cold_user_mask(self):
        return np.logical_and(self._cold_user_CBF_mask, self._cold_user_mask)


class ItemKNNCustomSimilarityRecommender(BaseItemSimilarityMatrixRecommender):
    RECOMMENDER_NAME = "ItemKNNCustomSimilarityRecommender"

    def fit(self, W_sparse, selectTopK=False, topK=100):
        assert W_sparse.shape[0] == W_sparse.shape[1], \
            "ItemKNNCustomSimilarityRecommender: W_sparse matrice is not square. Current shape is {}".format(W_sparse.shape)
        assert self.URM_train.shape[1] == W_sparse.shape[0], \
            "ItemKNNCustomSimilarityRecommender: URM_train and W_sparse matrices are not consistent. " \
            "The number of columns in URM_train must be equal to the rows in W_sparse. " \
            "Current shapes are: URM_train {}, W_sparse {}".format(self.URM_train.shape, W_sparse.shape)
        if selectTopK:  # similarityMatrixTopK (Base/Recommender_utils.py:55-122) on the device
            from .graph import sparse_column_topk
            W_sparse = sparse_column_topk(sps.csr_matrix(W_sparse, dtype=np.float32), topK)
        self.W_sparse = sps.csr_matrix(W_sparse, dtype=np.float32)
