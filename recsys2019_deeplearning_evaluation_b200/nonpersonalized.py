"""TopPop (Base/NonPersonalizedRecommender.py:14-64): the non-personalised baseline every experiment of the reference
starts with.  No hot loop of its own -- item popularity is an O(nnz) count; it is mirrored so that the usage example and
the evaluator run end to end on the device (its score block is one broadcast row)."""
import numpy as np

from .recommenders import BaseRecommender


class TopPop(BaseRecommender):
    RECOMMENDER_NAME = "TopPopRecommender"

    def fit(self):
        # np.ediff1d on the CSC pointer and NOT a sum: there may be values other than 0/1 (NonPersonalizedRecommender.py:25-26)
        self.item_pop = np.ediff1d(self.URM_train.tocsc().indptr)
        self._d_pop = None

    def _scores_device(self, d_users, items_to_compute=None):
        import torch
        if getattr(self, "_d_pop", None) is None:
            self._d_pop = torch.from_numpy(np.ascontiguousarray(self.item_pop, np.float32)).to(d_users.device)
        return self._d_pop.unsqueeze(0).expand(d_users.shape[0], self.n_items).contiguous()

    def _model_dict(self):
        return {"item_pop": self.item_pop}  # :52

    def _model_loaded(self):
        self._d_pop = None
