"""-m gpu: the CUDA path against the golden vectors of the unmodified reference (no /root/reference needed)."""
import numpy as np
import pytest

from golden_util import load_golden, same_sparse, tie_free
from oracle.similarity_oracle import SimilarityOracle, check_topk_against_dense

pytestmark = pytest.mark.gpu
URMS, CASES, KNN = load_golden()


@pytest.mark.parametrize("n", range(len(CASES)))
def test_cuda_matches_reference_golden(n):
    from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity_Cython
    values, kw, W = CASES[n]
    X = URMS[values]
    Wg = Compute_Similarity_Cython(X, **kw).compute_similarity()
    if tie_free(kw, values):
        # bit-exact index sets vs the reference; values within the north_star tolerance
        assert same_sparse(Wg, W["py"], rtol=1e-4)
        if kw["similarity"] not in ("adjusted", "pearson"):
            assert same_sparse(Wg, W["cy"], rtol=1e-4)
    check_topk_against_dense(Wg, SimilarityOracle(X, **kw), np.arange(150), rtol=1e-4)
