"""-m gpu: the euclidean similarity on the CUDA path (b200_sim_create_euclidean) against the reference's golden W
(tests/golden/euclid_golden.npz) and the fp64 oracle.  Every column is a candidate here (co-rated or not)."""
import numpy as np
import pytest
import scipy.sparse as sps

from golden_util import load_euclid_golden, load_golden, same_sparse
from oracle.similarity_oracle import EuclideanOracle, check_topk_against_dense
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

pytestmark = pytest.mark.gpu
URMS = load_golden()[0]
CASES = load_euclid_golden()
RTOL = 1e-4


def _cls():
    from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity_Euclidean
    return Compute_Similarity_Euclidean


def _check(X, cols=None, debug_cap=None, **kw):
    sim = _cls()(X, **kw)
    if debug_cap is not None:
        from recsys2019_deeplearning_evaluation_b200 import _lib
        _lib.check(_lib.load().b200_sim_debug_set_cap(sim._h, debug_cap))
    W = sim.compute_similarity()
    assert sps.isspmatrix_csr(W) and W.dtype == np.float32 and W.shape == (X.shape[1], X.shape[1])
    assert (W.data > 0).all() and W.diagonal().sum() == 0
    cols = np.arange(X.shape[1]) if cols is None else cols
    # every column but the target has a finite distance: exactly min(K, n - 1) neighbours each
    assert (np.diff(sps.csc_matrix(W).indptr)[cols] == min(sim.TopK, X.shape[1] - 1)).all()
    ties = check_topk_against_dense(W, EuclideanOracle(X, **kw), cols, rtol=RTOL)
    return W, sim, ties


@pytest.mark.parametrize("n", range(len(CASES)))
def test_cuda_matches_reference_golden(n):
    values, kw, Wref = CASES[n]
    W, sim, ties = _check(URMS[values], **kw)
    if values == "continuous":  # tie-free: the reference's index sets are reproducible
        assert ties == 0 and same_sparse(W, Wref, rtol=RTOL)


@pytest.mark.parametrize("mode", ["exp", "lin", "log"])
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("values", ["continuous", "binary"])
def test_modes_medium(mode, normalize, values):
    X = synth_urm(3000, 1200, 0.02, seed=31, values=values)
    _check(X, cols=np.arange(0, 1200, 7), topK=40, shrink=1, normalize=normalize, similarity_from_distance_mode=mode)


def test_signed_data_and_empty_columns():
    X = synth_urm(800, 300, 0.05, seed=33, values="continuous")
    X.data -= 0.5  # negative dot products: the zero-dot bound is not a floor any more
    X = X.tolil()
    X[:, 9] = 0
    X[:, 200] = 0
    X = sps.csr_matrix(X.tocsr(), dtype=np.float32)
    X.eliminate_zeros()
    for normalize in (False, True):
        W, sim, _ = _check(X, topK=20, shrink=0.5, normalize=normalize, similarity_from_distance_mode="lin")
        assert sim.signed_data


def test_prune_path_and_avg_row():
    """A tiny logical buffer prunes before every chunk; normalize_avg_row divides by n_rows."""
    X = synth_urm(4000, 6000, 0.004, seed=35, values="ratings")
    _check(X, cols=np.arange(0, 6000, 41), debug_cap=64, topK=30, shrink=0, normalize_avg_row=True,
           similarity_from_distance_mode="log")


def test_multiwindow():
    """More columns than one accumulator window: the floor from the norm-ordered prefix carries across windows."""
    X = synth_urm(20_000, 120_000, 0.0004, seed=8, values="continuous")
    W, sim, _ = _check(X, cols=np.arange(0, 120_000, 2999), topK=50, shrink=2, similarity_from_distance_mode="lin")
    assert sim.n_windows >= 2
    Xb = synth_urm(20_000, 230_000, 0.0003, seed=8, values="binary")
    W, sim, _ = _check(Xb, cols=np.arange(0, 230_000, 5999), topK=50, shrink=2, normalize=True, similarity_from_distance_mode="lin")
    assert sim.n_windows >= 2 and sim.binary_path


def test_dispatcher_column_range_and_errors():
    from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity
    X = synth_urm(500, 300, 0.03, seed=6, values="continuous")
    obj = Compute_Similarity(X, similarity="euclidean", topK=10, shrink=1)
    Wfull = obj.compute_similarity()
    Wpart = obj.compute_similarity(start_col=100, end_col=180)
    assert abs(Wpart[:, 100:180] - Wfull[:, 100:180]).max() < 1e-7
    assert Wpart[:, :100].nnz == 0 and Wpart[:, 180:].nnz == 0
    with pytest.raises(ValueError):
        _cls()(X, similarity_from_distance_mode="sqrt")
    with pytest.raises(ValueError):
        _cls()(X, row_weights=np.ones(499))
    with pytest.raises(NotImplementedError):
        _cls()(X, row_weights=np.ones(500))
