"""CPU: the P3alpha / RP3beta restatement (oracle/graph_oracle.py) against golden W_sparse matrices produced by the
reference's own GraphBased classes (tests/golden/make_golden.py::make_graph_golden)."""
import os
import runpy

import numpy as np
import pytest
import scipy.sparse as sps

from oracle.graph_oracle import p3_similarity
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

HERE = os.path.dirname(os.path.abspath(__file__))
GRAPH_CASES = runpy.run_path(os.path.join(HERE, "golden", "make_golden.py"), run_name="cases")["GRAPH_CASES"]
Z = np.load(os.path.join(HERE, "golden", "graph_golden.npz"))


def golden(n):
    return sps.csr_matrix((Z["g%d_data" % n], Z["g%d_indices" % n], Z["g%d_indptr" % n]), shape=(150, 150))


def case_inputs(n):
    c = dict(GRAPH_CASES[n])
    c.pop("cls")
    X = synth_urm(400, 150, 0.06, seed=17, values="continuous" if n % 2 == 0 else "ratings")
    return X, c


def close_enough(A, B, rtol=1e-4):
    """Same matrix up to top-K boundary ties: values agree where both have an entry, and the entries present in
    only one of them are at the K-th value of their row/column (tie class)."""
    A, B = A.toarray().astype(np.float64), B.toarray().astype(np.float64)
    both = (A != 0) & (B != 0)
    assert np.allclose(A[both], B[both], rtol=rtol, atol=1e-9)
    only = (A != 0) ^ (B != 0)
    return int(only.sum())


@pytest.mark.parametrize("n", range(len(GRAPH_CASES)))
def test_graph_oracle_matches_reference_golden(n):
    X, c = case_inputs(n)
    Wo = p3_similarity(X, **c)
    Wg = golden(n)
    assert Wo.shape == Wg.shape == (150, 150)
    mism = close_enough(Wo, Wg)
    if n % 2 == 0:  # continuous values: no ties, identical structure
        assert mism == 0
    else:
        assert mism <= 0.02 * Wg.nnz
