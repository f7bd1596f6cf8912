"""CPU: the selection logic of the sparse-candidate similarity kernel (csrc/sim_k1b.cuh) as an executable numpy model
(tools/k1b_model.py, float32 like the device code) against the fp64 oracle: nothing that belongs to the top-K may be
skipped by the table bootstrap, the coarse-tile dot thresholds or the count-1 prefix rule."""
import importlib.util
import os

import pytest

from recsys2019_deeplearning_evaluation_b200.synth import synth_urm

_spec = importlib.util.spec_from_file_location(
    "k1b_model", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "k1b_model.py"))
k1b = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(k1b)

CASES = [
    ((1500, 9000, 0.004), None, "cosine", 20, 10, {}),        # sparse: tables smaller than K, most count-1 cells enumerated
    ((1200, 900, 0.04), None, "jaccard", 30, 0, {}),          # dense: the floor sits above every count-1 cell
    ((2500, 9000, 0.012), None, "cosine", 40, 100, {}),       # three coarse tiles, shrink-dominated similarities
    ((1500, 5000, 0.02), 1.0, "dice", 25, 5, {}),             # skewed popularity: wide norm range inside a tile
    ((1200, 800, 0.03), None, "tversky", 15, 2, dict(tversky_alpha=0.7, tversky_beta=1.3)),
]


@pytest.mark.parametrize("shape,pop,similarity,K,shrink,kw", CASES)
def test_selection_logic_is_exact(shape, pop, similarity, K, shrink, kw):
    X = synth_urm(*shape, seed=5, values="binary", popularity=pop)
    stats = k1b.check_against_oracle(X, K, similarity, shrink, range(0, X.shape[1], max(1, X.shape[1] // 60)), **kw)
    assert stats["table_eval"] > 0
