"""CPU: the numerics claim behind the planned tensor-core IALS Gram (tensor-core status): a 3xTF32 Gram with a
truncating fp32 accumulator (tools/ials_slicing_study.py's model of the tensor core) plus ONE refinement step against the
exact operator reaches fp64-level solutions on the reference's all-positive factors."""
import importlib.util
import os

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "ials_study", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ials_slicing_study.py"))
study = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(study)


def test_tf32_split_and_truncation_models():
    x = np.array([1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, -3.14159274], np.float32)
    t = study.tf32(x)
    assert t[0] == np.float32(1.0 + 2.0 ** -10) and t[1] == np.float32(1.0)  # ties round away, 10 mantissa bits
    lo = study.tf32(x - t)
    assert np.all(np.abs((t.astype(np.float64) + lo) - x) <= np.abs(x) * 2.0 ** -21)
    y = study.rz32(np.array([1.0 + 2.0 ** -30, -(1.0 + 2.0 ** -30)]))
    assert y[0] == np.float32(1.0) and y[1] == np.float32(-1.0)  # toward zero


def test_one_refinement_step_reaches_fp64_level():
    rng = np.random.default_rng(3)
    f, n_users, plen = 64, 6000, 2400
    U = f ** -0.5 * rng.random((n_users, f))
    UU = U.T @ U
    idx = rng.choice(n_users, plen, replace=False)
    c = 1.0 + rng.integers(1, 6, plen).astype(np.float64)
    Yp = U[idx]
    B = UU + Yp.T @ ((c - 1.0)[:, None] * Yp) + 1e-3 * np.eye(f)
    rhs = Yp.T @ c
    x = np.linalg.solve(B, rhs)
    Bt = UU + study.gram_3xtf32_truncating(Yp, c - 1.0) + 1e-3 * np.eye(f)
    x0 = np.linalg.solve(Bt, rhs)
    x1 = x0 + np.linalg.solve(Bt, rhs - B @ x0)
    e0, e1 = (np.abs(v - x).max() / np.abs(x).max() for v in (x0, x1))
    assert 1e-7 < e0 < 1e-3      # the truncating 3xTF32 Gram alone sits near the 1e-4 parity bar
    assert e1 < 1e-6             # one matrix-free refinement step is enough
