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


def test_contraction_check_of_the_device_kernel_on_the_model():
    """The check of ials_v2.cuh (|r|^2 must shrink >= 1000-fold between the first and the second correction): only the
    PROFILE Gram is approximate (Y^T Y stays fp64), so both a well-conditioned system and the ill-conditioned shape with
    barely more rows than factors contract by many orders of magnitude; a factor that is really off (1 % error) is caught."""
    rng = np.random.default_rng(5)

    def ratio(n_other, f, plen, spoil=0.0):
        Y = f ** -0.5 * rng.random((n_other, f))
        YY = Y.T @ Y
        idx = rng.choice(n_other, plen, replace=False)
        c = 1.0 + 2.0 * rng.integers(1, 6, plen).astype(np.float64)
        Yp = Y[idx]
        B = YY + Yp.T @ ((c - 1.0)[:, None] * Yp) + 1e-3 * np.eye(f)
        b = Yp.T @ c
        G = study.gram_3xtf32_truncating(Yp, c - 1.0)
        Bt = YY + G * (1.0 + spoil) + 1e-3 * np.eye(f)  # the device factors the approximate profile Gram + the exact Y^T Y
        x = np.linalg.solve(Bt, b)
        r0 = b - B @ x
        x = x + np.linalg.solve(Bt, r0)
        r1 = b - B @ x
        xe = np.linalg.solve(B, b)
        return float(r1 @ r1) / float(r0 @ r0), np.linalg.cond(B), np.abs(x - xe).max() / np.abs(xe).max()

    for shape in ((6000, 64, 300), (66, 64, 60)):
        rr, cond, err = ratio(*shape)
        assert rr <= 1e-3 and err < 1e-6, (shape, rr, cond, err)
    assert ratio(66, 64, 60)[1] > 1e4  # the second shape is the ill-conditioned one
    rr, cond, err = ratio(66, 64, 60, spoil=0.3)
    assert rr > 1e-3, (rr, cond, err)  # a factor this wrong does not contract fast enough: the fp64 path redoes the half epoch
