"""Numerics study for a tensor-core IALS Gram (tensor-core status): how accurate must Y_p^T diag(c-1) Y_p be for
the reference's systems, and what does an error-free TF32 slicing deliver?  numpy emulation, no GPU.

For the per-row systems of the reference's golden IALS case (all-positive initial factors) it reports: the condition
numbers; the solution error when the Gram is computed (a) in fp32, (b) as 3xTF32 (hi/lo split, fp32 accumulation),
(c) with s slices of `bits` mantissa bits per operand (products of two slices are exact in fp32 for K <= 2^(24-2*bits)
terms, partial sums combined in fp64)."""
import os
import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ials_oracle import confidence  # noqa: E402
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm  # noqa: E402


def tf32(x):
    """round-to-nearest-away to 10 explicit mantissa bits (cvt.rna.tf32.f32)"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def slices(x, bits, n):
    """x ~ sum of n pieces, each with `bits` significant bits relative to a per-matrix scale (error-free split)"""
    scale = 2.0 ** np.ceil(np.log2(np.abs(x).max()))
    out, r = [], x / scale
    for k in range(n):
        q = np.round(r * 2.0 ** (bits * (k + 1))) / 2.0 ** (bits * (k + 1))
        out.append(q * scale)
        r = r - q
    return out


def gram_variants(Yp, w):
    Yw = Yp * w[:, None]
    exact = Yp.T @ Yw
    g32 = (Yp.astype(np.float32).T @ Yw.astype(np.float32)).astype(np.float64)
    ah, bh = tf32(Yp), tf32(Yw)
    al, bl = tf32(Yp - ah), tf32(Yw - bh)
    g3 = (ah.T.astype(np.float32) @ bh + ah.T @ bl + al.T @ bh).astype(np.float64)  # fp32 accumulation
    res = {"fp32": g32, "3xTF32": g3}
    for bits, n in ((7, 3), (7, 4)):
        A, B = slices(Yp, bits, n), slices(Yw, bits, n)
        acc = np.zeros_like(exact)
        for i in range(n):
            for j in range(n - i):  # drop the pieces below 2^-(bits*n)
                acc += A[i].T @ B[j]  # exact in an fp32 accumulator for short K; summed in fp64 here
        res["%dx%d-bit slices" % (n, bits)] = acc
    return exact, res


def main():
    f = 64
    X = synth_urm(500, 180, 0.05, seed=29, values="ratings")
    C = confidence(X, "linear", 2.0)
    np.random.seed(100)
    V = f ** -0.5 * np.random.random_sample((180, f))  # the reference's init (IALSRecommender.py:204-207)
    VV = V.T @ V
    conds, errs = [], {}
    for u in np.flatnonzero(np.diff(C.indptr) > 0)[:200]:
        s, e = C.indptr[u], C.indptr[u + 1]
        Yp, c = V[C.indices[s:e]], C.data[s:e].astype(np.float64)
        exact, res = gram_variants(Yp, c - 1.0)
        B = VV + exact + 1e-3 * np.eye(f)
        rhs = Yp.T @ c
        x = np.linalg.solve(B, rhs)
        conds.append(np.linalg.cond(B))
        for name, G in res.items():
            xe = np.linalg.solve(VV + G + 1e-3 * np.eye(f), rhs)
            errs.setdefault(name, []).append(np.abs(xe - x).max() / np.abs(x).max())
    print("cond(B): median %.2e  max %.2e" % (np.median(conds), np.max(conds)))
    for name, v in errs.items():
        print("%-18s solution error: median %.1e  max %.1e   (parity bar 1e-4)" % (name, np.median(v), np.max(v)))


if __name__ == "__main__":
    main()


def rz32(x):
    """fp64 -> fp32 with round-toward-zero (a model of the tensor core's accumulator truncation)"""
    y = np.asarray(x, np.float64).astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y


def gram_3xtf32_truncating(Yp, w, kstep=8):
    """3xTF32 Gram with an fp32 accumulator that truncates after every K = kstep block (one MMA)."""
    Yw = Yp * w[:, None]
    ah, bh = tf32(Yp), tf32(Yw)
    al, bl = tf32(Yp - ah), tf32(Yw - bh)
    acc = np.zeros((Yp.shape[1], Yp.shape[1]), np.float32)
    for k0 in range(0, Yp.shape[0], kstep):
        s = slice(k0, k0 + kstep)
        for a, b in ((ah, bh), (ah, bl), (al, bh)):
            acc = rz32(acc.astype(np.float64) + a[s].T.astype(np.float64) @ b[s].astype(np.float64))
    return acc.astype(np.float64)


def long_profile_study():
    """Item-side systems of a Netflix-shaped column: 5 650 users in the profile, f = 128."""
    rng = np.random.default_rng(3)
    f, n_users, plen = 128, 20000, 5650
    U = f ** -0.5 * rng.random((n_users, f))  # all-positive like the reference's factors in the first epochs
    UU = U.T @ U
    worst = {}
    for trial in range(4):
        idx = rng.choice(n_users, plen, replace=False)
        c = 1.0 + 1.0 * rng.integers(1, 6, plen).astype(np.float64)
        Yp = U[idx]
        exact = Yp.T @ ((c - 1.0)[:, None] * Yp)
        B = UU + exact + 1e-3 * np.eye(f)
        rhs = Yp.T @ c
        x = np.linalg.solve(B, rhs)
        Bt = UU + gram_3xtf32_truncating(Yp, c - 1.0) + 1e-3 * np.eye(f)
        xk = np.linalg.solve(Bt, rhs)
        errs = [np.abs(xk - x).max() / np.abs(x).max()]
        for it in range(3):  # iterative refinement against the exact operator (matrix-free fp64 on the device)
            xk = xk + np.linalg.solve(Bt, rhs - B @ xk)
            errs.append(np.abs(xk - x).max() / np.abs(x).max())
        worst[trial] = (np.linalg.cond(B), np.abs(Bt - B).max() / np.abs(B).max(), errs)
    for t, (cond, gerr, errs) in worst.items():
        print("profile %d: cond %.1e  Gram error %.1e  solution error after 0..3 refinement steps: %s" % (
            t, cond, gerr, "  ".join("%.1e" % e for e in errs)))


if __name__ == "__main__":
    long_profile_study()
