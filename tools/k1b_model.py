"""Executable model (numpy, float32 arithmetic) of the selection logic of csrc/sim_k1b.cuh, for columns of a binary matrix:
first-touch / second-touch bitmaps, table of repeated hits, bootstrap floor from per-slot lower bounds with the coarse-tile
norm bounds, permissive dot thresholds, and the count-1 prefix rule on the norm-sorted neighbour axis.  It checks the
ALGORITHM (that nothing which belongs to the top-K can be skipped) against oracle.similarity_oracle, not the CUDA
mechanics.  Used by tests/test_k1b_model.py; run directly for a quick report."""
import os
import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
f32 = np.float32
CB = 4096
LB_BASE = (127 - 40) << 6
HBINS = 4096


def sim_value(F, d, a, b, se, sdiv, ta, tb):
    d, a, b = f32(d), f32(a), f32(b)
    if F == "prod":
        return d / (a * b + se)
    if F == "nonorm":
        return d / sdiv
    if F == "jaccard":
        return d / (a + b - d + se)
    if F == "dice":
        return d / (a + b + se)
    return d / (d + (a - d) * ta + (b - d) * tb + se)


def dot_threshold(F, t, a, b_lo, b_hi, se, sdiv, ta, tb):
    r = f32(3.4e38)
    for b in (b_lo, b_hi):
        if F == "prod":
            v = t * (a * b + se)
        elif F == "nonorm":
            v = t * sdiv
        elif F == "jaccard":
            v = t * (a + b + se) / (f32(1) + t)
        elif F == "dice":
            v = t * (a + b + se)
        else:
            den = f32(1) - t * (f32(1) - ta - tb)
            v = t * (a * ta + b * tb + se) / den if den > f32(1e-6) else f32(0)
        r = min(r, f32(v))
    return r * (f32(1) - f32(1e-5)) if r > 0 else f32(0)


def lower_bound_scale(F, a, b_lo, b_hi, se, sdiv, ta, tb):
    b = max(b_lo, b_hi)
    if F == "prod":
        den = a * b + se
    elif F == "nonorm":
        den = sdiv
    elif F in ("jaccard", "dice"):
        den = a + b + se
    else:
        if ta < 0 or tb < 0:
            return f32(0)
        den = a * max(f32(0), f32(1) - ta - tb) + a * ta + b * tb + se
    return (f32(1) - f32(1e-5)) / den if den > 0 else f32(0)


def lb_bin(lb):
    return int(min(max((int(np.float32(lb).view(np.uint32)) >> 17) - LB_BASE, 0), HBINS - 1))


def bin_floor(b):
    return np.uint32((b + LB_BASE) << 17).view(np.float32)


def model_topk(X, K, formula, shrink=0.0, ta=1.0, tb=1.0, cols=None):
    """X: binary CSR (users x items).  Returns {col: set of selected neighbour (original) indices} and statistics."""
    X = sps.csr_matrix(X, dtype=np.float32)
    n = X.shape[1]
    Xc = X.tocsc()
    sq = np.asarray(X.multiply(X).sum(axis=0)).ravel()
    norm = sq.astype(np.float32) if formula in ("jaccard", "dice", "tversky") else np.sqrt(sq).astype(np.float32)
    order = np.lexsort((np.arange(n), norm))  # new numbering: ascending norm term, ties by original index
    old2new = np.empty(n, np.int64)
    old2new[order] = np.arange(n)
    B = norm[order]
    se, sdiv, ta, tb = f32(shrink + 1e-6), f32(shrink if shrink != 0 else 1.0), f32(ta), f32(tb)
    ncb = (n + CB - 1) // CB
    cbs = np.array([B[min(t * CB, n - 1)] for t in range(ncb + 1)], np.float32)
    out, stats = {}, dict(prefix_cells=0, table_eval=0, count1_total=0)
    for col_old in (range(n) if cols is None else cols):
        col = old2new[col_old]
        a = B[col]
        users = Xc.indices[Xc.indptr[col_old]:Xc.indptr[col_old + 1]]
        hits = np.concatenate([old2new[X.indices[X.indptr[u]:X.indptr[u + 1]]] for u in users]) if len(users) else np.zeros(0, np.int64)
        hits = hits[hits != col]
        cnt = np.bincount(hits, minlength=n)
        table = {int(j): int(cnt[j] - 1) for j in np.flatnonzero(cnt >= 2)}  # hits after the first
        bm1, bm2 = cnt >= 1, cnt >= 2
        # bootstrap
        thr = f32(0)
        if len(table) > 2 * K:
            hist = np.zeros(HBINS, np.int64)
            scale = [lower_bound_scale(formula, a, cbs[t], cbs[t + 1], se, sdiv, ta, tb) for t in range(ncb)]
            for j, c in table.items():
                lb = f32(c + 1) * scale[j // CB]
                if lb > 0:
                    hist[lb_bin(lb)] += 1
            cum = 0
            for b in range(HBINS - 1, -1, -1):
                if cum < K <= cum + hist[b]:
                    if b > 0:
                        thr = bin_floor(b)
                    break
                cum += hist[b]
        dthr = [dot_threshold(formula, thr, a, cbs[t], cbs[t + 1], se, sdiv, ta, tb) if thr > 0 else f32(0) for t in range(ncb)]
        cand = []  # (sim, -orig) keys
        for j, c in table.items():
            d = f32(c + 1)
            if d >= dthr[j // CB]:
                stats["table_eval"] += 1
                sv = sim_value(formula, d, a, B[j], se, sdiv, ta, tb)
                if sv > 0 and sv >= thr:
                    cand.append((sv, -int(order[j])))
        cand.sort(reverse=True)
        cand = cand[:K]
        if len(cand) == K:
            thr = max(thr, cand[-1][0])
        # count-1 prefix
        t_stop = ncb
        for t in range(ncb):
            if thr != 0 and not (sim_value(formula, 1, a, cbs[t], se, sdiv, ta, tb) >= thr):
                t_stop = t
                break
        j_end = min(n, t_stop * CB)
        ones = np.flatnonzero(bm1[:j_end] & ~bm2[:j_end])
        stats["prefix_cells"] += len(ones)
        stats["count1_total"] += int((bm1 & ~bm2).sum())
        for j in ones:
            sv = sim_value(formula, 1, a, B[j], se, sdiv, ta, tb)
            if sv > 0 and sv >= thr:
                cand.append((sv, -int(order[j])))
        cand.sort(reverse=True)
        out[col_old] = [(-o, s) for s, o in cand[:K]]
    return out, stats


def check_against_oracle(X, K, similarity, shrink, cols, **kw):
    from oracle.similarity_oracle import SimilarityOracle
    formula = {"cosine": "prod", "jaccard": "jaccard", "dice": "dice", "tversky": "tversky"}[similarity]
    got, stats = model_topk(X, K, formula, shrink, kw.get("tversky_alpha", 1.0), kw.get("tversky_beta", 1.0), cols)
    orc = SimilarityOracle(X, topK=K, shrink=shrink, similarity=similarity, **kw)
    V = orc.column_values(np.asarray(list(cols)))
    for k, c in enumerate(cols):
        v = V[:, k]
        top = SimilarityOracle.select_topk(v, K)
        sel = [i for i, _ in got[c]]
        assert len(sel) == len(top), (c, len(sel), len(top))
        if len(top) == 0:
            continue
        kth = v[top[-1]]
        must = set(np.flatnonzero(v > kth * (1 + 1e-5)).tolist())
        assert must <= set(sel), (c, len(must - set(sel)))
        assert all(v[i] >= kth * (1 - 1e-5) for i in sel), c
    return stats


if __name__ == "__main__":
    from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
    for shape, sim, K, shrink in (((3000, 9000, 0.004), "cosine", 20, 10), ((2000, 1500, 0.03), "jaccard", 30, 0),
                                  ((4000, 9000, 0.01), "cosine", 50, 100)):
        X = synth_urm(*shape, seed=3, values="binary")
        st = check_against_oracle(X, K, sim, shrink, range(0, X.shape[1], 37))
        print(shape, sim, "ok", st)
