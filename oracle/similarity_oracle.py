"""CPU restatement (numpy/scipy, fp64) of the reference's sparse column-similarity + top-K path.

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this; the product (recsys2019_deeplearning_evaluation_b200/) never does.

Follows Base/Similarity/Cython/Compute_Similarity_Cython.pyx (paths relative to the reference checkout):
  * data transforms                     pyx:221-232 (set kinds), :236-273 (pearson), :277-312 (adjusted)
  * column norms                        pyx:169-180
  * row weights applied at accumulate   pyx:184-194, :383-400
  * dot products, diagonal excluded     pyx:378-408
  * per-kind normalisation              pyx:476-507
  * top-K, zeros dropped, fp32 CSR      pyx:541-565, :603-611
Top-K semantics for signed similarities follow Base/Similarity/Compute_Similarity_Python.py:335-345 (zeros
outrank negatives and are then dropped), which is what the Cython code intends; the Cython class itself reads
stale scratch slots in that case (last row).  Ties at the K boundary: the reference's order
is numpy-introselect over touched-list insertion order (not reproducible); this oracle and the CUDA kernel
both use "larger value first, then ascending neighbour index".

Pinned against the compiled reference (oracle/_ref) in tests/test_oracle_similarity.py and against the
dense-control recipes of Base/Similarity/Compute_similarity_test.py:31-439.
"""
import numpy as np
import scipy.sparse as sps

KINDS = ("cosine", "adjusted", "asymmetric", "pearson", "jaccard", "tanimoto", "dice", "tversky")
SET_KINDS = ("jaccard", "tanimoto", "dice", "tversky")


def _transform(X, kind):
    """pyx:160-165.  X: CSR float32 -> CSR float64 (the reference mutates float32 data in place, so the centred
    values are rounded to float32 first: pyx:269, :308 operate on the float32 `dataMatrix.data`)."""
    X = sps.csr_matrix(X, dtype=np.float32, copy=True)
    X.sort_indices()
    if kind in SET_KINDS:
        X.data[:] = 1.0
    elif kind == "adjusted":
        cnt = np.diff(X.indptr)
        sums = np.asarray(X.sum(axis=1), dtype=np.float64).ravel()
        mean = np.divide(sums, cnt, out=np.zeros_like(sums), where=cnt > 0)
        X.data = (X.data.astype(np.float64) - np.repeat(mean, cnt)).astype(np.float32)
    elif kind == "pearson":
        Xc = X.tocsc()
        cnt = np.diff(Xc.indptr)
        sums = np.asarray(Xc.sum(axis=0), dtype=np.float64).ravel()
        mean = np.divide(sums, cnt, out=np.zeros_like(sums), where=cnt > 0)
        Xc.data = (Xc.data.astype(np.float64) - np.repeat(mean, cnt)).astype(np.float32)
        X = Xc.tocsr()
        X.sort_indices()
    return X.astype(np.float64)


class SimilarityOracle:
    def __init__(self, dataMatrix, topK=100, shrink=0, normalize=True, asymmetric_alpha=0.5, tversky_alpha=1.0,
                 tversky_beta=1.0, similarity="cosine", row_weights=None):
        if similarity not in KINDS:
            raise ValueError("unknown similarity %r" % (similarity,))
        self.kind = similarity
        self.n_rows, self.n_columns = dataMatrix.shape
        self.TopK = min(topK, self.n_columns)  # pyx:147
        self.shrink = int(shrink)  # C int, pyx:65
        self.normalize = bool(normalize) and similarity not in SET_KINDS  # pyx:128,132,136
        self.alpha, self.ta, self.tb = float(asymmetric_alpha), float(tversky_alpha), float(tversky_beta)
        X = _transform(dataMatrix, similarity)
        sq = np.asarray(X.multiply(X).sum(axis=0), dtype=np.float64).ravel()
        self.norm = sq if similarity in SET_KINDS else np.sqrt(sq)  # pyx:170-174
        if similarity == "asymmetric":  # pyx:176-180
            self.norm_a = np.power(self.norm + 1e-6, 2 * self.alpha)
            self.norm_b = np.power(self.norm + 1e-6, 2 * (1 - self.alpha))
        self.X = X.tocsc()
        if row_weights is not None:
            if len(row_weights) != self.n_rows:
                raise ValueError("row_weights length")
            self.Xw_T = sps.diags(np.asarray(row_weights, dtype=np.float64)).dot(X).T.tocsr()
        else:
            self.Xw_T = X.T.tocsr()

    def column_values(self, cols):
        """Dense (n_columns, len(cols)) float64 similarities of every column against `cols` (diagonal 0)."""
        cols = np.asarray(cols)
        D = np.asarray(self.Xw_T.dot(self.X[:, cols]).todense(), dtype=np.float64)
        D[cols, np.arange(len(cols))] = 0.0  # pyx:396
        s = self.shrink
        ni = self.norm[cols][None, :]
        nj = self.norm[:, None]
        if self.normalize:
            if self.kind == "asymmetric":
                den = self.norm_a[cols][None, :] * self.norm_b[:, None] + s + 1e-6  # pyx:480-481
            else:
                den = ni * nj + s + 1e-6  # pyx:484-485
            return D / den
        if self.kind in ("jaccard", "tanimoto"):
            return D / (ni + nj - D + s + 1e-6)  # pyx:490-491
        if self.kind == "dice":
            return D / (ni + nj + s + 1e-6)  # pyx:495-496
        if self.kind == "tversky":
            return D / (D + (ni - D) * self.ta + (nj - D) * self.tb + s + 1e-6)  # pyx:500-503
        return D / s if s != 0 else D  # pyx:505-507

    @staticmethod
    def select_topk(values, K):
        """values: dense float64 column (all n_columns, zeros included).  Returns the indices of the K largest
        (value desc, index asc), zeros dropped (pyx:541-565 with Compute_Similarity_Python.py:335-345 semantics)."""
        n = len(values)
        order = np.lexsort((np.arange(n), -values))[:K]
        return order[values[order] != 0.0]

    def compute_similarity(self, start_col=None, end_col=None, block=256):
        lo, hi = 0, self.n_columns
        if start_col is not None and 0 < start_col < self.n_columns:  # pyx:450-454
            lo = start_col
        if end_col is not None and lo < end_col < self.n_columns:
            hi = end_col
        rows, cols, vals = [], [], []
        for b0 in range(lo, hi, block):
            cc = np.arange(b0, min(hi, b0 + block))
            V = self.column_values(cc)
            for k, c in enumerate(cc):
                idx = self.select_topk(V[:, k], self.TopK)
                rows.append(idx)
                cols.append(np.full(len(idx), c, dtype=np.int64))
                vals.append(V[idx, k])
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        vals = np.concatenate(vals) if vals else np.zeros(0)
        return sps.csr_matrix((vals, (rows, cols)), shape=(self.n_columns, self.n_columns), dtype=np.float32)  # pyx:607-609


class EuclideanOracle:
    """fp64 restatement of Base/Similarity/Compute_Similarity_Euclidean.py:17-223 (the reference computes in the input
    dtype, fp32).  Pinned against the reference class itself through tests/golden/euclid_golden.npz
    (tests/test_oracle_similarity.py).  row_weights are not restated (:152 multiplies an n_columns vector by n_rows
    weights, only defined for square data)."""

    def __init__(self, dataMatrix, topK=100, shrink=0, normalize=False, normalize_avg_row=False,
                 similarity_from_distance_mode="lin", row_weights=None, **args):
        if similarity_from_distance_mode not in ("exp", "lin", "log"):
            raise ValueError("mode")  # :44-46
        assert row_weights is None
        self.n_rows, self.n_columns = dataMatrix.shape
        self.TopK = min(topK, self.n_columns)  # :26
        self.shrink, self.normalize, self.avg, self.mode = shrink, normalize, normalize_avg_row, similarity_from_distance_mode
        self.X = sps.csc_matrix(dataMatrix, dtype=np.float64)
        self.sq = np.asarray(self.X.multiply(self.X).sum(axis=0)).ravel()  # :112
        self.norm = np.sqrt(self.sq)  # :113

    def column_values(self, cols):
        """Dense (n_columns, len(cols)) float64 similarities of every column against `cols` (diagonal 0, :171)."""
        cols = np.asarray(cols)
        D = np.asarray(self.X.T.dot(self.X[:, cols]).todense(), dtype=np.float64)  # :135
        d2 = self.sq[:, None] + self.sq[cols][None, :] - 2.0 * D  # :144-148
        d2[cols, np.arange(len(cols))] = 0.0  # :149
        if self.normalize:  # :152-154
            den = self.norm[:, None] * self.norm[cols][None, :]
            d2 = np.divide(d2, den, out=d2.copy(), where=den != 0.0)
        if self.avg:
            d2 = d2 / self.n_rows  # :156-157
        d = np.where(d2 > 0.0, np.sqrt(np.maximum(d2, 0.0)), d2)  # :159-160
        with np.errstate(over="ignore"):
            g = np.exp(d) if self.mode == "exp" else (d if self.mode == "lin" else np.log(d + 1.0))
            S = 1.0 / (g + self.shrink + 1e-9)  # :162-169
        S[cols, np.arange(len(cols))] = 0.0  # :171
        return S

    def compute_similarity(self, start_col=None, end_col=None, block_size=100):
        lo, hi = 0, self.n_columns
        if start_col is not None and 0 < start_col < self.n_columns:  # :96-100
            lo = start_col
        if end_col is not None and lo < end_col < self.n_columns:
            hi = end_col
        rows, cols, vals = [], [], []
        for b0 in range(lo, hi, 256):
            cc = np.arange(b0, min(hi, b0 + 256))
            V = self.column_values(cc)
            for k, c in enumerate(cc):
                idx = SimilarityOracle.select_topk(V[:, k], self.TopK)  # :176-186, ties -> ascending index
                rows.append(idx)
                cols.append(np.full(len(idx), c, dtype=np.int64))
                vals.append(V[idx, k])
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        vals = np.concatenate(vals) if vals else np.zeros(0)
        return sps.csr_matrix((vals, (rows, cols)), shape=(self.n_columns, self.n_columns), dtype=np.float32)  # :219-221


def check_topk_against_dense(W, oracle, cols, rtol=1e-4, atol=1e-7):
    """Tie-aware parity check of a computed W (CSR/CSC, entries W[j, col]) against the fp64 dense values.
    For every column in `cols`: (1) every emitted value matches the oracle value at that (j, col) within
    rtol; (2) the emitted set is a valid top-K set: with kth = K-th largest oracle value, every oracle entry
    strictly above kth*(1+rtol) is present and no emitted entry is below kth*(1-rtol); (3) the count equals
    min(K, #non-zero candidates within tolerance of the rule).  Returns the number of boundary-tie columns."""
    Wc = sps.csc_matrix(W)
    K = oracle.TopK
    ties = 0
    cols = np.asarray(cols)
    for b0 in range(0, len(cols), 256):
        cc = cols[b0:b0 + 256]
        V = oracle.column_values(cc)
        for k, c in enumerate(cc):
            v = V[:, k]
            s, e = Wc.indptr[c], Wc.indptr[c + 1]
            got_idx, got_val = Wc.indices[s:e], Wc.data[s:e].astype(np.float64)
            assert len(np.unique(got_idx)) == len(got_idx), "duplicate neighbour in column %d" % c
            ref_at = v[got_idx]
            assert np.allclose(got_val, ref_at, rtol=rtol, atol=atol), "value mismatch in column %d: %r" % (
                c, np.abs(got_val - ref_at).max())
            order = np.lexsort((np.arange(len(v)), -v))
            top = order[:K]
            top = top[v[top] != 0.0]
            assert len(got_idx) == len(top), "column %d: %d neighbours, oracle %d" % (c, len(got_idx), len(top))
            if len(top) == 0:
                continue
            kth = v[top[-1]]
            tol = rtol * abs(kth) + atol
            must = set(order[:K][v[order[:K]] > kth + tol].tolist()) if kth > 0 else set(
                j for j in top.tolist() if v[j] > kth + tol)
            got = set(got_idx.tolist())
            assert must <= got, "column %d misses %d clear top-K members" % (c, len(must - got))
            assert (ref_at >= kth - tol).all(), "column %d holds an entry below the K-th value" % c
            if set(top.tolist()) != got:
                ties += 1
    return ties


def cosine_pair_values(Xc, jj, cc, shrink):
    """fp64 similarity of the column pairs (jj[k], cc[k]) of the CSC matrix Xc, cosine with shrink, normalize=True:
    dot / (|j| |c| + shrink + 1e-6) (pyx:378-408 for the dot, :170-174 and :484-485 for the normalisation)."""
    out = np.zeros(len(jj), np.float64)
    sq = {}

    def col(c):
        s, e = Xc.indptr[c], Xc.indptr[c + 1]
        return Xc.indices[s:e], Xc.data[s:e].astype(np.float64)

    for k, (j, c) in enumerate(zip(jj, cc)):
        ij, vj = col(j)
        ic, vc = col(c)
        for t, v in ((j, vj), (c, vc)):
            if t not in sq:
                sq[t] = float(np.sqrt((v * v).sum()))
        _, aj, ac = np.intersect1d(ij, ic, assume_unique=True, return_indices=True)
        out[k] = float((vj[aj] * vc[ac]).sum()) / (sq[j] * sq[c] + int(shrink) + 1e-6)
    return out


def compare_topk_with_reference(G, R, cols, K, pair_values=None, rtol=1e-4, atol=1e-7, max_pairs=20000):
    """Tie-aware comparison of two top-K results (scipy sparse, entry [j, c] = similarity of neighbour j to target c):
    G the result under test, R the reference's (Compute_Similarity_Cython.compute_similarity, pyx:413-611).
    Per target column in `cols`: the neighbour counts agree; the sorted value lists agree within rtol; common neighbours
    carry the same value; and the two index sets differ only by neighbours tied (within tolerance) with the K-th value --
    the reference's pick among equal values is numpy-introselect order (pyx:544-548), ours ascending index.  For the
    entries only G holds, `pair_values(jj, cc)` (an exact fp64 evaluation) confirms the value G reports.
    Returns {"cols", "ok", "max_rel", "tie_cols", "tie_pairs_checked", "failures": [...first few...]}."""
    G, R = sps.csc_matrix(G), sps.csc_matrix(R)
    fails, max_rel, tie_cols = [], 0.0, 0
    pj, pc, pv = [], [], []
    for c in np.asarray(cols):
        gi, gv = G.indices[G.indptr[c]:G.indptr[c + 1]], G.data[G.indptr[c]:G.indptr[c + 1]].astype(np.float64)
        ri, rv = R.indices[R.indptr[c]:R.indptr[c + 1]], R.data[R.indptr[c]:R.indptr[c + 1]].astype(np.float64)
        if len(gi) != len(ri):
            fails.append("column %d: %d neighbours, reference %d" % (c, len(gi), len(ri)))
            continue
        if len(ri) == 0:
            continue
        sg, sr = np.sort(gv)[::-1], np.sort(rv)[::-1]
        rel = np.abs(sg - sr) / np.maximum(np.abs(sr), 1e-30)
        max_rel = max(max_rel, float(rel.max()))
        if not np.allclose(sg, sr, rtol=rtol, atol=atol):
            fails.append("column %d: sorted values differ, max rel %.3e" % (c, rel.max()))
            continue
        og, orr = np.argsort(gi), np.argsort(ri)
        gi, gv, ri, rv = gi[og], gv[og], ri[orr], rv[orr]
        _, ag, ar = np.intersect1d(gi, ri, assume_unique=True, return_indices=True)
        if not np.allclose(gv[ag], rv[ar], rtol=rtol, atol=atol):
            fails.append("column %d: a common neighbour carries a different value" % c)
            continue
        if len(ag) == len(gi):
            continue
        tie_cols += 1
        kth = sr[-1]
        tol = rtol * abs(kth) + atol
        only_g = np.setdiff1d(np.arange(len(gi)), ag)
        only_r = np.setdiff1d(np.arange(len(ri)), ar)
        if len(ri) < K or (np.abs(gv[only_g] - kth) > tol).any() or (np.abs(rv[only_r] - kth) > tol).any():
            fails.append("column %d: index sets differ beyond ties at the K-th value" % c)
            continue
        if len(pj) < max_pairs:
            pj.extend(gi[only_g].tolist()); pc.extend([int(c)] * len(only_g)); pv.extend(gv[only_g].tolist())
    checked = 0
    if pair_values is not None and pj:
        exact = pair_values(np.asarray(pj), np.asarray(pc))
        checked = len(pj)
        bad = ~np.isclose(np.asarray(pv), exact, rtol=rtol, atol=atol)
        if bad.any():
            k = int(np.argmax(bad))
            fails.append("%d tie entries carry a wrong value, e.g. (%d, %d): %.7g vs exact %.7g" % (
                int(bad.sum()), pj[k], pc[k], pv[k], exact[k]))
    return {"cols": int(len(cols)), "ok": not fails, "max_rel": max_rel, "tie_cols": tie_cols,
            "tie_pairs_checked": checked, "failures": fails[:5]}
