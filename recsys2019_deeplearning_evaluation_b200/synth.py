"""Deterministic synthetic URMs of the BASELINE.json shapes (SURVEY.md 8(d) / Appendix E recipe)."""
import numpy as np
import scipy.sparse as sps

# name -> (n_users, n_items, density); BASELINE.json configs[0..4]
CONFIGS = {
    "C1": (10_000, 5_000, 0.01),
    "C2": (6_040, 3_706, 0.0447),
    "C3": (138_000, 27_000, 0.00535),
    "C4": (480_000, 17_700, 0.0118),
    "C5": (1_000_000, 200_000, 0.0005),
}


def synth_urm(n_users, n_items, density, seed=42, values="binary", popularity=None):
    """CSR float32, sorted indices, duplicate (u,i) draws merged.
    values: 'binary' (all 1), 'ratings' (1..5), 'continuous' (tie-free (0.001,1.001)).
    popularity: None = uniform items; a float s>0 = Zipf(s) item popularity (load-imbalance stress case)."""
    rng = np.random.default_rng(seed)
    lens = rng.binomial(n_items, density, size=n_users).astype(np.int64)
    total = int(lens.sum())
    if popularity is None:
        cols = rng.integers(0, n_items, size=total, dtype=np.int64)
    else:
        w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), float(popularity))
        cdf = np.cumsum(w / w.sum())
        cols = np.searchsorted(cdf, rng.random(total), side="left").astype(np.int64)
        np.minimum(cols, n_items - 1, out=cols)
    rows = np.repeat(np.arange(n_users, dtype=np.int64), lens)
    # merge duplicate (row, col) draws without scipy's slow COO path
    key = rows * n_items + cols
    key.sort()  # (np.unique is an order of magnitude slower than sort + adjacent-difference here)
    if len(key):
        keep = np.empty(len(key), bool)
        keep[0] = True
        np.not_equal(key[1:], key[:-1], out=keep[1:])
        key = key[keep]
    rows = (key // n_items).astype(np.int32)
    cols = (key % n_items).astype(np.int32)
    m = len(key)
    vrng = np.random.default_rng(1)
    if values == "binary":
        data = np.ones(m, np.float32)
    elif values == "ratings":
        data = vrng.integers(1, 6, m).astype(np.float32)
    elif values == "continuous":
        data = vrng.random(m, dtype=np.float32) + np.float32(1e-3)
    else:
        raise ValueError("values must be binary|ratings|continuous")
    indptr = np.zeros(n_users + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    M = sps.csr_matrix((data, cols, indptr), shape=(n_users, n_items), dtype=np.float32)
    M.has_sorted_indices = True
    return M


def synth_config(name, seed=42, values="binary", popularity=None):
    """The URM of a BASELINE.json config.  B200REC_SYNTH_CACHE=<dir> (e.g. /dev/shm) keeps the three CSR arrays on disk so
    that several processes of one measurement session (bench arms, profiler passes) do not regenerate them."""
    import os
    nu, ni, d = CONFIGS[name]
    cache = os.environ.get("B200REC_SYNTH_CACHE")
    if not cache:
        return synth_urm(nu, ni, d, seed=seed, values=values, popularity=popularity)
    stem = os.path.join(cache, "urm_%s_%d_%s_%s" % (name, seed, values, popularity))
    try:
        data, idx, ptr = (np.load(stem + "_%s.npy" % k) for k in ("data", "indices", "indptr"))
        M = sps.csr_matrix((data, idx, ptr), shape=(nu, ni), dtype=np.float32)
        M.has_sorted_indices = True
        return M
    except (OSError, ValueError):
        M = synth_urm(nu, ni, d, seed=seed, values=values, popularity=popularity)
        try:
            for k, a in (("data", M.data), ("indices", M.indices), ("indptr", M.indptr)):
                np.save(stem + "_%s.tmp.npy" % k, a)
                os.replace(stem + "_%s.tmp.npy" % k, stem + "_%s.npy" % k)
        except OSError:
            pass
        return M
