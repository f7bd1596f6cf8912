"""One-GPU sweep over the remaining BASELINE.json configs (SLIM-BPR C2, P3/RP3 + EASE C1, IALS/EASE Netflix-shape,
scoring) with the reference's CPU implementation timed on a bounded sample beside each.  Prints one JSON object per line.
    python tools/sweep_bench.py [--quick]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import torch

from recsys2019_deeplearning_evaluation_b200.synth import synth_config, synth_urm
from recsys2019_deeplearning_evaluation_b200 import recommenders as R
from recsys2019_deeplearning_evaluation_b200.slim_bpr_epoch import SLIM_BPR_Cython_Epoch
from oracle import ref_loader

quick = "--quick" in sys.argv
ref_loader.numpy_alias_shim()


def sync():
    torch.cuda.synchronize()


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        sync(); t = time.perf_counter(); fn(); sync(); ts.append(time.perf_counter() - t)
    return min(ts)


# ---- C2: SLIM-BPR epoch (6040 x 3706, 4.47 %), symmetric, adagrad -- BASELINE.json configs[1]
X = synth_config("C2")
for label, kw in (("sequential_glibc (reference semantics)", dict(sampler="glibc")), ("hogwild_philox", dict(sampler="philox", hogwild=True))):
    m = SLIM_BPR_Cython_Epoch(X, topK=200, symmetric=True, sgd_mode="adagrad", learning_rate=1e-4, random_seed=42, **kw)
    m.epochIteration_Cython(); sync()
    devs = []
    for _ in range(3):
        m.epochIteration_Cython(); devs.append(m.last_epoch_ms() * 1e-3)
    emit(bench="SLIM_BPR epoch C2", mode=label, samples_per_s=X.shape[0] / min(devs), ms_per_epoch=1e3 * min(devs))
    t = timed(lambda: m.get_S(), 2)
    emit(bench="SLIM_BPR get_S (row top-200 of S) C2", mode=label, seconds=t)
    m._dealloc()
mod = ref_loader.load("SLIM_BPR_Cython_Epoch")
if mod is not None:
    r = mod.SLIM_BPR_Cython_Epoch(X, train_with_sparse_weights=False, topK=200, symmetric=True, sgd_mode="adagrad", learning_rate=1e-4, random_seed=42)
    t = time.perf_counter(); r.epochIteration_Cython(); dt = time.perf_counter() - t
    emit(bench="SLIM_BPR epoch C2", mode="reference Cython (1 thread)", samples_per_s=X.shape[0] / dt, ms_per_epoch=1e3 * dt)

# ---- C1: P3alpha / RP3beta / EASE (10K x 5K, 1 %) -- the reference's own CPU-runnable case
X = synth_config("C1")
for name, cls, kw in (("P3alpha", R.P3alphaRecommender, dict(topK=200, alpha=1.0)), ("RP3beta", R.RP3betaRecommender, dict(topK=200, alpha=1.0, beta=0.6)),
                      ("EASE_R", R.EASE_R_Recommender, dict(topK=None, l2_norm=1e3, verbose=False)), ("ItemKNN cosine", R.ItemKNNCFRecommender, dict(topK=200, shrink=100))):
    rec = cls(X, verbose=False)
    rec.fit(**kw); sync()
    t = timed(lambda: rec.fit(**kw), 2)
    emit(bench="%s fit C1" % name, mode="b200 (host scipy in -> model out)", seconds=t, items_per_s=X.shape[1] / t)
if ref_loader.reference_python_available():
    ref_loader.ensure_import_path(); ref_loader.load("Compute_Similarity_Cython")
    import io, contextlib
    from GraphBased.RP3betaRecommender import RP3betaRecommender as RefRP3
    from EASE_R.EASE_R_Recommender import EASE_R_Recommender as RefEASE
    with contextlib.redirect_stdout(io.StringIO()):
        r = RefRP3(X); t = time.perf_counter(); r.fit(topK=200, alpha=1.0, beta=0.6); dt = time.perf_counter() - t
    emit(bench="RP3beta fit C1", mode="reference (numpy/scipy)", seconds=dt, items_per_s=X.shape[1] / dt)
    with contextlib.redirect_stdout(io.StringIO()):
        r = RefEASE(X); t = time.perf_counter(); r.fit(topK=None, l2_norm=1e3, verbose=False); dt = time.perf_counter() - t
    emit(bench="EASE_R fit C1", mode="reference (Cython Gram + LAPACK inverse)", seconds=dt, items_per_s=X.shape[1] / dt)

# ---- Netflix shape (C4: 480K x 17.7K, 1.18 %): EASE and IALS -- BASELINE.json configs[3]
if not quick:
    X = synth_config("C4")
    rec = R.EASE_R_Recommender(X, verbose=False)
    t = timed(lambda: rec.fit(topK=None, l2_norm=1e3, verbose=False), 1)
    n = X.shape[1]
    emit(bench="EASE_R fit C4", mode="b200", seconds=t, gram_gathered_entries=float(np.sum(np.diff(X.indptr).astype(np.float64) ** 2)),
         inverse_flops=float(2.0 * n ** 3), note="inverse = blocked Cholesky, CUDA-core fp32 GEMM")
    del rec
    torch.cuda.empty_cache()
    f = 128
    np.random.seed(0)
    rec = R.IALSRecommender(X, verbose=False)
    rec.fit(epochs=1, num_factors=f, alpha=1.0, reg=1e-3)
    t = timed(lambda: rec._run_epoch(0), 2)
    lens_u, lens_i = np.diff(X.indptr).astype(np.float64), np.diff(X.tocsc().indptr).astype(np.float64)
    flops = float(2 * f * f * (lens_u.sum() + lens_i.sum()) / 2 + (X.shape[0] + X.shape[1]) * (f ** 3 / 3.0))
    emit(bench="IALS epoch C4 f=%d" % f, mode="b200 fp64", seconds=t, row_solves_per_s=(X.shape[0] + X.shape[1]) / t, approx_flops=flops,
         tflops=flops / t / 1e12)
    # reference: _update_row on a sample of users and items
    from oracle.ials_oracle import update_row, confidence
    C = confidence(X); Ct = sps.csc_matrix(C)
    V = f ** -0.5 * np.random.random_sample((X.shape[1], f)); U = f ** -0.5 * np.random.random_sample((X.shape[0], f))
    VV = V.T @ V; UU = U.T @ U
    t0 = time.perf_counter()
    for u in range(300):
        update_row(C.indices[C.indptr[u]:C.indptr[u + 1]], C.data[C.indptr[u]:C.indptr[u + 1]], V, VV, 1e-3)
    tu = (time.perf_counter() - t0) / 300
    t0 = time.perf_counter()
    for i in range(60):
        update_row(Ct.indices[Ct.indptr[i]:Ct.indptr[i + 1]], Ct.data[Ct.indptr[i]:Ct.indptr[i + 1]], U, UU, 1e-3)
    ti = (time.perf_counter() - t0) / 60
    emit(bench="IALS epoch C4 f=%d" % f, mode="reference _update_row (numpy, sampled 300 users + 60 items, extrapolated)",
         seconds=tu * X.shape[0] + ti * X.shape[1], user_solves_per_s=1 / tu, item_solves_per_s=1 / ti)

# ---- scoring: 1000-user blocks (Evaluator.py:422) on C3
X = synth_config("C3")
rec = R.ItemKNNCFRecommender(X, verbose=False)
rec.fit(topK=200, shrink=100)
users = np.arange(1000)
rec.recommend(users, cutoff=20); sync()
t = timed(lambda: rec.recommend(users, cutoff=20), 3)
emit(bench="recommend(1000 users, cutoff 20) ItemKNN C3", mode="b200 (scores + seen mask + top-N on device)", users_per_s=1000 / t)
W = rec.W_sparse
t0 = time.perf_counter(); sc = X[users].dot(W).toarray(); dt = time.perf_counter() - t0
emit(bench="_compute_item_score(1000 users) ItemKNN C3", mode="reference formula (scipy SpGEMM + toarray)", users_per_s=1000 / dt)
