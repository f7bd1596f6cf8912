"""BASELINE.json configs[4] on ONE GPU: every baseline of the sweep at the C5 shape (1 M x 200 K, 0.05 %, binary), one JSON
object per line.  ItemKNN with every similarity, P3alpha / RP3beta, BPR-MF (both semantics), IALS, and one rank's share of a
column-sharded SLIM-BPR epoch (the dense S of 200 K items is 160 GB: this GPU holds 1/8 of the columns, like one rank of 8).
    python tools/c5_sweep.py [--quick]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200 import recommenders as R
from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity_Cython, Compute_Similarity_Euclidean
from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch
from recsys2019_deeplearning_evaluation_b200.dist import ShardedSLIM_BPR

quick = "--quick" in sys.argv


def sync():
    torch.cuda.synchronize()


def emit(**kw):
    print(json.dumps(kw), flush=True)


t = time.time()
X = synth_config("C5", values="binary")
n_users, n_items = X.shape
emit(bench="URM", shape=list(X.shape), nnz=int(X.nnz), gen_s=time.time() - t)

# ---- ItemKNN, every similarity of Compute_Similarity (topK 200, shrink 100)
for kind in ("cosine", "jaccard", "dice", "tversky", "asymmetric", "adjusted", "pearson"):
    t = time.perf_counter()
    sim = Compute_Similarity_Cython(X, topK=200, shrink=100, normalize=True, similarity=kind, asymmetric_alpha=0.5, tversky_alpha=1.0, tversky_beta=1.0)
    sync(); t_create = time.perf_counter() - t
    sim.compute_topk_device(0, n_items); sync()
    tab = sim.compute_topk_device(0, n_items); sync()
    ms = sim.last_kernel_ms()
    t = time.perf_counter(); W = sim.table_to_csr(tab); t_csr = time.perf_counter() - t
    emit(bench="ItemKNN %s C5" % kind, kernel_ms=ms, rows_per_s=n_items / (ms * 1e-3), create_s=t_create, csr_s=t_csr, nnz=int(W.nnz),
         binary_path=bool(sim.binary_path), windows=sim.n_windows)
    sim._dealloc(); del W, tab
if not quick:
    t = time.perf_counter()
    sim = Compute_Similarity_Euclidean(X, topK=200, shrink=100, normalize=True, similarity_from_distance_mode="lin")
    sync(); t_create = time.perf_counter() - t
    tab = sim.compute_topk_device(0, n_items); sync()
    emit(bench="ItemKNN euclidean C5", kernel_ms=sim.last_kernel_ms(), rows_per_s=n_items / (sim.last_kernel_ms() * 1e-3), create_s=t_create)
    sim._dealloc(); del tab

# ---- graph-based
for name, cls, kw in (("P3alpha", R.P3alphaRecommender, dict(topK=200, alpha=1.0)), ("RP3beta", R.RP3betaRecommender, dict(topK=200, alpha=1.0, beta=0.6))):
    rec = cls(X, verbose=False)
    sync(); t = time.perf_counter(); rec.fit(**kw); sync(); dt = time.perf_counter() - t
    emit(bench="%s fit C5" % name, seconds=dt, items_per_s=n_items / dt, nnz=int(rec.W_sparse.nnz))
    del rec

# ---- BPR-MF, 128 factors
for label, kw in (("reference semantics (mini-batch 1000, dataflow kernel)", dict(batch_size=1000, sampler="philox")),
                  ("hogwild", dict(batch_size=1000, sampler="philox", hogwild=True))):
    m = MatrixFactorization_Cython_Epoch(X, n_factors=128, algorithm_name="MF_BPR", learning_rate=1e-3, random_seed=42, sgd_mode="sgd", **kw)
    for _ in range(3):
        m.epochIteration_Cython()
    sync()
    ms = []
    for _ in range(5):
        m.epochIteration_Cython(); sync(); ms.append(m.last_epoch_ms())
    emit(bench="BPRMF f=128 C5", mode=label, samples_per_s=m.samples_last_epoch() / (min(ms) * 1e-3), ms_per_epoch=min(ms))
    m._dealloc()

# ---- IALS, 128 factors (tensor-core kernel)
np.random.seed(0)
rec = R.IALSRecommender(X, verbose=False)
rec.fit(epochs=1, num_factors=128, alpha=1.0, reg=1e-3)
sync(); t = time.perf_counter(); rec._run_epoch(1); sync(); dt = time.perf_counter() - t
emit(bench="IALS f=128 epoch C5", seconds=dt, row_solves_per_s=(n_users + n_items) / dt)
del rec
torch.cuda.empty_cache()

# ---- SLIM-BPR: one rank's share of a run sharded over 8 GPUs (columns [0, n_items / 8)), batches of 8192, no exchange timed
tr = ShardedSLIM_BPR(X, batch_size=8192, col_range=(0, n_items // 8), world_rank=(1, 0), learning_rate=1e-4, topK=200, random_seed=42, sgd_mode="adagrad")
tr.epochIteration_Cython(); sync()
t = time.perf_counter(); tr.epochIteration_Cython(); sync(); dt = time.perf_counter() - t
emit(bench="SLIM_BPR epoch C5, 1/8 of the columns of S on this GPU (20 GB slab)", seconds=dt, samples_per_s=n_users / dt, batches=(n_users + 8191) // 8192,
     note="partial + apply kernels of every batch; the all-reduce of 8192 partial sums per batch is not in this figure")
t = time.perf_counter(); idx, val = tr.local_row_topk(); sync(); dt = time.perf_counter() - t
emit(bench="SLIM_BPR local row top-200 of the slab C5", seconds=dt)
