"""torchrun --nproc-per-node N tools/mgpu_bpr_check.py : user-sharded Hogwild BPR descends like the single-GPU run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from recsys2019_deeplearning_evaluation_b200.dist import ShardedBPR
from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = synth_urm(20_000, 3_000, 0.01, seed=7, popularity=0.8)
kw = dict(n_factors=32, batch_size=1000, learning_rate=0.05, random_seed=3, sgd_mode="sgd", user_reg=1e-4, positive_reg=1e-4, negative_reg=1e-4)
rng = np.random.default_rng(0)
us = rng.integers(0, X.shape[0], 40000); us = us[np.diff(X.indptr)[us] > 0]
pos = np.array([X.indices[X.indptr[u] + rng.integers(0, X.indptr[u + 1] - X.indptr[u])] for u in us])
neg = rng.integers(0, X.shape[1], len(us))


def loss(U, V):
    x = np.einsum("ij,ij->i", U[us], V[pos] - V[neg])
    return float(np.mean(np.log1p(np.exp(-x))))


tr = ShardedBPR(X, **kw)
l0 = loss(tr.U0.cpu().numpy().astype(np.float64), tr.V.cpu().numpy().astype(np.float64))
for _ in range(60):
    tr.epoch()
tr.flush()
torch.cuda.synchronize()
U = tr.gather_user_factors().cpu().numpy().astype(np.float64)
V = tr.V.cpu().numpy().astype(np.float64)
ls = loss(U, V)
ok = True
if rank == 0:
    single = MatrixFactorization_Cython_Epoch(X, algorithm_name="MF_BPR", sampler="philox", hogwild=True, **kw)
    for _ in range(60):
        single.epochIteration_Cython()
    l1 = loss(single.get_USER_factors(), single.get_ITEM_factors())
    ratio = (l0 - ls) / (l0 - l1)
    ok = (l0 - l1 > 0.01) and 0.7 < ratio < 1.3
    print("[sharded BPR x%d] loss %.4f -> sharded %.4f, single-GPU %.4f, descent ratio %.3f : %s" % (world, l0, ls, l1, ratio, "OK" if ok else "FAIL"), flush=True)
# replicas stay identical
v = tr.V.double().sum()
vs = [torch.zeros_like(v) for _ in range(world)]
dist.all_gather(vs, v)
same = all(abs(float(a - vs[0])) < 1e-6 * abs(float(vs[0])) + 1e-9 for a in vs)
if rank == 0:
    print("item-factor replicas identical across ranks:", same, flush=True)
t = torch.tensor([1 if (ok and same) else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
