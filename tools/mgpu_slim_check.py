"""torchrun --nproc-per-node N tools/mgpu_slim_check.py : column-sharded SLIM-BPR over NCCL == one shard holding every column
(same stream, same batches), and the merged row top-K agrees; prints samples/s of the sharded epochs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from recsys2019_deeplearning_evaluation_b200.dist import ShardedSLIM_BPR

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = synth_urm(20_000, 3_000, 0.01, seed=7, popularity=0.8)
kw = dict(batch_size=2048, learning_rate=0.05, li_reg=1e-3, lj_reg=1e-3, topK=50, random_seed=3, sgd_mode="adagrad")
tr = ShardedSLIM_BPR(X, **kw)
for _ in range(2):
    tr.epochIteration_Cython()
torch.cuda.synchronize(); dist.barrier()
t = time.perf_counter()
for _ in range(5):
    tr.epochIteration_Cython()
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t
W = tr.get_S()
ok = True
if rank == 0:
    one = ShardedSLIM_BPR(X, col_range=(0, X.shape[1]), world_rank=(1, 0), **kw)
    for _ in range(7):
        one.epochIteration_Cython()
    mine = tr.slab().cpu().numpy()
    ref = one.slab()[:, tr.lo:tr.hi].cpu().numpy()
    same = np.allclose(mine, ref, rtol=1e-4, atol=2e-5)
    W1 = one.get_S()
    d = abs(W - W1)
    topk_same = d.nnz == 0 or d.max() < 1e-4
    ok = same and topk_same and abs(ref).max() > 0
    print("[sharded SLIM x%d] slab == single-shard run: %s, merged row top-K == single: %s, %.3e samples/s (batch %d)" % (
        world, same, topk_same, 5 * X.shape[0] / dt, kw["batch_size"]), flush=True)
f = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(f, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(f.item()) == 1 else 1)
