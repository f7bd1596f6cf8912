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
    mine = tr.slab().cpu().numpy().astype(np.float64)
    ref = one.slab()[:, tr.lo:tr.hi].cpu().numpy().astype(np.float64)
    # the two runs add the ranks' partial sums in a different order (fp32): the trajectories agree to rounding amplified
    # over 7 epochs of adagrad steps, not bit for bit
    rel = float(np.linalg.norm(mine - ref) / np.linalg.norm(ref))
    same = rel < 1e-3
    W1 = one.get_S()
    a, b = (W != 0).astype(np.int8), (W1 != 0).astype(np.int8)
    both = a.multiply(b)
    overlap = float(both.sum() / max(1, b.sum()))
    # cells near the K-th value may swap between the two runs (they differ by rounding): membership must agree on > 99 % of
    # the entries and the common entries must carry the same value
    common_diff = float(abs((W - W1).multiply(both)).max())
    topk_same = overlap > 0.99 and common_diff < 1e-3 * float(abs(W1).max())
    ok = same and topk_same and abs(ref).max() > 0
    print("[sharded SLIM x%d] slab vs single-shard run: rel. Frobenius %.2e (max abs %.2e, max |S| %.2e): %s; merged row top-K overlap %.4f: %s; %.3e samples/s (batch %d)" % (
        world, rel, float(abs(mine - ref).max()), float(abs(ref).max()), same, overlap, topk_same, 5 * X.shape[0] / dt, kw["batch_size"]), flush=True)
f = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(f, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(f.item()) == 1 else 1)
