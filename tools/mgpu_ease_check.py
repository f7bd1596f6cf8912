"""torchrun --nproc-per-node N tools/mgpu_ease_check.py : EASE_R with the Gram summed over user shards (NCCL all-reduce) ==
the single-GPU fit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from recsys2019_deeplearning_evaluation_b200.recommenders import EASE_R_Recommender
from recsys2019_deeplearning_evaluation_b200.dist import make_sharded_ease

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for values, n_items in (("binary", 1100), ("binary", 3000)):  # binary: the popularity diagonal of EASE_R is the Gram diagonal
    X = synth_urm(20_000, n_items, 0.01, seed=5, values=values, popularity=0.8)
    a = make_sharded_ease()(X, verbose=False)
    a.fit(topK=None, l2_norm=50.0, verbose=False)
    b = EASE_R_Recommender(X, verbose=False)
    b.fit(topK=None, l2_norm=50.0, verbose=False)
    A, B = np.asarray(a.W_sparse, np.float64), np.asarray(b.W_sparse, np.float64)
    rel = float(np.abs(A - B).max() / np.abs(B).max())
    same = rel < 1e-5  # integer-valued Gram entries sum exactly in fp32 below 2^24; the inverse is the same code on both
    print("[rank %d/%d] %s n_items=%d sharded Gram == single: %s (max rel %.2e)" % (rank, world, values, n_items, same, rel), flush=True)
    ok = ok and same
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
