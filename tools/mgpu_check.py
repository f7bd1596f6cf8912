"""torchrun --nproc-per-node N tools/mgpu_check.py : item-sharded similarity == single-GPU result, on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from recsys2019_deeplearning_evaluation_b200.similarity import Compute_Similarity_Cython
from recsys2019_deeplearning_evaluation_b200.dist import compute_similarity_sharded, SymmetricTopKTable

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for values, pop in (("continuous", None), ("binary", 1.0)):
    X = synth_urm(30_000, 9_000, 0.004, seed=5, values=values, popularity=pop)
    sim = Compute_Similarity_Cython(X, topK=40, shrink=10, similarity="cosine")
    W_full = sim.compute_similarity()
    table = SymmetricTopKTable(sim.n_columns, sim.K)
    for label, W_shard in (("nccl all-gather", compute_similarity_sharded(sim)),
                           ("peer stores", compute_similarity_sharded(sim, table=table)),
                           ("peer stores, second fill", compute_similarity_sharded(sim, table=table))):
        d = abs(W_shard - W_full)
        same = (W_shard.nnz == W_full.nnz) and (d.nnz == 0 or d.max() < 1e-6)
        print("[rank %d/%d] %s pop=%s nnz=%d %s: sharded==single: %s" % (rank, world, values, pop, W_full.nnz, label, same), flush=True)
        ok = ok and same
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
