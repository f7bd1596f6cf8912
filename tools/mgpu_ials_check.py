"""torchrun --nproc-per-node N tools/mgpu_ials_check.py : row-sharded IALS == single-GPU IALS.  Y^T Y is summed with fp64
atomics, so two runs agree to rounding, not bit for bit: 1e-9 on the fp64 kernel; 1e-6 on the tensor-core kernel (above 128
factors), whose fp32 factorisation turns a last-bit change of Y^T Y into a ~1e-8 relative change of the refined solution --
four orders below the 1e-4 parity bar."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from recsys2019_deeplearning_evaluation_b200.synth import synth_urm
from recsys2019_deeplearning_evaluation_b200.recommenders import IALSRecommender
from recsys2019_deeplearning_evaluation_b200.dist import make_sharded_ials

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = synth_urm(20_000, 3_000, 0.01, seed=5, values="ratings", popularity=0.8)
ok = True
for f in (64, 256):
    np.random.seed(7)
    a = make_sharded_ials()(X, verbose=False)
    a.fit(epochs=2, num_factors=f, alpha=2.0, reg=1e-2)
    np.random.seed(7)
    b = IALSRecommender(X, verbose=False)
    b.fit(epochs=2, num_factors=f, alpha=2.0, reg=1e-2)
    tol = 1e-9 if f <= 128 else 1e-6
    same = np.allclose(a.USER_factors, b.USER_factors, rtol=tol, atol=1e-3 * tol) and np.allclose(a.ITEM_factors, b.ITEM_factors, rtol=tol, atol=1e-3 * tol)
    print("[rank %d/%d] f=%d sharded==single: %s (max diff %.3e)" % (rank, world, f, same, float(np.abs(a.USER_factors - b.USER_factors).max())), flush=True)
    ok = ok and same
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
