"""Development timing of the MF path on one GPU: python tools/dev_mf_bench.py C3 [f] [epochs]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
f = int(sys.argv[2]) if len(sys.argv) > 2 else 128
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
X = synth_config(cfg)
print("URM", X.shape, X.nnz, flush=True)
torch.cuda.init()
for label, kw in [("minibatch bs=1000 glibc", dict(batch_size=1000, sampler="glibc")),
                  ("minibatch bs=1000 philox", dict(batch_size=1000, sampler="philox")),
                  ("minibatch bs=65536 philox", dict(batch_size=65536, sampler="philox")),
                  ("minibatch bs=1000 philox adagrad", dict(batch_size=1000, sampler="philox", sgd_mode="adagrad")),
                  ("hogwild bs=1000 philox", dict(batch_size=1000, sampler="philox", hogwild=True)),
                  ("hogwild bs=65536 philox", dict(batch_size=65536, sampler="philox", hogwild=True))]:
    kw.setdefault("sgd_mode", "sgd")
    t = time.time()
    m = MatrixFactorization_Cython_Epoch(X, n_factors=f, algorithm_name="MF_BPR", learning_rate=1e-3, random_seed=42, **kw)
    torch.cuda.synchronize(); tc = time.time() - t
    walls, kms = [], []
    for e in range(epochs):
        torch.cuda.synchronize(); t = time.time(); m.epochIteration_Cython(); torch.cuda.synchronize(); walls.append(time.time() - t); kms.append(m.last_epoch_ms())
    n = m.samples_last_epoch()
    print("%-34s create %.2fs  samples/epoch %d  wall %.2f ms -> %.3e samples/s   device %.3f ms -> %.3e samples/s (%.1f GB/s at %d B/sample)" % (
        label, tc, n, 1e3 * min(walls), n / min(walls), min(kms), n / (min(kms) * 1e-3), n * 6 * f * 4 / 1e9 / (min(kms) * 1e-3), 6 * f * 4), flush=True)
    m._dealloc()
