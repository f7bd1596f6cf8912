"""SLIM-BPR sequential (reference-semantics) epochs at the C2 shape (BASELINE.json configs[1]): device time per epoch.
    python tools/dev_slim_bench.py          (B200REC_SLIM_PROF=1 prints the per-phase cycle counters of the kernel)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200.slim_bpr_epoch import SLIM_BPR_Cython_Epoch

X = synth_config("C2")
for sym, mode in ((True, "adagrad"), (True, "sgd"), (False, "adam"), (False, "adagrad")):
    m = SLIM_BPR_Cython_Epoch(X, topK=200, symmetric=sym, sgd_mode=mode, learning_rate=1e-4, random_seed=42)
    m.epochIteration_Cython(); torch.cuda.synchronize()
    devs = []
    for _ in range(5):
        m.epochIteration_Cython(); devs.append(m.last_epoch_ms())
    print(json.dumps(dict(bench="SLIM_BPR sequential epoch C2", symmetric=sym, sgd_mode=mode, ms_per_epoch=min(devs), samples_per_s=X.shape[0] / (min(devs) * 1e-3),
                          us_per_sample=min(devs) * 1e3 / X.shape[0])), flush=True)
    m._dealloc()
