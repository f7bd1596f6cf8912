"""Development timing of one IALS epoch on one GPU: python tools/dev_ials_bench.py [C4] [f=128] [reps=2]
(B200REC_IALS_V2=0 / 1: tensor-core kernel never / for every factor count; default: from 128 factors)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200 import recommenders as R

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
f = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
X = synth_config(cfg)
np.random.seed(0)
rec = R.IALSRecommender(X, verbose=False)
rec.fit(epochs=1, num_factors=f, alpha=1.0, reg=1e-3)
torch.cuda.synchronize()
lens_u = np.diff(X.indptr).astype(np.float64)
flops = float(2 * 2 * f * f * lens_u.sum() + (X.shape[0] + X.shape[1]) * (2.0 * f ** 3 / 3.0 + 2.0 * f * f))
for r in range(reps):
    t = time.perf_counter(); rec._run_epoch(r + 1); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("IALS %s f=%d v2=%s epoch %.3f s  %.1f TFLOP/s (8(d) flops %.3e)  row-solves/s %.3e" % (
        cfg, f, os.environ.get("B200REC_IALS_V2", "default"), dt, flops / dt / 1e12, flops, (X.shape[0] + X.shape[1]) / dt), flush=True)
