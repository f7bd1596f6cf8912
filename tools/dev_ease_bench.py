import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200 import recommenders as R
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
X = synth_config(cfg)
rec = R.EASE_R_Recommender(X, verbose=False)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); rec.fit(topK=None, l2_norm=1e3, verbose=False); torch.cuda.synchronize()
    print("EASE_R fit %s: %.3f s" % (cfg, time.perf_counter() - t), flush=True)
