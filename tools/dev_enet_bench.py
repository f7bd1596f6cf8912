"""SLIM ElasticNet whole fit at the C2 shape (one launch of slim_enet_kernel): python tools/dev_enet_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recsys2019_deeplearning_evaluation_b200.synth import synth_config
from recsys2019_deeplearning_evaluation_b200 import recommenders as R

X = synth_config("C2", values="ratings")
rec = R.SLIMElasticNetRecommender(X, verbose=False)
torch.cuda.synchronize(); t = time.perf_counter()
rec.fit(l1_ratio=0.1, alpha=1e-3, positive_only=True, topK=100)
torch.cuda.synchronize()
print(json.dumps(dict(bench="SLIM ElasticNet fit C2", seconds=time.perf_counter() - t, nnz=int(rec.W_sparse.nnz))))
