"""AsySVD with factor counts that are not multiples of 4 (rows are padded to float4s on the device): CUDA vs the C oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle.sgd_oracle import MFOracle
from recsys2019_deeplearning_evaluation_b200.mf_epoch import MatrixFactorization_Cython_Epoch
from test_oracle_next_rows import asy_urm
X = asy_urm()
for f, mode in ((10, "adagrad"), (50, "adam"), (3, "sgd"), (17, "rmsprop")):
    kw = dict(n_factors=f, algorithm_name="ASY_SVD", batch_size=1, learning_rate=0.01, random_seed=42, sgd_mode=mode, use_bias=True,
              negative_interactions_quota=0.3, user_reg=1e-3, item_reg=2e-3, bias_reg=1e-3)
    g, o = MatrixFactorization_Cython_Epoch(X, **kw), MFOracle(X, **kw)
    g.epochIteration_Cython(); o.epochIteration_Cython()
    d = [float(np.abs(getattr(g, n)() - getattr(o, n)()).max()) for n in ("get_USER_factors", "get_ITEM_factors", "get_USER_bias", "get_ITEM_bias")]
    print("f=%d %s max abs diff Y/X/bu/bi: %s  %s" % (f, mode, " ".join("%.2e" % x for x in d), "OK" if max(d) < 1e-4 else "MISMATCH"), flush=True)
