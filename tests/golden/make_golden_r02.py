"""Generates tests/golden/next_rows_golden.npz by running the UNMODIFIED reference in the authoring container:
the compiled Cython classes from oracle/_ref (tree-sparse SLIM-BPR mode, AsySVD) and
SLIM_ElasticNet/SLIMElasticNetRecommender.py imported from /root/reference (with this container's scikit-learn).

    python tests/golden/make_golden_r02.py

The cases and input generators are the ones tests/test_oracle_next_rows.py uses."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import test_oracle_next_rows as T  # noqa: E402


def main():
    out = {}
    SL = ref_loader.load("SLIM_BPR_Cython_Epoch").SLIM_BPR_Cython_Epoch
    MF = ref_loader.load("MatrixFactorization_Cython_Epoch").MatrixFactorization_Cython_Epoch
    for n, case in enumerate(T.TREE_CASES):
        for e, S in enumerate(T.run_tree(SL, case)):
            out["tree%d_S%d" % (n, e)] = S
    for n, kw in enumerate(T.ASY_CASES):
        for k, a in enumerate(T.run_asy(MF, kw)):
            out["asy%d_%d" % (n, k)] = a
    ref_loader.ensure_import_path()
    from SLIM_ElasticNet.SLIMElasticNetRecommender import SLIMElasticNetRecommender
    for n, (values, l1_ratio, alpha, positive, topK) in enumerate(T.ENET_CASES):
        np.random.seed(n)  # sklearn's selection='random' draws from numpy's global generator (random_state=None)
        r = SLIMElasticNetRecommender(T.enet_urm(values), verbose=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r.fit(l1_ratio=l1_ratio, alpha=alpha, positive_only=positive, topK=topK)
        out["enet%d_W" % n] = r.W_sparse.toarray()
    np.savez_compressed(os.path.join(HERE, "next_rows_golden.npz"), **out)
    print("wrote next_rows_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
