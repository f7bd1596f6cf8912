"""Loader for the compiled, unmodified reference modules in oracle/_ref (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.

`load(name)` returns the extension module for one of
    Compute_Similarity_Cython | SLIM_BPR_Cython_Epoch | MatrixFactorization_Cython_Epoch
or None when oracle/_ref does not hold it (then the caller falls back to the oracle port and says so).

The modules import `Base.Recommender_utils` and `Utils.seconds_to_biggest_unit` at import time: when
/root/reference exists (authoring container) the real ones are used; on the GPU box (no /root/reference)
the restated helpers in oracle/ref_shims/ are used instead.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200REC_REFERENCE", "/root/reference")
_DOTTED = {
    "Compute_Similarity_Cython": "Base.Similarity.Cython.Compute_Similarity_Cython",
    "SLIM_BPR_Cython_Epoch": "SLIM_BPR.Cython.SLIM_BPR_Cython_Epoch",
    "MatrixFactorization_Cython_Epoch": "MatrixFactorization.Cython.MatrixFactorization_Cython_Epoch",
}


def numpy_alias_shim():
    """The reference predates numpy 1.24 and uses the removed aliases (np.int at
    MatrixFactorization_Cython_Epoch.pyx:712-716, Base/BaseRecommender.py:30)."""
    for n, t in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, n):
            setattr(np, n, t)


def reference_python_available():
    return os.path.isdir(os.path.join(REF, "Base"))


def ensure_import_path():
    numpy_alias_shim()
    p = REF if reference_python_available() else os.path.join(HERE, "ref_shims")
    if p not in sys.path:
        sys.path.insert(0, p)
    return p


_cache = {}


def load(name):
    if name in _cache:
        return _cache[name]
    hits = sorted(glob.glob(os.path.join(HERE, "_ref", name + "*.so")))
    if not hits:
        _cache[name] = None
        return None
    ensure_import_path()
    dotted = _DOTTED[name]
    if dotted in sys.modules:
        mod = sys.modules[dotted]
    else:
        loader = importlib.machinery.ExtensionFileLoader(dotted, hits[0])
        spec = importlib.util.spec_from_file_location(dotted, hits[0], loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        if reference_python_available():
            # so that the reference's own wrappers (SLIM_BPR_Cython.py:78, MatrixFactorization_Cython.py:54,
            # Compute_Similarity.py:105) find the compiled class when they import it by dotted name
            try:
                importlib.import_module(dotted.rsplit(".", 1)[0])
                sys.modules[dotted] = mod
            except Exception:
                pass
    _cache[name] = mod
    return mod
