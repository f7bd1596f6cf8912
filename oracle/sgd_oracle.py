"""ctypes front-end of oracle/sgd_oracle.c with the reference's constructor signatures.  TEST INFRASTRUCTURE ONLY
(tests/, __graft_entry__.smoke(), bench.py cpu_baseline).

MFOracle   ~ MatrixFactorization_Cython_Epoch  (MatrixFactorization/Cython/MatrixFactorization_Cython_Epoch.pyx:96-191)
SLIMOracle ~ SLIM_BPR_Cython_Epoch             (SLIM_BPR/Cython/SLIM_BPR_Cython_Epoch.pyx:88-134, dense/symmetric S)

Factor initialisation uses numpy's legacy global RNG exactly like the reference: np.random.seed(seed) then
U = normal(mean, std, (n_users, f)), V = normal(...) in that order (pyx:145-147, :177-178).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sps

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build_all())
        _lib.mf_epoch.restype = C.c_long
        _lib.slim_epoch.restype = C.c_long
        _lib.glibc_rand_next.restype = C.c_int
        _lib.sizeof_mf.restype = C.c_ulong
        _lib.sizeof_slim.restype = C.c_ulong
        assert _lib.sizeof_mf() == C.sizeof(MFState), (_lib.sizeof_mf(), C.sizeof(MFState))
        assert _lib.sizeof_slim() == C.sizeof(SLIMState), (_lib.sizeof_slim(), C.sizeof(SLIMState))
    return _lib


class GlibcRandState(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]


class Adapt(C.Structure):
    _fields_ = [("mode", C.c_int), ("gamma", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("b1_pow", C.c_double), ("b2_pow", C.c_double)]


P = C.c_void_p


class MFState(C.Structure):
    _fields_ = [("n_users", C.c_int), ("n_items", C.c_int), ("f", C.c_int), ("batch_size", C.c_int), ("algorithm", C.c_int),
                ("use_bias", C.c_int),
                ("lr", C.c_double), ("user_reg", C.c_double), ("item_reg", C.c_double), ("bias_reg", C.c_double),
                ("positive_reg", C.c_double), ("negative_reg", C.c_double), ("quota", C.c_double),
                ("ad", Adapt),
                ("indptr", P), ("indices", P), ("data", P), ("nnz", C.c_long),
                ("U", P), ("V", P), ("bu", P), ("bi", P), ("mu", P),
                ("accU", P), ("accV", P), ("accbu", P), ("accbi", P), ("accmu", P),
                ("cU", P), ("cV", P), ("cbu", P), ("cbi", P), ("cmu", P),
                ("m1U", P), ("m2U", P), ("m1V", P), ("m2V", P), ("m1bu", P), ("m2bu", P), ("m1bi", P), ("m2bi", P),
                ("m1mu", P), ("m2mu", P),
                ("items_list", P), ("users_list", P), ("items_flag", P), ("users_flag", P),
                ("n_items_touched", C.c_long), ("n_users_touched", C.c_long),
                ("rng", GlibcRandState),
                ("ext_u", P), ("ext_i", P), ("ext_j", P), ("ext_r", P), ("ext_pos", C.c_long),
                ("rec_u", P), ("rec_i", P), ("rec_j", P), ("rec_pos", C.c_long)]


class SLIMState(C.Structure):
    _fields_ = [("n_users", C.c_int), ("n_items", C.c_int), ("symmetric", C.c_int),
                ("lr", C.c_double), ("li_reg", C.c_double), ("lj_reg", C.c_double),
                ("ad", Adapt),
                ("indptr", P), ("indices", P),
                ("S", P), ("c", P), ("m1", P), ("m2", P),
                ("rng", GlibcRandState),
                ("ext_u", P), ("ext_i", P), ("ext_j", P), ("ext_pos", C.c_long),
                ("rec_u", P), ("rec_i", P), ("rec_j", P), ("rec_pos", C.c_long),
                ("exists", P), ("tree_topk", C.c_int)]


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


_MODES = {"sgd": 0, "adagrad": 1, "rmsprop": 2, "adam": 3}


class GlibcRand:
    """srand(seed); rand() -- pinned against libc in tests/test_oracle_sgd.py."""

    def __init__(self, seed):
        self.s = GlibcRandState()
        lib().glibc_rand_seed(C.byref(self.s), C.c_uint(seed))

    def rand(self):
        return lib().glibc_rand_next(C.byref(self.s))


class MFOracle:
    def __init__(self, URM_train, n_factors=1, algorithm_name=None, batch_size=1, negative_interactions_quota=0.5,
                 learning_rate=1e-3, use_bias=False, user_reg=0.0, item_reg=0.0, bias_reg=0.0, positive_reg=0.0,
                 negative_reg=0.0, verbose=False, print_step_seconds=300, random_seed=None, init_mean=0.0,
                 init_std_dev=0.1, sgd_mode="sgd", gamma=0.995, beta_1=0.9, beta_2=0.999,
                 init_factors=None, samples=None, record=0):
        if sgd_mode not in _MODES:
            raise ValueError("sgd_mode")
        if algorithm_name not in ("FUNK_SVD", "MF_BPR", "ASY_SVD"):
            raise ValueError("algorithm_name")
        if algorithm_name == "ASY_SVD":
            assert batch_size == 1, "Batch size other than 1 not supported for ASY_SVD"  # pyx:399
        X = sps.csr_matrix(URM_train, dtype=np.float32).sorted_indices()
        self.nu, self.ni = X.shape
        f = self.f = int(n_factors)
        self._keep = k = {}
        k["indptr"] = np.ascontiguousarray(X.indptr, np.int32)
        k["indices"] = np.ascontiguousarray(X.indices, np.int32)
        k["data"] = np.ascontiguousarray(X.data, np.float64)
        s = self.s = MFState()
        s.n_users, s.n_items, s.f, s.batch_size = self.nu, self.ni, f, int(batch_size)
        s.algorithm = {"MF_BPR": 0, "FUNK_SVD": 1, "ASY_SVD": 2}[algorithm_name]
        nuf = self.ni if algorithm_name == "ASY_SVD" else self.nu  # pyx:163-166: USER_factors has one row per ITEM
        s.use_bias = int(bool(use_bias))
        s.lr, s.user_reg, s.item_reg, s.bias_reg = learning_rate, user_reg, item_reg, bias_reg
        s.positive_reg, s.negative_reg, s.quota = positive_reg, negative_reg, negative_interactions_quota
        s.ad.mode, s.ad.gamma, s.ad.beta1, s.ad.beta2 = _MODES[sgd_mode], gamma, beta_1, beta_2
        s.ad.b1_pow, s.ad.b2_pow = beta_1, beta_2  # pyx:220-221
        s.indptr, s.indices, s.data, s.nnz = _p(k["indptr"]), _p(k["indices"]), _p(k["data"]), len(k["data"])
        lib()
        if random_seed is not None:  # pyx:145-147
            np.random.seed(seed=random_seed)
            _lib.glibc_rand_seed(C.byref(s.rng), C.c_uint(int(random_seed)))
        else:
            _lib.glibc_rand_seed(C.byref(s.rng), C.c_uint(1))
        if init_factors is None:
            k["U"] = np.random.normal(init_mean, init_std_dev, (nuf, f)).astype(np.float64)  # pyx:177-178
            k["V"] = np.random.normal(init_mean, init_std_dev, (self.ni, f)).astype(np.float64)
        else:
            k["U"] = np.array(init_factors[0], np.float64, copy=True)
            k["V"] = np.array(init_factors[1], np.float64, copy=True)

        def z(name, shape, dtype=np.float64):
            k[name] = np.zeros(shape, dtype)
            return _p(k[name])

        s.U, s.V = _p(k["U"]), _p(k["V"])
        s.accU, s.accV = z("accU", (nuf, f)), z("accV", (self.ni, f))
        s.bu, s.bi, s.mu = z("bu", self.nu), z("bi", self.ni), z("mu", 1)
        s.accbu, s.accbi, s.accmu = z("accbu", self.nu), z("accbi", self.ni), z("accmu", 1)
        if sgd_mode != "sgd":
            s.cU, s.cV, s.cbu, s.cbi, s.cmu = z("cU", (nuf, f)), z("cV", (self.ni, f)), z("cbu", self.nu), z("cbi", self.ni), z("cmu", 1)
            s.m1U, s.m2U, s.m1V, s.m2V = z("m1U", (nuf, f)), z("m2U", (nuf, f)), z("m1V", (self.ni, f)), z("m2V", (self.ni, f))
            s.m1bu, s.m2bu, s.m1bi, s.m2bi = z("m1bu", self.nu), z("m2bu", self.nu), z("m1bi", self.ni), z("m2bi", self.ni)
            s.m1mu, s.m2mu = z("m1mu", 1), z("m2mu", 1)
        s.items_list, s.users_list = z("il", 2 * int(batch_size), np.int64), z("ul", int(batch_size), np.int64)
        s.items_flag, s.users_flag = z("if", self.ni, np.int8), z("uf", self.nu, np.int8)
        if samples is not None:  # external (u, i, j|rating) stream instead of the glibc sampler
            k["eu"] = np.ascontiguousarray(samples[0], np.int32)
            k["ei"] = np.ascontiguousarray(samples[1], np.int32)
            s.ext_u, s.ext_i = _p(k["eu"]), _p(k["ei"])
            if s.algorithm == 0:
                k["ej"] = np.ascontiguousarray(samples[2], np.int32)
                s.ext_j = _p(k["ej"])
            else:
                k["er"] = np.ascontiguousarray(samples[2], np.float64)
                s.ext_r = _p(k["er"])
        if record:
            s.rec_u, s.rec_i, s.rec_j = z("ru", record, np.int32), z("ri", record, np.int32), z("rj", record, np.int32)

    def epochIteration_Cython(self):
        return int(lib().mf_epoch(C.byref(self.s)))

    def recorded(self):
        n = self.s.rec_pos
        return self._keep["ru"][:n].copy(), self._keep["ri"][:n].copy(), self._keep["rj"][:n].copy()

    def get_USER_factors(self):
        return self._keep["U"].copy()

    def get_ITEM_factors(self):
        return self._keep["V"].copy()

    def get_USER_bias(self):
        return self._keep["bu"].copy()

    def get_ITEM_bias(self):
        return self._keep["bi"].copy()

    def get_GLOBAL_bias(self):
        return np.array(self._keep["mu"][0])


class SLIMOracle:
    def __init__(self, URM_mask, train_with_sparse_weights=False, final_model_sparse_weights=True, learning_rate=0.01,
                 li_reg=0.0, lj_reg=0.0, topK=150, symmetric=True, verbose=False, random_seed=None, sgd_mode="adam",
                 gamma=0.995, beta_1=0.9, beta_2=0.999, samples=None, record=0):
        X = sps.csr_matrix(URM_mask, dtype=np.float32).sorted_indices()
        self.nu, self.ni = X.shape
        self.topK = min(topK, self.ni)
        self.tree = bool(train_with_sparse_weights)
        if self.tree:
            symmetric = False  # pyx:111-112
        self._keep = k = {}
        k["indptr"] = np.ascontiguousarray(X.indptr, np.int32)
        k["indices"] = np.ascontiguousarray(X.indices, np.int32)
        s = self.s = SLIMState()
        s.n_users, s.n_items, s.symmetric = self.nu, self.ni, int(bool(symmetric))
        s.lr, s.li_reg, s.lj_reg = learning_rate, li_reg, lj_reg
        s.ad.mode, s.ad.gamma, s.ad.beta1, s.ad.beta2 = _MODES[sgd_mode], gamma, beta_1, beta_2
        s.ad.b1_pow, s.ad.b2_pow = beta_1, beta_2
        s.indptr, s.indices = _p(k["indptr"]), _p(k["indices"])
        lib()
        _lib.glibc_rand_seed(C.byref(s.rng), C.c_uint(int(random_seed) if random_seed is not None else 1))
        k["S"] = np.zeros((self.ni, self.ni), np.float64)
        s.S = _p(k["S"])
        if self.tree:  # Sparse_Matrix_Tree_CSR restated as dense values + a cell-exists map (sgd_oracle.c slim_prune)
            k["exists"] = np.zeros((self.ni, self.ni), np.uint8)
            s.exists = _p(k["exists"])
            s.tree_topk = int(self.topK) if self.topK else 0
        if sgd_mode != "sgd":
            for n in ("c", "m1", "m2"):
                k[n] = np.zeros(self.ni, np.float64)
                setattr(s, n, _p(k[n]))
        if samples is not None:
            k["eu"], k["ei"], k["ej"] = (np.ascontiguousarray(a, np.int32) for a in samples)
            s.ext_u, s.ext_i, s.ext_j = _p(k["eu"]), _p(k["ei"]), _p(k["ej"])
        if record:
            for n in ("ru", "ri", "rj"):
                k[n] = np.zeros(record, np.int32)
            s.rec_u, s.rec_i, s.rec_j = _p(k["ru"]), _p(k["ri"]), _p(k["rj"])

    def epochIteration_Cython(self):
        return int(lib().slim_epoch(C.byref(self.s)))

    def recorded(self):
        n = self.s.rec_pos
        return self._keep["ru"][:n].copy(), self._keep["ri"][:n].copy(), self._keep["rj"][:n].copy()

    def get_S_tree(self):
        """get_S of the tree mode (pyx:340-388 with train_with_sparse_weights): the diagonal cells are touched, the rows are
        cut to their TopK IN PLACE (get_scipy_csr(TopK) replaces each row's list, pyx:762-763; zero-valued cells, the diagonal
        among them, count towards the row length and can take a place among the TopK) and the non-zero cells that are
        left are emitted (from_linked_list_to_python_list skips zeros, pyx:849-862)."""
        assert self.tree
        lib().slim_touch_diagonal(C.byref(self.s))
        if self.topK:
            lib().slim_prune(C.byref(self.s), C.c_long(int(self.topK)))
        S, E = self._keep["S"], self._keep["exists"]
        return sps.csr_matrix(np.where(E != 0, S, 0.0))

    def S_full(self):
        """The full item-item matrix (symmetric mode: mirrored lower triangle), diagonal as stored."""
        S = self._keep["S"]
        if self.s.symmetric:
            L = np.tril(S)
            return L + np.tril(S, -1).T
        return S.copy()
