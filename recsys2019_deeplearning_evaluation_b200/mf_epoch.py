"""Host-side mirror of MatrixFactorization/Cython/MatrixFactorization_Cython_Epoch.pyx:51-987 (BPR-MF, FunkSVD and AsySVD)
backed by libb200rec.so.  Same constructor signature, `epochIteration_Cython()` and getters (pyx:688-705).

Extra keyword arguments (not in the reference):
    sampler = "glibc"   sample stream replayed with glibc's srand(seed)/rand(): the reference's own stream
              "philox"  Philox4x32-10 drawn on the device (same acceptance rules, different numbers)
    hogwild = False     True drops the mini-batch barrier (every sample updates at once)

Declared differences (DESIGN.md "K2"): parameters and optimiser state are fp32 on the device (fp64 in pyx:78,
:177-178); `random_seed=None` starts glibc's default stream (seed 1) per object instead of continuing the process-wide
libc state.  ASY_SVD (pyx:396-578) runs on its own sequential kernel (csrc/asysvd.cu): batch size 1 only (pyx:399), the
reference's glibc sample stream only, `get_USER_factors()` is the n_items x f matrix Y like in the reference (pyx:163-166).
"""
import ctypes

import numpy as np
import scipy.sparse as sps

from . import _lib

_ALGO = {"MF_BPR": 0, "FUNK_SVD": 1}
_MODE = {"sgd": 0, "adagrad": 1, "rmsprop": 2, "adam": 3}
_SAMPLER = {"glibc": 0, "philox": 1}


class MatrixFactorization_Cython_Epoch:
    SGD_MODE_VALUES = ["sgd", "adam", "adagrad", "rmsprop"]
    ALGORITHM_NAME_VALUES = ["FUNK_SVD", "ASY_SVD", "MF_BPR"]

    def __init__(self, URM_train, n_factors=1, algorithm_name=None, batch_size=1, negative_interactions_quota=0.5,
                 learning_rate=1e-3, use_bias=False, user_reg=0.0, item_reg=0.0, bias_reg=0.0, positive_reg=0.0,
                 negative_reg=0.0, verbose=False, print_step_seconds=300, random_seed=None, init_mean=0.0,
                 init_std_dev=0.1, sgd_mode="sgd", gamma=0.995, beta_1=0.9, beta_2=0.999,
                 sampler="glibc", hogwild=False):
        self._h = ctypes.c_void_p()
        self._lib = _lib.load()
        if sgd_mode not in self.SGD_MODE_VALUES:  # pyx:108-109
            raise ValueError("Value for 'sgd_mode' not recognized. Acceptable values are {}, provided was '{}'".format(
                self.SGD_MODE_VALUES, sgd_mode))
        if algorithm_name not in self.ALGORITHM_NAME_VALUES:  # pyx:111-112
            raise ValueError("Value for 'algorithm_name' not recognized. Acceptable values are {}, provided was '{}'".format(
                self.ALGORITHM_NAME_VALUES, algorithm_name))
        if sampler not in _SAMPLER:
            raise ValueError("sampler must be 'glibc' or 'philox'")
        self._asy = algorithm_name == "ASY_SVD"
        if self._asy and (sampler != "glibc" or hogwild):
            raise ValueError("ASY_SVD runs the reference's sequential recursion on its glibc stream (sampler='glibc', hogwild=False)")
        X = sps.csr_matrix(URM_train, dtype=np.float32)  # check_matrix(URM_train, 'csr') + sorted_indices, pyx:116-117
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        self.n_users, self.n_items = X.shape
        self._nnz = int(X.nnz)
        self.n_factors = int(n_factors)
        self.batch_size = int(batch_size)
        self.algorithm_name = algorithm_name
        if random_seed is not None:  # pyx:145-147
            np.random.seed(seed=random_seed)
        # pyx:177-178: user factors first, then item factors, from numpy's legacy global RNG
        self.n_user_factor_rows = self.n_items if self._asy else self.n_users  # pyx:163-166
        U0 = np.random.normal(init_mean, init_std_dev, (self.n_user_factor_rows, self.n_factors)).astype(np.float64)
        V0 = np.random.normal(init_mean, init_std_dev, (self.n_items, self.n_factors)).astype(np.float64)
        indptr = np.ascontiguousarray(X.indptr, np.int32)
        indices = np.ascontiguousarray(X.indices, np.int32)
        data = np.ascontiguousarray(X.data, np.float32)
        self.use_bias = bool(use_bias)
        if self._asy:
            assert self.batch_size == 1, "Batch size other than 1 not supported for ASY_SVD"  # pyx:399
            _lib.check(self._lib.b200_asysvd_create(
                ctypes.byref(self._h), self.n_users, self.n_items, X.nnz, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
                self.n_factors, float(negative_interactions_quota), float(learning_rate), int(bool(use_bias)), float(user_reg),
                float(item_reg), float(bias_reg), _MODE[sgd_mode], float(gamma), float(beta_1), float(beta_2), _lib.ptr(U0),
                _lib.ptr(V0), int(random_seed is not None), int(random_seed) & 0xFFFFFFFF if random_seed is not None else 0))
            self._n_last = 0
            return
        _lib.check(self._lib.b200_mf_create(
            ctypes.byref(self._h), self.n_users, self.n_items, X.nnz, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
            self.n_factors, _ALGO[algorithm_name], self.batch_size, float(negative_interactions_quota), float(learning_rate),
            int(bool(use_bias)), float(user_reg), float(item_reg), float(bias_reg), float(positive_reg), float(negative_reg),
            _MODE[sgd_mode], float(gamma), float(beta_1), float(beta_2), _lib.ptr(U0), _lib.ptr(V0),
            int(random_seed is not None), int(random_seed) & 0xFFFFFFFF if random_seed is not None else 0,
            _SAMPLER[sampler], int(bool(hogwild))))

    def epochIteration_Cython(self):
        import torch
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self._asy:
            _lib.check(self._lib.b200_asysvd_epoch(self._h, st))
            self._n_last = self._nnz + 1
            return
        _lib.check(self._lib.b200_mf_epoch(self._h, st))

    def set_user_shard(self, user_lo, user_hi, samples_per_epoch=0, stream_id=0):
        assert not self._asy, "ASY_SVD is sequential: no user shards"
        _lib.check(self._lib.b200_mf_set_user_shard(self._h, int(user_lo), int(user_hi), int(samples_per_epoch), int(stream_id)))

    def device_factors(self):
        """(USER_factors, ITEM_factors) as torch CUDA float32 tensors that ALIAS the trainer's device memory."""
        import torch
        assert not self._asy, "ASY_SVD: read the factors through the getters"
        pu, pv = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self._lib.b200_mf_device_factors(self._h, ctypes.byref(pu), ctypes.byref(pv)))

        class _Arr:
            def __init__(s, ptr, shape):
                s.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

        dev = torch.device("cuda", torch.cuda.current_device())
        U = torch.as_tensor(_Arr(pu.value, (self.n_users, self.n_factors)), device=dev)
        V = torch.as_tensor(_Arr(pv.value, (self.n_items, self.n_factors)), device=dev)
        return U, V

    def samples_last_epoch(self):
        if self._asy:
            return self._n_last
        n = ctypes.c_int64()
        _lib.check(self._lib.b200_mf_samples_last_epoch(self._h, ctypes.byref(n)))
        return int(n.value)

    def last_epoch_ms(self):
        ms = ctypes.c_float()
        _lib.check((self._lib.b200_asysvd_last_epoch_ms if self._asy else self._lib.b200_mf_last_epoch_ms)(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def get_samples(self):
        """(u, i, j) for MF_BPR or (u, i, rating) for FUNK_SVD of the last epoch."""
        n = self.samples_last_epoch()
        u = np.empty(n, np.int32); i = np.empty(n, np.int32)
        if self.algorithm_name == "MF_BPR":
            j = np.empty(n, np.int32)
            _lib.check(self._lib.b200_mf_get_samples(self._h, _lib.ptr(u), _lib.ptr(i), _lib.ptr(j), None))
            return u, i, j
        r = np.empty(n, np.float32)
        if self._asy:
            _lib.check(self._lib.b200_asysvd_get_samples(self._h, _lib.ptr(u), _lib.ptr(i), _lib.ptr(r)))
            return u, i, r
        _lib.check(self._lib.b200_mf_get_samples(self._h, _lib.ptr(u), _lib.ptr(i), None, _lib.ptr(r)))
        return u, i, r

    def _get(self, which):
        U = np.empty((self.n_user_factor_rows, self.n_factors), np.float64) if which == 0 else None
        V = np.empty((self.n_items, self.n_factors), np.float64) if which == 1 else None
        bu = np.empty(self.n_users, np.float64) if which == 2 else None
        bi = np.empty(self.n_items, np.float64) if which == 3 else None
        mu = np.empty(1, np.float64) if which == 4 else None
        get = self._lib.b200_asysvd_get_factors if self._asy else self._lib.b200_mf_get_factors
        _lib.check(get(self._h, _lib.ptr(U), _lib.ptr(V), _lib.ptr(bu), _lib.ptr(bi), _lib.ptr(mu)))
        return (U, V, bu, bi, mu)[which]

    def get_USER_factors(self):
        return self._get(0)

    def get_ITEM_factors(self):
        return self._get(1)

    def get_USER_bias(self):
        return self._get(2)

    def get_ITEM_bias(self):
        return self._get(3)

    def get_GLOBAL_bias(self):
        return np.array(self._get(4)[0])

    def _dealloc(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            (self._lib.b200_asysvd_destroy if getattr(self, "_asy", False) else self._lib.b200_mf_destroy)(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self._dealloc()
        except Exception:
            pass

