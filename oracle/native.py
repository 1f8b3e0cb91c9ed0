"""ctypes front end of the C oracle (oracle/csrc/oracle.c).  TEST INFRASTRUCTURE ONLY.

``FlatL2C`` / ``IVFPQC`` expose the faiss-like surface the reference's evaluation loop uses (``add`` / ``train`` /
``search`` / ``ntotal`` / ``nprobe``), so ``oracle.seq_match.evaluate`` and the timed CPU baselines of ``bench.py`` run
on them unchanged.  Results equal the numpy oracles (tests/test_oracle_c.py); these are the versions that scale."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

import numpy as np

from . import cbuild
from .ivfpq_index import select_rows

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(cbuild.build())
        _lib.orc_version.restype = c_int
        _lib.orc_max_threads.restype = c_int
        _lib.orc_set_threads.argtypes = [c_int]
        _lib.orc_flat_search.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]
        _lib.orc_kmeans.argtypes = [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
        _lib.orc_pq_encode.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]
        _lib.orc_ivfpq_search.argtypes = [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        for f in (_lib.orc_flat_search, _lib.orc_kmeans, _lib.orc_pq_encode, _lib.orc_ivfpq_search):
            f.restype = c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def threads():
    return int(lib().orc_max_threads())


def kmeans(x, k, niter=25, seed=1234):
    """Lloyd k-means from a seeded choice of k training points (``select_rows``) -> (k, d) float32."""
    x = _f32(x)
    n, d = x.shape
    init = select_rows(n, k, seed)
    cent = np.ascontiguousarray(x[init[np.arange(k) % max(len(init), 1)]])
    rc = lib().orc_kmeans(_p(x), n, d, d, k, _p(cent), niter, None)
    assert rc == 0, rc
    return cent


class FlatL2C:
    def __init__(self, d=128):
        self.d = int(d)
        self._chunks, self._x, self.ntotal = [], None, 0

    def train(self, x):
        return None

    def add(self, x):
        x = _f32(x)
        self._chunks.append(x)
        self.ntotal += len(x)
        self._x = None

    def _data(self):
        if self._x is None:
            self._x = self._chunks[0] if len(self._chunks) == 1 else np.concatenate(self._chunks, 0)
            self._chunks = [self._x]
        return self._x

    def reconstruct_n(self, i0, n):
        return self._data()[i0:i0 + n]

    def search(self, q, k):
        q = _f32(q)
        x = self._data() if self.ntotal else np.zeros((0, self.d), np.float32)
        D = np.empty((len(q), k), np.float32)
        I = np.empty((len(q), k), np.int64)
        rc = lib().orc_flat_search(_p(q), len(q), _p(x), len(x), self.d, k, _p(D), _p(I))
        assert rc == 0, rc
        return D, I


class IVFPQC:
    def __init__(self, d=128, nlist=256, m=64, nbits=8):
        self.d, self.nlist, self.m, self.ksub = d, nlist, m, 1 << nbits
        self.dsub = d // m
        self.nprobe = 1
        self.coarse = self.pq = None
        self.codes = np.zeros((0, m), np.uint8)
        self.assign = np.zeros((0,), np.int32)
        self.is_trained = False
        self._lists = None

    @property
    def ntotal(self):
        return len(self.codes)

    def set_params(self, coarse, pq):
        self.coarse = _f32(coarse).reshape(self.nlist, self.d)
        self.pq = _f32(pq).reshape(self.m, self.ksub, self.dsub)
        self.is_trained = True

    def train(self, x, seed=1234):
        """Same definition as oracle.ivfpq_index.IVFPQ.train and csrc/ivfpq.cu ivfpq_train: seeded training subset,
        coarse k-means (seed + 1), final assignment, residual k-means per sub-space (seed + 2 + j)."""
        x = _f32(x)
        x = np.ascontiguousarray(x[select_rows(len(x), 256 * max(self.nlist, self.ksub), seed)])
        n = len(x)
        init = select_rows(n, self.nlist, seed + 1)
        coarse = np.ascontiguousarray(x[init[np.arange(self.nlist) % max(len(init), 1)]])
        a = np.empty(n, np.int32)
        assert lib().orc_kmeans(_p(x), n, self.d, self.d, self.nlist, _p(coarse), 25, _p(a)) == 0
        r = np.ascontiguousarray(x - coarse[a])
        pq = np.empty((self.m, self.ksub, self.dsub), np.float32)
        for j in range(self.m):
            sub = r[:, j * self.dsub:(j + 1) * self.dsub]          # strided view: row stride d floats
            init = select_rows(n, self.ksub, seed + 2 + j)
            cent = np.ascontiguousarray(sub[init[np.arange(self.ksub) % max(len(init), 1)]])
            base = r.ctypes.data + j * self.dsub * 4
            assert lib().orc_kmeans(c_void_p(base), n, self.d, self.dsub, self.ksub, _p(cent), 25, None) == 0
            pq[j] = cent
        self.coarse, self.pq, self.is_trained = coarse, pq, True

    def add(self, x):
        x = _f32(x)
        a = np.empty(len(x), np.int32)
        c = np.empty((len(x), self.m), np.uint8)
        rc = lib().orc_pq_encode(_p(x), len(x), self.d, _p(self.coarse), self.nlist, _p(self.pq), self.m, self.ksub, _p(a), _p(c))
        assert rc == 0, rc
        self.codes = np.concatenate([self.codes, c])
        self.assign = np.concatenate([self.assign, a])
        self._lists = None

    def search(self, q, k):
        q = _f32(q)
        if self._lists is None:
            order = np.argsort(self.assign, kind="stable").astype(np.int64)
            loff = np.searchsorted(self.assign[order], np.arange(self.nlist + 1)).astype(np.int64)
            self._lists = (order, loff)
        order, loff = self._lists
        D = np.empty((len(q), k), np.float32)
        I = np.empty((len(q), k), np.int64)
        rc = lib().orc_ivfpq_search(_p(q), len(q), self.d, k, int(self.nprobe), _p(self.coarse), self.nlist, _p(self.pq), self.m,
                                    self.ksub, _p(self.codes), _p(order), _p(loff), _p(D), _p(I))
        assert rc == 0, rc
        return D, I
