"""Oracle: exact flat index.  TEST INFRASTRUCTURE ONLY.

Restates what the reference gets from ``faiss.IndexFlatL2(d)``
(``eval/utils/get_index_faiss.py:58``) through ``index.add`` (``eval/eval_faiss.py:147-148``)
and ``index.search(q, k)`` (``eval/eval_faiss.py:211``): squared-L2 distances, the k smallest per
query row in ascending order, int64 labels = insertion order, ``-1`` / ``+inf`` padding when fewer
than k rows exist.  faiss (1.6.5, un-vendored) computes ``sum((q-x)^2)`` with SIMD for nq < 20
and ``|q|^2+|x|^2-2qx`` through BLAS otherwise; both are fp32 roundings of the value computed
here in fp64, which is why the parity contract is "indices identical outside score ties <= 1e-6".
Ties are broken by the lower label.  PARITY UNPINNED against faiss itself; cross-checked against
scipy.spatial.distance.cdist in tests/test_oracle_search.py.
"""
from __future__ import annotations

import numpy as np


class FlatL2:
    def __init__(self, d):
        self.d = int(d)
        self._chunks = []
        self.ntotal = 0
        self._x = None

    # faiss API surface used by the reference
    def train(self, x):          # IndexFlat needs no training (get_index_faiss.py:116 comment)
        return None

    def add(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        assert x.ndim == 2 and x.shape[1] == self.d
        self._chunks.append(x)
        self.ntotal += x.shape[0]
        self._x = None

    def _data(self):
        if self._x is None:
            self._x = self._chunks[0] if len(self._chunks) == 1 else np.concatenate(self._chunks, 0)
            self._chunks = [self._x]
        return self._x

    def reconstruct_n(self, i0, n):
        return self._data()[i0:i0 + n]

    def search(self, q, k, chunk=262144, fast=False):
        """fast=False: fp64 distances (the parity oracle).  fast=True: fp32 BLAS
        (``|q|^2+|x|^2-2qx``), the form used when this oracle is *timed* as the CPU baseline."""
        x = self._data() if self.ntotal else np.zeros((0, self.d), np.float32)
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float32))
        nq = q.shape[0]
        D = np.full((nq, k), np.inf, dtype=np.float64)
        I = np.full((nq, k), -1, dtype=np.int64)
        ft = np.float32 if fast else np.float64
        qf = q.astype(ft)
        qn = (qf * qf).sum(1)
        for s in range(0, self.ntotal, chunk):
            xb = x[s:s + chunk].astype(ft)
            dist = qn[:, None] + (xb * xb).sum(1)[None, :] - 2.0 * (qf @ xb.T)
            kk = min(k, dist.shape[1])
            if kk < dist.shape[1]:
                part = np.argpartition(dist, kk - 1, axis=1)[:, :kk]
            else:
                part = np.broadcast_to(np.arange(dist.shape[1]), (nq, dist.shape[1]))
            cd = np.concatenate([D, np.take_along_axis(dist, part, 1).astype(np.float64)], 1)
            ci = np.concatenate([I, part.astype(np.int64) + s], 1)
            # order by (distance, label); padding (-1, inf) sorts last
            key_i = np.where(ci < 0, np.iinfo(np.int64).max, ci)
            order = np.lexsort((key_i, cd), axis=1)[:, :k]
            D = np.take_along_axis(cd, order, 1)
            I = np.take_along_axis(ci, order, 1)
        return D.astype(np.float32), I
