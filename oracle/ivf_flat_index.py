"""Oracle: IVF-Flat index (index_type 'ivf').  TEST INFRASTRUCTURE ONLY.

Restates what the reference gets from ``faiss.IndexIVFFlat(faiss.IndexFlatL2(d), d, 400)``
(``eval/utils/get_index_faiss.py:63-66``), ``index.train`` (``:105-117``), ``index.nprobe = 40`` (``:120``),
``index.add`` / ``index.search`` (``eval/eval_faiss.py:147-148,211``), following the published algorithm of
faiss 1.6.5 (un-vendored; restated from its documentation):

* coarse quantizer: k-means (the same Lloyd restatement as ``oracle/ivfpq_index.py``) with nlist centroids;
* add: every row goes, uncompressed, to the inverted list of its nearest centroid (L2, ties -> lower id);
* search: the nprobe nearest lists of the query; exact squared-L2 distance to every row stored in them; the
  k smallest, ascending (ties -> lower label), labels = add order, -1 / +inf when fewer than k rows exist.

PARITY UNPINNED against faiss itself (k-means details differ); with the SAME centroids (``set_coarse``) the
search is deterministic and is compared id for id with the CUDA implementation.
"""
from __future__ import annotations

import numpy as np

from .ivfpq_index import kmeans


class IVFFlat:
    def __init__(self, d=128, nlist=400):
        self.d, self.nlist = d, nlist
        self.nprobe = 1
        self.coarse = None
        self.x = np.zeros((0, d), np.float32)
        self.assign = np.zeros((0,), np.int32)
        self.is_trained = False

    @property
    def ntotal(self):
        return len(self.x)

    def set_coarse(self, coarse):
        self.coarse = np.ascontiguousarray(coarse, np.float32).reshape(self.nlist, self.d)
        self.is_trained = True

    def train(self, x, seed=1234):
        self.coarse, _ = kmeans(np.ascontiguousarray(x, np.float32), self.nlist, seed=seed)
        self.is_trained = True

    def _coarse_dist(self, x):
        # direct differences (what the CUDA kernels compute), float32
        return ((x[:, None, :] - self.coarse[None, :, :]) ** 2).sum(-1, dtype=np.float32)

    def add(self, x, chunk=512):
        x = np.ascontiguousarray(x, np.float32)
        a = [self._coarse_dist(x[s:s + chunk]).argmin(1).astype(np.int32) for s in range(0, len(x), chunk)]
        self.x = np.concatenate([self.x, x])
        self.assign = np.concatenate([self.assign] + a)

    def search(self, q, k):
        q = np.ascontiguousarray(q, np.float32)
        nq = len(q)
        D = np.full((nq, k), np.inf, np.float32)
        I = np.full((nq, k), -1, np.int64)
        probes = np.argsort(self._coarse_dist(q), 1, kind="stable")[:, :self.nprobe]
        for i in range(nq):
            rows = np.nonzero(np.isin(self.assign, probes[i]))[0]
            if len(rows) == 0:
                continue
            ds = ((self.x[rows] - q[i]) ** 2).sum(1, dtype=np.float32)
            order = np.lexsort((rows, ds))[:k]
            D[i, :len(order)] = ds[order]
            I[i, :len(order)] = rows[order]
        return D, I
