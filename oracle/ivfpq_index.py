"""Oracle: IVF-PQ index.  TEST INFRASTRUCTURE ONLY.

Restates what the reference gets from ``faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, 256, 64, 8)``
(``eval/utils/get_index_faiss.py:69-74``), ``index.train`` (``:105-117``), ``index.nprobe = 40``
(``:120``), ``index.add`` / ``index.search`` (``eval/eval_faiss.py:147-148,211``), following the
published algorithm of faiss 1.6.5 (un-vendored; restated from its documentation):

* coarse quantizer: k-means (Lloyd, 25 iterations, at most 256 training points per centroid drawn by
  a seeded permutation) with nlist centroids, L2 assignment;
* product quantizer trained on the residuals x - c(x) (``by_residual=True``): M sub-spaces of
  d/M dims, 2^nbits centroids each, same k-means;
* add: code_m = argmin_c |r_m - pq_m[c]|^2 stored in the inverted list of the nearest coarse centroid;
* search: the nprobe nearest lists; for each, a look-up table T[m][c] = |(q - c_l)_m - pq_m[c]|^2 and the
  ADC distance sum_m T[m][code_m]; the k smallest over all probed lists, ascending, labels = add order.

PARITY UNPINNED against faiss itself (k-means initialisation / empty-cluster handling differ in
detail), which is why the contract for this index is the top-1 hit rate (within 0.1 pt), not ids.
Given the SAME centroids and codebooks (``set_params``) the search is deterministic and is compared
id-for-id with the CUDA implementation.
"""
from __future__ import annotations

import numpy as np


def kmeans(x, k, niter=25, seed=1234, max_points_per_centroid=256):
    """Lloyd iterations; returns (centroids (k,d) float32, objective per iteration)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(x)
    rng = np.random.default_rng(seed)
    if n > k * max_points_per_centroid:
        x = x[np.sort(rng.permutation(n)[:k * max_points_per_centroid])]
        n = len(x)
    cent = x[np.sort(rng.permutation(n)[:k])].copy()
    obj = []
    x2 = (x * x).sum(1)
    for _ in range(niter):
        d = x2[:, None] - 2.0 * (x @ cent.T) + (cent * cent).sum(1)[None, :]
        a = d.argmin(1)
        obj.append(float(d[np.arange(n), a].sum()))
        for c in range(k):
            m = a == c
            if m.any():
                cent[c] = x[m].mean(0)
    return cent, obj


class IVFPQ:
    def __init__(self, d=128, nlist=256, m=64, nbits=8):
        self.d, self.nlist, self.m, self.ksub = d, nlist, m, 1 << nbits
        self.dsub = d // m
        self.nprobe = 1
        self.coarse = None            # (nlist, d)
        self.pq = None                # (m, ksub, dsub)
        self.codes = np.zeros((0, m), np.uint8)
        self.assign = np.zeros((0,), np.int32)
        self.is_trained = False

    @property
    def ntotal(self):
        return len(self.codes)

    def set_params(self, coarse, pq):
        self.coarse = np.ascontiguousarray(coarse, np.float32).reshape(self.nlist, self.d)
        self.pq = np.ascontiguousarray(pq, np.float32).reshape(self.m, self.ksub, self.dsub)
        self.is_trained = True

    def train(self, x, seed=1234):
        x = np.ascontiguousarray(x, np.float32)
        self.coarse, _ = kmeans(x, self.nlist, seed=seed)
        r = x - self.coarse[self._assign(x)]
        self.pq = np.stack([kmeans(r[:, j * self.dsub:(j + 1) * self.dsub], self.ksub, seed=seed + 1 + j)[0]
                            for j in range(self.m)])
        self.is_trained = True

    def _assign(self, x):
        d = (x * x).sum(1)[:, None] - 2.0 * (x @ self.coarse.T) + (self.coarse * self.coarse).sum(1)[None, :]
        return d.argmin(1).astype(np.int32)

    def add(self, x, chunk=65536):
        x = np.ascontiguousarray(x, np.float32)
        for s in range(0, len(x), chunk):
            xb = x[s:s + chunk]
            a = self._assign(xb)
            r = (xb - self.coarse[a]).reshape(len(xb), self.m, self.dsub)
            # (n, m, ksub) distances
            d = ((r[:, :, None, :] - self.pq[None]) ** 2).sum(-1)
            self.codes = np.concatenate([self.codes, d.argmin(2).astype(np.uint8)])
            self.assign = np.concatenate([self.assign, a])

    def search(self, q, k):
        q = np.ascontiguousarray(q, np.float32)
        nq = len(q)
        D = np.full((nq, k), np.inf, np.float32)
        I = np.full((nq, k), -1, np.int64)
        dc = (q * q).sum(1)[:, None] - 2.0 * (q @ self.coarse.T) + (self.coarse * self.coarse).sum(1)[None, :]
        probes = np.argsort(dc, 1, kind="stable")[:, :self.nprobe]
        lists = [np.nonzero(self.assign == l)[0] for l in range(self.nlist)]
        for i in range(nq):
            ds, ids = [], []
            for l in probes[i]:
                rows = lists[l]
                if len(rows) == 0:
                    continue
                r = (q[i] - self.coarse[l]).reshape(self.m, 1, self.dsub)
                T = ((r - self.pq) ** 2).sum(-1).astype(np.float32)          # (m, ksub)
                ds.append(T[np.arange(self.m)[None, :], self.codes[rows]].sum(1, dtype=np.float32))
                ids.append(rows)
            if not ds:
                continue
            ds, ids = np.concatenate(ds), np.concatenate(ids)
            order = np.lexsort((ids, ds))[:k]
            D[i, :len(order)] = ds[order]
            I[i, :len(order)] = ids[order]
        return D, I
