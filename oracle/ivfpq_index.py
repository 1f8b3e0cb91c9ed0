"""Oracle: IVF-PQ index.  TEST INFRASTRUCTURE ONLY.

Restates what the reference gets from ``faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, 256, 64, 8)``
(``eval/utils/get_index_faiss.py:69-74``), ``index.train`` (``:105-117``), ``index.nprobe = 40``
(``:120``), ``index.add`` / ``index.search`` (``eval/eval_faiss.py:147-148,211``), following the
published algorithm of faiss 1.6.5 (un-vendored; restated from its documentation):

* training set: at most 256 points per centroid (256 * max(nlist, 2^nbits) rows) drawn by a seeded choice
  (``select_rows``, the same portable definition the CUDA library uses, so that "same seed" means "same rows");
* coarse quantizer: k-means (Lloyd, 25 iterations, initial centroids = seeded choice of training points, empty
  clusters keep their centroid) with nlist centroids, L2 assignment, ties to the lower id;
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


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def select_rows(n, cnt, seed):
    """Seeded choice of ``cnt`` of ``n`` rows, ascending: row i gets the key splitmix64(seed * 0xD1342543DE82EF95 + i),
    the rows with the ``cnt`` smallest (key, i) are taken.  Portable by construction (csrc/ivfpq.cu select_rows)."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed % (1 << 64)) * np.uint64(0xD1342543DE82EF95)
        keys = _splitmix64(base + np.arange(n, dtype=np.uint64))
    order = np.argsort(keys, kind="stable")[:min(cnt, n)]
    return np.sort(order)


def kmeans_batched(x, k, seeds, niter=25, chunk=8192):
    """G independent k-means problems at once: x (G, n, d) -> centroids ((G, k, d) float32 centroids, summed
    objective per iteration).  Lloyd iterations from a
    seeded choice of k training points per problem (``select_rows``); squared distances |c|^2 - 2 x.c in float64
    (the |x|^2 term does not change the arg-min), assignment ties to the lower id, centroid = float64 mean of its
    points, empty clusters keep their centroid.  torch-CPU (all host threads); chunked over the points."""
    import torch
    x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).double()
    G, n, d = x.shape
    cent = torch.stack([x[g, torch.from_numpy(select_rows(n, k, int(seeds[g]))[np.arange(k) % max(min(n, k), 1)])]
                        for g in range(G)])                                   # (G, k, d)
    offs = (torch.arange(G) * k)[:, None]
    x2 = (x * x).sum(2)                                                       # only for the reported objective
    obj = []
    for _ in range(niter):
        c2 = (cent * cent).sum(2)                                             # (G, k)
        sums = torch.zeros((G * k, d), dtype=torch.float64)
        cnt = torch.zeros(G * k, dtype=torch.float64)
        tot = 0.0
        for s0 in range(0, n, chunk):
            xb = x[:, s0:s0 + chunk]                                          # (G, b, d)
            dist = c2[:, None, :] - 2.0 * torch.bmm(xb, cent.transpose(1, 2))
            val, a = dist.min(2)                                              # first (lowest) index among equal minima
            tot = tot + float((val + x2[:, s0:s0 + chunk]).sum())
            a = a + offs
            sums.index_add_(0, a.reshape(-1), xb.reshape(-1, d))
            cnt.index_add_(0, a.reshape(-1), torch.ones(a.numel(), dtype=torch.float64))
        nz = cnt > 0
        new = cent.reshape(G * k, d).clone()
        new[nz] = sums[nz] / cnt[nz, None]
        cent = new.reshape(G, k, d)
        obj.append(tot)
    return cent.float().numpy(), obj


def kmeans(x, k, niter=25, seed=1234):
    """One k-means problem (see ``kmeans_batched``); returns (centroids (k, d) float32, objective per iteration)."""
    cent, obj = kmeans_batched(np.asarray(x, np.float32)[None], k, [seed], niter)
    return cent[0], obj


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
        x = x[select_rows(len(x), 256 * max(self.nlist, self.ksub), seed)]      # faiss: <= 256 points per centroid
        self.coarse, _ = kmeans(x, self.nlist, seed=seed + 1)
        r = x - self.coarse[self._assign(x)]
        sub = np.ascontiguousarray(r.reshape(len(r), self.m, self.dsub).transpose(1, 0, 2))     # (m, n, dsub)
        self.pq, _ = kmeans_batched(sub, self.ksub, [seed + 2 + j for j in range(self.m)])
        self.is_trained = True

    def _assign(self, x):
        d = (x * x).sum(1)[:, None] - 2.0 * (x @ self.coarse.T) + (self.coarse * self.coarse).sum(1)[None, :]
        return d.argmin(1).astype(np.int32)

    def add(self, x, chunk=16384):
        import torch
        x = np.ascontiguousarray(x, np.float32)
        codes, assign = [self.codes], [self.assign]
        pq = torch.from_numpy(self.pq)                                         # (m, ksub, dsub)
        p2 = (pq * pq).sum(2)                                                  # (m, ksub)
        for s in range(0, len(x), chunk):
            xb = x[s:s + chunk]
            a = self._assign(xb)
            r = torch.from_numpy((xb - self.coarse[a]).reshape(len(xb), self.m, self.dsub)).transpose(0, 1)   # (m, n, dsub)
            # arg-min over the ksub codewords of |r - c|^2 = |r|^2 - 2 r.c + |c|^2 (the |r|^2 term is common)
            d = p2[:, None, :] - 2.0 * torch.bmm(r, pq.transpose(1, 2))        # (m, n, ksub)
            codes.append(d.argmin(2).T.numpy().astype(np.uint8))
            assign.append(a)
        self.codes, self.assign = np.concatenate(codes), np.concatenate(assign)

    def reconstruct(self, rows):
        """xhat = coarse[list] + concat_m pq[m][code_m] (what the ADC distance is measured against)."""
        rows = np.asarray(rows)
        sub = self.pq[np.arange(self.m)[None, :], self.codes[rows]]          # (n, m, dsub)
        return self.coarse[self.assign[rows]] + sub.reshape(len(rows), self.d)

    def search_fast(self, q, k):
        """The same answer list-major, for scale: ADC(q, code) = |q - xhat|^2, so for every list one BLAS product of
        the queries that probe it with the list's reconstructions replaces nq * nprobe look-up-table scans.  The
        distances agree with ``search`` to fp32 rounding (~1e-6); used where hit rates are compared at >= 200 k rows
        and as the timed IVF-PQ CPU baseline.  torch-CPU, all host threads."""
        import torch
        q = np.ascontiguousarray(q, np.float32)
        nq = len(q)
        dc = (q * q).sum(1)[:, None] - 2.0 * (q @ self.coarse.T) + (self.coarse * self.coarse).sum(1)[None, :]
        probes = np.argsort(dc, 1, kind="stable")[:, :self.nprobe]
        order = np.argsort(self.assign, kind="stable")
        bounds = np.searchsorted(self.assign[order], np.arange(self.nlist + 1))
        candD = torch.full((nq, self.nprobe * k), float("inf"))
        candI = torch.full((nq, self.nprobe * k), -1, dtype=torch.int64)
        qt = torch.from_numpy(q)
        for l in range(self.nlist):
            rows = order[bounds[l]:bounds[l + 1]]
            qi, slot = np.nonzero(probes == l)
            if len(rows) == 0 or len(qi) == 0:
                continue
            xh = torch.from_numpy(self.reconstruct(rows).astype(np.float32))
            ql = qt[qi]
            d = (ql * ql).sum(1)[:, None] - 2.0 * (ql @ xh.T) + (xh * xh).sum(1)[None, :]
            kk = min(k, len(rows))
            dv, di = torch.topk(d, kk, dim=1, largest=False)
            cols = torch.from_numpy(slot)[:, None] * k + torch.arange(kk)[None, :]
            candD[torch.from_numpy(qi)[:, None], cols] = dv
            candI[torch.from_numpy(qi)[:, None], cols] = torch.from_numpy(rows)[di]
        D = np.full((nq, k), np.inf, np.float32)
        I = np.full((nq, k), -1, np.int64)
        cd, ci = candD.numpy(), candI.numpy()
        key_i = np.where(ci < 0, np.iinfo(np.int64).max, ci)
        sel = np.lexsort((key_i, cd), axis=1)[:, :k]
        D[:] = np.take_along_axis(cd, sel, 1)
        I[:] = np.take_along_axis(ci, sel, 1)
        return D, I

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


class IVFPQR(IVFPQ):
    """Oracle of ``faiss.IndexIVFPQR(flat, d, 256, 64, 8, M_refine=4, nbits_refine=4)`` (``get_index_faiss.py:75-85``),
    restated from faiss' published algorithm: a second product quantizer (4 sub-spaces of d/4 dims, 16 centroids each:
    2 bytes per row) is trained on and encodes the residual x - xhat of the IVF-PQ reconstruction; ``search`` takes
    ``k * k_factor`` (default 4) candidates from the IVF-PQ search and re-ranks them by |q - (xhat + rhat)|^2.
    PARITY UNPINNED against faiss (same caveats as ``IVFPQ``)."""

    M_REFINE, KSUB_REFINE, K_FACTOR = 4, 16, 4

    def __init__(self, d=128, nlist=256, m=64, nbits=8):
        super().__init__(d, nlist, m, nbits)
        self.rdsub = d // self.M_REFINE
        self.rpq = None                                   # (4, 16, d/4)
        self.rcodes = np.zeros((0, self.M_REFINE), np.uint8)

    def set_refine(self, rpq):
        self.rpq = np.ascontiguousarray(rpq, np.float32).reshape(self.M_REFINE, self.KSUB_REFINE, self.rdsub)

    def _first_level(self, x):
        """(list, PQ code, xhat) of rows x -- the IVF-PQ encoding ``add`` performs."""
        a = self._assign(x)
        r = (x - self.coarse[a]).reshape(len(x), self.m, self.dsub)
        d = ((r[:, :, None, :] - self.pq[None]) ** 2).sum(-1)
        c = d.argmin(2).astype(np.uint8)
        xhat = self.coarse[a] + self.pq[np.arange(self.m)[None, :], c].reshape(len(x), self.d)
        return a, c, xhat

    def train(self, x, seed=1234):
        x = np.ascontiguousarray(x, np.float32)
        super().train(x, seed=seed)
        xt = x[select_rows(len(x), 256 * max(self.nlist, self.ksub), seed)]
        _, _, xhat = self._first_level(xt)
        r2 = np.ascontiguousarray((xt - xhat).reshape(len(xt), self.M_REFINE, self.rdsub).transpose(1, 0, 2))
        self.rpq, _ = kmeans_batched(r2, self.KSUB_REFINE, [seed + 100 + j for j in range(self.M_REFINE)])

    def add(self, x, chunk=8192):
        x = np.ascontiguousarray(x, np.float32)
        rc = [self.rcodes]
        for s0 in range(0, len(x), chunk):
            xb = x[s0:s0 + chunk]
            _, _, xhat = self._first_level(xb)
            r2 = (xb - xhat).reshape(len(xb), self.M_REFINE, 1, self.rdsub)
            rc.append(((r2 - self.rpq[None]) ** 2).sum(-1).argmin(2).astype(np.uint8))
        super().add(x)
        self.rcodes = np.concatenate(rc)

    def search(self, q, k):
        q = np.ascontiguousarray(q, np.float32)
        _, I1 = super().search(q, k * self.K_FACTOR)
        D = np.full((len(q), k), np.inf, np.float32)
        I = np.full((len(q), k), -1, np.int64)
        for i in range(len(q)):
            ids = I1[i][I1[i] >= 0]
            if len(ids) == 0:
                continue
            rec = self.reconstruct(ids) + self.rpq[np.arange(self.M_REFINE)[None, :], self.rcodes[ids]].reshape(len(ids), self.d)
            d = ((q[i][None] - rec) ** 2).sum(1, dtype=np.float32)
            order = np.lexsort((ids, d))[:k]
            D[i, :len(order)] = d[order]
            I[i, :len(order)] = ids[order]
        return D, I

