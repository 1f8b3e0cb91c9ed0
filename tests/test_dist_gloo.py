"""world_size-2 gloo test of the row-sharded sequence matcher's HOST logic (SURVEY §8 e): shard
bounds, halo ownership, label offsets, the all-gather / merge / max-reduce orchestration of
nafp_b200.dist.sharded_seq_match.  The rank-local work is done by a numpy stand-in built on the oracle
(the CUDA kernels cannot run here); the result must equal the unsharded oracle evaluation."""
import os
import socket

import numpy as np
import pytest

from nafp_b200 import synth
from nafp_b200.dist import SEQ_MAXC, SeqPlan, TorchComm, shard_with_halo, sharded_seq_match
from oracle import seq_match
from oracle.flat_index import FlatL2

N_DUMMY, N_DB, MAX_LEN, K = 4000, 590, 9, 20
SEQ_LENS = [1, 3, 5, 9]
TEST_IDS = np.array([0, 7, 250, 581, 585, 589], dtype=np.int64)


class CpuOps:
    """numpy stand-in for nafp_b200.dist.GpuOps with the same contracts."""

    def __init__(self, x_all, rank, world, query):
        import torch
        self.torch = torch
        self.lo, self.hi, self.hi_halo = shard_with_halo(len(x_all), rank, world, MAX_LEN - 1)
        self.x = x_all[self.lo:self.hi_halo]
        self.n_global = len(x_all)
        self.query = query
        self.index = FlatL2(128)
        self.index.add(self.x[:self.hi - self.lo])         # halo rows are not searched

    def plan(self, query, test_ids):
        need = np.zeros(len(query), bool)
        for i in test_ids:
            need[i:i + MAX_LEN] = True
        uniq = np.nonzero(need)[0]
        pos = np.cumsum(need) - 1
        rowmap = np.full(len(test_ids) * MAX_LEN, -1, np.int32)
        for t, i in enumerate(test_ids):
            for j in range(MAX_LEN):
                if i + j < len(query):
                    rowmap[t * MAX_LEN + j] = pos[i + j]
        return SeqPlan(self.torch.from_numpy(query[uniq]), rowmap, uniq)

    def local_topk(self, qrows, k):
        D, I = self.index.search(qrows.numpy(), k)
        I = np.where(I >= 0, I + self.lo, -1)
        return self.torch.from_numpy(D), self.torch.from_numpy(I)

    def merge(self, D_all, I_all):
        W, n, k = D_all.shape
        D = D_all.numpy().transpose(1, 0, 2).reshape(n, W * k)
        I = I_all.numpy().transpose(1, 0, 2).reshape(n, W * k)
        D = np.where(I < 0, np.inf, D)
        order = np.lexsort((np.where(I < 0, np.iinfo(np.int64).max, I), D), axis=1)[:, :k]
        return self.torch.from_numpy(np.take_along_axis(D, order, 1)), self.torch.from_numpy(np.take_along_axis(I, order, 1))

    def cand_scores(self, query, plan, test_ids, seq_lens, k, I):
        I = I.numpy()
        cid = np.full((len(test_ids), SEQ_MAXC), -1, np.int64)
        csc = np.full((len(test_ids), len(seq_lens), SEQ_MAXC), -np.inf, np.float32)
        nc = np.zeros(len(test_ids), np.int32)
        for t, tid in enumerate(test_ids):
            lq = min(MAX_LEN, len(self.query) - int(tid))
            cands = {}
            for j in range(lq):
                for lab in I[plan.rowmap[t * MAX_LEN + j]]:
                    if lab >= 0 and lab - j >= 0:
                        cands[lab - j] = min(cands.get(lab - j, 99), j)
            keys = sorted(cands)
            nc[t] = len(keys)
            for u, c in enumerate(keys):
                cid[t, u] = c
                if not (self.lo <= c < self.hi):
                    continue
                avail = min(MAX_LEN, self.n_global - c, self.hi_halo - c)
                terms = min(avail, lq)
                d = [float(np.dot(query[int(tid) + j], self.x[c - self.lo + j])) for j in range(terms)]
                for li, sl in enumerate(seq_lens):
                    m = min(sl, terms)
                    if m > 0 and cands[c] < min(sl, lq):
                        csc[t, li, u] = np.float32(np.mean(d[:m]))
        return self.torch.from_numpy(cid), self.torch.from_numpy(csc), self.torch.from_numpy(nc)

    def top(self, cand_ids, cand_scores, n_cand, n_len):
        cid, csc = cand_ids.numpy(), cand_scores.numpy()
        pred = np.full((len(cid), n_len, 10), -1, np.int64)
        for t in range(len(cid)):
            for li in range(n_len):
                s = csc[t, li, :n_cand[t]]
                order = [u for u in np.argsort(-s, kind="stable") if s[u] > -np.inf][:10]
                pred[t, li, :len(order)] = cid[t, order]
        return pred, None


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dummy, db, query = synth.synth_search_set(N_DUMMY, N_DB, seed=8)
    ops = CpuOps(np.concatenate([dummy, db]), rank, world, query)
    pred, _ = sharded_seq_match(ops, TorchComm(), query, TEST_IDS, SEQ_LENS, K)
    if rank == 0:
        np.save(out, pred)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_match_equals_unsharded_oracle(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "pred.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    pred = np.load(out)
    dummy, db, query = synth.synth_search_set(N_DUMMY, N_DB, seed=8)
    idx = FlatL2(128)
    idx.add(dummy)
    idx.add(db)
    _, ref = seq_match.evaluate(idx, query, np.concatenate([dummy, db]), N_DUMMY, TEST_IDS, SEQ_LENS, K)
    np.testing.assert_array_equal(pred, ref)
