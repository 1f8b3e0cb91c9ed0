"""GPU: the sharded index / device-resident matcher path that bench.py and the multi-GPU host use
(world = 1 here; the two-rank orchestration is covered on CPU by test_dist_gloo.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sharded_index_world1_equals_plain_index():
    from nafp_b200 import synth
    from nafp_b200.dist import ShardedFlatIndex
    from nafp_b200.eval.utils.get_index import Index
    from oracle import seq_match
    from oracle.flat_index import FlatL2
    dummy, db, query = synth.synth_search_set(40000, 1180, seed=6)
    s = ShardedFlatIndex(len(dummy) + len(db), 0, 1, max_len=19, device=0)
    s.add_from([dummy, db])
    ids = np.array([0, 3, 500, 1100, 1170, 1179], dtype=np.int64)
    lens = [1, 3, 5, 9, 11, 19]
    pred, _ = s.seq_match(query, ids, lens, 20)
    o = FlatL2(128)
    o.add(dummy)
    o.add(db)
    _, ref = seq_match.evaluate(o, query, np.concatenate([dummy, db]), len(dummy), ids, lens, 20)
    np.testing.assert_array_equal(pred, ref)
    p = Index(0, 128)
    p.add(dummy)
    p.add(db)
    pred2, _ = p.seq_match(query, ids, lens, 20)
    np.testing.assert_array_equal(pred2, ref)


def test_halo_rows_are_scored_but_not_searched():
    """A shard that holds rows [0, 30000) + an 18-row halo: labels carry the offset, halo rows never
    appear as search hits, yet sequences that run into the halo are scored in full."""
    from nafp_b200 import synth
    from nafp_b200.eval.utils.get_index import Index
    dummy = synth.synth_fp_db(30018, seed=9)
    idx = Index(0, 128)
    idx.add(dummy)
    idx.set_search_rows(30000)
    idx.set_label_offset(1000)
    D, I = idx.search(dummy[29990:30018], 5)
    assert (I[:10, 0] == np.arange(29990, 30000) + 1000).all()      # owned rows find themselves (+offset)
    assert (I[10:] < 30000 + 1000).all()                            # halo rows are never returned
    idx.set_search_rows(-1)
    D2, I2 = idx.search(dummy[30010:30012], 1)
    assert (I2[:, 0] == np.array([30010, 30011]) + 1000).all()


def test_device_synth_rows_are_deterministic_unit_norm(ctx):
    import ctypes
    from nafp_b200._lib import check, lib
    n = 5900
    out = np.empty((n, 128), np.float32)
    part = np.empty((1000, 128), np.float32)
    d = ctx.malloc(out.nbytes)
    check(lib.nafp_synth_fp_rows(ctx.h, 11, 0, n, 59, 0.5, d))
    ctx.d2h(out, d)
    check(lib.nafp_synth_fp_rows(ctx.h, 11, 2345, 1000, 59, 0.5, d))          # any slice regenerates identically
    ctx.d2h(part, d)
    ctx.sync()
    ctx.free(d)
    np.testing.assert_array_equal(part, out[2345:3345])
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    c = (out[:-1] * out[1:]).sum(1).reshape(-1)
    same_track = (np.arange(n - 1) % 59) != 58
    assert 0.4 < c[same_track].mean() < 0.6 and abs(c[~same_track].mean()) < 0.1      # AR(1) rho = 0.5 inside tracks


def _two_rank_match(shards, query, ids, lens, k):
    """sharded_seq_match of nafp_b200.dist for two shards that live on ONE GPU: the collectives are replaced by
    their definitions (stack = all-gather, element-wise maximum = max all-reduce)."""
    import torch
    dev = torch.device("cuda", 0)
    ops = [s.ops(len(query)) for s in shards]
    q = torch.from_numpy(np.ascontiguousarray(query, np.float32)).to(dev)
    t_ids = torch.from_numpy(np.asarray(ids, np.int64)).to(dev)
    t_sl = torch.from_numpy(np.asarray(lens, np.int32)).to(dev)
    plans = [o.plan(q, t_ids) for o in ops]
    loc = [o.local_topk(p.qrows, k) for o, p in zip(ops, plans)]
    _, I = ops[0].merge(torch.stack([d for d, _ in loc]), torch.stack([i for _, i in loc]))
    cs = [o.cand_scores(q, p, t_ids, t_sl, k, I) for o, p in zip(ops, plans)]
    assert (cs[0][0] == cs[1][0]).all() and (cs[0][2] == cs[1][2]).all()        # same candidates on every rank
    scores = torch.maximum(cs[0][1], cs[1][1])
    pid, _ = ops[0].top(cs[0][0], scores, cs[0][2], len(lens))
    torch.cuda.synchronize()
    return pid.cpu().numpy()


@pytest.mark.parametrize("kind,nlist,nprobe", [("ivfpq", 256, 40), ("ivfpq", 256, 2), ("ivf", 400, 40), ("ivf", 400, 3)])
def test_two_shards_of_an_ivf_index_equal_the_unsharded_index(kind, nlist, nprobe):
    """SURVEY §8 e for the IVF types: quantizers replicated, codes / lists sharded by the same contiguous row blocks
    (+ halo), per-shard top-k merged -- the predictions equal those of one index over all rows.  nprobe 2 / 3 sends
    most query rows through the list-scan fallback, which must not return halo rows either."""
    from nafp_b200 import synth
    from nafp_b200.dist import ShardedFlatIndex
    from nafp_b200.eval.utils.get_index import IVF_FLAT, IVFPQ, Index
    itype = IVFPQ if kind == "ivfpq" else IVF_FLAT
    dummy, db, query = synth.synth_search_set(30000, 1180, seed=16)
    plain = Index(itype, 128, nlist=nlist)
    plain.train(dummy, seed=5)
    plain.add(dummy)
    plain.add(db)
    plain.nprobe = nprobe
    shards = []
    for r in range(2):
        s = ShardedFlatIndex(len(dummy) + len(db), r, 2, max_len=9, device=0, index_type=itype, nlist=nlist)
        if itype == IVFPQ:
            s.index.set_ivfpq_params(*plain.ivfpq_params())
        else:
            s.index.set_ivf_coarse(plain.ivf_coarse())
        s.index.nprobe = nprobe
        s.add_from([dummy, db])
        assert s.index.ntotal == s.hi_halo - s.lo
        shards.append(s)
    ids = np.arange(0, 1150, 29, dtype=np.int64)
    lens = [1, 3, 5, 9]
    ref, _ = plain.seq_match(query, ids, lens, 20)
    got = _two_rank_match(shards, query, ids, lens, 20)
    np.testing.assert_array_equal(got[:, :, 0], ref[:, :, 0])
    assert (got == ref).mean() >= 0.995        # beyond the top-1: equal up to ties between equal scores


def test_ivf_halo_rows_are_not_returned_by_the_list_scan():
    from nafp_b200 import synth
    from nafp_b200.eval.utils.get_index import IVF_FLAT, Index
    x = synth.synth_fp_db(30018, seed=19)
    idx = Index(IVF_FLAT, 128, nlist=64)
    idx.train(x[:20000], seed=3)
    idx.add(x)
    idx.set_search_rows(30000)
    idx.set_label_offset(500)
    idx.nprobe = 1                               # too few probed rows among the 64 nearest: exact list scan
    D, I = idx.search(x[29990:30018], 20)
    assert (I[:10, 0] == np.arange(29990, 30000) + 500).all()
    assert (I < 30000 + 500).all()
