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
