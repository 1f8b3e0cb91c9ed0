"""Oracle pinning, flat index + sequence matcher."""
import os

import numpy as np

from nafp_b200 import synth
from oracle import seq_match
from oracle.flat_index import FlatL2

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _index(parts):
    idx = FlatL2(128)
    for p in parts:
        idx.add(p)
    return idx


def test_flat_l2_matches_scipy_cdist():
    from scipy.spatial.distance import cdist
    dummy, db, query = synth.synth_search_set(5000, 590, seed=1)
    idx = _index([dummy, db])
    D, I = idx.search(query[:19], 20)
    allx = np.concatenate([dummy, db])
    Dc = cdist(query[:19].astype(np.float64), allx.astype(np.float64), 'sqeuclidean')
    Ic = np.argsort(Dc, 1, kind='stable')[:, :20]
    assert (I == Ic).all()
    assert np.abs(D - np.take_along_axis(Dc, Ic, 1)).max() < 1e-6
    Df, If = idx.search(query[:19], 20, fast=True)     # the timed fp32 BLAS form agrees up to ties
    assert (If == I).mean() > 0.995


def test_flat_l2_padding_and_chunking():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((7, 128)).astype(np.float32)
    idx = _index([x])
    D, I = idx.search(x[:2], 10)
    assert (I[:, 7:] == -1).all() and np.isinf(D[:, 7:]).all() and I[0, 0] == 0 and I[1, 0] == 1
    big = rng.standard_normal((3000, 128)).astype(np.float32)
    a = _index([big]).search(big[:5], 8, chunk=512)
    b = _index([big]).search(big[:5], 8, chunk=100000)
    assert (a[1] == b[1]).all()


def test_icassp_test_ids_fixture():
    ids = np.load(os.path.join(os.path.dirname(GOLD), "..", "neural-audio-fp_b200", "eval", "test_ids_icassp2021.npy"))
    assert ids.shape == (2000,) and ids.dtype == np.int64
    assert ids.min() == 13 and ids.max() == 29492      # 8 rows from the end -> truncated sequences
    assert len(np.unique(ids // 59)) == 491


def test_seq_match_semantics():
    dummy, db, query = synth.synth_search_set(3000, 590, seed=3)
    idx = _index([dummy, db])
    recon = np.concatenate([dummy, db])
    # clamping at the end of the query set (eval_faiss.py:208): 5 rows left for a length-9 request
    p, s = seq_match.match_one(idx, query, recon, 585, 9)
    assert p[0] == 3000 + 585
    q = query[585:590]
    assert abs(s[0] - np.mean([q[j] @ recon[3585 + j] for j in range(5)])) < 1e-6
    raw, pred = seq_match.evaluate(idx, query, recon, 3000, np.array([0, 100, 585]), [1, 5, 9])
    assert raw.shape == (3, 12) and pred.shape == (3, 3, 10)
    rates = seq_match.hit_rates(raw, 3)
    assert rates.shape == (4, 3) and (rates[3] >= rates[0]).all()
    assert seq_match.hit_flags(np.array([7, 3, 9]), 8) == (0, 1, 0, 0)
    assert seq_match.hit_flags(np.array([7, 8, 9]), 8) == (0, 1, 1, 1)


def test_golden_search_fixture():
    g = np.load(os.path.join(GOLD, "search.npz"))
    dummy, db, query = synth.synth_search_set(30000, 1180, seed=5)
    idx = _index([dummy, db])
    D, I = idx.search(query[:40], 20)
    assert (I == g["I"]).all() and np.abs(D - g["D"]).max() < 1e-6
    raw, pred = seq_match.evaluate(idx, query, np.concatenate([dummy, db]), len(dummy), g["test_ids"], list(g["seq_lens"]), 20)
    assert (raw == g["raw"]).all() and (pred == g["pred"]).all()


def test_scalable_matcher_forms_equal_the_literal_loop():
    """oracle.seq_match.evaluate(fast_scores=True, batch_search=True) -- the forms bench.py times / runs at scale --
    give the predictions and hit flags of the literal restatement of eval_faiss.py:204-243."""
    from nafp_b200 import synth
    from oracle import native, seq_match
    from oracle.flat_index import FlatL2
    dummy, db, query = synth.synth_search_set(6000, 590, seed=21)
    o, c = FlatL2(128), native.FlatL2C(128)
    for idx in (o, c):
        idx.add(dummy)
        idx.add(db)
    recon = np.concatenate([dummy, db])
    ids = np.r_[np.arange(0, 560, 11), 585]                  # 585 + 19 runs past the end of the query set
    r1, p1 = seq_match.evaluate(o, query, recon, len(dummy), ids, [1, 3, 9, 19], 20)
    r2, p2 = seq_match.evaluate(c, query, recon, len(dummy), ids, [1, 3, 9, 19], 20, fast_scores=True)
    r3, p3 = seq_match.evaluate(c, query, recon, len(dummy), ids, [1, 3, 9, 19], 20, fast_scores=True, batch_search=True)
    assert (p1 == p2).all() and (p1 == p3).all() and (r1 == r2).all() and (r1 == r3).all()
