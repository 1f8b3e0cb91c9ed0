"""GPU parity of the in-training mini search (model/utils/mini_search_subroutines.py) against the
oracle, through the C ABI.  Distances are fp32 on both sides (the reference computes them in fp32 TensorFlow):
tolerance 2e-5 absolute; ranks come from comparing those sums, so an accuracy may differ by the few targets whose
competing scores tie to <= 1e-6 -- the gate is 0.1 pt on every accuracy and 1e-3 relative on the mean rank."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n_q, n_aug, n_d, d, seed, noise):
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((n_d, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    rho = 0.6                                    # neighbouring rows correlate like consecutive segments of a track
    for i in range(1, n_d):
        if i % 59:
            db[i] = rho * db[i - 1] + np.sqrt(1 - rho * rho) * db[i]
            db[i] /= np.linalg.norm(db[i])
    q = db[:n_q, None, :] + noise * rng.standard_normal((n_q, n_aug, d)).astype(np.float32) / np.sqrt(d)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q.astype(np.float32), db


@pytest.mark.parametrize("shape", [(300, 2, 420, 128), (150, 1, 150, 1024), (19, 3, 70, 8)])
def test_distances_and_conv_eye_match_oracle(ctx, shape):
    from nafp_b200.model.utils import mini_search_subroutines as gm
    from oracle import mini_search as om
    q, db = _data(*shape, seed=1, noise=1.5)
    for kw in (dict(), dict(return_dotprod=True), dict(squared=False)):
        got = gm.pairwise_distances_for_eval(q, db, ctx=ctx, **kw)
        ref = om.pairwise_distances_for_eval(q, db, **kw)
        assert got.shape == ref.shape == (shape[1], shape[0], shape[2], 1)
        assert np.abs(got - ref).max() < (2e-3 if kw.get("squared") is False else 2e-5)      # sqrt near 0 amplifies
    d2 = om.pairwise_distances_for_eval(q, db)
    for s in (1, 3, 19):
        got = gm.conv_eye_func(d2, s, ctx=ctx)
        ref = om.conv_eye_func(d2, s)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-5


@pytest.mark.parametrize("mode", ["argmin", "argmax"])
@pytest.mark.parametrize("shape,noise,off", [((300, 2, 420, 128), 3.5, 0), ((128, 1, 128, 1024), 9.0, 0), ((100, 2, 180, 64), 2.0, 11)])
def test_mini_search_eval_matches_oracle(ctx, mode, shape, noise, off):
    from nafp_b200.model.utils import mini_search_subroutines as gm
    from oracle import mini_search as om
    q, db = _data(*shape, seed=5, noise=noise)
    if off:
        db = np.concatenate([db[-off:], db])              # ground truth of query i is row i + off
    scopes = [1, 3, 5, 9, 11, 19]
    (g1, g3, g10), gr = gm.mini_search_eval(q, db, scopes, mode, display=False, gt_id_offset=off, ctx=ctx)
    (o1, o3, o10), orank = om.mini_search_eval(q, db, scopes, mode, gt_id_offset=off)
    assert 5 < o1[0] < 99.9                                # the test means something: scope 1 is neither trivial nor hopeless
    for g, o in ((g1, o1), (g3, o3), (g10, o10)):
        assert np.abs(g - o).max() <= 0.1, (g, o)
    assert np.allclose(gr, orank, rtol=1e-3, atol=1e-3), (gr, orank)


def test_mini_search_edge_cases(ctx):
    from nafp_b200._lib import NafpError
    from nafp_b200.model.utils import mini_search_subroutines as gm
    q, db = _data(20, 1, 20, 16, seed=2, noise=0.0)
    (t1, t3, t10), mr = gm.mini_search_eval(q, db, [1, 19, 20], display=False, ctx=ctx)
    assert (t1 == 100).all() and (mr == 0).all()           # scope == matrix size: a single target
    with pytest.raises(NafpError):
        gm.mini_search_eval(q, db, [21], display=False, ctx=ctx)      # scope larger than the matrix (TF would raise too)
    with pytest.raises(NotImplementedError):
        gm.mini_search_eval(q, db, [1], mode='nearest', ctx=ctx)
    # ground truth outside the valid range contributes nothing (np.where finds no rank)
    (t1, _, _), mr = gm.mini_search_eval(q, db, [1], display=False, gt_id_offset=100, ctx=ctx)
    assert t1[0] == 0 and mr[0] == 0
