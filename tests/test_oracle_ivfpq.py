"""Oracle pinning, IVF-PQ: internal consistency of the numpy restatement."""
import numpy as np

from nafp_b200 import synth
from oracle.flat_index import FlatL2
from oracle.ivfpq_index import IVFPQ, kmeans


def test_kmeans_objective_decreases_and_is_seeded():
    x = synth.synth_fp_db(6000, seed=3)
    c1, obj1 = kmeans(x, 32, niter=10, seed=5)
    c2, _ = kmeans(x, 32, niter=10, seed=5)
    assert (c1 == c2).all()
    assert all(b <= a + 1e-3 for a, b in zip(obj1, obj1[1:]))
    assert obj1[-1] < 0.9 * obj1[0]


def test_ivfpq_recall_against_exact_search():
    dummy, db, query = synth.synth_search_set(6000, 590, seed=4)
    idx = IVFPQ(128, nlist=16, m=64)
    idx.train(dummy)
    idx.add(dummy)
    idx.add(db)
    idx.nprobe = 16                      # all lists: only the PQ approximation separates it from exact search
    D, I = idx.search(query[:40], 20)
    flat = FlatL2(128)
    flat.add(dummy)
    flat.add(db)
    _, Ie = flat.search(query[:40], 20)
    assert (I[:, 0] == Ie[:, 0]).mean() >= 0.9          # 64-byte codes on 128-d unit vectors are near-lossless for top-1
    assert (np.diff(D, axis=1) >= 0).all() and (I >= 0).all()
    idx.nprobe = 2
    D2, I2 = idx.search(query[:40], 20)
    assert 0.15 <= (I2[:, 0] == Ie[:, 0]).mean() <= (I[:, 0] == Ie[:, 0]).mean()   # fewer probes, lower recall
    # labels are insertion order; -1 padding when the probed lists hold fewer than k rows
    small = IVFPQ(128, nlist=16, m=64)
    small.set_params(idx.coarse, idx.pq)
    small.add(dummy[:5])
    small.nprobe = 16
    Ds, Is = small.search(query[:2], 20)
    assert (Is[:, 5:] == -1).all() and np.isinf(Ds[:, 5:]).all() and set(Is[0, :5]) == set(range(5))


def test_ivf_flat_all_lists_equals_exact_search_and_fewer_probes_lose_recall():
    """IVF-Flat restatement (index_type 'ivf'): probing every list IS the exact search; probing a few can only
    lose rows; labels are insertion order with -1 padding."""
    from oracle.ivf_flat_index import IVFFlat
    dummy, db, query = synth.synth_search_set(5000, 590, seed=14)
    idx = IVFFlat(128, nlist=20)
    idx.train(dummy)
    idx.add(dummy)
    idx.add(db)
    assert idx.ntotal == 5590 and idx.assign.min() >= 0 and idx.assign.max() < 20
    flat = FlatL2(128)
    flat.add(dummy)
    flat.add(db)
    De, Ie = flat.search(query[:40], 20)
    idx.nprobe = 20
    D, I = idx.search(query[:40], 20)
    assert (I == Ie).all()
    np.testing.assert_allclose(D, De, atol=2e-6)
    idx.nprobe = 2
    D2, I2 = idx.search(query[:40], 20)
    assert (np.diff(D2, axis=1) >= 0).all()
    assert (D2 >= D - 1e-6).all()                         # a subset of the rows: every rank can only get worse
    assert 0.15 <= (I2[:, 0] == Ie[:, 0]).mean() <= 1.0
    for r in range(5):                                    # every returned row lives in a probed list
        probes = np.argsort(idx._coarse_dist(query[r:r + 1]), 1, kind="stable")[0, :2]
        assert np.isin(idx.assign[I2[r]], probes).all()
    small = IVFFlat(128, nlist=20)
    small.set_coarse(idx.coarse)
    small.add(dummy[:5])
    small.nprobe = 20
    Ds, Is = small.search(query[:2], 20)
    assert (Is[:, 5:] == -1).all() and np.isinf(Ds[:, 5:]).all() and set(Is[0, :5]) == set(range(5))


def test_ivfpqr_refinement_improves_the_ranking():
    """IndexIVFPQR oracle: the 2-byte refinement codes shrink the reconstruction error and never hurt top-1 recall."""
    from oracle.ivfpq_index import IVFPQR
    dummy, db, query = synth.synth_search_set(6000, 590, seed=14)
    rr = IVFPQR(128, nlist=16, m=64)
    rr.train(dummy, seed=3)
    assert rr.rpq.shape == (4, 16, 32)
    rr.add(dummy)
    rr.add(db)
    rr.nprobe = 16
    assert rr.rcodes.shape == (6590, 4) and rr.rcodes.max() < 16
    x = np.concatenate([dummy, db])
    rows = np.arange(0, 6590, 13)
    e1 = np.linalg.norm(x[rows] - rr.reconstruct(rows), axis=1)
    rec2 = rr.reconstruct(rows) + rr.rpq[np.arange(4)[None, :], rr.rcodes[rows]].reshape(len(rows), 128)
    e2 = np.linalg.norm(x[rows] - rec2, axis=1)
    assert e2.mean() < e1.mean()
    flat = FlatL2(128)
    flat.add(dummy)
    flat.add(db)
    _, Ie = flat.search(query[:60], 1)
    D, I = rr.search(query[:60], 20)
    D0, I0 = IVFPQ.search(rr, query[:60], 20)
    assert (np.diff(D, axis=1) >= 0).all() and (I >= 0).all()
    assert (I[:, 0] == Ie[:, 0]).mean() >= (I0[:, 0] == Ie[:, 0]).mean() - 0.02
    # the refined list is a re-ordering of (a subset of) the 4k first-level candidates
    _, I4 = IVFPQ.search(rr, query[:60], 80)
    assert all(set(I[r]) <= set(I4[r]) for r in range(60))


def test_list_major_formulation_and_its_rounding_bound():
    """The algebra and the error bound csrc/ivfpq_lm.cu rests on, on the oracle's own quantities: with xhat = c_l + d
    (d = the decoded PQ residual), |q - xhat|^2 = |q - c_l|^2 - 2 (q . d - h) with h = 0.5 |d|^2 + c_l . d, and the score the
    tensor cores compute from bf16-rounded q and d differs from the exact q . d by at most E = |q| max|d| 2^-8 * 1.03 -- the
    slack the thresholds and the proof of the CUDA path subtract.  Also: -h split into three bf16 pieces (how the kernel
    feeds it to the MMA) reproduces h to fp32 accuracy."""
    import torch
    from oracle.ivfpq_index import IVFPQ
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6000, 128)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    o = IVFPQ(128, 32, 64, 8)
    o.train(x[:4000], seed=3)
    o.add(x)
    q = x[:64] + 0.3 * rng.standard_normal((64, 128)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    rows = np.arange(0, 6000, 7)
    d = o.pq[np.arange(64)[None, :], o.codes[rows]].reshape(len(rows), 128).astype(np.float64)      # decoded residuals
    c = o.coarse[o.assign[rows]].astype(np.float64)
    xhat = c + d
    np.testing.assert_allclose(xhat, o.reconstruct(rows), atol=1e-6)
    h = 0.5 * (d * d).sum(1) + (c * d).sum(1)
    q64 = q.astype(np.float64)
    lhs = ((q64[:, None, :] - xhat[None, :, :]) ** 2).sum(2)
    g = ((q64[:, None, :] - c[None, :, :]) ** 2).sum(2)
    rhs = g - 2.0 * (q64 @ d.T - h[None, :])
    np.testing.assert_allclose(lhs, rhs, atol=1e-10)
    # bf16 operands, exact accumulation: the worst case of what the MMA adds up
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float64).numpy()
    err = np.abs(bf(q) @ bf(d).T - q64 @ d.T)
    E = np.linalg.norm(q64, axis=1)[:, None] * np.sqrt((d * d).sum(1).max()) * (1.03 / 256.0)
    assert (err <= E).all() and err.max() > 0.02 * E.min()          # a real bound, not a vacuous one
    # -h = hi + mid + lo in bf16
    nh = (-h).astype(np.float32)
    hi = bf(nh)
    mid = bf((nh - hi).astype(np.float32))
    lo = bf((nh - hi - mid).astype(np.float32))
    assert np.abs((hi + mid + lo) - nh.astype(np.float64)).max() <= 2.0 ** -22 * np.abs(nh).max()
