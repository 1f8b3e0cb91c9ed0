"""Oracle of the in-training mini search against independent implementations available here:
scipy cdist, torch conv2d with an identity kernel (the reference's own 'convolution trick'), and known answers."""
import numpy as np
import pytest


def _data(n_q=60, n_aug=2, n_d=80, d=16, seed=0, noise=0.4):
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((n_d, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = db[:n_q, None, :] + noise * rng.standard_normal((n_q, n_aug, d)).astype(np.float32) / np.sqrt(d)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q.astype(np.float32), db


def test_pairwise_distances_match_cdist():
    from scipy.spatial.distance import cdist
    from oracle import mini_search as ms
    q, db = _data()
    d2 = ms.pairwise_distances_for_eval(q, db)
    assert d2.shape == (2, 60, 80, 1)
    for a in range(2):
        assert np.abs(d2[a, :, :, 0] - cdist(q[:, a], db, "sqeuclidean")).max() < 1e-5
        assert np.abs(ms.pairwise_distances_for_eval(q, db, squared=False)[a, :, :, 0] - cdist(q[:, a], db)).max() < 1e-3
    dot = ms.pairwise_distances_for_eval(q, db, return_dotprod=True)
    assert np.abs(dot[1, :, :, 0] - q[:, 1] @ db.T).max() < 1e-6


def test_conv_eye_is_a_conv2d_with_identity_kernel():
    import torch
    import torch.nn.functional as F
    from oracle import mini_search as ms
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 20, 30, 1)).astype(np.float32)
    for s in (1, 3, 5, 19):
        got = ms.conv_eye_func(x, s)[..., 0]
        w = torch.eye(s).reshape(1, 1, s, s)
        ref = F.conv2d(torch.from_numpy(x[..., 0])[:, None], w)[:, 0].numpy()      # 'valid', as mini_search_subroutines.py:110-119
        assert got.shape == (3, 20 - s + 1, 30 - s + 1)
        assert np.abs(got - ref).max() < 1e-5


def test_mini_search_eval_known_answers():
    from oracle import mini_search as ms
    q, db = _data(noise=0.0)
    (t1, t3, t10), mr = ms.mini_search_eval(q, db, scopes=(1, 3, 5))
    assert (t1 == 100).all() and (t3 == 100).all() and (mr == 0).all()          # query == db rows: always rank 0
    (t1, _, t10), mr = ms.mini_search_eval(q, db, scopes=(1, 3), mode='argmax')
    assert (t1 == 100).all() and (mr == 0).all()
    # offset ground truth: shifting the database by 7 rows moves every hit with it
    db7 = np.concatenate([np.roll(db, 3, axis=1)[:7], db])
    (t1, _, _), mr = ms.mini_search_eval(q, db7, scopes=(1, 3), gt_id_offset=7)
    assert (t1 == 100).all()
    # noisy queries: longer scopes never hurt on average, accuracies are percentages
    q2, db2 = _data(noise=2.5, seed=3)
    (a1, a3, a10), mr2 = ms.mini_search_eval(q2, db2, scopes=(1, 5, 11))
    assert ((0 <= a1) & (a1 <= a3) & (a3 <= a10) & (a10 <= 100)).all() and a1[0] < a1[2]
    with pytest.raises(NotImplementedError):
        ms.mini_search_eval(q, db, mode='nearest')
