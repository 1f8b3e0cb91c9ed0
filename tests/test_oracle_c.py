"""The C oracle (oracle/csrc/oracle.c, the versions that scale) against the numpy oracles it restates, and the
portable seeded row selection shared by both oracles and the CUDA library."""
import numpy as np
import pytest


def _rows(n, seed, d=128):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def test_select_rows_known_answer():
    """csrc/ivfpq.cu select_rows uses the same definition (splitmix64 keys): this vector pins it."""
    from oracle.ivfpq_index import select_rows
    assert select_rows(1000, 10, 1234).tolist() == [44, 46, 128, 181, 230, 247, 600, 730, 906, 917]
    assert select_rows(5, 10, 7).tolist() == [0, 1, 2, 3, 4]
    a, b = select_rows(100000, 256, 1), select_rows(100000, 256, 2)
    assert len(set(a.tolist())) == 256 and (np.diff(a) > 0).all() and len(set(a.tolist()) & set(b.tolist())) < 10


def test_flat_search_c_equals_numpy_oracle():
    from oracle.flat_index import FlatL2
    from oracle.native import FlatL2C
    x, q = _rows(30000, 1), _rows(37, 2)
    x[100] = x[200]                              # exact duplicate: ties go to the lower label in both
    q[3] = x[200]
    c, o = FlatL2C(128), FlatL2(128)
    for part in (x[:12345], x[12345:]):
        c.add(part)
        o.add(part)
    for k in (1, 20, 128):
        Dc, Ic = c.search(q, k)
        Do, Io = o.search(q, k)
        assert np.abs(Dc - Do).max() < 1e-5
        same = Ic == Io
        assert same.mean() > 0.97
        for r, col in np.argwhere(~same):        # only inside ties: the two rows are equally far (<= 1e-6)
            assert abs(Dc[r, col] - Do[r, col]) < 1e-6
    assert (c.search(q, 20)[1][3, :2] == [100, 200]).all()
    e = FlatL2C(128)
    e.add(x[:7])
    D, I = e.search(q[:2], 20)                   # fewer rows than k: -1 / +inf padding like faiss
    assert (I[:, 7:] == -1).all() and np.isinf(D[:, 7:]).all() and (I[:, :7] >= 0).all()


def test_ivfpq_c_equals_numpy_oracle():
    from oracle.ivfpq_index import IVFPQ
    from oracle.native import IVFPQC
    x, q = _rows(20000, 3), _rows(40, 4)
    q[:20] = x[:20] + 0.05 * _rows(20, 5)
    c = IVFPQC(128, 32, 64, 8)
    c.train(x[:8000], seed=11)
    o = IVFPQ(128, 32, 64, 8)
    o.train(x[:8000], seed=11)
    # same algorithm, same seeded rows: the two trainings agree up to fp rounding of near-tied assignments
    assert np.abs(c.coarse - o.coarse).max() < 2e-2 and np.abs(c.coarse - o.coarse).mean() < 2e-4
    assert np.abs(c.pq - o.pq).mean() < 2e-3
    o.set_params(c.coarse, c.pq)                 # identical quantizers from here on
    c.add(x)
    o.add(x)
    assert (c.assign == o.assign).mean() > 0.9995 and (c.codes == o.codes).mean() > 0.9995
    o.codes, o.assign = c.codes, c.assign
    for nprobe in (1, 8, 32):
        c.nprobe = o.nprobe = nprobe
        Dc, Ic = c.search(q, 20)
        Do, Io = o.search(q, 20)
        assert np.abs(Dc - Do)[np.isfinite(Do)].max() < 5e-6 and (np.isfinite(Dc) == np.isfinite(Do)).all()
        assert (Ic == Io).mean() > 0.995
        Df, If = o.search_fast(q, 20)            # the BLAS formulation used at scale
        assert np.abs(Df - Do)[np.isfinite(Do)].max() < 5e-6 and (If == Io).mean() > 0.99


def test_kmeans_c_known_answer():
    from oracle import native
    rng = np.random.default_rng(0)
    centers = np.array([[0, 0], [10, 0], [0, 10], [10, 10]], np.float32)
    x = np.concatenate([c + 0.1 * rng.standard_normal((200, 2)).astype(np.float32) for c in centers])
    got = native.kmeans(x, 4, niter=25, seed=3)
    d = np.abs(got[:, None, :] - centers[None]).sum(2)
    # Lloyd from seeded points may merge two blobs; every centroid must at least sit on a blob mean or between two
    assert (d.min(1) < 0.2).sum() >= 2 and np.isfinite(got).all()
