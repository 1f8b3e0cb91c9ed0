"""Shared helpers for the parity tests."""
import numpy as np


def assert_topk_matches(D_gpu, I_gpu, D_ref, I_ref, x_all, q, tie=1e-6, dtol=2e-5):
    """Indices must be identical except where the reference scores tie within `tie`
    (BASELINE north_star: 'bit-exact outside score ties of <= 1e-6'); distances within dtol."""
    assert I_gpu.shape == I_ref.shape
    np.testing.assert_allclose(D_gpu, D_ref, rtol=0, atol=dtol)
    bad = np.argwhere(I_gpu != I_ref)
    for r, c in bad:
        ig = I_gpu[r, c]
        assert ig >= 0, f"row {r} rank {c}: gpu returned padding, ref {I_ref[r, c]}"
        d_true = float(((q[r].astype(np.float64) - x_all[ig].astype(np.float64)) ** 2).sum())
        assert abs(d_true - float(D_ref[r, c])) <= tie * 4, (
            f"row {r} rank {c}: gpu id {ig} (d={d_true}) vs ref id {I_ref[r, c]} (d={D_ref[r, c]}) is not a tie")
    return len(bad)
