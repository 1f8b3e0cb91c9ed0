"""GPU parity: fused log-mel kernel (a1) and FingerPrinter encoder (a2-a4) against the oracle,
through the C ABI.  Tolerances are the ones BASELINE.json states for the fingerprints
(cosine >= 0.9999, max abs error <= 1e-3); the log-mel is held to 1e-4."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def models(ctx):
    from nafp_b200.model import weights as W
    from nafp_b200.model.fp import FingerPrinter, Melspec
    w = W.init_weights(7, randomize_affine=True)
    return Melspec(ctx), FingerPrinter(ctx).load(w), w


def _audio(n, seed=1, hop=700):
    from nafp_b200 import synth
    tr = synth.synth_track(seed).astype(np.float32) / 32768.0
    return np.stack([tr[i * hop:i * hop + 8000] for i in range(n)]).astype(np.float32)


def test_logmel_matches_oracle_groups_and_edge_cases(models):
    from oracle import melspec
    m_pre = models[0]
    x = _audio(260)
    x[7] = 0.0                 # digital silence
    x[8] *= 1e-3               # very quiet
    x[130] = np.clip(x[130] * 20, -1, 1)      # clipped / loud
    y = m_pre(x[:, None, :], group_size=125)          # groups 125, 125, 10 (ragged tail)
    ref = melspec.melspec_layer(x[:, None, :], group_size=125)
    assert y.shape == (260, 256, 32, 1) and y.dtype == np.float32
    assert np.abs(y - ref).max() < 1e-4
    for g0 in (0, 125, 250):
        assert y[g0:g0 + 125].max() == 0.0            # batch-global max of every group
    y1 = m_pre(x[:3, None, :])                        # one group, tiny batch
    assert np.abs(y1 - melspec.melspec_layer(x[:3, None, :])).max() < 1e-4
    assert m_pre(x[:0, None, :]).shape == (0, 256, 32, 1)


def test_encoder_matches_oracle_layer_by_layer(models):
    from oracle import fingerprinter as ofp
    from oracle import melspec
    _, m_fp, w = models
    x = _audio(9, seed=2)
    mel = melspec.melspec_layer(x[:, None, :], group_size=9, dtype=np.float32)
    emb = m_fp(mel)
    ref, acts = ofp.fingerprinter(mel, w, return_all=True)
    for l in range(16):
        a = m_fp.activation(l, 9)
        assert a.shape == acts[l].shape
        err = np.abs(a - acts[l])
        assert err.mean() < 3e-3 and err.max() < 3e-2, (l, err.mean(), err.max())   # fp16 operands, values O(1)
    cos = (emb * ref).sum(1)
    assert cos.min() >= 0.9999 and np.abs(emb - ref).max() <= 1e-3, (cos.min(), np.abs(emb - ref).max())
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)


def test_fused_test_step_float_and_pcm16(models):
    from nafp_b200 import synth
    from oracle import fingerprinter as ofp
    from oracle import melspec
    _, m_fp, w = models
    pcm = np.stack([synth.synth_track(5)[i * 4000:i * 4000 + 8000] for i in range(13)])       # int16
    x = (pcm / 2 ** 15).astype(np.float32)                                                    # the reference's scaling
    ref = ofp.fingerprinter(melspec.melspec_layer(x[:, None, :], group_size=5), w)            # groups 5, 5, 3
    for emb in (m_fp.fingerprint(x[:, None, :], group_size=5), m_fp.fingerprint(pcm, group_size=5)):
        cos = (emb * ref).sum(1)
        assert cos.min() >= 0.9999 and np.abs(emb - ref).max() <= 1e-3
    np.testing.assert_array_equal(m_fp.fingerprint(x, 5), m_fp.fingerprint(pcm, 5))            # same arithmetic on the device


def test_many_segments_chunking_property(models):
    """More segments than one encoder pass (1000): a segment's fingerprint depends only on its own
    samples and on its group's maximum, not on where the group sits in the call."""
    _, m_fp, _ = models
    x = _audio(125, seed=3, hop=1000)
    big = np.tile(x, (17, 1))[:2100]               # 16 full groups + a partial one
    emb = m_fp.fingerprint(big, group_size=125)
    one = m_fp.fingerprint(x, group_size=125)
    for g in range(16):       # LayerNorm sums use fixed-order partial slots (no atomics): bit-reproducible
        np.testing.assert_array_equal(emb[g * 125:(g + 1) * 125], one)
    assert np.isfinite(emb).all() and np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)


def test_golden_fixture(models):
    from golden.make_golden import extractor_inputs
    m_pre, m_fp, _ = models
    g = np.load(os.path.join(GOLD, "extractor.npz"))
    x = extractor_inputs()
    assert np.abs(m_pre(x[:, None, :], group_size=3)[..., 0] - g["mel"]).max() < 1e-4
    emb = m_fp.fingerprint(x, group_size=3)
    assert (emb * g["emb"]).sum(1).min() >= 0.9999 and np.abs(emb - g["emb"]).max() <= 1e-3


def test_encoder_requires_weights(ctx):
    from nafp_b200._lib import NafpError
    from nafp_b200.model.fp import FingerPrinter
    import ctypes
    from nafp_b200._lib import Context
    fresh = Context(0)                 # a context that never saw nafp_weights_load
    with pytest.raises(NafpError):
        FingerPrinter(fresh)(np.zeros((1, 256, 32, 1), np.float32))


def test_track_windows_cut_on_the_gpu_equal_host_cut_segments(tmp_path):
    """nafp_fingerprint_pcm16_tracks_host (whole-track sample runs + one window per segment, what generate.py sends)
    returns bit for bit the fingerprints of the host-cut (n_seg, 8000) int16 rows -- short file (zero padding),
    file boundaries inside a group, partial last group."""
    from nafp_b200 import synth
    from nafp_b200._lib import Context
    from nafp_b200.model import dataset, fp as FP, weights as W
    paths = []
    for i, n in enumerate([30000, 8000, 5000, 44123, 240000]):
        p = str(tmp_path / f"t{i}.wav")
        synth.write_wav(p, synth.synth_track(i, n_samples=n))
        paths.append(p)
    seq = dataset.SegmentSequence(paths, bsz=7)
    m_fp = FP.FingerPrinter(Context.get(0)).load(W.init_weights(7, randomize_affine=True))
    rows = seq.get_pcm_range(0, len(seq))
    ref = m_fp.fingerprint(rows, group_size=7)
    pcm, off, valid = seq.get_track_block(0, len(seq))
    got = m_fp.fingerprint_tracks(pcm, off, valid, group_size=7)
    assert got.shape == ref.shape == (seq.n_samples, 128)
    np.testing.assert_array_equal(got, ref)
    assert len(pcm) < 0.6 * rows.size


def test_melspec_maxnorm_feature_branch(ctx, models):
    """MODEL.FEAT = 'melspec_maxnorm' (melspectrogram.py:110-111): x = (x - min/2) / |min/2 + 1e-10| over the batch
    tensor after the max subtraction and the clamp -- standalone log-mel and through the fused test_step."""
    from nafp_b200.model.fp import FingerPrinter, Melspec
    from oracle import fingerprinter as ofp
    from oracle import melspec
    _, _, w = models
    x = _audio(23, seed=4)
    x[5] = 0.0
    ref = melspec.melspec_layer(x[:, None, :], group_size=10, segment_norm=True)          # groups 10, 10, 3
    got = Melspec(ctx, segment_norm=True)(x[:, None, :], group_size=10)
    assert np.abs(got - ref).max() < 1e-4
    assert abs(got[:10].min() + 1.0) < 1e-5 and abs(got[:10].max() - 1.0) < 1e-5          # the branch maps [min, 0] to [-1, 1]
    m_fp = FingerPrinter(ctx, segment_norm=True).load(w)
    emb = m_fp.fingerprint(x, group_size=10)
    eref = ofp.fingerprinter(ref, w)
    assert (emb * eref).sum(1).min() >= 0.9999 and np.abs(emb - eref).max() <= 1e-3
    # and the default feature is untouched by it
    plain = Melspec(ctx)(x[:, None, :], group_size=10)
    assert np.abs(plain - melspec.melspec_layer(x[:, None, :], group_size=10)).max() < 1e-4
