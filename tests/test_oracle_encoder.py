"""Oracle pinning, FingerPrinter encoder: known answers the reference holds
(parameter counts, output shapes) and agreement with torch.nn.functional."""
import os

import numpy as np
import pytest

from nafp_b200.model import arch, weights
from oracle import fingerprinter as ofp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_param_count_known_answers():
    # model/fp/nnfp.py:270-274 -- "Total params: 19,224,576" is the 2 s-input (256,63,1) model;
    # the 1 s model built by get_fingerprinter (nnfp.py:248) has 16,939,008.
    assert ofp.param_count((256, 63, 1)) == 19_224_576
    assert ofp.param_count((256, 32, 1)) == 16_939_008
    assert arch.n_params((256, 63, 1)) == 19_224_576 and arch.n_params((256, 32, 1)) == 16_939_008
    w = weights.init_weights(0)
    assert sum(v.size for v in w.values()) == 16_939_008


def test_same_padding_table():
    # SURVEY §8 layer table: (lo, hi) on the convolved axis
    assert ofp.same_pad(32, 3, 2) == (16, 0, 1)
    assert ofp.same_pad(256, 3, 2) == (128, 0, 1)
    assert ofp.same_pad(2, 3, 1) == (2, 1, 1)
    assert ofp.same_pad(1, 3, 2) == (1, 1, 1)
    assert ofp.same_pad(2, 3, 2) == (1, 0, 1)
    rows = ofp.layer_table()
    assert [r[4] for r in rows][:4] == [(256, 16, 128), (128, 16, 128), (128, 8, 128), (64, 8, 128)]
    assert rows[-1][4] == (1, 1, 1024)
    specs = arch.conv_specs()
    assert [(s.f_out, s.t_out, s.c_out) for s in specs] == [r[4] for r in rows]
    assert arch.FLOPS_PER_SEGMENT == 607_199_232


def _torch_encoder(y, w):
    import torch
    import torch.nn.functional as F
    xt = torch.from_numpy(y.astype(np.float64)).permute(0, 3, 1, 2)
    for i, (sa, sb) in enumerate(ofp.FRONT_STRIDES):
        for tag, s in (('a', sa), ('b', sb)):
            k = torch.from_numpy(w[f'conv{i}_{tag}_w'].astype(np.float64)).permute(3, 2, 0, 1)
            kh, kw = k.shape[2:]
            _, flo, fhi = ofp.same_pad(xt.shape[2], kh, s[0])
            _, tlo, thi = ofp.same_pad(xt.shape[3], kw, s[1])
            xt = F.conv2d(F.pad(xt, (tlo, thi, flo, fhi)), k, torch.from_numpy(w[f'conv{i}_{tag}_b'].astype(np.float64)), stride=s)
            g = torch.from_numpy(w[f'ln{i}_{tag}_g'].astype(np.float64)).permute(2, 0, 1)
            b = torch.from_numpy(w[f'ln{i}_{tag}_b'].astype(np.float64)).permute(2, 0, 1)
            xt = F.layer_norm(F.elu(xt), xt.shape[1:], g, b, eps=1e-3)
    flat = xt.permute(0, 2, 3, 1).reshape(xt.shape[0], -1)
    xs = flat.reshape(-1, 128, flat.shape[1] // 128)
    h = F.elu(torch.einsum('bqs,qsu->bqu', xs, torch.from_numpy(w['div_w1'].astype(np.float64))) +
              torch.from_numpy(w['div_b1'].astype(np.float64)))
    o = torch.einsum('bqu,quo->bqo', h, torch.from_numpy(w['div_w2'].astype(np.float64)))[..., 0] + \
        torch.from_numpy(w['div_b2'].astype(np.float64))[:, 0]
    return F.normalize(o, dim=1).numpy()


def test_encoder_matches_torch_functional():
    rng = np.random.default_rng(3)
    y = rng.standard_normal((2, 256, 32, 1)) * 0.5 - 0.7
    w = weights.init_weights(11, randomize_affine=True)
    e = ofp.fingerprinter(y, w)
    assert e.shape == (2, 128)
    assert np.allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-12)
    assert np.abs(e - _torch_encoder(y, w)).max() < 1e-12
    e32 = ofp.fingerprinter(y, w, dtype=np.float32)
    assert np.abs(e32 - e).max() < 5e-6


def test_two_second_input_shape_like_reference_smoke():
    # model/fp/nnfp.py:261-268 builds the model on (3,256,63,1) too; only shapes are asserted there
    w = weights.init_weights(1, input_shape=(256, 63, 1))
    y = np.random.default_rng(4).standard_normal((1, 256, 63, 1)).astype(np.float32)
    flat = ofp.front_conv(y.astype(np.float64), {k: v.astype(np.float64) for k, v in w.items()})
    assert flat.shape == (1, 1024)      # T: 63 -> 32 -> 16 -> 8 -> 4 -> 4 -> 2 -> 2 -> 1


def test_golden_extractor_fingerprint():
    g = np.load(os.path.join(GOLD, "extractor.npz"))
    w = weights.init_weights(7, randomize_affine=True)
    e = ofp.fingerprinter(g["mel"][..., None], w)
    assert np.abs(e - g["emb"]).max() < 2e-6


def test_torch_cpu_extractor_equals_the_fp64_oracle():
    """oracle/torch_ref.py (the TIMED torch-CPU form used as the generation cpu_baseline) == the parity oracles."""
    from nafp_b200 import synth
    from nafp_b200.model import weights as W
    from oracle import fingerprinter as ofp
    from oracle import melspec, torch_ref
    w = W.init_weights(7, randomize_affine=True)
    tr = synth.synth_track(3).astype(np.float32) / 32768
    x = np.stack([tr[i * 4000:i * 4000 + 8000] for i in range(7)])[:, None, :]
    m_ref = melspec.melspec_layer(x, group_size=4)
    m_t = torch_ref.melspec_torch(x, group_size=4)
    assert m_t.shape == m_ref.shape and np.abs(m_t - m_ref).max() < 1e-5
    e_ref = ofp.fingerprinter(m_ref, w)
    e_t = torch_ref.TorchFingerPrinter(w)(m_t)
    assert np.abs(e_t - e_ref).max() < 1e-5 and (e_t * e_ref).sum(1).min() > 0.999999
