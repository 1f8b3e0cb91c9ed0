"""Oracle: log-mel front end.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates ``model/fp/melspec/melspectrogram.py:10-112`` of the reference:

* ``:59-65``  zero-pad n_fft//2 on both sides,
* ``:82-89``  kapre ``STFT(n_fft, hop, pad_begin=False, pad_end=False)`` which is
  ``tf.signal.stft(frame_length=n_fft, frame_step=hop, fft_length=n_fft, window=hann(periodic))``,
* ``:90-92``  kapre ``Magnitude`` (``|X|``, not power),
* ``:93-98``  kapre ``ApplyFilterbank(type='mel')`` = librosa 0.8.1 ``filters.mel(htk=False,
  norm='slaney')`` transposed, applied with a dense tensordot,
* ``:104-109`` ``+0.06``, ``log10(max(., amin))``, minus the max over the WHOLE batch tensor,
  clamp at ``-dynamic_range``,
* ``:110-111`` optional ``melspec_maxnorm`` branch,
* ``:112``    ``Permute((3,2,1))`` -> (B, n_mels, T, 1).

PARITY UNPINNED against kapre/librosa/TF themselves (not installable here); the pieces are
cross-checked against torch.stft and torchaudio's Slaney filterbank in tests/test_oracle_melspec.py.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------- mel filterbank
def _hz_to_mel_slaney(f):
    """librosa 0.8.1 ``hz_to_mel(htk=False)`` (Slaney / Auditory-toolbox scale)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(log_t, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)
    return mels


def _mel_to_hz_slaney(m):
    """librosa 0.8.1 ``mel_to_hz(htk=False)``."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(fs=8000, n_fft=1024, n_mels=256, f_min=300.0, f_max=4000.0):
    """(n_mels, n_fft//2+1) float32 -- librosa 0.8.1 ``filters.mel(sr, n_fft, n_mels, fmin, fmax,
    htk=False, norm='slaney')``, which kapre 0.3.5 ``backend.filterbank_mel`` calls with those
    defaults for ``ApplyFilterbank(type='mel')`` (``melspectrogram.py:93-98``)."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, fs / 2.0, n_freq)
    mel_pts = np.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


# ----------------------------------------------------------------------------- STFT
def hann_periodic(n):
    """``tf.signal.hann_window(n, periodic=True)``: 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def stft_magnitude(x, n_fft=1024, hop=256, dtype=np.float64):
    """x: (B, L) already padded.  Returns |STFT| of shape (B, n_frames, n_fft//2+1);
    n_frames = 1 + (L - n_fft)//hop (``pad_end=False``)."""
    x = np.asarray(x, dtype=dtype)
    B, L = x.shape
    n_frames = 1 + (L - n_fft) // hop
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = x[:, idx] * hann_periodic(n_fft).astype(dtype)[None, None, :]
    spec = np.fft.rfft(frames.astype(np.float64), n=n_fft, axis=-1)
    return np.abs(spec).astype(dtype)


# ----------------------------------------------------------------------------- full layer
def melspec_layer(x, group_size=None, fs=8000, n_fft=1024, hop=256, n_mels=256, f_min=300.0,
                  f_max=4000.0, amin=1e-10, dynamic_range=80.0, segment_norm=False,
                  dtype=np.float64):
    """Restatement of ``Melspec_layer.call`` (``melspectrogram.py:102-112``).

    x: (B, 1, T) or (B, T) float.  ``group_size``: the reference evaluates one *batch* at a
    time (``model/generate.py:176-181``: consecutive ``TS_BATCH_SZ`` segments, last one partial)
    and subtracts the max over the whole batch tensor (``:108``); rows are processed here in
    consecutive groups of ``group_size`` (None = all rows are one batch).

    Returns (B, n_mels, n_frames, 1) in ``dtype``.
    """
    x = np.asarray(x)
    if x.ndim == 3:
        x = x[:, 0, :]
    B = x.shape[0]
    pad = n_fft // 2
    xp = np.pad(x.astype(dtype), ((0, 0), (pad, pad)))
    mag = stft_magnitude(xp, n_fft, hop, dtype)                       # (B, T, F)
    fb = mel_filterbank(fs, n_fft, n_mels, f_min, f_max).astype(dtype)  # (M, F)
    mel = mag @ fb.T                                                   # (B, T, M)
    y = mel + dtype(0.06)
    y = np.log(np.maximum(y, dtype(amin))) / dtype(math.log(10))
    if group_size is None:
        group_size = max(B, 1)
    out = np.empty_like(y)
    for s in range(0, B, group_size):
        g = y[s:s + group_size]
        g = g - g.max()
        g = np.maximum(g, dtype(-dynamic_range))
        if segment_norm:  # melspec_maxnorm branch (``:110-111``)
            g = (g - g.min() / 2) / np.abs(g.min() / 2 + dtype(1e-10))
        out[s:s + group_size] = g
    # (B, T, M) -> (B, M, T, 1): reference tensor is (B,1,T,M) permuted by (3,2,1).
    return np.ascontiguousarray(out.transpose(0, 2, 1))[..., None]
