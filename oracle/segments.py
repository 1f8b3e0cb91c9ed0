"""Oracle: segment enumeration, loading and batching.  TEST INFRASTRUCTURE ONLY.

Restates the no-augmentation path the reference's generator takes for fingerprint generation:

* ``model/utils/audio_utils.py:140-218`` ``get_fns_seg_list(segment_mode='all')``:
  n_segs = (n_frames - seg + hop) // hop if n_frames > seg else 1,
* ``model/utils/audio_utils.py:221-264`` ``load_audio``: start = floor(seg_idx * hop_sec * fs),
  read ``seg`` frames of int16, ``/ 2**15`` (float64), zero-pad to ``seg`` samples,
* ``model/utils/dataloader_keras.py:132-141,186-193,223-228,303-311`` with
  ``drop_the_last_non_full_batch=False`` and ``n_anchor == bsz`` (``model/dataset.py:203-215``):
  batches are consecutive runs of ``bsz`` segments across file boundaries, last one partial,
  cast to float32 and shaped (n, 1, T).

Uses only the standard-library ``wave`` module, like the reference.
PINNED: bit-identical to the reference's ``get_fns_seg_list`` + ``load_audio`` outputs on six WAV files
of awkward lengths (``tests/golden/ref_segments.npz``, ``tests/test_reference_golden.py``).
"""
from __future__ import annotations

import wave

import numpy as np


def n_segments(n_frames, fs=8000, duration=1.0, hop=0.5):
    seg = fs * duration
    hp = fs * hop
    if n_frames > seg:
        return int((n_frames - seg + hp) // hp)
    return 1


def seg_list(filenames, fs=8000, duration=1.0, hop=0.5):
    """[(filename, seg_idx)] in file order -- the (filename, seg_idx) part of fns_event_seg_list."""
    out = []
    for fn in filenames:
        with wave.open(fn, "r") as w:
            if w.getframerate() != fs:
                raise ValueError(f"Sample rate should be {fs} but got {w.getframerate()}")
            n_frames = w.getnframes()
        for s in range(n_segments(n_frames, fs, duration, hop)):
            out.append((fn, s))
    return out


def load_segment(filename, seg_idx, fs=8000, duration=1.0, hop=0.5):
    start = int(np.floor(seg_idx * hop * fs))
    n = int(np.floor(duration * fs))
    with wave.open(filename, "r") as w:
        w.setpos(start)
        raw = w.readframes(n)
    x = np.frombuffer(raw, dtype=np.int16) / 2 ** 15
    out = np.zeros(int(duration * fs))
    out[:len(x)] = x
    return out


def batches(filenames, bsz=125, fs=8000, duration=1.0, hop=0.5):
    """Yield (n<=bsz, 1, T) float32 batches exactly as ``genUnbalSequence.__getitem__`` would."""
    segs = seg_list(filenames, fs, duration, hop)
    for s in range(0, len(segs), bsz):
        xs = [load_segment(fn, si, fs, duration, hop) for fn, si in segs[s:s + bsz]]
        yield np.expand_dims(np.vstack(xs), 1).astype(np.float32)
