"""Host-side segmenter for fingerprint generation.

Mirrors the no-augmentation data path of the reference: ``Dataset.get_test_dummy_db_ds`` /
``get_test_query_db_ds`` / ``get_custom_db_ds`` (``model/dataset.py:182-323``) build a
``genUnbalSequence(bsz == n_anchor == TS_BATCH_SZ, shuffle=False, drop_the_last_non_full_batch=False)``
whose batches are consecutive runs of TS_BATCH_SZ one-second segments (0.5 s hop) across file
boundaries (``model/utils/dataloader_keras.py:132-141,186-193,223-228``); segment enumeration and
loading follow ``model/utils/audio_utils.py:140-264``.

Differences by design (results identical): every WAV file is read ONCE (the reference re-opens the
file for each segment), the segments of a file are cut with one strided copy (``get_pcm_range``: the
per-segment Python loop was 4x slower than the GPU that consumes its output), and batches can be handed
over as int16 PCM so that the ``/ 2**15`` scaling (``audio_utils.py:243-244``) happens on the GPU.
"""
from __future__ import annotations

import glob
import wave

import numpy as np


def n_segments(n_frames, fs=8000, duration=1.0, hop=0.5):
    """``audio_utils.py:173-177``."""
    seg, hp = fs * duration, fs * hop
    if n_frames > seg:
        return int((n_frames - seg + hp) // hp)
    return 1


def _wav_info(path, fs):
    with wave.open(path, 'r') as w:
        if w.getframerate() != fs:
            raise ValueError('Sample rate should be {} but got {}'.format(str(fs), str(w.getframerate())))
        if w.getsampwidth() != 2 or w.getnchannels() != 1:
            raise NotImplementedError(f'{path}: only 16-bit mono PCM is supported (like the reference reader)')
        return w.getnframes()


class SegmentSequence:
    """Indexable sequence of batches, the role ``genUnbalSequence`` plays in ``generate.py:170-181``."""

    def __init__(self, filenames, bsz=125, duration=1.0, hop=0.5, fs=8000):
        self.filenames = list(filenames)
        self.bsz = int(bsz)
        self.fs, self.duration, self.hop = int(fs), float(duration), float(hop)
        self.seg_len = int(self.duration * self.fs)
        # (file index, seg_idx) for every segment, in file order -- fns_event_seg_list
        self.file_nseg = [n_segments(_wav_info(fn, self.fs), self.fs, self.duration, self.hop) for fn in self.filenames]
        self.file_first = np.concatenate([[0], np.cumsum(self.file_nseg)]).astype(np.int64)
        self.n_samples = int(self.file_first[-1])
        self._cache_idx = -1
        self._cache_pcm = None

    def __len__(self):
        return int(np.ceil(self.n_samples / float(self.bsz)))

    def _file_pcm(self, fi):
        if fi != self._cache_idx:
            with wave.open(self.filenames[fi], 'r') as w:
                raw = w.readframes(w.getnframes())
            self._cache_pcm = np.frombuffer(raw, dtype=np.int16)
            self._cache_idx = fi
        return self._cache_pcm

    def _segment_pcm(self, fi, seg_idx, out):
        pcm = self._file_pcm(fi)
        start = int(np.floor(seg_idx * self.hop * self.fs))
        x = pcm[start:start + self.seg_len]
        out[:len(x)] = x
        out[len(x):] = 0

    def get_pcm(self, idx):
        """Batch ``idx`` as int16 (n, seg_len); n == bsz except for the last batch."""
        lo, hi = idx * self.bsz, min((idx + 1) * self.bsz, self.n_samples)
        if lo >= hi:
            raise IndexError(idx)
        out = np.empty((hi - lo, self.seg_len), dtype=np.int16)
        fi = int(np.searchsorted(self.file_first, lo, side='right') - 1)
        for r, g in enumerate(range(lo, hi)):
            while g >= self.file_first[fi + 1]:
                fi += 1
            self._segment_pcm(fi, g - int(self.file_first[fi]), out[r])
        return out

    def get_pcm_range(self, b_lo, b_hi):
        """Batches ``b_lo .. b_hi-1`` as one int16 array (n, seg_len) -- the same rows as the concatenation of
        ``get_pcm(b)``, cut file by file: the segments s0..s1 of a file are rows of a strided view of its samples
        (start = floor(s * hop * fs), ``audio_utils.py:246-247``), copied in one assignment."""
        lo, hi = b_lo * self.bsz, min(b_hi * self.bsz, self.n_samples)
        if lo >= hi:
            raise IndexError((b_lo, b_hi))
        out = np.empty((hi - lo, self.seg_len), dtype=np.int16)
        fi = int(np.searchsorted(self.file_first, lo, side='right') - 1)
        g = lo
        while g < hi:
            while g >= self.file_first[fi + 1]:
                fi += 1
            s0 = g - int(self.file_first[fi])
            s1 = min(int(self.file_nseg[fi]), s0 + (hi - g))
            pcm = self._file_pcm(fi)
            starts = np.floor(np.arange(s0, s1) * self.hop * self.fs).astype(np.int64)
            need = int(starts[-1]) + self.seg_len
            if len(pcm) < need:                       # a file shorter than one segment: zero padded (audio_utils.py:249-252)
                pcm = np.concatenate([pcm, np.zeros(need - len(pcm), np.int16)])
            win = np.lib.stride_tricks.sliding_window_view(pcm, self.seg_len)
            out[g - lo:g - lo + (s1 - s0)] = win[starts]
            g += s1 - s0
        return out

    def get_track_block(self, b_lo, b_hi):
        """Batches ``b_lo .. b_hi-1`` WITHOUT cutting the segments: (pcm, seg_off, seg_valid) -- the sample runs of
        the files the batches touch, back to back (int16), and for every segment its start offset in ``pcm`` (int64)
        and its number of real samples (int32, < seg_len only for a file shorter than one segment).  The GPU cuts
        the overlapping windows (``nafp_fingerprint_pcm16_tracks_host``): every sample is uploaded once."""
        lo, hi = b_lo * self.bsz, min(b_hi * self.bsz, self.n_samples)
        if lo >= hi:
            raise IndexError((b_lo, b_hi))
        runs, offs, valid = [], [], []
        base = 0
        fi = int(np.searchsorted(self.file_first, lo, side='right') - 1)
        g = lo
        while g < hi:
            while g >= self.file_first[fi + 1]:
                fi += 1
            s0 = g - int(self.file_first[fi])
            s1 = min(int(self.file_nseg[fi]), s0 + (hi - g))
            pcm = self._file_pcm(fi)
            starts = np.floor(np.arange(s0, s1) * self.hop * self.fs).astype(np.int64)
            first, last = int(starts[0]), min(int(starts[-1]) + self.seg_len, len(pcm))
            runs.append(pcm[first:last])
            offs.append(starts - first + base)
            valid.append(np.minimum(self.seg_len, len(pcm) - starts).astype(np.int32))
            base += last - first
            g += s1 - s0
        return np.concatenate(runs), np.concatenate(offs), np.concatenate(valid)

    def __getitem__(self, idx):
        """(Xa, Xp) like the reference generator: Xa float32 (n, 1, T), Xp empty."""
        xa = (self.get_pcm(idx) / 2 ** 15).astype(np.float32)[:, None, :]
        return xa, np.zeros((0, 1, self.seg_len), dtype=np.float32)


class Dataset:
    """File selection of ``model/dataset.py:35-323`` for the generation splits."""

    def __init__(self, cfg):
        self.source_root_dir = cfg['DIR']['SOURCE_ROOT_DIR']
        self.datasel_test_dummy_db = cfg['DATA_SEL']['TEST_DUMMY_DB']
        self.datasel_test_query_db = cfg['DATA_SEL']['TEST_QUERY_DB']
        self.ts_batch_sz = cfg['BSZ']['TS_BATCH_SZ']
        self.dur = cfg['MODEL']['DUR']
        self.hop = cfg['MODEL']['HOP']
        self.fs = cfg['MODEL']['FS']

    def _seq(self, fps):
        return SegmentSequence(fps, self.ts_batch_sz, self.dur, self.hop, self.fs)

    def get_test_dummy_db_ds(self):
        fps = sorted(glob.glob(self.source_root_dir + 'test-dummy-db-100k-full/' + '**/*.wav', recursive=True))
        sel = str(self.datasel_test_dummy_db)
        if sel in ['10k_full', '10k_30s']:
            fps = fps[:10000]
        elif sel == '100k_full_icassp':
            pass
        elif sel.isnumeric():
            fps = fps[:int(sel)]
        else:
            raise NotImplementedError(sel)
        return self._seq(fps)

    def get_test_query_db_ds(self):
        if self.datasel_test_query_db == 'unseen_icassp':
            q = sorted(glob.glob(self.source_root_dir + 'test-query-db-500-30s/' + 'query/**/*.wav', recursive=True))
            d = sorted(glob.glob(self.source_root_dir + 'test-query-db-500-30s/' + 'db/**/*.wav', recursive=True))
            return self._seq(q), self._seq(d)
        if self.datasel_test_query_db == 'unseen_syn':
            raise NotImplementedError("'unseen_syn' synthesises queries with the training-time augmentation "
                                      "chain, which is outside the inference hot path")
        raise NotImplementedError(self.datasel_test_query_db)

    def get_custom_db_ds(self, source_root_dir):
        fps = sorted(glob.glob(source_root_dir + '/**/*.wav', recursive=True))
        return self._seq(fps)
