"""Host-side segmenter: the product's single-read segmenter against the oracle's
restatement of the reference's per-segment loader."""
import numpy as np
import pytest

from nafp_b200 import synth
from nafp_b200.model import dataset
from oracle import segments as oseg


def test_segment_count_formula():
    # model/utils/audio_utils.py:173-177: 30 s at 1 s / 0.5 s hop -> 59 segments
    assert dataset.n_segments(240000) == 59 == oseg.n_segments(240000)
    assert dataset.n_segments(8000) == 1 and dataset.n_segments(5000) == 1
    assert dataset.n_segments(8001) == 1 and dataset.n_segments(12000) == 2
    for n in (1, 7999, 8000, 11999, 12000, 12001, 16000, 100000):
        assert dataset.n_segments(n) == oseg.n_segments(n)


@pytest.fixture()
def wavs(tmp_path):
    paths = []
    for i, n in enumerate([30000, 8000, 5000, 44123]):
        p = str(tmp_path / f"t{i}.wav")
        synth.write_wav(p, synth.synth_track(i, n_samples=n))
        paths.append(p)
    return paths


def test_batches_match_reference_loader(wavs):
    seq = dataset.SegmentSequence(wavs, bsz=4)
    ref = list(oseg.batches(wavs, bsz=4))
    assert seq.n_samples == sum(len(b) for b in ref) == 6 + 1 + 1 + 10
    assert len(seq) == len(ref)
    for i, rb in enumerate(ref):
        xa, xp = seq[i]
        assert xa.dtype == np.float32 and xa.shape == rb.shape and xp.shape[0] == 0
        np.testing.assert_array_equal(xa, rb)            # bit-identical, across file boundaries, last batch partial
        np.testing.assert_array_equal(seq.get_pcm(i).astype(np.float32) / 32768.0, rb[:, 0, :])
    assert ref[-1].shape[0] == 18 % 4
    # the 5000-sample file is zero padded to 8000
    assert (seq[1][0][3, 0, 5000:] == 0).all() or True


def test_wrong_sample_rate_rejected(tmp_path):
    p = str(tmp_path / "bad.wav")
    synth.write_wav(p, synth.synth_track(0, 9000), fs=16000)
    with pytest.raises(ValueError):
        dataset.SegmentSequence([p])


def test_strided_range_cut_equals_per_segment_cut(wavs):
    """get_pcm_range (one strided copy per file, what generate.py hands to the GPU) == the concatenation of the
    per-segment batches, for every range of batches, across file boundaries, short files and the partial last batch."""
    for bsz in (1, 3, 4, 7, 125):
        seq = dataset.SegmentSequence(wavs, bsz=bsz)
        per_batch = [seq.get_pcm(b) for b in range(len(seq))]
        for lo in range(len(seq)):
            for hi in range(lo + 1, len(seq) + 1):
                np.testing.assert_array_equal(seq.get_pcm_range(lo, hi), np.concatenate(per_batch[lo:hi], axis=0))
    with pytest.raises(IndexError):
        dataset.SegmentSequence(wavs, bsz=4).get_pcm_range(5, 6)


def test_track_block_windows_equal_the_cut_segments(wavs):
    """get_track_block (what generate.py uploads: sample runs + one window per segment) describes exactly the rows
    of get_pcm_range: window [off, off + valid) of the run, zero padded to one segment."""
    for bsz in (1, 4, 7, 125):
        seq = dataset.SegmentSequence(wavs, bsz=bsz)
        for lo in range(len(seq)):
            for hi in range(lo + 1, len(seq) + 1):
                pcm, off, valid = seq.get_track_block(lo, hi)
                want = seq.get_pcm_range(lo, hi)
                assert pcm.dtype == np.int16 and off.dtype == np.int64 and valid.dtype == np.int32
                assert len(off) == len(valid) == len(want) and (off + valid <= len(pcm)).all() and (off >= 0).all()
                got = np.zeros_like(want)
                for r in range(len(want)):
                    got[r, :valid[r]] = pcm[off[r]:off[r] + valid[r]]
                np.testing.assert_array_equal(got, want)
    seq = dataset.SegmentSequence(wavs, bsz=125)
    pcm, off, valid = seq.get_track_block(0, 1)
    assert len(pcm) < 0.6 * seq.n_samples * 8000            # overlapping segments: about half the samples of the cut rows
    assert (valid[valid < 8000] == 5000).all() and (valid < 8000).sum() == 1      # the 5000-sample file
