"""Golden vectors produced by the REFERENCE's own code, run in the build container.

Two pieces of the hot path are plain numpy in the reference and can run offline:

* the segmenter -- ``model/utils/audio_utils.py`` (``get_fns_seg_list`` / ``load_audio``; only ``wave`` + numpy),
* the evaluation loop -- ``eval/eval_faiss.py`` (offset compensation, candidate scoring, the four hit
  metrics, ``raw_score.npy``).  It imports ``faiss`` and ``curses``; both are replaced by minimal
  stand-ins *before* the import: ``curses`` by no-ops, ``faiss.IndexFlatL2`` by an exact numpy search with
  faiss' documented semantics (squared L2, ascending, ties to the lower id).  The loop itself, its
  memmap loader and ``get_index`` run UNMODIFIED from ``/root/reference``.

Writes ``ref_segments.npz`` and ``ref_eval_flat.npz`` next to this file; the inputs are regenerated
from seeds by the tests.  Run from the repo root (needs /root/reference, i.e. the build container):

    python tests/golden/make_reference_golden.py
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NAFP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from nafp_b200 import synth  # noqa: E402

SEG_WAV_LENGTHS = [30000, 8000, 5000, 44123, 8001, 12000]
EVAL = dict(n_dummy=60000, n_db=1180, seed=21, q_seed=22, sigma_lo=2.0, sigma_hi=4.5, seq_lens="1 3 5 9 11 19", k_probe=20,
            test_ids=sorted(set(list(range(0, 1161, 9)) + [1, 58, 59, 60, 1159, 1160, 1161])))     # <= n_query - 19


def eval_set():
    """(dummy, db, query) of the evaluation fixture; the tests rebuild the same arrays from the seeds."""
    dummy, db, _ = synth.synth_search_set(EVAL["n_dummy"], EVAL["n_db"], seed=EVAL["seed"])
    query = synth.synth_fp_queries(db, EVAL["q_seed"], EVAL["sigma_lo"], EVAL["sigma_hi"])
    return dummy, db, query


def segment_wavs(dirname):
    paths = []
    for i, n in enumerate(SEG_WAV_LENGTHS):
        p = os.path.join(dirname, f"t{i}.wav")
        synth.write_wav(p, synth.synth_track(100 + i, n_samples=n))
        paths.append(p)
    return paths


def write_emb_dir(dirname):
    dummy, db, query = eval_set()
    for name, arr in (("dummy_db", dummy), ("db", db), ("query", query)):
        mm = np.memmap(os.path.join(dirname, name + ".mm"), dtype="float32", mode="w+", shape=arr.shape)
        mm[:] = arr
        mm.flush()
        np.save(os.path.join(dirname, name + "_shape.npy"), np.asarray(arr.shape))
    return dummy, db, query


class _FlatL2:
    """numpy stand-in for faiss.IndexFlatL2 (exact; float64 distances, ties -> lower id)."""

    def __init__(self, d):
        self.d = d
        self.x = np.zeros((0, d), np.float32)
        self.nprobe = 1

    @property
    def ntotal(self):
        return len(self.x)

    def train(self, x):
        pass

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, np.float32)])

    def search(self, q, k):
        q = np.asarray(q, np.float64)
        x = self.x.astype(np.float64)
        d = (q * q).sum(1)[:, None] - 2.0 * q @ x.T + (x * x).sum(1)[None, :]
        order = np.lexsort((np.broadcast_to(np.arange(len(x)), d.shape), d), axis=1)[:, :k]
        D = np.take_along_axis(d, order, 1).astype(np.float32)
        I = order.astype(np.int64)
        if order.shape[1] < k:
            pad = k - order.shape[1]
            D = np.concatenate([D, np.full((len(q), pad), np.inf, np.float32)], 1)
            I = np.concatenate([I, np.full((len(q), pad), -1, np.int64)], 1)
        return D, I


def install_stubs():
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatL2 = _FlatL2
    sys.modules["faiss"] = faiss

    class _Scr:
        def __getattr__(self, name):
            return lambda *a, **k: None

    curses = types.ModuleType("curses")
    curses.initscr = lambda: _Scr()
    for fn in ("start_color", "use_default_colors", "init_pair", "endwin"):
        setattr(curses, fn, lambda *a, **k: None)
    curses.color_pair = lambda n: 0
    curses.COLOR_GREEN = curses.COLOR_CYAN = curses.COLOR_BLACK = 0
    curses.wrapper = lambda f, *a, **k: None
    sys.modules["curses"] = curses


def main():
    assert os.path.isdir(REF), f"{REF} not found: run in the build container"
    sys.path.insert(0, REF)
    tmp = tempfile.mkdtemp(prefix="nafp_golden_")
    try:
        # ---- segmenter: reference get_fns_seg_list + load_audio, as the generator calls them
        from model.utils import audio_utils as ref_audio
        wav_dir = os.path.join(tmp, "wav")
        os.makedirs(wav_dir)
        paths = segment_wavs(wav_dir)
        seg_list = ref_audio.get_fns_seg_list(paths, "all", 8000, 1.0, 0.5)
        segs = []
        for fn, seg_idx, _, _ in seg_list:
            # model/utils/dataloader_keras.py:303-311 (no augmentation): start = seg_idx * hop, no offset
            segs.append(ref_audio.load_audio(fn, seg_idx * 0.5, 0.0, 1.0, 0.0, fs=8000, amp_mode="normal"))
        np.savez_compressed(os.path.join(HERE, "ref_segments.npz"),
                            file_index=np.asarray([paths.index(s[0]) for s in seg_list], np.int32),
                            seg_index=np.asarray([s[1] for s in seg_list], np.int32),
                            offset_min=np.asarray([s[2] for s in seg_list], np.int64),
                            offset_max=np.asarray([s[3] for s in seg_list], np.int64),
                            audio=np.asarray(segs, np.float32))
        print(f"ref_segments.npz: {len(segs)} segments from {len(paths)} files")

        # ---- evaluation loop: reference eval_faiss (unmodified) on a small synthetic emb_dir
        install_stubs()
        emb = os.path.join(tmp, "emb") + "/"
        os.makedirs(emb)
        write_emb_dir(emb)
        ids_path = os.path.join(tmp, "ids.npy")
        np.save(ids_path, np.asarray(EVAL["test_ids"], np.int64))
        from eval import eval_faiss as ref_eval
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            ref_eval.eval_faiss.callback(emb, None, "l2", True, 10000000, ids_path, EVAL["seq_lens"], EVAL["k_probe"], 5)
        finally:
            os.chdir(cwd)
        raw = np.load(os.path.join(emb, "raw_score.npy"))
        ids = np.load(os.path.join(emb, "test_ids.npy"))
        np.savez_compressed(os.path.join(HERE, "ref_eval_flat.npz"), raw_score=raw, test_ids=ids)
        print("ref_eval_flat.npz: raw_score", raw.shape, "top-1 exact %", 100 * raw[:, :6].mean(0))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
