"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference itself (TensorFlow / kapre / faiss) cannot run offline, so these fixtures pin the
ORACLE's outputs on seeded inputs; the CPU tests check the oracle still reproduces them and the GPU
tests check the CUDA path against them.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nafp_b200 import synth  # noqa: E402
from nafp_b200.model import weights as W  # noqa: E402
from oracle import fingerprinter as ofp  # noqa: E402
from oracle import melspec as omel  # noqa: E402
from oracle import seq_match as oseq  # noqa: E402
from oracle.flat_index import FlatL2  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def extractor_inputs():
    tr = synth.synth_track(21).astype(np.float32) / 32768.0
    x = np.stack([tr[i * 4000:i * 4000 + 8000] for i in range(5)]).astype(np.float32)
    x[3] *= 0.01          # a quiet segment inside the second group
    return x


def main():
    x = extractor_inputs()
    mel = omel.melspec_layer(x[:, None, :], group_size=3)            # groups of 3 + 2
    w = W.init_weights(7, randomize_affine=True)
    emb = ofp.fingerprinter(mel, w)
    np.savez_compressed(os.path.join(HERE, "extractor.npz"), mel=mel[..., 0].astype(np.float32),
                        emb=emb.astype(np.float32))
    dummy, db, query = synth.synth_search_set(30000, 1180, seed=5)
    idx = FlatL2(128)
    idx.add(dummy)
    idx.add(db)
    D, I = idx.search(query[:40], 20)
    ids = np.array([0, 17, 400, 1161, 1175, 1179], dtype=np.int64)
    lens = [1, 3, 5, 9, 11, 19]
    raw, pred = oseq.evaluate(idx, query, np.concatenate([dummy, db]), len(dummy), ids, lens, 20)
    np.savez_compressed(os.path.join(HERE, "search.npz"), D=D, I=I, test_ids=ids, seq_lens=np.array(lens), raw=raw,
                        pred=pred)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
