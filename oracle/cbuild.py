"""Builds oracle/csrc/oracle.c -> oracle/_build/liboracle.so with gcc (TEST INFRASTRUCTURE).

Compiled on the machine that uses it (this container, and again on the GPU box when the source hash or the machine
differs): `-mavx2 -mfma` only, no `-march=native`, so that a library built here still runs there."""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")
FLAGS = ["-O3", "-mavx2", "-mfma", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-Wall"]


def _digest():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    with open(SRC, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    tag = OUT + ".srchash"
    want = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(tag) and open(tag).read().strip() == want:
        return OUT
    r = subprocess.run(["gcc", *FLAGS, SRC, "-o", OUT, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on oracle.c:\n" + r.stdout + r.stderr)
    with open(tag, "w") as f:
        f.write(want)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
