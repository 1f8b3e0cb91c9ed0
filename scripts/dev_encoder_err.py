import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
from nafp_b200._lib import Context
from nafp_b200 import synth
from nafp_b200.model import weights as W, fp as FP
from oracle import melspec, fingerprinter as ofp
ctx = Context.get(0)
B = 100
xs = []
for t in range(10):
    tr = synth.synth_track(100 + t, 48000).astype(np.float32) / 32768.0
    xs += [tr[i * 4000: i * 4000 + 8000] for i in range(10)]
x = np.stack(xs).astype(np.float32)
w = W.init_weights(7, randomize_affine=False)
m_fp = FP.FingerPrinter(ctx).load(w)
emb = m_fp.fingerprint(x, 25)
mel = np.concatenate([melspec.melspec_layer(x[i:i + 25, None, :], group_size=25) for i in range(0, B, 25)])
ref = ofp.fingerprinter(mel, w)
d = np.abs(emb - ref)
print("segments", B, "max abs err %.3e  rms %.3e  min cos %.7f" % (d.max(), np.sqrt((d ** 2).mean()), (emb * ref).sum(1).min()))
