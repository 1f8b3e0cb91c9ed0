"""Developer check of the encoder against the oracle, layer by layer."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200._lib import Context, lib, check, ptr
from nafp_b200 import synth
from nafp_b200.model import weights as W, fp as FP
from oracle import melspec, fingerprinter as ofp
ctx = Context.get(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
tr = synth.synth_track(1).astype(np.float32) / 32768.0
x = np.stack([tr[i * 700: i * 700 + 8000] for i in range(B)]).astype(np.float32)
w = W.init_weights(7, randomize_affine=True)
mel = melspec.melspec_layer(x[:, None, :], group_size=B, dtype=np.float32)
m_fp = FP.FingerPrinter(ctx).load(w)
emb = m_fp(mel)
ref, acts = ofp.fingerprinter(mel, w, dtype=np.float64, return_all=True)
for l in range(16):
    a = m_fp.activation(l, B)
    d = np.abs(a - acts[l])
    print(f"layer {l:2d} shape {a.shape} max abs err {d.max():.4e} mean {d.mean():.3e}  ref rms {np.sqrt((acts[l]**2).mean()):.3f}")
print("emb max abs err", np.abs(emb - ref).max(), "min cosine", (emb * ref).sum(1).min(), "norms", np.linalg.norm(emb, axis=1)[:3])
emb2 = m_fp.fingerprint(x, B)
print("fused vs split max diff", np.abs(emb2 - emb).max(), "fused vs ref", np.abs(emb2 - ref).max())
# timing
for n in (1000, 4000, 8000):
    xb = np.tile(x, (n // B + 1, 1))[:n]
    for _ in range(2): m_fp.fingerprint(xb, 125)
    t0 = time.time(); m_fp.fingerprint(xb, 125); dt = time.time() - t0
    print(f"e2e host->host {n} segs in {dt*1e3:.1f} ms -> {n/dt:.0f} seg/s")
    xd = ctx.malloc(xb.nbytes); ctx.h2d(xd, xb); ed = ctx.malloc(n * 512)
    for _ in range(2): check(lib.nafp_fingerprint(ctx.h, xd, n, 125, ed))
    ctx.sync(); ctx.timer_start()
    for _ in range(5): check(lib.nafp_fingerprint(ctx.h, xd, n, 125, ed))
    ms = ctx.timer_stop() / 5
    print(f"device {n} segs in {ms:.2f} ms -> {n/ms*1e3:.0f} seg/s = {n/ms*1e3*0.6072/1e3:.1f} TFLOP/s")
    ctx.free(xd); ctx.free(ed)
