"""Developer experiment: are the first encoder kernels faster per segment when a sub-chunk's activations fit in L2?
Runs device-resident fingerprint passes of n segments (n = 16, 32, 64, 1000); use under
ncu --cache-control none --metrics gpu__time_duration.sum to read the per-kernel times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200._lib import Context, lib, check
from nafp_b200 import synth
from nafp_b200.model import weights as W, fp as FP
ctx = Context.get(0)
tr = synth.synth_track(1).astype(np.float32) / 32768.0
x = np.stack([tr[i * 700: i * 700 + 8000] for i in range(8)]).astype(np.float32)
m_fp = FP.FingerPrinter(ctx).load(W.init_weights(7, randomize_affine=True))
xb = np.tile(x, (126, 1))[:1000]
xd = ctx.malloc(xb.nbytes); ctx.h2d(xd, xb); ed = ctx.malloc(1000 * 512)
for n in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 1000]:
    for _ in range(3): check(lib.nafp_fingerprint(ctx.h, xd, n, n, ed))
    ctx.sync(); ctx.timer_start()
    for _ in range(10): check(lib.nafp_fingerprint(ctx.h, xd, n, n, ed))
    ms = ctx.timer_stop() / 10
    print(f"n={n}: {ms*1e3:.1f} us per pass, {ms*1e3/n:.2f} us per segment")
