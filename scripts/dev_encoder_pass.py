"""Developer helper for the ncu launch list: two fingerprint passes of N segments (one warm, one listed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200._lib import Context, lib, check
from nafp_b200.model import weights as W, fp as FP

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = Context.get(0)
m_fp = FP.FingerPrinter(ctx).load(W.init_weights(7))
xd = ctx.malloc(n * 32000)
check(lib.nafp_synth_audio(ctx.h, 5, 0, n, xd))
ed = ctx.malloc(n * 512)
for _ in range(2):
    check(lib.nafp_fingerprint(ctx.h, xd, n, 125, ed))
ctx.sync()
print("done", n)
