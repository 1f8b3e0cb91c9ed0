"""Developer helper: fingerprint throughput (device-resident fp32 audio, CUDA events on the ctx stream) and the
per-stage split, for N segments per pass."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200._lib import Context, lib, check
from nafp_b200.model import weights as W, fp as FP

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = Context.get(0)
m_fp = FP.FingerPrinter(ctx).load(W.init_weights(7))
xd = ctx.malloc(n * 32000)
check(lib.nafp_synth_audio(ctx.h, 5, 0, n, xd))
ed = ctx.malloc(n * 512)
md = ctx.malloc(n * 32768)
for _ in range(3):
    check(lib.nafp_fingerprint(ctx.h, xd, n, 125, ed))
ctx.sync()
ms = ctypes.c_float()
check(lib.nafp_timer_start(ctx.h))
for _ in range(reps):
    check(lib.nafp_fingerprint(ctx.h, xd, n, 125, ed))
check(lib.nafp_timer_stop(ctx.h, ctypes.byref(ms)))
t_all = ms.value / reps
check(lib.nafp_timer_start(ctx.h))
for _ in range(reps):
    check(lib.nafp_logmel_forward(ctx.h, xd, n, 125, md))
check(lib.nafp_timer_stop(ctx.h, ctypes.byref(ms)))
t_mel = ms.value / reps
check(lib.nafp_timer_start(ctx.h))
for _ in range(reps):
    check(lib.nafp_encoder_forward(ctx.h, md, n, ed))
check(lib.nafp_timer_stop(ctx.h, ctypes.byref(ms)))
t_enc = ms.value / reps
print(f"n={n}: fingerprint {t_all:.3f} ms = {n / t_all * 1e3:,.0f} seg/s = {n / t_all * 1e3 * 607.2e6 / 1e12:.1f} TFLOP/s; "
      f"logmel(+finish) {t_mel:.3f} ms = {n / t_mel * 1e3:,.0f} seg/s = {n * 64768 / t_mel / 1e6:.1f} GB/s; encoder {t_enc:.3f} ms")
