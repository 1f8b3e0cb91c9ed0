"""Developer check of the log-mel kernel against the oracle + timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200._lib import Context, lib, check, ptr
from nafp_b200 import synth
from oracle import melspec
ctx = Context.get(0)
tr = synth.synth_track(1).astype(np.float32) / 32768.0
B = 250
x = np.stack([tr[i * 900: i * 900 + 8000] for i in range(B)]).astype(np.float32)
x[7] = 0; x[8] *= 1e-3
xd = ctx.malloc(x.nbytes); ctx.h2d(xd, x)
out = np.zeros((B, 256, 32), np.float32); od = ctx.malloc(out.nbytes)
check(lib.nafp_logmel_forward(ctx.h, xd, B, 125, od)); ctx.d2h(out, od); ctx.sync()
ref = melspec.melspec_layer(x[:, None, :], group_size=125)[..., 0]
ref32 = melspec.melspec_layer(x[:, None, :], group_size=125, dtype=np.float32)[..., 0]
print("max abs err vs fp64 oracle", np.abs(out - ref).max(), "fp32 oracle vs fp64", np.abs(ref32 - ref).max(), "range", ref.min(), ref.max())
n = 148 * 2 * 16
xb = ctx.malloc(n * 32000); ob = ctx.malloc(n * 32768)
for _ in range(3): check(lib.nafp_logmel_forward(ctx.h, xb, n, 125, ob))
ctx.sync(); ctx.timer_start()
for _ in range(10): check(lib.nafp_logmel_forward(ctx.h, xb, n, 125, ob))
ms = ctx.timer_stop() / 10
print(f"{n} segs: {ms*1e3:.1f} us -> {n/ms*1e3/1e6:.2f} M seg/s, {n*64768/ms/1e6:.0f} GB/s algorithmic")
