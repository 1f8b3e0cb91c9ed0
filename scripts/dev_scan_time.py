"""Developer timing of flat_scan_kernel on a device-generated database (not the contract bench).
usage: dev_scan_time.py [db_rows] [n_query_rows]   -> scan ms/launch, GB/s, whole-search ms, stats, checksum of labels"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nafp_b200._lib import Context, check, lib
from nafp_b200.eval.utils.get_index import Index

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = Context.get(0)
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=dev))
idx = Index(0, 128)
idx.reserve(n + 29500)
buf = torch.empty((4_000_000, 128), dtype=torch.float32, device=dev)
r = 0
while r < n:
    m = min(4_000_000, n - r)
    check(lib.nafp_synth_fp_rows(ctx.h, 11, r, m, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
    idx.add_dev(buf.data_ptr(), m)
    r += m
check(lib.nafp_synth_fp_rows(ctx.h, 13, 0, 29500, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
idx.add_dev(buf.data_ptr(), 29500)
db = buf[:29500].clone()
g = torch.Generator(device=dev); g.manual_seed(5)
q = db[torch.arange(nq, device=dev) * 23 % 29500] + 0.35 / 128 ** 0.5 * torch.randn((nq, 128), device=dev, generator=g)
q = torch.nn.functional.normalize(q, dim=1).contiguous()
D = torch.empty((nq, 20), dtype=torch.float32, device=dev); I = torch.empty((nq, 20), dtype=torch.int64, device=dev)
for _ in range(3):
    idx.search_dev(q.data_ptr(), nq, 20, D.data_ptr(), I.data_ptr())
torch.cuda.synchronize()
idx.last_search_stats(); idx.profile_scans(True)
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    idx.search_dev(q.data_ptr(), nq, 20, D.data_ptr(), I.data_ptr())
e1.record(); torch.cuda.synchronize()
ms, launches = idx.profile_scans(False)
st = idx.last_search_stats()
N = idx.ntotal
per = ms / max(launches, 1)
tiles_per_cta = (N + 127) // 128 / 148
print(f"N={N} nq={nq}: scan {per*1e3:.1f} us/launch = {N*256/per/1e6:.0f} GB/s, {per*1e6/tiles_per_cta:.0f} ns/tile/CTA; "
      f"search {e0.elapsed_time(e1)/reps:.3f} ms; stats/rep { {k: v/reps for k, v in st.items()} }; "
      f"top1 ok {(I[:, 0] == (n + torch.arange(nq, device=dev) * 23 % 29500)).float().mean().item():.4f} "
      f"checksum {int(I.sum().item())} {float(D.double().sum().item()):.6f}")
