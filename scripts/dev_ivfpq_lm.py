"""Developer check of the list-major IVF-PQ path: same answers as the LUT kernel (same quantizers), then timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200 import synth
from nafp_b200._lib import Context
from nafp_b200.eval.utils.get_index import Index, IVFPQ

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nq_time = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = Context.get(0)
t0 = time.time()
dummy = synth.synth_fp_db(n, seed=11)
db = synth.synth_fp_db(29500, 11, start_track=n // 59 + 2)
query = synth.synth_fp_queries(db, 12)
print("synth", round(time.time() - t0, 1), flush=True)

def make(path, params=None):
    os.environ["NAFP_IVFPQ_PATH"] = path
    g = Index(IVFPQ, 128, nlist=256, pq_m=64, pq_nbits=8, ctx=ctx)
    if params is None:
        g.train(dummy[:100000], seed=1234)
    else:
        g.set_ivfpq_params(*params)
    g.add(dummy); g.add(db)
    g.nprobe = 40
    return g

t0 = time.time(); a = make("lm"); print("build lm", round(time.time() - t0, 1), flush=True)
b = make("lut", a.ivfpq_params())
q = query[:512]
Da, Ia = a.search(q, 20)
print("lm search done", flush=True)
Db, Ib = b.search(q, 20)
same = (Ia == Ib)
print("ids identical: %.5f  max|dD| %.3g  stats lm %s" % (same.mean(), np.abs(Da - Db)[np.isfinite(Db)].max(), a.last_search_stats()), flush=True)
if not same.all():
    r, c = np.argwhere(~same)[0]
    print("first mismatch row", r, "col", c, Ia[r, c], Ib[r, c], Da[r, c], Db[r, c])
# timing: device-resident queries
qs = np.concatenate([query] * (nq_time // len(query) + 1))[:nq_time]
qd = ctx.malloc(qs.nbytes); ctx.h2d(qd, qs)
Dd = ctx.malloc(nq_time * 20 * 4); Id = ctx.malloc(nq_time * 20 * 8)
for g, name, reps in ((a, "lm", 5), (b, "lut", 1)):
    nqq = nq_time if name == "lm" else min(nq_time, 256)
    for _ in range(2): g.search_dev(qd.value, nqq, 20, Dd.value, Id.value)
    ctx.sync(); g.last_search_stats()
    ctx.timer_start()
    for _ in range(reps): g.search_dev(qd.value, nqq, 20, Dd.value, Id.value)
    ms = ctx.timer_stop() / reps
    print(f"{name}: N={g.ntotal} nq={nqq}: {ms:.2f} ms/search  {nqq/ms*1e3:.0f} query rows/s  stats {g.last_search_stats()}", flush=True)
