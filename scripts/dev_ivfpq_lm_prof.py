"""One list-major IVF-PQ search for profilers: build (train on 100k rows), 2 searches of nq rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200 import synth
from nafp_b200._lib import Context
from nafp_b200.eval.utils.get_index import Index, IVFPQ
n = int(sys.argv[1]); nq = int(sys.argv[2])
ctx = Context.get(0)
dummy = synth.synth_fp_db(n, seed=11)
db = synth.synth_fp_db(29500, 11, start_track=n // 59 + 2)
query = synth.synth_fp_queries(db, 12)
g = Index(IVFPQ, 128, nlist=256, pq_m=64, pq_nbits=8, ctx=ctx)
g.train(dummy[:100000], seed=1234); g.add(dummy); g.add(db); g.nprobe = 40
qs = np.concatenate([query] * (nq // len(query) + 1))[:nq]
qd = ctx.malloc(qs.nbytes); ctx.h2d(qd, qs)
Dd = ctx.malloc(nq * 20 * 4); Id = ctx.malloc(nq * 20 * 8)
for _ in range(2): g.search_dev(qd.value, nq, 20, Dd.value, Id.value)
ctx.sync()
print(g.last_search_stats())
