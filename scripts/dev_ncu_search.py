import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nafp_b200 import synth
from nafp_b200._lib import Context
from nafp_b200.eval.utils.get_index import Index
n = int(sys.argv[1]); nq = int(sys.argv[2]); reps = int(sys.argv[3])
ctx = Context.get(0)
dummy = synth.synth_fp_db(n, seed=11); db = synth.synth_fp_db(29500, 11, start_track=n // 59 + 2); query = synth.synth_fp_queries(db, 12)
idx = Index(0, 128); idx.add(dummy); idx.add(db)
q = np.concatenate([query[i * 59: i * 59 + 19] for i in range(60)])[:nq]
qd = ctx.malloc(q.nbytes); ctx.h2d(qd, q)
Dd = ctx.malloc(nq * 20 * 4); Id = ctx.malloc(nq * 20 * 8)
for _ in range(reps): idx.search_dev(qd.value, nq, 20, Dd.value, Id.value)
ctx.sync()
