"""Small invocation of every kernel family of libnafp for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
Sizes are tiny (the sanitizer serialises and instruments every access); correctness is checked elsewhere."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nafp_b200 import synth
from nafp_b200._lib import Context
from nafp_b200.eval.utils.get_index import IVF_FLAT, IVFPQ, Index
from nafp_b200.model import weights as W
from nafp_b200.model.fp import FingerPrinter, Melspec
from nafp_b200.model.utils import mini_search_subroutines as ms

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = Context.get(0)
if which in ("all", "extractor"):
    tr = synth.synth_track(3).astype(np.float32) / 32768.0
    x = np.stack([tr[i * 4000:i * 4000 + 8000] for i in range(7)])[:, None, :]
    mel = Melspec(ctx)(x, group_size=4)
    fp = FingerPrinter(ctx).load(W.init_weights(7, randomize_affine=True))
    emb = fp.fingerprint(x, group_size=4)
    emb2 = fp(mel)
    print("extractor", emb.shape, float(np.abs(emb - emb2).max()))
if which in ("all", "search"):
    dummy, db, query = synth.synth_search_set(12000, 295, seed=2)
    g = Index(0, 128, ctx=ctx)
    g.add(dummy)
    g.add(db)
    D, I = g.search(query[:40], 20)
    pred, _ = g.seq_match(query, np.array([0, 100, 250], np.int64), [1, 3, 19], 20)
    print("flat", I.shape, pred[:, :, 0].tolist())
if which in ("all", "ivf"):
    dummy, db, query = synth.synth_search_set(6000, 295, seed=4)
    for kind, nlist in ((IVFPQ, 32), (IVF_FLAT, 40)):
        g = Index(kind, 128, nlist=nlist, pq_m=64, pq_nbits=8, ctx=ctx)
        g.train(dummy[:3000], seed=5)
        g.add(dummy)
        g.add(db)
        g.nprobe = 8
        D, I = g.search(query[:24], 20)
        D2, I2 = g.search(query[:6], 40)            # k > 32: the LUT / list-scan kernels
        print("ivf", kind, I.shape, I2.shape)
if which in ("all", "mini"):
    rng = np.random.default_rng(0)
    dbm = rng.standard_normal((90, 128)).astype(np.float32)
    q = dbm[:60, None, :] + 0.5 * rng.standard_normal((60, 2, 128)).astype(np.float32)
    print("mini", ms.mini_search_eval(q, dbm, [1, 3, 19], display=False, ctx=ctx)[0][0])
ctx.sync()
print("sanitize_smoke done")
