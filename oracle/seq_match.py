"""Oracle: sequence-offset matcher and hit metrics.  TEST INFRASTRUCTURE ONLY.

Literal restatement of the hot loop of ``eval/eval_faiss.py:204-243`` (reference) around any
index object exposing ``search(q, k) -> (D, I)``:

* ``:208``      q = query[test_id : test_id + sl]          (numpy clamps at the array end)
* ``:211``      _, I = index.search(q, k_probe)
* ``:215-216``  I[offset, :] -= offset
* ``:219``      candidates = np.unique(I[I >= 0])
* ``:222-229``  score[c] = mean(diag(q . recon[c : c+sl].T))
* ``:232``      pred_ids = candidates[argsort(-score)[:10]]
* ``:236-243``  top1_exact / top1_near / top3 / top10 against gt = test_id + n_dummy (``:189``)

``recon`` is the reference's ``fake_recon_index`` (``:167-171``) = [dummy_db; db] rows.
PINNED: reproduces ``raw_score.npy`` written by the reference's own loop on a 61,180-row fixture
(``tests/golden/ref_eval_flat.npz``, ``tests/test_reference_golden.py``).
The only deviation: ``argsort`` is made stable so that ties resolve to the lower candidate id
(the reference's quicksort order on exact ties is unspecified).
"""
from __future__ import annotations

import numpy as np


def seq_scores(q, recon, candidates, sl):
    """``eval_faiss.py:222-229``: mean of the main diagonal of q . recon[c:c+sl].T (fp32 dot)."""
    scores = np.zeros(len(candidates))
    for ci, cid in enumerate(candidates):
        scores[ci] = np.mean(np.diag(np.dot(q, recon[cid:cid + sl, :].T)))
    return scores


def seq_scores_fast(q, recon, candidates, sl):
    """The same scores in one gather + one einsum (fp32 products, fp64 mean): the form used where the CPU path is
    TIMED or run at scale; ``seq_scores`` stays the literal restatement (tests/test_oracle_search.py compares them)."""
    if len(candidates) == 0:
        return np.zeros(0)
    n = len(recon)
    rows = candidates[:, None] + np.arange(len(q))[None, :]                    # (n_cand, sl')
    valid = rows < n                                                           # numpy slicing clamps at the end (:223-229)
    g = recon[np.minimum(rows, n - 1)]                                         # (n_cand, sl', d)
    dots = np.einsum("csd,sd->cs", g, q, dtype=np.float32).astype(np.float64) * valid
    return dots.sum(1) / np.maximum(valid.sum(1), 1)


def match_one(index, query, recon, test_id, sl, k_probe=20, n_pred=10, fast_scores=False, table=None):
    q = np.asarray(query[test_id:test_id + sl, :])
    if table is None:
        _, I = index.search(q, k_probe)
    else:                       # search results of every query row, computed in one batch (``evaluate(batch_search=True)``)
        I = table[test_id:test_id + len(q)]
    I = np.array(I, dtype=np.int64, copy=True)
    for offset in range(len(I)):
        I[offset, :] -= offset
    candidates = np.unique(I[np.where(I >= 0)])
    scores = (seq_scores_fast if fast_scores else seq_scores)(q, recon, candidates, sl)
    order = np.argsort(-scores, kind="stable")[:n_pred]
    return candidates[order], scores[order]


def hit_flags(pred_ids, gt_id):
    """(top1_exact, top1_near, top3_exact, top10_exact) -- ``eval_faiss.py:236-243``."""
    if len(pred_ids) == 0:
        return 0, 0, 0, 0
    return (int(gt_id == pred_ids[0]),
            int(pred_ids[0] in [gt_id - 1, gt_id, gt_id + 1]),
            int(gt_id in pred_ids[:3]),
            int(gt_id in pred_ids[:10]))


def evaluate(index, query, recon, n_dummy, test_ids, test_seq_len, k_probe=20, fast_scores=False, batch_search=False):
    """Returns (raw_score (n_test, 4*n_len) int, pred (n_test, n_len, 10) int64 padded with -1).

    raw_score column blocks = [top1_exact | top1_near | top3_exact | top10_exact], the layout of
    ``raw_score.npy`` (``eval_faiss.py:271-273``).  ``fast_scores`` / ``batch_search`` are the scalable forms of the same
    computation (vectorised candidate scoring; one search call for all query rows the test ids touch -- a row's
    neighbours do not depend on which sequence asks for them), used by the timed / large CPU legs."""
    test_ids = np.asarray(test_ids, dtype=np.int64)
    table = None
    if batch_search:
        need = np.zeros(len(query), bool)
        for t in test_ids:
            need[t:t + max(test_seq_len)] = True
        rows = np.nonzero(need)[0]
        _, I_rows = index.search(np.asarray(query[rows]), k_probe)
        table = np.full((len(query), k_probe), -1, np.int64)
        table[rows] = I_rows
    n_test, n_len = len(test_ids), len(test_seq_len)
    hits = np.zeros((4, n_test, n_len), dtype=int)
    pred = np.full((n_test, n_len, 10), -1, dtype=np.int64)
    gt_ids = test_ids + n_dummy
    for ti, test_id in enumerate(test_ids):
        for si, sl in enumerate(test_seq_len):
            p, _ = match_one(index, query, recon, int(test_id), int(sl), k_probe, fast_scores=fast_scores, table=table)
            pred[ti, si, :len(p)] = p
            hits[:, ti, si] = hit_flags(p, gt_ids[ti])
    raw = np.concatenate([hits[0], hits[1], hits[2], hits[3]], axis=1)
    return raw, pred


def hit_rates(raw, n_len):
    """100 * mean per column block -> (4, n_len): rows Top1 exact / Top1 near / Top3 / Top10."""
    return 100.0 * raw.reshape(raw.shape[0], 4, n_len).mean(axis=0)
