"""In-training mini search on B200 -- mirror of the reference's model/utils/mini_search_subroutines.py
(same function names, arguments and return values; numpy arrays in and out instead of tf tensors).

    pairwise_distances_for_eval(emb_que, emb_db, return_dotprod, squared)   reference :29-90
    conv_eye_func(x, s)                                                     reference :93-120
    mini_search_eval(query, db, scopes, mode, display, gt_id_offset)        reference :123-236

The arithmetic runs in libnafp.so (csrc/mini_search.cu); there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from ..._lib import Context, check, lib, ptr


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def pairwise_distances_for_eval(emb_que, emb_db, return_dotprod=False, squared=True, ctx=None):
    """(nQ, nAug, d) x (nD, d) -> (nAug, nQ, nD, 1) float32 pairwise squared-L2 distances (or dot products)."""
    ctx = ctx or Context.get(0)
    q, db = _f32(emb_que), _f32(emb_db)
    if q.ndim != 3 or db.ndim != 2 or q.shape[2] != db.shape[1]:
        raise ValueError(f"expected (nQ, nAug, d) and (nD, d), got {q.shape} and {db.shape}")
    n_q, n_aug, d = q.shape
    out = np.empty((n_aug, n_q, db.shape[0]), np.float32)
    check(lib.nafp_pairwise_dists_host(ctx.h, ptr(q), ptr(db), n_q, n_aug, db.shape[0], d, int(bool(return_dotprod)),
                                       int(bool(squared)), ptr(out)))
    return out[..., None]


def conv_eye_func(x, s, ctx=None):
    """(nAug, nQ, nD, 1) -> (nAug, nQ-s+1, nD-s+1, 1): sums over s consecutive diagonal elements."""
    ctx = ctx or Context.get(0)
    x = _f32(x)
    if x.ndim != 4 or x.shape[3] != 1:
        raise ValueError(f"expected (nAug, nQ, nD, 1), got {x.shape}")
    n_aug, n_q, n_d = x.shape[:3]
    s = int(s)
    out = np.empty((n_aug, n_q - s + 1, n_d - s + 1), np.float32)
    check(lib.nafp_conv_eye_host(ctx.h, ptr(x), n_aug, n_q, n_d, s, ptr(out)))
    return out[..., None]


def mini_search_eval(query, db, scopes=[1, 3, 5, 9, 11, 19], mode='argmin', display=True, gt_id_offset=0, ctx=None):
    """Returns ((top1_acc, top3_acc, top10_acc), mean_rank), arrays over `scopes`, accuracies in percent."""
    ctx = ctx or Context.get(0)
    if mode == 'argmin':
        argmax = 0
    elif mode.lower() == 'argmax':
        argmax = 1
    else:
        raise NotImplementedError(mode)
    q, d_ = _f32(query), _f32(db)
    if q.ndim != 3 or d_.ndim != 2 or q.shape[2] != d_.shape[1]:
        raise ValueError(f"expected (nQ, nAug, d) and (nD, d), got {q.shape} and {d_.shape}")
    sc = np.ascontiguousarray(scopes, dtype=np.int32)
    n = len(sc)
    top1, top3, top10, mean_rank = (np.zeros(n, np.float64) for _ in range(4))
    check(lib.nafp_mini_search_host(ctx.h, ptr(q), ptr(d_), q.shape[0], q.shape[1], d_.shape[0], q.shape[2], ptr(sc), n,
                                    argmax, int(gt_id_offset), ptr(top1), ptr(top3), ptr(top10), ptr(mean_rank)))
    if display:
        color_cyan = '\033[36m'
        color_def = '\033[0m'
        line_int = '{:^6}\t' * len(scopes)
        line_float = '{:>4.2f}\t' * len(scopes)
        print(color_cyan + 'Scope:\t', line_int.format(*scopes), color_def)
        print(color_cyan + 'T1acc:\t' + color_def, line_float.format(*top1))
        print(color_cyan + 'mRank:\t' + color_def, line_float.format(*mean_rank))
    return (top1, top3, top10), mean_rank
