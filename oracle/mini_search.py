"""CPU restatement of the reference's in-training mini search (TEST INFRASTRUCTURE, never on the product path).

Follows /root/reference/model/utils/mini_search_subroutines.py:
    pairwise_distances_for_eval  :29-90    (tf.matmul / reduce_sum / maximum -> numpy, same operation order)
    conv_eye_func                :93-120   (Conv2D with np.eye(s) kernel, 'valid' -> explicit diagonal sums)
    mini_search_eval             :123-236  (np.argsort + the reference's own counting loops, kept verbatim in structure)
TensorFlow is not installable here, so the two tf functions are restated; mini_search_eval's numpy part is the
reference's algorithm step for step (argsort, np.where rank lookup, top-k membership).
"""
from __future__ import annotations

import numpy as np


def pairwise_distances_for_eval(emb_que, emb_db, return_dotprod=False, squared=True, dtype=np.float32):
    q = np.asarray(emb_que, dtype)
    db = np.asarray(emb_db, dtype)
    dot = np.matmul(q, db.T)                     # (nQ, nAug, nD)          :65-66
    dot = np.transpose(dot, (1, 0, 2))           # (nAug, nQ, nD)          :67-68
    if return_dotprod:
        return dot[..., None]
    que_sq = np.sum(np.square(q), axis=2).T      # (nAug, nQ)              :75-76
    db_sq = np.sum(np.square(db), axis=1).reshape(1, -1)
    dists = que_sq[:, :, None] + db_sq[:, None, :] - dtype(2.0) * dot       # :80-81
    dists = np.maximum(dists, 0.0)[..., None]
    if not squared:
        mask = (dists == 0.0).astype(dtype)
        dists = np.sqrt(dists + mask * dtype(1e-16)) * (1.0 - mask)
    return dists


def conv_eye_func(x, s):
    x = np.asarray(x)
    n_a, n_q, n_d, _ = x.shape
    out = np.zeros((n_a, n_q - s + 1, n_d - s + 1), x.dtype)
    for t in range(s):
        out += x[:, t:t + n_q - s + 1, t:t + n_d - s + 1, 0]
    return out[..., None]


def mini_search_eval(query, db, scopes=(1, 3, 5, 9, 11, 19), mode='argmin', gt_id_offset=0, dtype=np.float32):
    n_augs = query.shape[1]
    n_scopes = len(scopes)
    if mode == 'argmin':
        all_dists = pairwise_distances_for_eval(query, db, squared=True, dtype=dtype)
    elif mode.lower() == 'argmax':
        all_dists = pairwise_distances_for_eval(query, db, return_dotprod=True, dtype=dtype)
    else:
        raise NotImplementedError(mode)
    mean_rank = np.zeros(n_scopes)
    top1_acc, top3_acc, top10_acc = np.zeros(n_scopes), np.zeros(n_scopes), np.zeros(n_scopes)
    for i, s in enumerate(scopes):
        conv_dists = np.squeeze(conv_eye_func(all_dists, s), 3)
        srt = np.argsort(conv_dists, axis=2, kind='stable')
        if mode.lower() == 'argmax':
            srt = srt[:, :, ::-1]
        n_targets = conv_dists.shape[1]
        _sum_rank = 0
        for target_id in range(n_targets):
            gt_id = target_id + gt_id_offset
            _, _rank = np.where(srt[:, target_id, :] == gt_id)
            _sum_rank += np.sum(_rank) / n_augs
        mean_rank[i] = _sum_rank / n_targets
        c1 = c3 = c10 = 0
        for target_id in range(n_targets):
            gt_id = target_id + gt_id_offset
            c1 += np.sum(srt[:, target_id, 0] == gt_id) / n_augs
            c3 += np.sum(srt[:, target_id, :3] == gt_id) / n_augs
            c10 += np.sum(srt[:, target_id, :10] == gt_id) / n_augs
        top1_acc[i] = c1 / n_targets
        top3_acc[i] = c3 / n_targets
        top10_acc[i] = c10 / n_targets
    return (top1_acc * 100., top3_acc * 100., top10_acc * 100.), mean_rank
