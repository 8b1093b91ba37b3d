"""Oracle: merge of per-shard top-k lists (K5).  New with sharding (the reference is single-node and has
no such step); defined as: per query, the k_out best of the union by (score desc, id asc).
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
from __future__ import annotations

import numpy as np


def merge_topk(scores: np.ndarray, ids: np.ndarray, counts, k_out: int):
    """scores [G,B,k] f64, ids [G,B,k] i64, counts [G,B] (valid prefix per list) or None
    -> scores [B,k_out] (-inf padded), ids [B,k_out] (-1 padded), count [B]."""
    G, B, k = scores.shape
    o_s = np.full((B, k_out), -np.inf)
    o_i = np.full((B, k_out), -1, dtype=np.int64)
    o_c = np.zeros((B,), dtype=np.int32)
    for b in range(B):
        s, i = [], []
        for g in range(G):
            n = k if counts is None else int(counts[g, b])
            s.append(scores[g, b, :n])
            i.append(ids[g, b, :n])
        s, i = np.concatenate(s), np.concatenate(i)
        order = np.lexsort((i, -s))[:k_out]
        o_s[b, :order.size], o_i[b, :order.size], o_c[b] = s[order], i[order], order.size
    return o_s, o_i, o_c
