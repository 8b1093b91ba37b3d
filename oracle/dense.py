"""Oracle: exact dense top-k (semantic channel).  parity unpinned — see oracle/__init__.py.

Restates rag2_semantic_search (database/migrations/20260114_rag2_schema.sql:377-410):
ORDER BY embedding <=> q ASC LIMIT n with similarity = 1 - cosine distance.  On L2-normalised
vectors (src/voice_agent/rag2/embedder.py:31-37) the order is by dot product descending; exact
scan, not HNSW.  Definition used for parity: score = fp64 dot of the bf16-rounded inputs, order
(score desc, id asc).
"""
from __future__ import annotations

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even), returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def dense_topk(Q: np.ndarray, X: np.ndarray, k: int, id_base: int = 0, block: int = 65536, tags=None, want=None):
    """Q [B,D], X [N,D] already bf16-representable (fp32 or fp64 storage).
    Returns ids [B,k] int64 (-1 padded), scores [B,k] float64.
    tags [N] / want [B]: the collection predicate of rag2_semantic_search (20260114_rag2_schema.sql:404-406) —
    query q only sees chunks with tags == want[q] (want < 0: all); slots past the eligible count are (-1, -inf)."""
    Q64 = np.asarray(Q, dtype=np.float64)
    B, N = Q64.shape[0], X.shape[0]
    kk = min(k, N)
    best_s = np.full((B, 0), 0.0)
    best_i = np.zeros((B, 0), dtype=np.int64)
    for s in range(0, N, block):
        Xb = np.asarray(X[s:s + block], dtype=np.float64)
        S = Q64 @ Xb.T
        if want is not None:
            w = np.asarray(want)[:, None]
            S = np.where((w < 0) | (np.asarray(tags[s:s + block])[None, :] == w), S, -np.inf)
        ids = np.broadcast_to(np.arange(s, s + Xb.shape[0], dtype=np.int64), S.shape)
        best_s = np.concatenate([best_s, S], axis=1)
        best_i = np.concatenate([best_i, ids], axis=1)
        if best_s.shape[1] > kk:
            order = np.lexsort((best_i, -best_s), axis=1)[:, :kk]
            best_s = np.take_along_axis(best_s, order, 1)
            best_i = np.take_along_axis(best_i, order, 1)
    order = np.lexsort((best_i, -best_s), axis=1)[:, :kk]
    best_s = np.take_along_axis(best_s, order, 1)
    best_i = np.take_along_axis(best_i, order, 1)
    out_i = np.full((B, k), -1, dtype=np.int64)
    out_s = np.full((B, k), -np.inf)
    out_i[:, :kk] = np.where(np.isneginf(best_s), -1, best_i + id_base)
    out_s[:, :kk] = best_s
    return out_i, out_s
