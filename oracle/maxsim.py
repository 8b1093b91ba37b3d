"""Oracle: late-interaction MaxSim.  parity unpinned — the reference has no MaxSim (its "late
interaction" label is an HTTP cross-encoder, src/voice_agent/retrieval/reranker.py:287-354); the
interface restated is _rerank_batch_native(query, documents) -> one float per document, input order.

    score(q, c) = sum_{i < q_len} max_{j < d_len} <Qtok[q,i,:], Dtok[c,j,:]>   in fp64 on bf16-rounded inputs;
    empty document (d_len == 0) or empty query -> 0;  candidate id < 0 -> -inf.
"""
from __future__ import annotations

import numpy as np


def maxsim(Qtok: np.ndarray, Dtok: np.ndarray, cand: np.ndarray, q_len=None, d_len=None) -> np.ndarray:
    """Qtok [B,Tq,d], Dtok [n_docs,Td,d], cand [B,C] -> [B,C] float64."""
    B, Tq, _ = Qtok.shape
    n_docs, Td, _ = Dtok.shape
    C = cand.shape[1]
    out = np.zeros((B, C), dtype=np.float64)
    for b in range(B):
        ql = Tq if q_len is None else int(min(max(q_len[b], 0), Tq))
        q = Qtok[b, :ql].astype(np.float64)
        for c in range(C):
            doc = int(cand[b, c])
            if doc < 0 or doc >= n_docs:
                out[b, c] = -np.inf
                continue
            dl = Td if d_len is None else int(min(max(d_len[doc], 0), Td))
            if dl == 0 or ql == 0:
                out[b, c] = 0.0
                continue
            S = q @ Dtok[doc, :dl].astype(np.float64).T
            out[b, c] = S.max(axis=1).sum()
    return out
