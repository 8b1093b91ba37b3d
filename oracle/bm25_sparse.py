"""Second, independent CPU restatement of the lexical channel: BM25 as a sparse matrix product (scipy.sparse),
accumulated in fp64.  TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).

BASELINE.md §3 names scipy.sparse as the CPU statement of BM25.  oracle/bm25.py walks posting lists term by term in
fp32 (the definition the kernel reproduces bit for bit); this file shares none of that code path: the corpus is a
CSR matrix W [V, N] of impacts, a batch of queries is a CSR matrix Qs [B, V] of idf weights, and the scores are
Qs @ W — a different algorithm and a different summation order, so a misreading shared by the kernel and
oracle/bm25.py would show up here (within the 1e-3 relative tolerance north_star states for fp32 against fp64).
parity unpinned: the reference ranks with Postgres ts_rank_cd, which is not in its tree.

    idf[t] = ln(1 + (N - df_t + 0.5) / (df_t + 0.5)),  impact(t, d) = tf (k1 + 1) / (tf + k1 (1 - b + b len_d / avgdl))
    (database interface: rag2_lexical_search, database/migrations/20260114_rag2_schema.sql:341-374)
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import scipy.sparse as sp


class SparseBM25:
    def __init__(self, doc: np.ndarray, term: np.ndarray, tf: np.ndarray, doc_len: np.ndarray, V: int,
                 k1: float = 1.2, b: float = 0.75, avgdl: float | None = None, n_docs_global: int | None = None,
                 df_global: np.ndarray | None = None):
        n = int(doc_len.shape[0])
        self.n_docs, self.V = n, V
        dl = doc_len.astype(np.float64)
        if avgdl is None:
            avgdl = float(dl.mean())
        tf64 = tf.astype(np.float64)
        # impacts are DEFINED as fp32 values (the index stores them so); the product and sum run in fp64
        imp = (tf64 * (k1 + 1.0) / (tf64 + k1 * (1.0 - b + b * dl[doc] / avgdl))).astype(np.float32).astype(np.float64)
        self.W = sp.csr_matrix((imp, (term.astype(np.int64), doc.astype(np.int64))), shape=(V, n))
        df = np.bincount(term, minlength=V).astype(np.float64) if df_global is None else df_global.astype(np.float64)
        N = float(n_docs_global or n)
        self.idf = np.log(1.0 + (N - df + 0.5) / (df + 0.5)).astype(np.float32).astype(np.float64)

    def scores(self, queries: Sequence[Sequence[int]]) -> np.ndarray:
        """Dense [B, N] fp64 score rows (OR semantics; a repeated term counts each time, like oracle/bm25.py)."""
        rows, cols, vals = [], [], []
        for q, terms in enumerate(queries):
            for t in terms:
                t = int(t)
                if 0 <= t < self.V:
                    rows.append(q); cols.append(t); vals.append(self.idf[t])
        Qs = sp.csr_matrix((vals, (rows, cols)), shape=(len(queries), self.V))   # duplicates are summed
        return np.asarray((Qs @ self.W).todense())

    def topk(self, queries: Sequence[Sequence[int]], k: int, batch: int = 32) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """ids [B,k] (-1 padded), fp64 scores [B,k], counts [B]; order (score desc, id asc); score > 0 only."""
        B = len(queries)
        out_i = np.full((B, k), -1, dtype=np.int64)
        out_s = np.zeros((B, k), dtype=np.float64)
        out_c = np.zeros(B, dtype=np.int32)
        for s in range(0, B, batch):
            S = self.scores(queries[s:s + batch])
            for j in range(S.shape[0]):
                row = S[j]
                kk = min(k, int((row > 0).sum()))
                if kk == 0:
                    continue
                cand = np.argpartition(-row, min(kk + 64, row.size - 1))[: kk + 64] if row.size > kk + 64 else np.arange(row.size)
                cand = cand[row[cand] > 0]
                order = np.lexsort((cand, -row[cand]))[:kk]
                sel = cand[order]
                out_i[s + j, :kk] = sel
                out_s[s + j, :kk] = row[sel]
                out_c[s + j] = kk
        return out_i, out_s, out_c
