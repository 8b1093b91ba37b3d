"""Oracle: BM25 top-k over an inverted index (lexical channel).  parity unpinned — see __init__.py.

Interface restated from rag2_lexical_search (database/migrations/20260114_rag2_schema.sql:341-374,
called at src/voice_agent/rag2/retrieval.py:273-292): keywords in, top-`limit` rows by descending
score out.  The reference ranks with Postgres ts_rank_cd, which is not in the tree; BASELINE.json's
north_star asks for BM25, defined here:

    idf[t]      = fp32( ln(1 + (N - df_t + 0.5) / (df_t + 0.5)) )                       (fp64 math, then rounded)
    impact(t,d) = fp32( tf*(k1+1) / (tf + k1*(1 - b + b*len_d/avgdl)) )                 (fp64 math, then rounded)
    score(d)    = fp32 sum over the query's terms IN ORDER of fp32(idf[t] * impact(t,d)) (fp32 mul, fp32 add)
    eligible: score > 0;  order: (score desc, doc id asc);  k1 = 1.2, b = 0.75 by default.

The summation order is part of the definition, which makes the ranking bit-reproducible.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def idf_table(df: np.ndarray, n_docs: int) -> np.ndarray:
    df = df.astype(np.float64)
    return np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)


def impacts(tf: np.ndarray, doc_len: np.ndarray, avgdl: float, k1: float = 1.2, b: float = 0.75) -> np.ndarray:
    tf = tf.astype(np.float64)
    norm = k1 * (1.0 - b + b * doc_len.astype(np.float64) / float(avgdl))
    return (tf * (k1 + 1.0) / (tf + norm)).astype(np.float32)


class CsrIndex:
    """Term-major CSR: postings of term t are doc[indptr[t]:indptr[t+1]] (ascending) with impact imp[...]."""

    def __init__(self, indptr: np.ndarray, doc: np.ndarray, imp: np.ndarray, idf: np.ndarray, n_docs: int):
        self.indptr, self.doc, self.imp, self.idf, self.n_docs = indptr, doc, imp, idf, n_docs

    @staticmethod
    def from_coo(doc: np.ndarray, term: np.ndarray, tf: np.ndarray, doc_len: np.ndarray, V: int,
                 k1: float = 1.2, b: float = 0.75, avgdl: float | None = None, idf: np.ndarray | None = None,
                 n_docs_global: int | None = None) -> "CsrIndex":
        n_docs = int(doc_len.shape[0])
        order = np.lexsort((doc, term))
        doc, term, tf = doc[order], term[order], tf[order]
        counts = np.bincount(term, minlength=V)
        indptr = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        if avgdl is None:
            avgdl = float(doc_len.astype(np.float64).mean())
        if idf is None:
            idf = idf_table(counts, n_docs_global or n_docs)
        imp = impacts(tf, doc_len[doc], avgdl, k1, b)
        return CsrIndex(indptr, doc.astype(np.int64), imp, idf.astype(np.float32), n_docs)


def bm25_topk(index: CsrIndex, queries: Sequence[Sequence[int]], k: int, id_base: int = 0, tags=None, want=None,
              require_all: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns ids [B,k] int64 (-1 padded), scores [B,k] float32, count [B] int32.
    tags [n_docs] / want [B]: the collection predicate of rag2_lexical_search (20260114_rag2_schema.sql:368-370) —
    query q only sees docs with tags == want[q] (want < 0: all).
    require_all: the `tsv @@ plainto_tsquery(...)` predicate (:369) — only docs that contain EVERY distinct term of
    the query; a repeated term counts and scores once; an unknown term (outside [0,V)) means no doc matches."""
    B = len(queries)
    out_i = np.full((B, k), -1, dtype=np.int64)
    out_s = np.zeros((B, k), dtype=np.float32)
    out_c = np.zeros((B,), dtype=np.int32)
    V = index.indptr.shape[0] - 1
    for qi, terms in enumerate(queries):
        acc = np.zeros(index.n_docs, dtype=np.float32)
        hits = np.zeros(index.n_docs, dtype=np.int32)
        need, dead, seen = 0, False, set()
        for t in terms:
            t = int(t)
            if t < 0 or t >= V:
                dead = True
                continue
            if require_all:
                if t in seen:
                    continue
                seen.add(t)
            need += 1
            lo, hi = index.indptr[t], index.indptr[t + 1]
            d = index.doc[lo:hi]
            contrib = (index.idf[t] * index.imp[lo:hi]).astype(np.float32)  # fp32 multiply
            acc[d] = acc[d] + contrib                                        # fp32 add; docs unique per list
            hits[d] += 1
        ok = acc > 0
        if require_all:
            ok &= (hits == need) & (need > 0) & (not dead)
        if want is not None and int(want[qi]) >= 0:
            ok &= np.asarray(tags) == int(want[qi])
        hit = np.nonzero(ok)[0]
        if hit.size == 0:
            continue
        order = np.lexsort((hit, -acc[hit].astype(np.float64)))[:k]
        sel = hit[order]
        n = sel.size
        out_i[qi, :n] = sel + id_base
        out_s[qi, :n] = acc[sel]
        out_c[qi] = n
    return out_i, out_s, out_c


def bm25_scores_fp64(index: CsrIndex, terms: Sequence[int], tf_csr=None) -> np.ndarray:
    """fp64 accumulation of the same fp32 factors — used to show the 1e-3 tolerance north_star states."""
    acc = np.zeros(index.n_docs, dtype=np.float64)
    V = index.indptr.shape[0] - 1
    for t in terms:
        t = int(t)
        if t < 0 or t >= V:
            continue
        lo, hi = index.indptr[t], index.indptr[t + 1]
        acc[index.doc[lo:hi]] += index.idf[t].astype(np.float64) * index.imp[lo:hi].astype(np.float64)
    return acc
