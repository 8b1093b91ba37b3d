"""Oracle-side restatement of BM25 top-k WITH MaxScore-style pruning — the algorithm DESIGN.md §8 names for round 2,
stated on the CPU first so that its exactness can be pinned against oracle/bm25.py before any kernel exists.
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Same definition as oracle/bm25.py (fp32 products, fp32 sum in query order, score > 0 eligible, order (score desc,
id asc)); what changes is the work.  Docs are visited in blocks of ascending id with a running threshold tau = the
k-th best score so far.  Per block:
  1. terms sorted by their upper bound ub_t = max posting contribution; the longest prefix whose bounds add up to
     at most tau is NON-ESSENTIAL: a doc holding only such terms cannot beat tau, so those lists are not streamed;
  2. the essential lists of the block are streamed into partial sums;
  3. a touched doc survives if partial + (bounds of the non-essential terms) can still beat tau;
  4. survivors are re-scored exactly — every term looked up by binary search, fp32 adds in query order;
  5. the top-k and tau are updated.
Docs arrive in ascending id, so a later doc that merely ties with the k-th best loses the tie: strict '>' is exact.
All bounds carry a relative slack that covers the fp32 rounding of the in-order sum (n terms: (1 + 2^-24)^(n-1)).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

from .bm25 import CsrIndex

_SLACK = 1.0 + 1e-5   # >> (1 + 2^-24)^32: an upper bound in real arithmetic also bounds the fp32 in-order sum


def bm25_topk_pruned(index: CsrIndex, queries: Sequence[Sequence[int]], k: int, id_base: int = 0, block: int = 32768,
                     stats: Dict[str, int] | None = None, ne_budget: float = 1.0
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """ne_budget in (0, 1]: the non-essential bounds may add up to at most ne_budget * tau.  1.0 skips the most
    postings; a smaller budget streams more lists but leaves far fewer survivors to look up (exact either way)."""
    B = len(queries)
    out_i = np.full((B, k), -1, dtype=np.int64)
    out_s = np.zeros((B, k), dtype=np.float32)
    out_c = np.zeros((B,), dtype=np.int32)
    V = index.indptr.shape[0] - 1
    st = stats if stats is not None else {}
    for key in ("postings", "streamed", "lookups", "survivors", "touched"):
        st.setdefault(key, 0)
    for qi, raw_terms in enumerate(queries):
        terms = [int(t) for t in raw_terms if 0 <= int(t) < V and index.indptr[int(t) + 1] > index.indptr[int(t)]]
        if not terms:
            continue
        docs = [index.doc[index.indptr[t]:index.indptr[t + 1]] for t in terms]
        con = [(index.idf[t] * index.imp[index.indptr[t]:index.indptr[t + 1]]).astype(np.float32) for t in terms]
        ub = np.array([float(c.max()) for c in con], dtype=np.float64)
        st["postings"] += sum(d.size for d in docs)
        by_ub = np.argsort(ub, kind="stable")
        best_s = np.zeros((0,), dtype=np.float32)    # current top-k, sorted by (score desc, id asc)
        best_i = np.zeros((0,), dtype=np.int64)
        tau = 0.0
        for b0 in range(0, index.n_docs, block):
            b1 = min(index.n_docs, b0 + block)
            seg = [(int(np.searchsorted(d, b0)), int(np.searchsorted(d, b1))) for d in docs]
            # 1. non-essential prefix under the current tau
            acc, n_ne = 0.0, 0
            for j in by_ub:
                if (acc + ub[j]) * _SLACK <= tau * ne_budget:
                    acc += ub[j]
                    n_ne += 1
                else:
                    break
            ne = set(int(j) for j in by_ub[:n_ne])
            ub_ne = acc
            # 2. stream the essential lists of the block
            part = np.zeros(b1 - b0, dtype=np.float64)
            hit = np.zeros(b1 - b0, dtype=bool)
            for j in range(len(terms)):
                if j in ne:
                    continue
                lo, hi = seg[j]
                st["streamed"] += hi - lo
                d = docs[j][lo:hi] - b0
                part[d] += con[j][lo:hi].astype(np.float64)
                hit[d] = True
            touched = np.nonzero(hit)[0]
            st["touched"] += int(touched.size)
            # 3. survivors: can still beat tau
            surv = touched[(part[touched] + ub_ne) * _SLACK > tau]
            st["survivors"] += int(surv.size)
            if surv.size == 0:
                continue
            # 4. exact scores: every term looked up, fp32 adds in QUERY order
            sc = np.zeros(surv.size, dtype=np.float32)
            gd = surv + b0
            for j in range(len(terms)):
                lo, hi = seg[j]
                d = docs[j][lo:hi]
                pos = np.searchsorted(d, gd)
                st["lookups"] += int(surv.size)
                ok = (pos < d.size) & (d[np.minimum(pos, d.size - 1)] == gd) if d.size else np.zeros(surv.size, bool)
                add = np.zeros(surv.size, dtype=np.float32)
                add[ok] = con[j][lo:hi][pos[ok]]
                sc = (sc + add).astype(np.float32)          # adding 0 leaves an fp32 value unchanged
            keep = sc > np.float32(tau) if best_s.size >= k else sc > 0
            if not keep.any():
                continue
            # 5. merge into the top-k (ids ascending inside the block, and every id above the earlier blocks')
            all_s = np.concatenate([best_s, sc[keep]])
            all_i = np.concatenate([best_i, gd[keep]])
            order = np.lexsort((all_i, -all_s.astype(np.float64)))[:k]
            best_s, best_i = all_s[order], all_i[order]
            if best_s.size >= k:
                tau = float(best_s[k - 1])
        n = best_s.size
        out_i[qi, :n] = best_i + id_base
        out_s[qi, :n] = best_s
        out_c[qi] = n
    return out_i, out_s, out_c
