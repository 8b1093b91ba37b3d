"""Oracle: weighted RRF fusion, safety threshold, conformal denoise (numpy / plain Python floats).

Restates, in its own words, the arithmetic of
  A  RAG2Retriever._retrieve_candidates merge + _fuse_rrf   src/voice_agent/rag2/retrieval.py:203-271, :358-376
     RAG2Retriever._apply_safety                            src/voice_agent/rag2/retrieval.py:461-495
  B  RRFFusion.fuse (+ _compute_rrf_scores, _apply_safety_threshold, _apply_conformal_denoising)
                                                            triple-hybrid-rag/src/triple_hybrid_rag/core/fusion.py:52-247
  C  HybridSearcher._rrf_fusion                             src/voice_agent/retrieval/hybrid_search.py:460-501
on plain integer ids.  Python floats are IEEE fp64 and every operation below is a single correctly
rounded operation in the reference's order, so agreement with the reference is bit-for-bit
(pinned by tests/test_oracle_fusion.py against tests/golden/fusion_golden.json.gz.gz).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

RAG2, LIB, RAG1 = 0, 1, 2
TIE_INSERTION, TIE_CHUNK_ID = 0, 1


def percentile_linear(values: Sequence[float], q: float) -> float:
    """numpy.percentile(values, q), method='linear' (numpy 2.x _function_base_impl.py:
    _QuantileMethods['linear'] -> (n-1)*quantile, _get_indexes, _get_gamma, _lerp)."""
    a = sorted(values)
    n = len(a)
    quant = q / 100.0
    vi = (n - 1) * quant
    if vi >= n - 1:
        lo = hi = n - 1
        prev = -1.0
    elif vi < 0:
        lo = hi = 0
        prev = 0.0
    else:
        lo = int(math.floor(vi))
        hi = lo + 1
        prev = float(lo)
    t = vi - prev
    lo_v, hi_v = a[lo], a[hi]
    diff = hi_v - lo_v
    r = lo_v + diff * t
    if t >= 0.5:
        r = hi_v - diff * (1 - t)
    return r


def fuse(variant: int,
         lists: Sequence[Optional[Sequence[int]]],
         weights: Sequence[float] = (0.7, 0.8, 1.0),
         rrf_k: int = 60,
         raw: Sequence[Optional[Sequence[float]]] = (None, None, None),
         safety_thr: float = 0.0, alpha: float = 0.0, denoise: bool = False, top_k: int = 0,
         tie_mode: int = TIE_INSERTION) -> List[Dict]:
    """Fuse three ranked id lists (lexical, semantic, graph order).  Returns rows
    {id, rrf, ranks:(l,s,g) 0=absent, raw:(l,s,g)} in final order."""
    first_seen: Dict[int, int] = {}
    last_rank: Dict[int, List[int]] = {}
    occ: Dict[int, List[int]] = {}
    merged_raw: Dict[int, List[float]] = {}
    rrf1: Dict[int, float] = {}
    e = 0
    for c in range(3):
        ids = lists[c]
        if ids is None:
            continue
        for pos, cid in enumerate(ids):
            cid = int(cid)
            if cid not in first_seen:
                first_seen[cid] = e
                last_rank[cid] = [0, 0, 0]
                occ[cid] = [0, 0, 0]
                merged_raw[cid] = [0.0, 0.0, 0.0]
                rrf1[cid] = 0.0
            last_rank[cid][c] = pos + 1
            occ[cid][c] += 1
            if raw[c] is not None:
                s = float(raw[c][pos])
                if variant == LIB:
                    merged_raw[cid][c] = s
                elif variant == RAG1:
                    merged_raw[cid][c] = max(merged_raw[cid][c], s)
            if variant == RAG1:
                rrf1[cid] = rrf1[cid] + 1.0 / (rrf_k + pos + 1)
            e += 1
    rows = []
    for cid, ins in first_seen.items():
        if variant == RAG2:
            score = 0.0
            for c in range(3):
                if last_rank[cid][c]:
                    score = score + weights[c] / (rrf_k + last_rank[cid][c])
        elif variant == LIB:
            score = 0.0
            for c in range(3):
                if occ[cid][c]:
                    s = weights[c] * (1.0 / (rrf_k + last_rank[cid][c]))
                    for _ in range(occ[cid][c]):
                        score = score + s
        else:
            score = rrf1[cid]
        rows.append({"id": cid, "rrf": score, "ranks": tuple(last_rank[cid]),
                     "raw": tuple(merged_raw[cid]), "_ins": ins})
    if tie_mode == TIE_CHUNK_ID:
        rows.sort(key=lambda r: (-r["rrf"], r["id"]))
    else:
        rows.sort(key=lambda r: (-r["rrf"], r["_ins"]))
    if variant == LIB:
        if safety_thr > 0:
            rows = [r for r in rows if max(r["raw"][1], r["raw"][0], r["raw"][2]) >= safety_thr]
        if denoise and len(rows) >= 3:
            thr = percentile_linear([r["rrf"] for r in rows], (1 - alpha) * 100)
            rows = [r for r in rows if r["rrf"] >= thr]
    if top_k and top_k > 0:
        rows = rows[:top_k]
    for r in rows:
        r.pop("_ins")
    return rows


def apply_safety(rrf: Sequence[float], rerank: Sequence[Optional[float]], threshold: float, alpha: float,
                 top_k: int) -> Tuple[List[int], bool, Optional[str], float]:
    """RAG2Retriever._apply_safety on parallel score arrays; returns (kept indices, refused, reason, max)."""
    n = len(rrf)
    if n == 0:
        return [], True, "No candidates after reranking", 0.0
    eff = [(rerank[i] if (rerank[i] is not None and rerank[i] != 0.0) else rrf[i]) for i in range(n)]
    mx = max(eff)
    if mx < threshold:
        return [], True, f"Max score {mx:.2f} below threshold {threshold}", mx
    floor_ = alpha * mx
    kept = [i for i in range(n) if eff[i] >= floor_]
    return kept[:top_k], False, None, mx
