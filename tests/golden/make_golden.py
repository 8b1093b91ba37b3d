"""Generate golden vectors by running the REFERENCE's own code (this container only).

    PYTHONDONTWRITEBYTECODE=1 RAG2_GRAPH_ENABLED=true \
    PYTHONPATH=/root/reference/src:/root/reference/triple-hybrid-rag/src \
    python tests/golden/make_golden.py

Writes tests/golden/fusion_golden.json.gz.  /root/reference does not exist on the GPU box, so the
vectors are committed; floats are stored as float.hex() strings (bit-exact).
Reference entry points exercised (unmodified, through their public/own call paths):
  RAG2Retriever._retrieve_candidates + _fuse_rrf + _apply_safety  (src/voice_agent/rag2/retrieval.py)
  RRFFusion.fuse                                                  (triple-hybrid-rag/.../core/fusion.py)
  HybridSearcher._rrf_fusion                                      (src/voice_agent/retrieval/hybrid_search.py)
"""
import asyncio
import json
import os
import random
import sys
from pathlib import Path
from unittest.mock import AsyncMock, patch
from uuid import UUID

os.environ.setdefault("RAG2_GRAPH_ENABLED", "true")

from voice_agent.config import SETTINGS  # noqa: E402
from voice_agent.rag2.query_planner import QueryPlan  # noqa: E402
from voice_agent.rag2.retrieval import RAG2Retriever, RetrievalCandidate  # noqa: E402
from voice_agent.retrieval.hybrid_search import HybridSearcher, SearchConfig  # noqa: E402
from voice_agent.retrieval.hybrid_search import SearchResult as R1Result  # noqa: E402
from triple_hybrid_rag.config import RAGConfig  # noqa: E402
from triple_hybrid_rag.core.fusion import RRFFusion  # noqa: E402
from triple_hybrid_rag.types import QueryPlan as LibPlan  # noqa: E402
from triple_hybrid_rag.types import SearchResult as LibResult  # noqa: E402

rng = random.Random(20261018)
hx = float.hex


def ragged_lists(pool, lens, dup_prob=0.0):
    out = []
    for n in lens:
        ids = rng.sample(range(pool), min(n, pool))
        if dup_prob and len(ids) > 3 and rng.random() < dup_prob:
            ids[rng.randrange(1, len(ids))] = ids[0]  # duplicate inside one channel
        out.append(ids)
    return out


def row(cid):
    return {"child_id": str(cid), "parent_id": f"p{cid}", "document_id": f"d{cid}", "text": f"t{cid}",
            "page": 1, "modality": "text"}


def run_rag2(lists, weights):
    r = RAG2Retriever(org_id="golden", graph_enabled=True)
    assert r.graph_enabled, "set RAG2_GRAPH_ENABLED=true"
    plan = QueryPlan(original_query="q", keywords=["k"] if lists[0] is not None else [],
                     semantic_query_text="q", requires_graph=lists[2] is not None,
                     cypher_query="MATCH (e) RETURN e" if lists[2] is not None else None, weights=weights)

    async def go():
        with patch.object(r, "_lexical_search", new_callable=AsyncMock) as ml, \
                patch.object(r, "_semantic_search", new_callable=AsyncMock) as ms, \
                patch.object(r, "_graph_search", new_callable=AsyncMock) as mg:
            ml.return_value = [row(c) for c in (lists[0] or [])]
            ms.return_value = [row(c) for c in (lists[1] or [])]
            mg.return_value = [row(c) for c in (lists[2] or [])]
            cands = await r._retrieve_candidates(plan, None)
        return r._fuse_rrf(cands, plan.weights)

    fused = asyncio.run(go())
    return [{"id": int(c.child_id), "rrf": hx(c.rrf_score),
             "ranks": [c.lexical_rank or 0, c.semantic_rank or 0, c.graph_rank or 0]} for c in fused]


def run_safety(rrf, rerank, thr, alpha, top_k):
    r = RAG2Retriever(org_id="golden")
    cands = [RetrievalCandidate(child_id=str(i), parent_id="p", document_id="d", text="", page=1, modality="text",
                                rrf_score=rrf[i], rerank_score=rerank[i]) for i in range(len(rrf))]
    old = SETTINGS.rag2_safety_threshold, SETTINGS.rag2_denoise_alpha
    SETTINGS.rag2_safety_threshold, SETTINGS.rag2_denoise_alpha = thr, alpha
    try:
        final, refused, reason, mx = r._apply_safety(cands, top_k)
    finally:
        SETTINGS.rag2_safety_threshold, SETTINGS.rag2_denoise_alpha = old
    return {"kept": [int(c.child_id) for c in final], "refused": refused, "reason": reason, "max": hx(float(mx))}


def run_lib(lists, raws, weights, thr, alpha, denoise, top_k):
    cfg = RAGConfig(rag_safety_threshold=thr, rag_denoise_alpha=alpha, rag_denoise_enabled=denoise)
    f = RRFFusion(cfg)
    names = ["lexical_score", "semantic_score", "graph_score"]
    rs = []
    for c in range(3):
        cur = []
        for pos, cid in enumerate(lists[c] or []):
            x = LibResult(chunk_id=UUID(int=cid))
            setattr(x, names[c], raws[c][pos])
            cur.append(x)
        rs.append(cur)
    plan = LibPlan(weights=weights) if weights is not None else None
    out = f.fuse(rs[0], rs[1], rs[2], query_plan=plan, top_k=top_k)
    return [{"id": x.chunk_id.int, "rrf": hx(x.rrf_score),
             "raw": [hx(float(x.lexical_score)), hx(float(x.semantic_score)), hx(float(x.graph_score))]} for x in out]


def run_rag1(lists, rrf_k):
    s = HybridSearcher(org_id="golden", embedder=object(), config=SearchConfig(rrf_k=rrf_k))
    rl = [[R1Result(chunk_id=str(c), content="", modality="text", source_document="", page=1, chunk_index=0)
           for c in ids] for ids in lists if ids is not None]
    out = s._rrf_fusion(rl)
    return [{"id": int(x.chunk_id), "rrf": hx(x.rrf_score)} for x in out]


cases = {"rag2": [], "safety": [], "lib": [], "rag1": []}
W = {"lexical": 0.7, "semantic": 0.8, "graph": 1.0}

# SURVEY App. A golden A and B (ties: insertion order L10, S20, G40)
fixed = [
    ([[1, 2, 3], [2, 4, 1], [4, 5]], W),
    ([list(range(1001, 1011)), list(range(2001, 2021)), list(range(3001, 3041))], W),
    ([[], [7, 8, 9], None], W),
    ([None, [5], None], W),
    ([[3, 3, 4], [4, 3], [3]], W),  # duplicates inside a channel: last rank wins
]
for lists, w in fixed:
    cases["rag2"].append({"lists": lists, "weights": [w["lexical"], w["semantic"], w["graph"]],
                          "out": run_rag2(lists, dict(w))})
for i in range(160):
    pool = rng.choice([40, 120, 400, 5000])
    lens = [rng.randint(0, 50), rng.randint(1, 100), rng.randint(0, 50)]
    lists = ragged_lists(pool, lens, dup_prob=0.15)
    if rng.random() < 0.15:
        lists[0] = None
    if rng.random() < 0.25:
        lists[2] = None
    w = dict(W) if rng.random() < 0.6 else {"lexical": round(rng.uniform(0.1, 2.0), 3),
                                            "semantic": round(rng.uniform(0.1, 2.0), 3),
                                            "graph": round(rng.uniform(0.1, 2.0), 3)}
    cases["rag2"].append({"lists": lists, "weights": [w["lexical"], w["semantic"], w["graph"]],
                          "out": run_rag2(lists, w)})

# _apply_safety (App. A golden C + random)
saf = [([0.02] * 6, [0.9, 0.7, 0.54, 0.5399999, 0.0, 0.3], 0.6, 0.6, 5),
       ([0.02], [0.59], 0.6, 0.6, 5), ([], [], 0.6, 0.6, 5),
       ([0.03, 0.02, 0.01], [None, None, None], 0.6, 0.6, 5),
       ([0.03, 0.02, 0.01], [None, None, None], 0.0, 0.0, 2)]
for _ in range(80):
    n = rng.randint(1, 60)
    rrf = [rng.uniform(0.001, 0.05) for _ in range(n)]
    rer = [rng.choice([None, 0.0, rng.uniform(0, 1), rng.uniform(0.5, 1)]) for _ in range(n)]
    saf.append((rrf, rer, rng.choice([0.0, 0.3, 0.6]), rng.choice([0.0, 0.5, 0.6, 0.9]), rng.randint(1, 20)))
for rrf, rer, thr, alpha, top_k in saf:
    cases["safety"].append({"rrf": [hx(x) for x in rrf], "rerank": [None if x is None else hx(x) for x in rer],
                            "thr": thr, "alpha": alpha, "top_k": top_k, "out": run_safety(rrf, rer, thr, alpha, top_k)})

# RRFFusion.fuse (App. A golden D + random)
libfixed = [([[1, 2, 3], [2, 4, 1, 5, 6], [4, 7]],
             [[0.9, 0.3, 0.05], [0.82, 0.71, 0.65, 0.58, 0.61], [1.0, 1.0]], None, 0.6, 0.6, True, None)]
for lists, raws, w, thr, alpha, dn, tk in libfixed:
    cases["lib"].append({"lists": lists, "raw": [[hx(x) for x in r] for r in raws], "weights": w, "thr": thr,
                         "alpha": alpha, "denoise": dn, "top_k": tk,
                         "out": run_lib(lists, raws, w, thr, alpha, dn, tk)})
for i in range(160):
    pool = rng.choice([30, 150, 3000])
    lens = [rng.randint(0, 50), rng.randint(0, 100), rng.randint(0, 50)]
    lists = ragged_lists(pool, lens, dup_prob=0.1)
    raws = [[rng.choice([rng.uniform(0, 1), rng.uniform(0.5, 1), 0.0]) for _ in l] for l in lists]
    w = None if rng.random() < 0.5 else {"lexical": round(rng.uniform(0.1, 2.0), 3),
                                         "semantic": round(rng.uniform(0.1, 2.0), 3),
                                         "graph": round(rng.uniform(0.1, 2.0), 3)}
    thr = rng.choice([0.0, 0.3, 0.6])
    alpha = rng.choice([0.6, 0.5, 0.9, 0.05, 1.0, 0.0])
    dn = rng.random() < 0.8
    tk = rng.choice([None, 5, 20, 50])
    cases["lib"].append({"lists": lists, "raw": [[hx(x) for x in r] for r in raws],
                         "weights": None if w is None else [w["lexical"], w["semantic"], w["graph"]],
                         "thr": thr, "alpha": alpha, "denoise": dn, "top_k": tk,
                         "out": run_lib(lists, raws, w, thr, alpha, dn, tk)})

# HybridSearcher._rrf_fusion (vector, bm25[, image])
for i in range(60):
    pool = rng.choice([30, 200, 4000])
    lens = [rng.randint(1, 50), rng.randint(0, 50), rng.randint(0, 5)]
    lists = ragged_lists(pool, lens)
    if rng.random() < 0.5:
        lists[2] = None
    k = rng.choice([60, 60, 10, 1])
    cases["rag1"].append({"lists": lists, "rrf_k": k, "out": run_rag1(lists, k)})

import gzip  # noqa: E402
out = Path(__file__).with_name("fusion_golden.json.gz")
with gzip.GzipFile(out, "wb", mtime=0) as fh:
    fh.write(json.dumps(cases, separators=(",", ":")).encode())
print({k: len(v) for k, v in cases.items()}, "->", out, out.stat().st_size, "bytes")
