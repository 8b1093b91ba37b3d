"""Golden vectors for the WHOLE reference pipeline: the unmodified RAG2Retriever.retrieve() driven with
fixed channel outputs (the way the reference's own tests drive it, tests/test_rag2_triple_hybrid.py:44-74)
and a fixed reranker response.  This container only (/root/reference is not on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 RAG2_GRAPH_ENABLED=true PYTHONPATH=/root/reference/src \
    python tests/golden/make_retrieve_golden.py

Writes tests/golden/retrieve_golden.json.gz; floats as float.hex() (bit-exact).
"""
import asyncio
import gzip
import json
import os
import random
from pathlib import Path
from unittest.mock import AsyncMock, patch

os.environ.setdefault("RAG2_GRAPH_ENABLED", "true")

from voice_agent.config import SETTINGS  # noqa: E402
from voice_agent.rag2.query_planner import QueryPlan  # noqa: E402
from voice_agent.rag2.retrieval import RAG2Retriever  # noqa: E402
from voice_agent.retrieval.reranker import Qwen3VLReranker  # noqa: E402

rng = random.Random(77001)
hx = float.hex
KNOBS = ("rag2_rerank_top_k", "rag2_safety_threshold", "rag2_denoise_alpha", "rag2_rerank_enabled")


def row(cid):
    return {"child_id": str(cid), "parent_id": f"p{cid % 7}", "document_id": f"d{cid}", "text": f"text {cid}",
            "page": 1 + cid % 5, "modality": "table" if cid % 11 == 0 else "text"}


def run(case):
    lists, weights = case["lists"], case["weights"]
    old = {k: getattr(SETTINGS, k) for k in KNOBS}
    for k in KNOBS:
        setattr(SETTINGS, k, case["settings"][k])
    try:
        r = RAG2Retriever(org_id="golden", graph_enabled=True)
        plan = QueryPlan(original_query="q", keywords=["k"] if lists[0] is not None else [], semantic_query_text="q",
                         requires_graph=lists[2] is not None,
                         cypher_query="MATCH (e) RETURN e" if lists[2] is not None else None,
                         weights={"lexical": weights[0], "semantic": weights[1], "graph": weights[2]})
        scores = {str(k): v for k, v in case["rerank"].items()}

        async def native(self, query, documents):  # the reranker's response: one float per document, input order
            return [scores.get(d.split()[-1], 0.5) for d in documents]

        async def identity(c):
            return c

        async def go():
            with patch.object(r, "_lexical_search", new_callable=AsyncMock) as ml, \
                    patch.object(r, "_semantic_search", new_callable=AsyncMock) as ms, \
                    patch.object(r, "_graph_search", new_callable=AsyncMock) as mg, \
                    patch.object(r.query_planner, "plan_async", new_callable=AsyncMock) as mp, \
                    patch.object(r, "_expand_to_parents", side_effect=identity), \
                    patch.object(Qwen3VLReranker, "_rerank_batch_native", native):
                ml.return_value = [row(c) for c in (lists[0] or [])]
                ms.return_value = [row(c) for c in (lists[1] or [])]
                mg.return_value = [row(c) for c in (lists[2] or [])]
                mp.return_value = plan
                return await r.retrieve("q", top_k=case["top_k"], skip_rerank=case["skip_rerank"])

        res = asyncio.run(go())
    finally:
        for k, v in old.items():
            setattr(SETTINGS, k, v)
    return {"refused": res.refused, "reason": res.refusal_reason, "max": hx(float(res.max_rerank_score)),
            "success": res.success, "timings": sorted(res.timings),
            "contexts": [{"id": int(c.child_id), "rrf": hx(c.rrf_score),
                          "ranks": [c.lexical_rank or 0, c.semantic_rank or 0, c.graph_rank or 0],
                          "rerank": None if c.rerank_score is None else hx(float(c.rerank_score)),
                          "modality": c.modality, "page": c.page, "parent_id": c.parent_id} for c in res.contexts]}


cases = []
for i in range(120):
    pool = rng.choice([60, 300, 4000])
    lens = [rng.randint(0, 50), rng.randint(1, 100), rng.randint(0, 50)]
    lists = [rng.sample(range(pool), min(n, pool)) for n in lens]
    if rng.random() < 0.15:
        lists[0] = None
    if rng.random() < 0.25:
        lists[2] = None
    if i == 0:
        lists = [None, [], None]            # nothing found -> "No candidates found"
    w = [0.7, 0.8, 1.0] if rng.random() < 0.6 else [round(rng.uniform(0.1, 2.0), 3) for _ in range(3)]
    ids = sorted({c for l in lists if l for c in l})
    mode = rng.choice(["high", "mixed", "low", "zeros"])
    rer = {}
    for c in ids:
        if mode == "high":
            rer[c] = rng.uniform(0.55, 1.0)
        elif mode == "mixed":
            rer[c] = rng.choice([rng.uniform(0, 1), rng.uniform(0.6, 0.99), 0.0])
        elif mode == "low":
            rer[c] = rng.uniform(0.0, 0.59)
        else:
            rer[c] = 0.0
    settings = {"rag2_rerank_top_k": rng.choice([20, 20, 50, 100]),
                "rag2_safety_threshold": rng.choice([0.6, 0.6, 0.0, 0.3]),
                "rag2_denoise_alpha": rng.choice([0.6, 0.6, 0.0, 0.9]),
                "rag2_rerank_enabled": rng.random() < 0.9}
    case = {"lists": lists, "weights": w, "rerank": rer, "settings": settings,
            "top_k": rng.choice([5, 5, 10, 50]), "skip_rerank": rng.random() < 0.2}
    case["out"] = run(case)
    case["rerank"] = {str(k): hx(v) for k, v in rer.items()}
    cases.append(case)

out = Path(__file__).with_name("retrieve_golden.json.gz")
with gzip.GzipFile(out, "wb", mtime=0) as fh:
    fh.write(json.dumps(cases, separators=(",", ":")).encode())
n_ref = sum(c["out"]["refused"] for c in cases)
print(len(cases), "cases,", n_ref, "refused ->", out, out.stat().st_size, "bytes")
