"""Golden vectors for the tool boundary: the reference's own `_search_knowledge_base_rag2`
(src/voice_agent/tools/crm_knowledge.py:69-182), unmodified, run with a stub in place of RAG2Retriever that returns
prepared RetrievalResults (refused, empty, with / without parent text, falsy scores, tables).  This container only:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference/src python tests/golden/make_tool_golden.py

Writes tests/golden/tool_golden.json: [{"query", "category", "limit", "result": <fields of the RetrievalResult>,
"response": <the dictionary the reference's tool returned>}].
"""
import json
import random
from pathlib import Path
from unittest.mock import MagicMock, patch

import voice_agent.rag2.retrieval as ref_retrieval
import voice_agent.tools.crm_knowledge as tool
from voice_agent.rag2.retrieval import RetrievalCandidate, RetrievalResult

rng = random.Random(4242)


def candidate(i):
    rr = rng.choice([None, 0.0, round(rng.random(), 6), rng.random()])
    return dict(child_id=f"c{i}", parent_id=f"p{i % 3}", document_id=f"d{i % 2}", text=f"child text {i}",
                page=rng.choice([None, 1 + i % 7]), modality=rng.choice(["text", "table", "image_caption"]),
                lexical_rank=rng.choice([None, 1 + i]), semantic_rank=rng.choice([None, 2 + i]),
                graph_rank=rng.choice([None, 3 + i]), rrf_score=rng.choice([0.0, rng.random() / 30]),
                parent_text=rng.choice([None, "", f"parent text {i % 3}"]),
                section_heading=rng.choice([None, "", f"heading {i % 3}"]), rerank_score=rr)


cases = []
for n in (0, 1, 3, 5):
    for refused in (False, True):
        res = dict(success=True, contexts=[candidate(i) for i in range(n)],
                   max_rerank_score=rng.choice([0.0, 0.123456789, 0.87654321]), refused=refused,
                   refusal_reason="Max score 0.1000 below threshold 0.6" if refused else None,
                   timings={"planning": 0.0123456, "retrieval": 0.2, "rerank": 1.23456789e-3} if n else {})
        cases.append({"query": f"pergunta {n}", "category": rng.choice([None, "contracts"]), "limit": 5, "result": res})

out = []
for case in cases:
    r = case["result"]
    result = RetrievalResult(success=r["success"], contexts=[RetrievalCandidate(**c) for c in r["contexts"]],
                             max_rerank_score=r["max_rerank_score"], refused=r["refused"],
                             refusal_reason=r["refusal_reason"], timings=dict(r["timings"]))

    class Stub:
        def __init__(self, org_id, graph_enabled=False):
            pass

        async def retrieve(self, query, collection=None, top_k=None):
            return result

    with patch.object(tool, "get_supabase_client", return_value=MagicMock()), \
            patch.object(ref_retrieval, "RAG2Retriever", Stub):
        resp = tool._search_knowledge_base_rag2(case["query"], case["category"], case["limit"], org_id="org")
    out.append({**case, "response": resp})

path = Path(__file__).with_name("tool_golden.json")
path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n", encoding="utf-8")
print(f"wrote {len(out)} cases to {path}")
