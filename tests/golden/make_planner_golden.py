"""Golden vectors for the plan a retriever WITHOUT an LLM planner uses: the reference's QueryPlanner.plan when its
LLM call fails (src/voice_agent/rag2/query_planner.py:178-187), run unmodified with a client that raises.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference/src python tests/golden/make_planner_golden.py

Writes tests/golden/planner_golden.json: [{"query", "collection", "plan": dataclasses.asdict(QueryPlan)}] and
tests/golden/settings_golden.json: the reference's default rag2_* retrieval knobs (src/voice_agent/config.py:280-314).
"""
import dataclasses
import json
from pathlib import Path
from unittest.mock import MagicMock

from voice_agent.rag2.query_planner import QueryPlanner

planner = QueryPlanner()
client = MagicMock()
client.chat.completions.create.side_effect = RuntimeError("no network")
planner._client = client
queries = ["Qual é o prazo do contrato?", "  espaços   múltiplos\tentre palavras ", "", "uma", "política de reembolso 2024 — anexo B"]
out = [{"query": q, "collection": c, "plan": dataclasses.asdict(planner.plan(q, c))}
       for q in queries for c in (None, "contracts")]
path = Path(__file__).with_name("planner_golden.json")
path.write_text(json.dumps(out, indent=1, sort_keys=True, ensure_ascii=False) + "\n", encoding="utf-8")
print(f"wrote {len(out)} cases to {path}")

from voice_agent.config import Settings  # noqa: E402
defaults = Settings()   # the class defaults (no .env in this container)
knobs = ("rag2_graph_enabled", "rag2_rerank_enabled", "rag2_safety_threshold", "rag2_denoise_alpha", "rag2_lexical_weight",
         "rag2_semantic_weight", "rag2_graph_weight", "rag2_lexical_top_k", "rag2_semantic_top_k", "rag2_graph_top_k",
         "rag2_rerank_top_k", "rag2_final_top_k")
Path(__file__).with_name("settings_golden.json").write_text(
    json.dumps({k: getattr(defaults, k) for k in knobs}, indent=1, sort_keys=True) + "\n", encoding="utf-8")
