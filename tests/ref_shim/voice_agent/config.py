"""`from voice_agent.config import SETTINGS` -> the drop-in's knobs object (the reference's tests mutate
SETTINGS.rag2_* in place, tests/test_rag2_triple_hybrid.py:851-860).  RAG2_GRAPH_ENABLED is read from the
environment like the reference's pydantic Settings does (src/voice_agent/config.py:283)."""
import os

from triple_hybrid_rag_b200.retriever import SETTINGS

__thr_shim__ = True
SETTINGS.rag2_graph_enabled = os.environ.get("RAG2_GRAPH_ENABLED", "").lower() in ("1", "true", "yes")
