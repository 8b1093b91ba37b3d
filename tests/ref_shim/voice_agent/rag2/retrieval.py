"""`from voice_agent.rag2.retrieval import RAG2Retriever, RetrievalCandidate, RetrievalResult, retrieve` -> the
drop-in.  The reference's tests build `RAG2Retriever(org_id=..., graph_enabled=...)` with nothing else (its clients
are lazy, src/voice_agent/rag2/retrieval.py:79-116); here the engine of the one GPU is supplied the same way."""
from unittest.mock import MagicMock

import voice_agent.config  # noqa: F401  (applies RAG2_GRAPH_ENABLED before a retriever reads it)
from triple_hybrid_rag_b200 import retriever as _r
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.retriever import RetrievalCandidate, RetrievalResult  # noqa: F401

__thr_shim__ = True
_engine = None


def _shared_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(0)       # raises without a B200: there is no CPU fallback
    return _engine


class RAG2Retriever(_r.StandaloneGpuRAG2Retriever):
    def __init__(self, org_id, embedder=None, query_planner=None, graph_enabled=False, **kw):
        kw.setdefault("engine", _shared_engine())
        super().__init__(org_id, embedder=embedder or MagicMock(), query_planner=query_planner,
                         graph_enabled=graph_enabled, **kw)


async def retrieve(org_id, query, **kwargs):
    """retrieval.py:498-505: the module-level convenience function (resolves RAG2Retriever through this module, as
    the reference does, so that `patch('voice_agent.rag2.retrieval.RAG2Retriever')` takes effect)."""
    import voice_agent.rag2.retrieval as me
    retriever = me.RAG2Retriever(org_id=org_id)
    return await retriever.retrieve(query, **kwargs)
