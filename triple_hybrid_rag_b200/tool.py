"""The tool boundary above the path: what `_search_knowledge_base_rag2` does around
`RAG2Retriever.retrieve` (reference: src/voice_agent/tools/crm_knowledge.py:69-182), with the candidate-retrieval
step of CONCURRENT calls coalesced into one GPU batch.

The reference builds a retriever per voice turn and runs one query through it (`crm_knowledge.py:105-124`): N
concurrent turns are N scans of the corpus.  `CoalescedRetriever.retrieve` is the same pipeline — plan, channels,
fusion, parent expansion, rerank, safety, the same RetrievalResult — except that steps 2-3 (K1 + K2 + K3) of all calls
that arrive within `max_wait_ms` share one `retrieve_batch` launch; expansion, rerank (K4) and safety (K3) then run
per call on the caller's own candidates.  `search_knowledge_base_rag2` formats the result into the tool's response
dictionary, key for key the reference's.  Organisation look-up and the hybrid / legacy fall-backs stay in the
reference's tool layer (CRM glue: out of scope)."""
from __future__ import annotations

import asyncio
import time
from typing import Any, Dict, List, Optional

import torch

from . import _lib
from .frontend import CoalescingFrontEnd
from .retriever import QueryPlan, RetrievalResult


class CoalescedRetriever:
    def __init__(self, retriever, max_batch: int = 256, max_wait_ms: float = 2.0):
        """retriever: a GpuRAG2Retriever with an index, an embedder (embed_query) and a query planner."""
        self.r = retriever
        self._fe = CoalescingFrontEnd(self._batch, max_batch=max_batch, max_wait_ms=max_wait_ms)

    def _batch(self, queries, vectors, keywords, graph_ids, collections):
        # the reference's tie rule (stable sort: first-seen order lexical -> semantic -> graph) and its channel depths
        cfg = self.r._cfg
        return self.r.retrieve_batch(queries, vectors, keywords, graph_ids=graph_ids, collections=collections,
                                     top_k=max(cfg.rag2_rerank_top_k, 1), k_sem=cfg.rag2_semantic_top_k,
                                     k_lex=cfg.rag2_lexical_top_k, tie_mode=_lib.TIE_INSERTION)

    async def retrieve(self, query: str, collection: Optional[str] = None, top_k: Optional[int] = None,
                       skip_planning: bool = False, skip_rerank: bool = False) -> RetrievalResult:
        """RAG2Retriever.retrieve (retrieval.py:118-201) with the channels + fusion of concurrent calls batched."""
        r, cfg = self.r, self.r._cfg
        timings: Dict[str, float] = {}
        top_k = top_k or cfg.rag2_final_top_k
        t0 = time.time()
        if skip_planning:
            plan = QueryPlan(original_query=query, keywords=query.split(), semantic_query_text=query)
        else:
            plan = await r.query_planner.plan_async(query, collection)
        timings["planning"] = time.time() - t0
        t0 = time.time()
        vec = torch.as_tensor(r.embedder.embed_query(plan.semantic_query_text), dtype=torch.float32)
        graph = None
        if r.graph_enabled and plan.requires_graph and plan.cypher_query:
            rows = await r._graph_search(cypher=plan.cypher_query, keywords=plan.keywords, collection=collection,
                                         limit=plan.graph_top_k)
            graph = [x["child_id"] for x in rows]
        fused = await self._fe.retrieve_candidates(query, vec, plan.keywords, graph_ids=graph, collection=collection)
        timings["retrieval"] = time.time() - t0
        timings["fusion"] = 0.0            # inside the batch: K3 ran with K1 and K2
        if not fused:
            return RetrievalResult(success=True, contexts=[], refused=True, refusal_reason="No candidates found",
                                   query_plan=plan, timings=timings)
        t0 = time.time()
        expanded = await r._expand_to_parents(fused[:cfg.rag2_rerank_top_k])
        timings["expansion"] = time.time() - t0
        if not skip_rerank and cfg.rag2_rerank_enabled:
            t0 = time.time()
            with r.engine.lock:
                reranked = await r._rerank(query, expanded)
            timings["rerank"] = time.time() - t0
        else:
            reranked = expanded
        t0 = time.time()
        with r.engine.lock:
            final, refused, reason, max_score = r._apply_safety(reranked, top_k)
        timings["safety"] = time.time() - t0
        return RetrievalResult(success=True, contexts=final, max_rerank_score=max_score, refused=refused,
                               refusal_reason=reason, query_plan=plan, timings=timings)

    async def drain(self):
        await self._fe.drain()

    def close(self):
        self._fe.close()


def format_tool_response(query: str, category: Optional[str], result: RetrievalResult) -> Dict[str, Any]:
    """The response dictionary of the `search_knowledge_base` tool for a RAG 2.0 result (crm_knowledge.py:126-182)."""
    if result.refused:
        return {"success": True, "query": query, "category": category, "result_count": 0,
                "search_type": "rag2_triple_hybrid", "refused": True, "refusal_reason": result.refusal_reason,
                "results": []}
    results: List[Dict[str, Any]] = []
    for i, ctx in enumerate(result.contexts):
        results.append({
            "chunk_id": ctx.child_id, "parent_id": ctx.parent_id, "document_id": ctx.document_id, "category": category,
            "title": ctx.section_heading or "", "content": ctx.parent_text if ctx.parent_text else ctx.text,
            "source_document": None, "page": ctx.page, "chunk_index": None, "modality": ctx.modality,
            "relevance_rank": i + 1,
            "similarity_score": round(ctx.rrf_score, 4) if ctx.rrf_score else None,
            "rerank_score": round(ctx.rerank_score, 4) if ctx.rerank_score else None,
            "ocr_confidence": None, "is_table": ctx.modality == "table", "table_context": None, "alt_text": None,
            "lexical_rank": ctx.lexical_rank, "semantic_rank": ctx.semantic_rank, "graph_rank": ctx.graph_rank})
    return {"success": True, "query": query, "category": category, "result_count": len(results),
            "search_type": "rag2_triple_hybrid",
            "max_rerank_score": round(result.max_rerank_score, 4) if result.max_rerank_score else None,
            "timings_ms": {k: round(v * 1000, 2) for k, v in result.timings.items()}, "results": results}


def search_knowledge_base_rag2(query: str, category: Optional[str] = None, limit: int = 5, *, retriever) -> Dict[str, Any]:
    """`_search_knowledge_base_rag2(query, category, limit)` (crm_knowledge.py:69-182) over a GPU retriever: runs
    `retriever.retrieve(query, collection=category, top_k=limit)` to completion (the tool handler is synchronous,
    :111-124) and formats the tool's response.  `retriever`: a GpuRAG2Retriever, or a CoalescedRetriever shared by the
    concurrent handlers of one process (then call it from the event loop with `await retriever.retrieve(...)` and
    `format_tool_response` instead: a blocking wrapper cannot coalesce with itself)."""
    coro = retriever.retrieve(query=query, collection=category, top_k=limit)
    try:
        loop = asyncio.get_event_loop()
        if loop.is_running():
            raise RuntimeError("search_knowledge_base_rag2 is the blocking tool handler; inside a running event loop "
                               "await retriever.retrieve(...) and call format_tool_response")
    except RuntimeError as e:
        if "blocking tool handler" in str(e):
            coro.close()
            raise
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
    return format_tool_response(query, category, loop.run_until_complete(coro))
