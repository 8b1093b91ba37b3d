"""Drop-in for the reference's RAG 2.0 retriever call surface, scored on the GPU.

Mirrors `RAG2Retriever` (reference: src/voice_agent/rag2/retrieval.py:66-495): same constructor
arguments, same coroutine/method names and argument meaning, same result dataclasses, same error
conventions (a refusal is data, channel failures degrade to empty lists) — so the reference's own
tests read the same against this class (they patch `_lexical_search`, `_semantic_search`,
`_graph_search`, `_expand_to_parents`, `_apply_safety` by name; see tests/test_gpu_retriever.py).
What changes is where the arithmetic runs:

    reference                                         here
    _lexical_search  -> Postgres RPC rag2_lexical_search   K2  thr_bm25_topk over the resident CSR index
    _semantic_search -> pgvector RPC rag2_semantic_search  K1  thr_dense_topk (exact, tcgen05)
    _graph_search    -> PuppyGraph / SQL fallback          still external: a callable that returns ids
    _fuse_rrf        -> Python floats                      K3  thr_fuse_ranked (bit-identical fp64)
    _rerank          -> HTTP cross-encoder                 K4  thr_maxsim over a resident token store
    _apply_safety    -> Python floats                      K3  thr_safety (bit-identical fp64)

There is no CPU fallback: every scoring method needs the Engine (libthr.so on a B200).
`retrieve_batch` is the batched entry the reference lacks (one query per call there).
"""
from __future__ import annotations

import re
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import Engine
from .index import BM25Index, pack_queries


# ---- result / plan types: field-for-field the reference's (retrieval.py:26-63, query_planner.py:23-50) ----
@dataclass
class RetrievalCandidate:
    child_id: str
    parent_id: str
    document_id: str
    text: str
    page: int
    modality: str
    lexical_rank: Optional[int] = None
    semantic_rank: Optional[int] = None
    graph_rank: Optional[int] = None
    rrf_score: float = 0.0
    parent_text: Optional[str] = None
    section_heading: Optional[str] = None
    rerank_score: Optional[float] = None


@dataclass
class QueryPlan:
    original_query: str
    keywords: List[str] = field(default_factory=list)
    lexical_top_k: int = 50
    semantic_query_text: str = ""
    semantic_top_k: int = 100
    cypher_query: Optional[str] = None
    graph_top_k: int = 50
    weights: Dict[str, float] = field(default_factory=lambda: {"lexical": 0.7, "semantic": 0.8, "graph": 1.0})
    intent: str = "general"
    requires_graph: bool = False


@dataclass
class RetrievalResult:
    success: bool
    contexts: List[RetrievalCandidate]
    max_rerank_score: float = 0.0
    refused: bool = False
    refusal_reason: Optional[str] = None
    query_plan: Optional[QueryPlan] = None
    timings: Dict[str, float] = field(default_factory=dict)


class Rag2Settings:
    """The RAG 2.0 knobs of the reference's Settings (src/voice_agent/config.py:280-314), read at call
    time (the reference's tests mutate them in place)."""
    rag2_graph_enabled: bool = False
    rag2_rerank_enabled: bool = True
    rag2_safety_threshold: float = 0.6
    rag2_denoise_alpha: float = 0.6
    rag2_lexical_weight: float = 0.7
    rag2_semantic_weight: float = 0.8
    rag2_graph_weight: float = 1.0
    rag2_lexical_top_k: int = 50
    rag2_semantic_top_k: int = 100
    rag2_graph_top_k: int = 50
    rag2_rerank_top_k: int = 20
    rag2_final_top_k: int = 5


SETTINGS = Rag2Settings()

_WORD = re.compile(r"\w+", re.UNICODE)


def tokenize(text: str) -> List[str]:
    """Lower-cased word tokens.  (Postgres' to_tsvector('portuguese') also stems and drops stop words;
    that linguistic front end is outside the scoring path — SURVEY.md §8f row 1.)"""
    return _WORD.findall(text.lower())


class ResidentIndex:
    """Everything the scoring path keeps in HBM for one tenant: the bf16 embedding matrix, the BM25
    inverted index, optionally a per-chunk token store for MaxSim, plus the host-side row metadata the
    retriever returns (the reference reads those columns from rag_child_chunks / rag_parent_chunks)."""

    def __init__(self, engine: Engine, chunks: Sequence[Dict[str, Any]], embeddings: torch.Tensor,
                 parents: Optional[Dict[str, Dict[str, Any]]] = None, blk_docs: int = 1024,
                 token_store: Optional[torch.Tensor] = None, token_lens: Optional[torch.Tensor] = None):
        """chunks: rows with child_id, parent_id, document_id, text, page, modality (row i <-> embeddings[i])."""
        if len(chunks) != embeddings.shape[0]:
            raise ValueError("one embedding row per chunk")
        self.engine = engine
        dev = engine.device
        self.rows = list(chunks)
        self.id_of = {r["child_id"]: i for i, r in enumerate(self.rows)}
        self.parents = dict(parents or {})
        self.collections = [r.get("collection") for r in self.rows]
        # dense
        X = embeddings.to(torch.float32)
        X = X / X.norm(dim=1, keepdim=True).clamp_min(1e-30)
        self.X = X.to(torch.bfloat16).to(dev).contiguous()
        engine.dense_index_set(self.X)
        # lexical
        vocab: Dict[str, int] = {}
        d_l, t_l, f_l, lens = [], [], [], []
        for i, r in enumerate(self.rows):
            toks = tokenize(r.get("text", ""))
            lens.append(max(len(toks), 1))
            tf: Dict[int, int] = {}
            for w in toks:
                t = vocab.setdefault(w, len(vocab))
                tf[t] = tf.get(t, 0) + 1
            for t, c in tf.items():
                d_l.append(i); t_l.append(t); f_l.append(c)
        self.vocab = vocab
        V = max(len(vocab), 1)
        self.bm25 = BM25Index.build(torch.tensor(d_l, dtype=torch.int64), torch.tensor(t_l, dtype=torch.int64),
                                    torch.tensor(f_l, dtype=torch.int32), torch.tensor(lens, dtype=torch.int64), V,
                                    blk_docs=blk_docs).to(dev)
        engine.bm25_index_set(self.bm25.skip, self.bm25.postings, self.bm25.idf, self.bm25.n_docs, self.bm25.blk_docs, V)
        # late-interaction token store (optional)
        self.token_store = None if token_store is None else token_store.to(torch.bfloat16).to(dev).contiguous()
        self.token_lens = None if token_lens is None else token_lens.to(torch.int32).to(dev).contiguous()
        self._set_tags()

    def _set_tags(self) -> None:
        """Collection names -> uint16 tags resident next to the indexes: the reference's `collection` predicate
        (20260114_rag2_schema.sql:368-370, :404-406) is evaluated inside K1 / K2 (thr_*_topk_tagged)."""
        names = sorted({c for c in self.collections if c is not None})
        if len(names) >= 0xffff:
            raise ValueError("at most 65534 collections per resident index")
        self.tag_of = {c: i for i, c in enumerate(names)}
        tags = torch.tensor([self.tag_of.get(c, 0xffff) for c in self.collections], dtype=torch.int32)
        self.tags = tags.to(torch.uint16).to(self.engine.device).contiguous()
        self.engine.dense_tags_set(self.tags)
        self.engine.bm25_tags_set(self.tags)

    def want(self, collection: Optional[str]) -> Optional[torch.Tensor]:
        """The per-query filter argument of the tagged kernels for one query (None: no filter).  A collection no
        chunk carries maps to a tag no chunk carries, i.e. to an empty result, like the SQL predicate."""
        if collection is None:
            return None
        return torch.tensor([self.tag_of.get(collection, 0xfffe)], dtype=torch.int32, device=self.engine.device)

    # ---- persistence: start a retriever without re-reading / re-embedding the corpus ----
    def save(self, path) -> None:
        torch.save({"format": "thr-resident-v1", "rows": self.rows, "parents": self.parents, "vocab": self.vocab,
                    "X": self.X.cpu(), "token_store": None if self.token_store is None else self.token_store.cpu(),
                    "token_lens": None if self.token_lens is None else self.token_lens.cpu(),
                    "bm25": {"skip": self.bm25.skip.cpu(), "postings": self.bm25.postings.cpu(),
                             "idf": self.bm25.idf.cpu(), "df": self.bm25.df.cpu(), "n_docs": self.bm25.n_docs,
                             "blk_docs": self.bm25.blk_docs, "V": self.bm25.V, "nnz": self.bm25.nnz,
                             "k1": self.bm25.k1, "b": self.bm25.b, "avgdl": self.bm25.avgdl}}, path)

    @classmethod
    def load(cls, engine: Engine, path) -> "ResidentIndex":
        d = torch.load(path, map_location="cpu", weights_only=True)
        if d.get("format") != "thr-resident-v1":
            raise ValueError(f"{path}: not a ResidentIndex file")
        self = cls.__new__(cls)
        dev = engine.device
        self.engine = engine
        self.rows = d["rows"]
        self.id_of = {r["child_id"]: i for i, r in enumerate(self.rows)}
        self.parents = d["parents"]
        self.collections = [r.get("collection") for r in self.rows]
        self.vocab = d["vocab"]
        self.X = d["X"].to(dev).contiguous()
        engine.dense_index_set(self.X)
        b = d["bm25"]
        self.bm25 = BM25Index(b["skip"], b["postings"], b["idf"], b["df"], int(b["n_docs"]), int(b["blk_docs"]),
                              int(b["V"]), int(b["nnz"]), float(b["k1"]), float(b["b"]), float(b["avgdl"])).to(dev)
        engine.bm25_index_set(self.bm25.skip, self.bm25.postings, self.bm25.idf, self.bm25.n_docs, self.bm25.blk_docs,
                              self.bm25.V)
        self.token_store = None if d["token_store"] is None else d["token_store"].to(dev).contiguous()
        self.token_lens = None if d["token_lens"] is None else d["token_lens"].to(dev).contiguous()
        self._set_tags()
        return self

    def row_dict(self, i: int, **extra) -> Dict[str, Any]:
        r = self.rows[i]
        d = {"child_id": r["child_id"], "parent_id": r["parent_id"], "document_id": r["document_id"],
             "text": r.get("text", ""), "page": r.get("page", 1), "modality": r.get("modality", "text")}
        d.update(extra)
        return d


class GpuRAG2Retriever:
    """Same call surface as the reference's RAG2Retriever (retrieval.py:66-495)."""

    def __init__(self, org_id: str, embedder: Any = None, query_planner: Any = None, graph_enabled: bool = False, *,
                 index: Optional[ResidentIndex] = None, engine: Optional[Engine] = None,
                 graph_search: Optional[Callable[..., Sequence[str]]] = None,
                 token_encoder: Optional[Callable[[str], torch.Tensor]] = None):
        """org_id / embedder / query_planner / graph_enabled: as in the reference (retrieval.py:79-101).
        index: the tenant's ResidentIndex.  graph_search(cypher=, keywords=, collection=, limit=) -> ranked
        child ids (the graph engine stays external).  token_encoder(query) -> [Tq, 128] token embeddings."""
        self.org_id = org_id
        self.embedder = embedder
        self.query_planner = query_planner
        self.graph_enabled = graph_enabled and SETTINGS.rag2_graph_enabled
        self.index = index
        self.engine = engine or (index.engine if index is not None else None)
        self._graph_search_fn = graph_search
        self.token_encoder = token_encoder

    def _need_engine(self) -> Engine:
        if self.engine is None:
            raise RuntimeError("GpuRAG2Retriever needs an Engine (libthr.so on a B200); there is no CPU fallback")
        return self.engine

    # ---- pipeline (reference: retrieval.py:118-201) ---------------------------------------------
    async def retrieve(self, query: str, collection: Optional[str] = None, top_k: Optional[int] = None,
                       skip_planning: bool = False, skip_rerank: bool = False) -> RetrievalResult:
        timings: Dict[str, float] = {}
        top_k = top_k or SETTINGS.rag2_final_top_k
        t0 = time.time()
        if skip_planning:
            plan = QueryPlan(original_query=query, keywords=query.split(), semantic_query_text=query)
        else:
            plan = await self.query_planner.plan_async(query, collection)
        timings["planning"] = time.time() - t0

        t0 = time.time()
        candidates = await self._retrieve_candidates(plan, collection)
        timings["retrieval"] = time.time() - t0
        if not candidates:
            return RetrievalResult(success=True, contexts=[], refused=True, refusal_reason="No candidates found",
                                   query_plan=plan, timings=timings)

        t0 = time.time()
        fused = self._fuse_rrf(candidates, plan.weights)
        timings["fusion"] = time.time() - t0

        t0 = time.time()
        expanded = await self._expand_to_parents(fused[:SETTINGS.rag2_rerank_top_k])
        timings["expansion"] = time.time() - t0

        if not skip_rerank and SETTINGS.rag2_rerank_enabled:
            t0 = time.time()
            reranked = await self._rerank(query, expanded)
            timings["rerank"] = time.time() - t0
        else:
            reranked = expanded

        t0 = time.time()
        final, refused, reason, max_score = self._apply_safety(reranked, top_k)
        timings["safety"] = time.time() - t0
        return RetrievalResult(success=True, contexts=final, max_rerank_score=max_score, refused=refused,
                               refusal_reason=reason, query_plan=plan, timings=timings)

    async def _retrieve_candidates(self, plan: QueryPlan, collection: Optional[str]) -> List[RetrievalCandidate]:
        """Union of the channel lists keyed by child_id; 1-based rank per channel; first-seen order
        lexical -> semantic -> graph (reference: retrieval.py:203-271)."""
        merged: Dict[str, RetrievalCandidate] = {}

        def absorb(rows: Sequence[Dict[str, Any]], attr: str):
            for rank, r in enumerate(rows, 1):
                cid = r["child_id"]
                c = merged.get(cid)
                if c is None:
                    c = merged[cid] = RetrievalCandidate(child_id=cid, parent_id=r["parent_id"],
                                                         document_id=r["document_id"], text=r["text"],
                                                         page=r.get("page", 1), modality=r.get("modality", "text"))
                setattr(c, attr, rank)

        if plan.keywords:
            absorb(await self._lexical_search(keywords=plan.keywords, collection=collection,
                                              limit=plan.lexical_top_k), "lexical_rank")
        absorb(await self._semantic_search(query_text=plan.semantic_query_text, collection=collection,
                                           limit=plan.semantic_top_k), "semantic_rank")
        if self.graph_enabled and plan.requires_graph and plan.cypher_query:
            absorb(await self._graph_search(cypher=plan.cypher_query, keywords=plan.keywords, collection=collection,
                                            limit=plan.graph_top_k), "graph_rank")
        return list(merged.values())

    # ---- channels ---------------------------------------------------------------------------------
    async def _lexical_search(self, keywords: List[str], collection: Optional[str], limit: int) -> List[Dict[str, Any]]:
        """Reference: retrieval.py:273-292 (query = the keywords joined by spaces, top-`limit` rows, best first).
        Rows carry `rank` = the BM25 score, as the RPC returns its ts_rank_cd."""
        eng, ix = self._need_engine(), self.index
        terms = [ix.vocab[w] for w in tokenize(" ".join(keywords)) if w in ix.vocab][:32]
        if not terms:
            return []
        qt, qo = pack_queries([terms], eng.device)
        ids, sc, cnt = eng.bm25_topk(qt, qo, min(limit, 256), want=ix.want(collection))   # predicate inside K2
        eng.sync()
        n = int(cnt[0])
        return [ix.row_dict(i, rank=r) for i, r in zip(ids[0, :n].tolist(), sc[0, :n].tolist())]

    async def _semantic_search(self, query_text: str, collection: Optional[str], limit: int) -> List[Dict[str, Any]]:
        """Reference: retrieval.py:294-314 (embed the query, top-`limit` by cosine similarity, best first)."""
        eng, ix = self._need_engine(), self.index
        q = torch.as_tensor(self.embedder.embed_query(query_text), dtype=torch.float32).reshape(1, -1)
        q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)
        k = min(limit, 228, len(ix.rows))
        ids, sc, cnt, _ = eng.dense_topk(q.to(torch.bfloat16).to(eng.device), k, want=ix.want(collection))  # inside K1
        eng.sync()
        n = int(cnt[0])
        return [ix.row_dict(i, similarity=v) for i, v in zip(ids[0, :n].tolist(), sc[0, :n].tolist())]

    async def _graph_search(self, cypher: str, keywords: List[str], collection: Optional[str],
                            limit: int) -> List[Dict[str, Any]]:
        """Reference: retrieval.py:316-356.  The graph engine stays external: its ranked child ids are an
        input; failures degrade to an empty list exactly as the reference's try/except does."""
        if self._graph_search_fn is None:
            return []
        try:
            ids = self._graph_search_fn(cypher=cypher, keywords=keywords, collection=collection, limit=limit)
            out = []
            for cid in list(ids)[:limit]:
                i = self.index.id_of.get(cid)
                if i is not None:
                    out.append(self.index.row_dict(i))
            return out
        except Exception:
            return []

    # ---- K3: fusion -------------------------------------------------------------------------------
    def _fuse_rrf(self, candidates: List[RetrievalCandidate], weights: Dict[str, float], k: int = 60
                  ) -> List[RetrievalCandidate]:
        """Reference: retrieval.py:358-376.  Sets c.rrf_score in place (bit-identical fp64) and returns the
        same objects in the order of Python's stable descending sort."""
        if not candidates:
            return []
        eng = self._need_engine()
        n = len(candidates)
        ranks = torch.tensor([[c.lexical_rank or 0, c.semantic_rank or 0, c.graph_rank or 0] for c in candidates],
                             dtype=torch.int32, device=eng.device)
        w = torch.tensor([[weights.get("lexical", 0.7), weights.get("semantic", 0.8), weights.get("graph", 1.0)]],
                         dtype=torch.float64, device=eng.device)
        off = torch.tensor([0, n], dtype=torch.int32, device=eng.device)
        rrf, order = eng.fuse_ranked(off, ranks, w, rrf_k=k)
        eng.sync()
        for c, s in zip(candidates, rrf.tolist()):
            c.rrf_score = s
        return [candidates[i] for i in order.tolist()]

    async def _expand_to_parents(self, candidates: List[RetrievalCandidate]) -> List[RetrievalCandidate]:
        """Reference: retrieval.py:378-403 (attach parent text / heading where the parent is known)."""
        if not candidates:
            return []
        parents = self.index.parents if self.index is not None else {}
        for c in candidates:
            p = parents.get(c.parent_id)
            if p is not None:
                c.parent_text = p["text"]
                c.section_heading = p.get("section_heading")
        return candidates

    # ---- K4: rerank -------------------------------------------------------------------------------
    async def _rerank_batch_native(self, query: str, documents: List[str]) -> List[float]:
        """Stands where Qwen3VLReranker._rerank_batch_native(query, documents) -> List[float] does
        (reference: src/voice_agent/retrieval/reranker.py:287-354): one score in [0, 1] per document, in
        input order.  The reference posts document TEXTS to an HTTP model; here `documents` are child ids
        of rows in the resident token store.  Score = MaxSim averaged over the query tokens, mapped from
        [-1, 1] to [0, 1], so the 0.6 safety threshold keeps its meaning; unknown ids get the reference's
        neutral 0.5."""
        eng, ix = self._need_engine(), self.index
        if ix is None or ix.token_store is None or self.token_encoder is None:
            raise RuntimeError("no token store / token encoder: late-interaction rerank unavailable")
        qtok = self.token_encoder(query).to(torch.float32)
        qtok = qtok / qtok.norm(dim=-1, keepdim=True).clamp_min(1e-30)
        qtok = qtok.to(torch.bfloat16).to(eng.device).unsqueeze(0).contiguous()
        rows = [ix.id_of.get(d, -1) for d in documents]
        cand = torch.tensor([rows], dtype=torch.int64, device=eng.device)
        raw = eng.maxsim(qtok, ix.token_store, cand, d_len=ix.token_lens)
        eng.sync()
        tq = qtok.shape[1]
        return [0.5 if r < 0 else min(1.0, max(0.0, 0.5 * (float(s) / tq + 1.0))) for r, s in zip(rows, raw[0].tolist())]

    async def _rerank(self, query: str, candidates: List[RetrievalCandidate]) -> List[RetrievalCandidate]:
        """Reference: retrieval.py:405-459: set c.rerank_score, return sorted by (rerank_score or 0)
        descending (stable); any failure returns the candidates unreranked."""
        if not candidates:
            return []
        try:
            scores = await self._rerank_batch_native(query, [c.child_id for c in candidates])
            for c, s in zip(candidates, scores):
                c.rerank_score = s
            return sorted(candidates, key=lambda c: c.rerank_score or 0, reverse=True)
        except Exception:
            return candidates

    # ---- K3: safety -------------------------------------------------------------------------------
    def _apply_safety(self, candidates: List[RetrievalCandidate], top_k: int
                      ) -> Tuple[List[RetrievalCandidate], bool, Optional[str], float]:
        """Reference: retrieval.py:461-495 -> (final, refused, reason, max_score)."""
        if not candidates:
            return [], True, "No candidates after reranking", 0.0
        eng = self._need_engine()
        dev = eng.device
        n = len(candidates)
        rrf = torch.tensor([c.rrf_score for c in candidates], dtype=torch.float64, device=dev)
        rer = torch.tensor([c.rerank_score if c.rerank_score is not None else 0.0 for c in candidates],
                           dtype=torch.float64, device=dev)
        has = torch.tensor([c.rerank_score is not None for c in candidates], dtype=torch.uint8, device=dev)
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        threshold = SETTINGS.rag2_safety_threshold
        keep, refused, mx = eng.safety(off, rrf, rer, has, threshold, SETTINGS.rag2_denoise_alpha, top_k)
        eng.sync()
        max_score = float(mx[0])
        if bool(refused[0]):
            return [], True, f"Max score {max_score:.2f} below threshold {threshold}", max_score
        kept = keep.tolist()
        return [c for c, k in zip(candidates, kept) if k], False, None, max_score

    # ---- batched entry (new) ----------------------------------------------------------------------
    def retrieve_batch(self, queries: Sequence[str], query_vectors: torch.Tensor,
                       keywords: Sequence[Sequence[str]], graph_ids: Optional[Sequence[Sequence[str]]] = None,
                       top_k: int = 100, k_sem: int = 100, k_lex: int = 50, weights: Optional[Dict[str, float]] = None,
                       collections: Optional[Sequence[Optional[str]]] = None) -> List[List[RetrievalCandidate]]:
        """B queries in one pass of K1 + K2 + K3 (no planner, no rerank): returns per query the fused
        candidates (rrf_score and channel ranks set), ties by chunk id.  query_vectors [B, D].
        collections: per query, the collection its semantic and lexical hits must belong to (None: any)."""
        from .pipeline import TripleHybridSearcher
        eng, ix = self._need_engine(), self.index
        s = TripleHybridSearcher(eng)
        s.has_dense = s.has_bm25 = True
        B = len(queries)
        Q = query_vectors.to(torch.float32)
        Q = (Q / Q.norm(dim=1, keepdim=True).clamp_min(1e-30)).to(torch.bfloat16).to(eng.device)
        terms = [[ix.vocab[w] for w in tokenize(" ".join(kw)) if w in ix.vocab][:32] for kw in keywords]
        qt, qo = pack_queries(terms, eng.device)
        g = None
        if graph_ids is not None:
            width = max(1, max(len(x) for x in graph_ids))
            g = torch.full((B, width), -1, dtype=torch.int64)
            for b, lst in enumerate(graph_ids):
                row = [ix.id_of[c] for c in lst if c in ix.id_of]
                g[b, :len(row)] = torch.tensor(row, dtype=torch.int64)
            g = g.to(eng.device)
        w = weights or {}
        wt = torch.tensor([[w.get("lexical", 0.7), w.get("semantic", 0.8), w.get("graph", 1.0)]] * B,
                          dtype=torch.float64, device=eng.device)
        want = None
        if collections is not None and any(c is not None for c in collections):
            want = torch.tensor([-1 if c is None else ix.tag_of.get(c, 0xfffe) for c in collections], dtype=torch.int32,
                                device=eng.device)
        out = s.search(Q, qt, qo, g, weights=wt, k_sem=min(k_sem, len(ix.rows)), k_lex=k_lex, top_k=top_k, want=want)
        eng.sync()
        ids, rrf, rk, cnt = out.ids.tolist(), out.rrf.tolist(), out.ranks.tolist(), out.count.tolist()
        res = []
        for b in range(B):
            lst = []
            for j in range(cnt[b]):
                d = ix.row_dict(ids[b][j])
                lst.append(RetrievalCandidate(**d, lexical_rank=rk[b][j][0] or None, semantic_rank=rk[b][j][1] or None,
                                              graph_rank=rk[b][j][2] or None, rrf_score=rrf[b][j]))
            res.append(lst)
        return res


async def retrieve(org_id: str, query: str, **kwargs: Any) -> RetrievalResult:
    """Convenience wrapper with the reference's signature (retrieval.py:498-505); the tenant's index and
    engine come through kwargs (`index=`, `embedder=`, ...), retrieval arguments through the rest."""
    ctor = {k: kwargs.pop(k) for k in ("embedder", "query_planner", "graph_enabled", "index", "engine",
                                       "graph_search", "token_encoder") if k in kwargs}
    return await GpuRAG2Retriever(org_id=org_id, **ctor).retrieve(query, **kwargs)
