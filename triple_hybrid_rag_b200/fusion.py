"""Drop-ins for the reference's two fusion entry points, running on K3 (thr_fuse).

* GpuRRFFusion mirrors the library class RRFFusion
  (triple-hybrid-rag/src/triple_hybrid_rag/core/fusion.py:24-247): same constructor, same
  `fuse(lexical_results, semantic_results, graph_results, query_plan=None, top_k=None)`, same mutation of the
  first-seen result object per chunk id, same safety-threshold and conformal-denoising filters.  Scores are
  bit-identical to the reference's fp64 arithmetic (variant THR_FUSE_LIB, `w * (1.0 / (60 + rank))`,
  numpy's linear percentile), ties keep first-occurrence order like Python's stable sort.
* rag1_rrf_fusion mirrors HybridSearcher._rrf_fusion (src/voice_agent/retrieval/hybrid_search.py:460-501):
  unweighted `1.0 / (k + rank0 + 1)` over up to three lists, best raw scores kept per chunk.

Result objects are duck-typed (the reference's SearchResult dataclasses work as they are); chunk ids may be
UUIDs or strings.  Like the reference, feed ids that are unique inside one channel list.  There is no CPU
path: without an Engine (sm_100 device + libthr.so) construction fails.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .engine import Engine

RRF_K = 60  # fusion.py:22


def _default_config() -> SimpleNamespace:
    # defaults of triple_hybrid_rag.config.RAGConfig (config.py:138,191-200)
    return SimpleNamespace(rag_lexical_weight=0.7, rag_semantic_weight=0.8, rag_graph_weight=1.0,
                           rag_safety_threshold=0.6, rag_denoise_enabled=True, rag_denoise_alpha=0.6)


def _csr(rows: Sequence[Sequence], dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
    off = [0]
    flat: List = []
    for r in rows:
        flat.extend(r)
        off.append(len(flat))
    vals = torch.tensor(flat if flat else [0], dtype=dtype)[: len(flat)] if flat else torch.zeros((0,), dtype=dtype)
    if vals.numel() == 0:
        vals = torch.zeros((1,), dtype=dtype)  # a valid pointer for an empty list
    return vals.to(device), torch.tensor(off, dtype=torch.int32, device=device)


class GpuRRFFusion:
    """RRFFusion on the GPU.  `engine` defaults to a new Engine on device 0."""

    def __init__(self, config: Any = None, engine: Optional[Engine] = None):
        self.config = config or _default_config()
        self.engine = engine or Engine(0)
        self.default_weights = {"lexical": self.config.rag_lexical_weight,
                                "semantic": self.config.rag_semantic_weight,
                                "graph": self.config.rag_graph_weight}
        self.safety_threshold = self.config.rag_safety_threshold
        self.denoise_enabled = self.config.rag_denoise_enabled
        self.denoise_alpha = self.config.rag_denoise_alpha

    # ---- the reference's call ----------------------------------------------------------------
    def fuse(self, lexical_results: List[Any], semantic_results: List[Any], graph_results: List[Any],
             query_plan: Any = None, top_k: Optional[int] = None) -> List[Any]:
        return self.fuse_batch([(lexical_results, semantic_results, graph_results)],
                               [query_plan] if query_plan is not None else None, top_k)[0]

    def fuse_batch(self, batch: Sequence[Tuple[List[Any], List[Any], List[Any]]],
                   query_plans: Optional[Sequence[Any]] = None, top_k: Optional[int] = None) -> List[List[Any]]:
        """Many independent fusions in one launch (one CTA per query)."""
        eng, dev = self.engine, self.engine.device
        B = len(batch)
        if B == 0:
            return []
        firsts: List[List[Any]] = []          # per query: first-seen object per dense id
        per_ch = [[], [], []]                  # per channel: per query id lists / raw score lists
        per_raw = [[], [], []]
        fields = ("lexical_score", "semantic_score", "graph_score")
        weights = []
        for b, lists in enumerate(batch):
            index: Dict[str, int] = {}
            first: List[Any] = []
            for c in range(3):
                ids, raw = [], []
                for r in lists[c]:
                    key = str(r.chunk_id)
                    i = index.get(key)
                    if i is None:
                        i = index[key] = len(first)
                        first.append(r)
                    ids.append(i)
                    raw.append(float(getattr(r, fields[c])))
                per_ch[c].append(ids)
                per_raw[c].append(raw)
            firsts.append(first)
            plan = query_plans[b] if query_plans is not None else None
            w = plan.weights if plan else self.default_weights        # fusion.py:76
            weights.append([w.get("lexical", self.default_weights["lexical"]),
                            w.get("semantic", self.default_weights["semantic"]),
                            w.get("graph", self.default_weights["graph"])])
        lists_dev = []
        for c in range(3):
            ids, off = _csr(per_ch[c], torch.int64, dev)
            raw, _ = _csr(per_raw[c], torch.float64, dev)
            lists_dev.append((ids, off, raw))
        longest = max(len(f) for f in firsts)
        o_ids, o_rrf, _, o_raw, o_cnt = eng.fuse(
            _lib.FUSE_LIB, B, lists_dev, torch.tensor(weights, dtype=torch.float64, device=dev), rrf_k=RRF_K,
            safety_thr=float(self.safety_threshold), alpha=float(self.denoise_alpha),
            denoise=bool(self.denoise_enabled), top_k=int(top_k or 0), max_out=max(longest, 1),
            tie_mode=_lib.TIE_INSERTION, want_raw=True)
        eng.sync()
        o_ids, o_rrf, o_raw, o_cnt = o_ids.cpu(), o_rrf.cpu(), o_raw.cpu(), o_cnt.cpu()
        out: List[List[Any]] = []
        for b in range(B):
            in_ch = [set(per_ch[c][b]) for c in range(3)]
            res = []
            for j in range(int(o_cnt[b])):
                i = int(o_ids[b, j])
                r = firsts[b][i]
                r.rrf_score = float(o_rrf[b, j])
                r.lexical_score, r.semantic_score, r.graph_score = (float(x) for x in o_raw[b, j])
                r.final_score = r.rrf_score                            # fusion.py:137
                meta = getattr(r, "metadata", None)
                if isinstance(meta, dict):
                    meta["source_channels"] = [n for c, n in enumerate(("lexical", "semantic", "graph")) if i in in_ch[c]]
                res.append(r)
            out.append(res)
        return out

    def fuse_two_channels(self, results_a: List[Any], results_b: List[Any], weight_a: float = 1.0,
                          weight_b: float = 1.0, top_k: Optional[int] = None) -> List[Any]:
        """fusion.py:249-291: two lists, no filters; sets rrf_score and final_score."""
        eng, dev = self.engine, self.engine.device
        index: Dict[str, int] = {}
        first: List[Any] = []
        ids2 = []
        for lst in (results_a, results_b):
            ids = []
            for r in lst:
                key = str(r.chunk_id)
                i = index.get(key)
                if i is None:
                    i = index[key] = len(first)
                    first.append(r)
                ids.append(i)
            ids2.append(ids)
        lists_dev = [(*_csr([ids2[0]], torch.int64, dev), None), (*_csr([ids2[1]], torch.int64, dev), None), None]
        w = torch.tensor([[weight_a, weight_b, 0.0]], dtype=torch.float64, device=dev)
        o_ids, o_rrf, _, _, o_cnt = eng.fuse(_lib.FUSE_LIB, 1, lists_dev, w, rrf_k=RRF_K, safety_thr=0.0, alpha=0.0,
                                             denoise=False, top_k=int(top_k or 0), max_out=max(len(first), 1),
                                             tie_mode=_lib.TIE_INSERTION)
        eng.sync()
        o_ids, o_rrf = o_ids.cpu(), o_rrf.cpu()
        res = []
        for j in range(int(o_cnt[0])):
            r = first[int(o_ids[0, j])]
            r.rrf_score = r.final_score = float(o_rrf[0, j])
            res.append(r)
        return res

    @staticmethod
    def normalize_scores(results: List[Any], score_field: str = "final_score") -> List[Any]:
        """fusion.py:293-318 (host-side min-max over at most a few hundred objects; not on the hot path)."""
        if not results:
            return results
        scores = [getattr(r, score_field, 0.0) for r in results]
        lo, hi = min(scores), max(scores)
        for r in results:
            setattr(r, score_field, 1.0 if hi == lo else (getattr(r, score_field, 0.0) - lo) / (hi - lo))
        return results


def rag1_rrf_fusion(engine: Engine, results_lists: Sequence[List[Any]], k: int = 60) -> List[Any]:
    """HybridSearcher._rrf_fusion (hybrid_search.py:460-501) for up to three result lists: unweighted RRF with
    0-based ranks on K3 (variant THR_FUSE_RAG1); the best similarity_score / bm25_score per chunk are kept on
    the first-seen object and retrieval_method becomes "hybrid", as in the reference."""
    lists = [l for l in results_lists]
    if len(lists) > 3:
        raise ValueError("rag1_rrf_fusion: at most three result lists (vector, bm25, image)")
    dev = engine.device
    index: Dict[Any, int] = {}
    first: List[Any] = []
    dev_lists: List[Optional[tuple]] = []
    for lst in lists:
        ids = []
        for r in lst:
            i = index.get(r.chunk_id)
            if i is None:
                i = index[r.chunk_id] = len(first)
                first.append(r)
            else:                                         # hybrid_search.py:486-491
                keep = first[i]
                if r.similarity_score > keep.similarity_score:
                    keep.similarity_score = r.similarity_score
                if r.bm25_score > keep.bm25_score:
                    keep.bm25_score = r.bm25_score
            ids.append(i)
        dev_lists.append((*_csr([ids], torch.int64, dev), None))
    dev_lists += [None] * (3 - len(dev_lists))
    if not first:
        return []
    w = torch.ones((1, 3), dtype=torch.float64, device=dev)
    o_ids, o_rrf, _, _, o_cnt = engine.fuse(_lib.FUSE_RAG1, 1, dev_lists, w, rrf_k=int(k), max_out=len(first),
                                            tie_mode=_lib.TIE_INSERTION)
    engine.sync()
    o_ids, o_rrf = o_ids.cpu(), o_rrf.cpu()
    out = []
    for j in range(int(o_cnt[0])):
        r = first[int(o_ids[0, j])]
        r.rrf_score = float(o_rrf[0, j])
        r.retrieval_method = "hybrid"
        out.append(r)
    return out
