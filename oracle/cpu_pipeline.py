"""CPU port of the triple-hybrid step, timed as the reported baseline (bench.py cpu_baseline and
`--impl reference`).  TEST/BENCH INFRASTRUCTURE — never imported by the product package.

The reference's own path cannot run offline (Postgres RPCs + HTTP services), so this is kind "port",
stated the way BASELINE.md §3 names it:
  dense  = torch.matmul fp32 on the bf16-rounded matrices + torch.topk, all host threads;
  BM25   = scipy.sparse: Qs [B, V] (idf weights) @ W [V, n] (fp32 impacts), top-k from the sparse result rows — the
           independent statement of oracle/bm25_sparse.py, query batches spread over the host threads;
  fusion = the reference's OWN RAG2Retriever._fuse_rrf when the reference package is importable (PYTHONPATH reaches
           /root/reference/src), else the oracle's restatement of it (pinned bit-exact against reference-made goldens).
The corpus is held as shards of consecutive chunks (dense rows + the shard's postings); a step scores the batch
against every shard it is given and merges the shard lists, so the same code times one shard (the bounded
cpu_baseline sample) or the whole corpus (`--impl reference`).
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import torch

from . import fusion as of


@dataclass
class CpuShard:
    lo: int                    # first global chunk id
    X: torch.Tensor            # [n, D] fp32 (bf16-representable)
    W: sp.csr_matrix           # [V, n] fp32 impacts


def make_shard(lo: int, X: torch.Tensor, doc: np.ndarray, term: np.ndarray, tf: np.ndarray, doc_len: np.ndarray,
               V: int, avgdl: float, k1: float = 1.2, b: float = 0.75) -> CpuShard:
    tf64 = tf.astype(np.float64)
    imp = (tf64 * (k1 + 1.0) / (tf64 + k1 * (1.0 - b + b * doc_len.astype(np.float64)[doc] / avgdl))).astype(np.float32)
    W = sp.csr_matrix((imp, (term.astype(np.int32), doc.astype(np.int32))), shape=(V, int(doc_len.shape[0])))
    return CpuShard(lo, X.float().contiguous(), W)


def dense_topk_fp32(Q: torch.Tensor, X: torch.Tensor, k: int, chunk: int = 65536):
    """Q [B,D] fp32, X [n,D] fp32 -> (scores [B,k], ids [B,k]); chunked so S never exceeds B x chunk."""
    best_s = None
    best_i = None
    for s in range(0, X.shape[0], chunk):
        S = Q @ X[s:s + chunk].T
        v, i = torch.topk(S, min(k, S.shape[1]), dim=1)
        i = i + s
        if best_s is None:
            best_s, best_i = v, i
        else:
            cs, ci = torch.cat([best_s, v], 1), torch.cat([best_i, i], 1)
            v2, sel = torch.topk(cs, min(k, cs.shape[1]), dim=1)
            best_s, best_i = v2, torch.gather(ci, 1, sel)
    return best_s, best_i


def bm25_topk_sparse(W: sp.csr_matrix, idf: np.ndarray, queries: Sequence[Sequence[int]], k: int, threads: int):
    """scipy.sparse BM25 top-k: ids [B,k] (-1 padded), scores [B,k] fp32, counts [B]."""
    B, V = len(queries), W.shape[0]
    out_i = np.full((B, k), -1, dtype=np.int64)
    out_s = np.zeros((B, k), dtype=np.float32)
    out_c = np.zeros(B, dtype=np.int32)

    def work(ab):
        a, b_ = ab
        rows, cols, vals = [], [], []
        for q in range(a, b_):
            for t in queries[q]:
                t = int(t)
                if 0 <= t < V:
                    rows.append(q - a); cols.append(t); vals.append(idf[t])
        Qs = sp.csr_matrix((np.asarray(vals, dtype=np.float32), (rows, cols)), shape=(b_ - a, V))
        S = (Qs @ W).tocsr()
        for q in range(a, b_):
            lo_, hi_ = S.indptr[q - a], S.indptr[q - a + 1]
            data, idx = S.data[lo_:hi_], S.indices[lo_:hi_]
            if data.size == 0:
                continue
            kk = min(k, data.size)
            part = np.argpartition(-data, kk - 1)[:kk] if data.size > kk else np.arange(data.size)
            order = part[np.lexsort((idx[part], -data[part]))]
            out_i[q, :kk], out_s[q, :kk], out_c[q] = idx[order], data[order], kk

    n = max(1, min(threads, B))
    cuts = [B * i // n for i in range(n + 1)]
    parts = [(cuts[i], cuts[i + 1]) for i in range(n) if cuts[i + 1] > cuts[i]]
    if len(parts) == 1:
        work(parts[0])
    else:
        with ThreadPoolExecutor(max_workers=len(parts)) as ex:
            list(ex.map(work, parts))
    return out_i, out_s, out_c


def _reference_fuse():
    """The reference's own fusion when importable: (callable(lists) -> None, label)."""
    try:
        from voice_agent.rag2.retrieval import RAG2Retriever, RetrievalCandidate
    except Exception:
        return None
    r = RAG2Retriever.__new__(RAG2Retriever)      # _fuse_rrf uses no instance state (retrieval.py:358-376)

    def fuse(lex, sem, gr, top_k):
        merged: Dict[int, "RetrievalCandidate"] = {}
        for attr, ids in (("lexical_rank", lex), ("semantic_rank", sem), ("graph_rank", gr)):
            for rank, cid in enumerate(ids, 1):
                c = merged.get(cid)
                if c is None:
                    c = merged[cid] = RetrievalCandidate(child_id=str(cid), parent_id="", document_id="", text="", page=1,
                                                         modality="text")
                setattr(c, attr, rank)
        return r._fuse_rrf(list(merged.values()), {"lexical": 0.7, "semantic": 0.8, "graph": 1.0})[:top_k]
    return fuse


def step(Q: torch.Tensor, shards: List[CpuShard], idf: np.ndarray, queries: Sequence[Sequence[int]],
         graph: np.ndarray, k: int, top_k: int, threads: Optional[int] = None) -> Dict[str, float]:
    """One batch over the given shards; returns per-stage seconds (+ which fusion code ran)."""
    threads = threads or torch.get_num_threads()
    B = Q.shape[0]
    t0 = time.perf_counter()
    d_s, d_i = None, None
    for sh in shards:
        v, i = dense_topk_fp32(Q, sh.X, k)
        i = i + sh.lo
        if d_s is None:
            d_s, d_i = v, i
        else:
            cs, ci = torch.cat([d_s, v], 1), torch.cat([d_i, i], 1)
            d_s, sel = torch.topk(cs, min(k, cs.shape[1]), dim=1)
            d_i = torch.gather(ci, 1, sel)
    t1 = time.perf_counter()
    l_i = np.full((B, 0), -1, dtype=np.int64)
    l_s = np.zeros((B, 0), dtype=np.float32)
    for sh in shards:
        bi, bs, _ = bm25_topk_sparse(sh.W, idf, queries, k, threads)
        bi = np.where(bi >= 0, bi + sh.lo, -1)
        ci, cs = np.concatenate([l_i, bi], 1), np.concatenate([l_s, bs], 1)
        cs_key = np.where(ci >= 0, cs, -np.inf)
        order = np.lexsort((ci, -cs_key.astype(np.float64)), axis=1)[:, :k]
        l_i, l_s = np.take_along_axis(ci, order, 1), np.take_along_axis(cs, order, 1)
    t2 = time.perf_counter()
    ref_fuse = _reference_fuse()
    d_np = d_i.numpy()
    for b in range(B):
        lex = [int(x) for x in l_i[b] if x >= 0]
        if ref_fuse is not None:
            ref_fuse(lex, [int(x) for x in d_np[b]], [int(x) for x in graph[b]], top_k)
        else:
            of.fuse(of.RAG2, [lex, [int(x) for x in d_np[b]], [int(x) for x in graph[b]]], top_k=top_k,
                    tie_mode=of.TIE_CHUNK_ID)
    t3 = time.perf_counter()
    return {"dense": t1 - t0, "bm25": t2 - t1, "fuse": t3 - t2,
            "fusion_code": "reference RAG2Retriever._fuse_rrf" if ref_fuse is not None else "oracle restatement of _fuse_rrf"}
