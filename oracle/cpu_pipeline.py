"""CPU port of the triple-hybrid step, timed as the reported baseline (bench.py cpu_baseline and
`--impl reference`).  TEST/BENCH INFRASTRUCTURE — never imported by the product package.

The reference's own path cannot run offline (Postgres RPCs + HTTP services), so this is
kind "port": dense = torch.matmul fp32 on the bf16-rounded matrices + torch.topk on all host threads
(BASELINE.md §3), BM25 = the oracle's restatement with the batch spread over the host threads, fusion = the oracle's
restatement (one thread: pure Python, like the reference's own fusion).
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import bm25 as ob
from . import fusion as of


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


def bm25_topk_threads(index: ob.CsrIndex, queries: Sequence[Sequence[int]], k: int, threads: int):
    """The oracle's BM25 over the batch, queries spread over `threads` host threads (numpy releases the GIL
    in the gather / multiply / scatter / sort calls that dominate)."""
    if threads <= 1 or len(queries) < 2:
        return ob.bm25_topk(index, queries, k)
    from concurrent.futures import ThreadPoolExecutor
    n = len(queries)
    cuts = [n * i // threads for i in range(threads + 1)]
    parts = [(cuts[i], cuts[i + 1]) for i in range(threads) if cuts[i + 1] > cuts[i]]
    with ThreadPoolExecutor(max_workers=len(parts)) as ex:
        res = list(ex.map(lambda ab: ob.bm25_topk(index, queries[ab[0]:ab[1]], k), parts))
    return (np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]),
            np.concatenate([r[2] for r in res]))


def step(Q: torch.Tensor, X: torch.Tensor, index: ob.CsrIndex, queries: Sequence[Sequence[int]],
         graph: np.ndarray, k: int, top_k: int) -> Dict[str, float]:
    """One batch over the SAMPLE corpus; returns per-stage seconds."""
    t0 = time.perf_counter()
    d_sc, d_ids = dense_topk_fp32(Q, X, k)
    t1 = time.perf_counter()
    l_ids, l_sc, l_cnt = bm25_topk_threads(index, list(queries), k, torch.get_num_threads())
    t2 = time.perf_counter()
    d_ids = d_ids.numpy()
    for b in range(Q.shape[0]):
        of.fuse(of.RAG2, [[int(x) for x in l_ids[b, :l_cnt[b]]], [int(x) for x in d_ids[b]],
                          [int(x) for x in graph[b]]], top_k=top_k, tie_mode=of.TIE_CHUNK_ID)
    t3 = time.perf_counter()
    return {"dense": t1 - t0, "bm25": t2 - t1, "fuse": t3 - t2}
