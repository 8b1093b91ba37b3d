"""Second, independent CPU restatement of MaxSim: one batched contraction (torch.einsum) in fp64, as BASELINE.md §3
names it.  TEST / BENCH INFRASTRUCTURE.  oracle/maxsim.py loops over (query, candidate) pairs with numpy; this one
shares no code with it.  parity unpinned — the reference has no MaxSim (src/voice_agent/retrieval/reranker.py:287-354
is an HTTP cross-encoder); the interface restated is one score per (query, document), input order.

    score(q, c) = sum_i max_j <Qtok[q, i], Dtok[c, j]>      (full lengths; fp64 on bf16-rounded inputs)
"""
from __future__ import annotations

import torch


def maxsim_einsum(Qtok: torch.Tensor, Dtok: torch.Tensor, cand: torch.Tensor, dtype=torch.float64,
                  chunk: int = 125) -> torch.Tensor:
    """Qtok [B,Tq,d], Dtok [n_docs,Td,d], cand [B,C] -> [B,C] (dtype)."""
    B, C = cand.shape
    out = torch.empty((B, C), dtype=dtype)
    Q = Qtok.to(dtype)
    for b in range(B):
        for s in range(0, C, chunk):
            D = Dtok[cand[b, s:s + chunk]].to(dtype)                 # [c, Td, d]
            sim = torch.einsum("qd,ctd->cqt", Q[b], D)               # [c, Tq, Td]
            out[b, s:s + chunk] = sim.amax(dim=2).sum(dim=1)
    return out
