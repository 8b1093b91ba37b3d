"""Synthetic corpora and queries with fixed seeds, generated block-wise so that a
shard's content does not depend on how many GPUs hold the corpus.

torch generators differ between CPU and CUDA: "the same corpus" means generated once on one device
and copied.  Tests generate on the CPU and copy to the GPU; bench.py generates on the GPU and copies
the CPU-baseline sample back.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

DENSE_BLOCK_ROWS = 1_048_576
BM25_V = 100_000


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


DENSE_SUB_ROWS = 131_072  # generator granularity: any aligned range can be produced independently


def dense_rows(lo: int, hi: int, D: int, device="cpu") -> torch.Tensor:
    """Corpus rows [lo, hi): unit-norm Gaussian rows rounded to bf16.  Sub-block s (rows
    [s*DENSE_SUB_ROWS, +DENSE_SUB_ROWS)) is drawn from its own generator seeded 1234 + s, so a shard
    is the same whatever the GPU count."""
    out = torch.empty((hi - lo, D), dtype=torch.bfloat16, device=device)
    s0, s1 = lo // DENSE_SUB_ROWS, (hi - 1) // DENSE_SUB_ROWS
    for sb in range(s0, s1 + 1):
        base = sb * DENSE_SUB_ROWS
        a, b = max(lo, base), min(hi, base + DENSE_SUB_ROWS)
        g = _gen(device, 1234 + sb)
        x = torch.randn((b - base, D), generator=g, dtype=torch.float32, device=device)  # stream prefix
        x = x[a - base:]
        x = x / x.norm(dim=1, keepdim=True)
        out[a - lo:b - lo] = x.to(torch.bfloat16)
    return out


def dense_block(block: int, rows: int, D: int, device="cpu") -> torch.Tensor:
    """Rows [block*DENSE_BLOCK_ROWS, +rows) of the corpus."""
    return dense_rows(block * DENSE_BLOCK_ROWS, block * DENSE_BLOCK_ROWS + rows, D, device)


def dense_queries(B: int, D: int, X: torch.Tensor, noise: float = 0.5, n_plant: int = 0) -> torch.Tensor:
    """Half random unit vectors, half planted: normalize(X[j] + noise * unit_random), j < n_plant
    (default: all of X) — gives a clear top-1 and a realistic score spread."""
    dev = X.device
    g = _gen(dev, 4321)
    q = torch.randn((B, D), generator=g, dtype=torch.float32, device=dev)
    gj = _gen("cpu", 4322)
    j = torch.randint(0, n_plant or X.shape[0], (B,), generator=gj).to(dev)
    planted = X[j].float() + noise * q / q.norm(dim=1, keepdim=True)
    use = (torch.arange(B, device=dev) % 2 == 1).unsqueeze(1)
    q = torch.where(use, planted, q)
    q = q / q.norm(dim=1, keepdim=True)
    return q.to(torch.bfloat16)


def zipf_cdf(V: int = BM25_V, device="cpu") -> torch.Tensor:
    r = torch.arange(1, V + 1, dtype=torch.float64, device=device)
    p = 1.0 / r
    return torch.cumsum(p / p.sum(), 0)


def bm25_doc_lens(block: int, rows: int, device="cpu") -> torch.Tensor:
    """L_d = clip(Poisson(200), 16, 512) (reference child chunks are ~200 tokens, config.py:299)."""
    g = _gen(device, 2024 + block)
    lam = torch.full((rows,), 200.0, dtype=torch.float32, device=device)
    return torch.poisson(lam, generator=g).clamp_(16, 512).to(torch.int64)


def bm25_block_coo(block: int, rows: int, V: int = BM25_V, device="cpu", doc_base: int = 0
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(doc, term, tf, doc_len) for docs [doc_base, doc_base+rows): iid Zipf(s=1) tokens."""
    L = bm25_doc_lens(block, rows, device)
    g = _gen(device, 7_000_000 + block)
    total = int(L.sum().item())
    cdf = zipf_cdf(V, device)
    u = torch.rand((total,), generator=g, dtype=torch.float64, device=device)
    term = torch.searchsorted(cdf, u).clamp_(max=V - 1)
    del u
    doc = torch.repeat_interleave(torch.arange(rows, device=device), L)
    key = doc * V + term
    del doc, term
    key, _ = torch.sort(key)
    uk, counts = torch.unique_consecutive(key, return_counts=True)
    del key
    return (uk // V) + doc_base, uk % V, counts.to(torch.int32), L


def bm25_queries(B: int, V: int = BM25_V, min_rank: int = 100, seed: int = 2025) -> List[List[int]]:
    """q_len ~ U{3..8}; distinct terms, Zipf restricted to ranks >= min_rank (mimics stop-word removal)."""
    g = _gen("cpu", seed)
    lo = min(min_rank, V - 9)
    r = torch.arange(lo + 1, V + 1, dtype=torch.float64)
    p = 1.0 / r
    out = []
    lens = torch.randint(3, 9, (B,), generator=g)
    for i in range(B):
        t = torch.multinomial(p, int(lens[i]), replacement=False, generator=g) + lo
        out.append([int(x) for x in t])
    return out


def graph_lists(dense_ids: torch.Tensor, bm25_ids: torch.Tensor, n_total: int, length: int = 50,
                seed: int = 77) -> torch.Tensor:
    """Per query: length/2 ids sampled from that query's dense/BM25 results + the rest uniform, shuffled."""
    g = _gen("cpu", seed)
    B = dense_ids.shape[0]
    out = torch.empty((B, length), dtype=torch.int64)
    d, l = dense_ids.cpu(), bm25_ids.cpu()
    for b in range(B):
        pool = torch.unique(torch.cat([d[b][d[b] >= 0], l[b][l[b] >= 0]]))
        half = min(length // 2, pool.numel())
        pick = pool[torch.randperm(pool.numel(), generator=g)[:half]]
        chosen = set(int(x) for x in pick)
        rest = []
        while len(rest) < length - half:
            x = int(torch.randint(0, n_total, (1,), generator=g))
            if x not in chosen:
                chosen.add(x)
                rest.append(x)
        row = torch.cat([pick, torch.tensor(rest, dtype=torch.int64)])
        out[b] = row[torch.randperm(length, generator=g)]
    return out


def maxsim_tokens(B: int, C: int, Tq: int = 32, Td: int = 128, d: int = 128, device="cpu"
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Qtok [B,Tq,d], token store Dtok [B*C,Td,d] (unit-norm tokens, bf16), cand [B,C] = distinct rows."""
    g = _gen(device, 99)
    q = torch.randn((B, Tq, d), generator=g, dtype=torch.float32, device=device)
    q = (q / q.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    D = torch.empty((B * C, Td, d), dtype=torch.bfloat16, device=device)
    step = max(1, 65536 // Td)
    for s in range(0, B * C, step):
        n = min(step, B * C - s)
        x = torch.randn((n, Td, d), generator=g, dtype=torch.float32, device=device)
        D[s:s + n] = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    cand = torch.arange(B * C, dtype=torch.int64, device=device).view(B, C)
    return q, D, cand
