"""Resident indexes for the hot path.

BM25Index builds the inverted index libthr's K2 kernel reads (layout in include/thr.h): postings in
term-major CSR order (doc ascending inside a term) stored as {uint32 doc, float32 impact}, plus a
skip table skip[t * n_blk + r] = first posting of term t whose doc lies in doc range r (ranges of
`blk_docs` docs, at most 2048: the skip granularity; the kernel accumulates spans of 15 such ranges).  Building is torch plumbing (sort / bincount / cumsum) and runs on whatever device
the inputs live on; it is the step before the hot path.  The BM25 formula is
the one oracle/bm25.py states; idf is always computed with numpy on the host so that both sides
use the same libm.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch


def bm25_idf(df: torch.Tensor, n_docs: int) -> torch.Tensor:
    d = df.detach().cpu().numpy().astype(np.float64)
    idf = np.log(1.0 + (n_docs - d + 0.5) / (d + 0.5)).astype(np.float32)
    return torch.from_numpy(idf)


def bm25_impacts(tf: torch.Tensor, dl: torch.Tensor, avgdl: float, k1: float, b: float) -> torch.Tensor:
    tf = tf.to(torch.float64)
    norm = k1 * (1.0 - b + b * dl.to(torch.float64) / float(avgdl))
    return (tf * (k1 + 1.0) / (tf + norm)).to(torch.float32)


@dataclass
class BM25Index:
    skip: torch.Tensor       # int64 [V * n_blk + 1]
    postings: torch.Tensor   # int32 [nnz + 2, 2]: {doc (u32 bits), impact (f32 bits)}, 16 B tail padding
    idf: torch.Tensor        # float32 [V]
    df: torch.Tensor         # int64 [V] (this shard)
    n_docs: int
    blk_docs: int
    V: int
    nnz: int
    k1: float = 1.2
    b: float = 0.75
    avgdl: float = 0.0

    @property
    def n_blk(self) -> int:
        return (self.n_docs + self.blk_docs - 1) // self.blk_docs

    @staticmethod
    def build(doc: torch.Tensor, term: torch.Tensor, tf: torch.Tensor, doc_len: torch.Tensor, V: int,
              blk_docs: int = 2048, k1: float = 1.2, b: float = 0.75, avgdl: Optional[float] = None,
              idf: Optional[torch.Tensor] = None, n_docs_global: Optional[int] = None) -> "BM25Index":
        """doc/term/tf: COO of (local doc id, term id, term frequency); doc_len [n_docs]."""
        dev = doc.device
        n_docs = int(doc_len.shape[0])
        R = int(blk_docs)
        n_blk = (n_docs + R - 1) // R
        doc = doc.to(torch.int64)
        term = term.to(torch.int64)
        if avgdl is None:
            avgdl = float(doc_len.to(torch.float64).mean().item())
        imp = bm25_impacts(tf, doc_len.to(dev)[doc], avgdl, k1, b)
        order = torch.argsort(term * n_docs + doc)          # (term, doc)
        counts = torch.bincount(term * n_blk + doc // R, minlength=V * n_blk)
        skip = torch.zeros(V * n_blk + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=skip[1:])
        nnz = int(doc.numel())
        post = torch.zeros((nnz + 2, 2), dtype=torch.int32, device=dev)
        post[:nnz, 0] = doc[order].to(torch.int32)  # u32 bit pattern; shards hold < 2^31 docs here
        post[:nnz, 1] = imp[order].view(torch.int32)
        df = torch.bincount(term, minlength=V)
        if idf is None:
            idf = bm25_idf(df, n_docs_global or n_docs)
        return BM25Index(skip, post, idf.to(torch.float32).to(dev), df, n_docs, R, V, nnz, k1, b, avgdl)

    @staticmethod
    def concat(parts: Sequence["BM25Index"], idf: Optional[torch.Tensor] = None,
               n_docs_global: Optional[int] = None) -> "BM25Index":
        """Merge indexes of consecutive doc ranges (every part but the last must hold a whole number
        of ranges; doc ids inside each part are local to the part and get rebased here).  Term-major
        order means a term's postings are the concatenation, in part order, of its per-part lists."""
        p0 = parts[0]
        dev = p0.postings.device
        V = p0.V
        for i, p in enumerate(parts):
            assert p.blk_docs == p0.blk_docs and p.V == V
            if i + 1 < len(parts):
                assert p.n_docs % p.blk_docs == 0, "only the last part may end with a partial range"
        counts = torch.cat([(p.skip[1:] - p.skip[:-1]).view(V, p.n_blk) for p in parts], dim=1)
        n_blk = counts.shape[1]
        skip = torch.zeros(V * n_blk + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts.reshape(-1), 0, out=skip[1:])
        del counts
        nnz = int(skip[-1].item())
        post = torch.zeros((nnz + 2, 2), dtype=torch.int32, device=dev)
        cursor = skip[:-1].view(V, n_blk)[:, 0].clone()     # where each term's next part goes
        df = torch.zeros(V, dtype=torch.int64, device=dev)
        ar_v = torch.arange(V, device=dev)
        base_doc = 0
        for p in parts:
            df_p = p.df.to(dev)
            term_of = torch.repeat_interleave(ar_v, df_p)
            local_start = p.skip[:-1].view(V, p.n_blk)[:, 0]
            dest = cursor[term_of] + (torch.arange(p.nnz, device=dev) - local_start[term_of])
            del term_of
            src = p.postings[:p.nnz]
            post[dest, 0] = src[:, 0] + base_doc
            post[dest, 1] = src[:, 1]
            del dest
            cursor += df_p
            df += df_p
            base_doc += p.n_docs
        if idf is None:
            idf = bm25_idf(df, n_docs_global or base_doc)
        return BM25Index(skip, post, idf.to(torch.float32).to(dev), df, base_doc, p0.blk_docs, V, nnz,
                         p0.k1, p0.b, p0.avgdl)

    def to(self, device) -> "BM25Index":
        return BM25Index(self.skip.to(device), self.postings.to(device), self.idf.to(device), self.df.to(device),
                         self.n_docs, self.blk_docs, self.V, self.nnz, self.k1, self.b, self.avgdl)

    def save(self, path) -> None:
        """Persist the index (one torch file; tensors are moved to the CPU).  SURVEY.md 8f row 1: the resident
        index needs its own checkpoint so that a retriever can start without re-reading the corpus."""
        torch.save({"format": "thr-bm25-v1", "skip": self.skip.cpu(), "postings": self.postings.cpu(),
                    "idf": self.idf.cpu(), "df": self.df.cpu(), "n_docs": self.n_docs, "blk_docs": self.blk_docs,
                    "V": self.V, "nnz": self.nnz, "k1": self.k1, "b": self.b, "avgdl": self.avgdl}, path)

    @staticmethod
    def load(path, device="cpu") -> "BM25Index":
        d = torch.load(path, map_location="cpu", weights_only=True)
        if d.get("format") != "thr-bm25-v1":
            raise ValueError(f"{path}: not a BM25Index file")
        return BM25Index(d["skip"], d["postings"], d["idf"], d["df"], int(d["n_docs"]), int(d["blk_docs"]),
                         int(d["V"]), int(d["nnz"]), float(d["k1"]), float(d["b"]), float(d["avgdl"])).to(device)

    def algorithmic_bytes(self, queries: Sequence[Sequence[int]]) -> int:
        """sum over queries and terms of df_t * 8 B of postings + 8 B of skip entry per
        (term, range)."""
        df = self.df.cpu()
        tot = 0
        for q in queries:
            for t in q:
                if 0 <= t < self.V:
                    tot += int(df[t]) * 8 + 8 * self.n_blk
        return tot


def pack_queries(queries: Sequence[Sequence[int]], device) -> tuple:
    """Ragged term-id lists -> (q_terms int32, q_off int32 [B+1]) on `device`."""
    off = [0]
    flat: List[int] = []
    for q in queries:
        flat.extend(int(t) for t in q)
        off.append(len(flat))
    return (torch.tensor(flat if flat else [0], dtype=torch.int32, device=device)[: max(len(flat), 1)],
            torch.tensor(off, dtype=torch.int32, device=device))
