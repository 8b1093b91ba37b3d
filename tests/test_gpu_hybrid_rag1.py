"""RAG 1.0 twin: GpuHybridSearcher._vector_search / _bm25_search / _rrf_fusion / search on the same
kernels, at the halfvec(4000) width (padded to 4032), against the CPU oracles and the reference's RRF formula
(hybrid_search.py:478: 1.0 / (k + rank0 + 1), pinned by the RAG1 goldens in tests/test_gpu_fusion_dropin.py)."""
import asyncio

import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from oracle import dense as od
from oracle import fusion as of
from triple_hybrid_rag_b200.hybrid import GpuHybridSearcher, SearchConfig
from triple_hybrid_rag_b200.retriever import ResidentIndex, tokenize

pytestmark = pytest.mark.gpu
WORDS = [f"w{i}" for i in range(300)]


def test_rag1_hybrid_search_d4000(engine):
    n, D = 900, 4000
    g = np.random.default_rng(17)
    p = 1.0 / np.arange(1, len(WORDS) + 1)
    p /= p.sum()
    chunks = [{"child_id": f"k{i}", "parent_id": "p", "document_id": f"doc{i % 5}.pdf", "page": 1 + i % 4, "modality": "text",
               "text": " ".join(g.choice(WORDS, size=int(g.integers(10, 50)), p=p)), "collection": "faq" if i % 2 else "manual",
               "chunk_index": i, "title": f"t{i}"} for i in range(n)]
    emb = torch.from_numpy(g.standard_normal((n, D)).astype(np.float32))
    ix = ResidentIndex(engine, chunks, emb, blk_docs=256)
    assert ix.dim == 4000 and ix.X.shape[1] == 4032
    qv = emb[77] + 0.5 * torch.from_numpy(g.standard_normal(D).astype(np.float32))

    class Emb:
        async def embed_query(self, q):
            return qv.tolist(), None          # the reference returns (text_embedding, image_embedding)

    hs = GpuHybridSearcher("org", ix, SearchConfig(top_k_retrieve=40, top_k_final=15), embedder=Emb())
    vec = asyncio.run(hs._vector_search(qv.tolist()))
    qn = (qv / qv.norm()).to(torch.bfloat16).float().numpy()[None]
    wi, ws = od.dense_topk(qn, ix.X[:, :4000].float().cpu().numpy(), 40)
    assert [r.chunk_id for r in vec] == [f"k{i}" for i in wi[0]] and vec[0].chunk_id == "k77"
    assert np.allclose([r.similarity_score for r in vec], ws[0], rtol=1e-3) and vec[0].retrieval_method == "vector"

    lex = asyncio.run(hs._bm25_search("w2 w9"))
    d_l, t_l, f_l, lens = [], [], [], []
    for i, c in enumerate(chunks):
        toks = tokenize(c["text"])
        lens.append(len(toks))
        for t in set(toks):
            d_l.append(i); t_l.append(ix.vocab[t]); f_l.append(toks.count(t))
    orc = ob.CsrIndex.from_coo(np.array(d_l), np.array(t_l), np.array(f_l), np.array(lens), len(ix.vocab))
    bi, bs, bc = ob.bm25_topk(orc, [[ix.vocab["w2"], ix.vocab["w9"]]], 40, require_all=True)
    assert bc[0] > 0 and [r.chunk_id for r in lex] == [f"k{i}" for i in bi[0, :bc[0]]]
    assert np.array_equal(np.array([r.bm25_score for r in lex], dtype=np.float32), bs[0, :bc[0]])
    assert asyncio.run(hs._bm25_search("w2 nosuchword")) == []

    out = asyncio.run(hs.search("w2 w9"))
    rows = of.fuse(of.RAG1, [[int(i) for i in wi[0]], [int(i) for i in bi[0, :bc[0]]], None], rrf_k=60)[:15]
    assert [(r.chunk_id, r.rrf_score.hex()) for r in out] == [(f"k{x['id']}", x["rrf"].hex()) for x in rows]
    assert all(r.retrieval_method == "hybrid" for r in out)
    faq = asyncio.run(hs.search("w2 w9", category="faq"))
    assert faq and all(r.category == "faq" for r in faq)
