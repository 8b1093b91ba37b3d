"""Drop-in boundary: GpuRAG2Retriever against (1) golden outputs of the reference's own
RAG2Retriever.retrieve() (tests/golden/make_retrieve_golden.py), (2) the known answers the reference's
tests assert (tests/test_rag2_triple_hybrid.py, tests/test_rag2_retrieval.py), (3) the CPU oracle on a
small resident corpus.  Channel methods are patched by name exactly as the reference's tests patch them."""
import asyncio
import gzip
import json
from pathlib import Path
from unittest.mock import AsyncMock, MagicMock, patch

import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from oracle import dense as od
from oracle import fusion as of
from oracle import maxsim as om
from triple_hybrid_rag_b200 import retriever as R
from triple_hybrid_rag_b200.retriever import (GpuRAG2Retriever, QueryPlan, ResidentIndex, RetrievalCandidate,
                                              RetrievalResult, SETTINGS)

pytestmark = pytest.mark.gpu
fh = float.fromhex
KNOBS = ("rag2_rerank_top_k", "rag2_safety_threshold", "rag2_denoise_alpha", "rag2_rerank_enabled",
         "rag2_graph_enabled")


@pytest.fixture()
def settings():
    old = {k: getattr(SETTINGS, k) for k in KNOBS}
    yield SETTINGS
    for k, v in old.items():
        setattr(SETTINGS, k, v)


def row(cid):
    return {"child_id": str(cid), "parent_id": f"p{cid % 7}", "document_id": f"d{cid}", "text": f"text {cid}",
            "page": 1 + cid % 5, "modality": "table" if cid % 11 == 0 else "text"}


def cand(cid, l=None, s=None, g=None, rrf=0.0, rerank=None):
    return RetrievalCandidate(child_id=str(cid), parent_id="p", document_id="d", text="t", page=1, modality="text",
                              lexical_rank=l, semantic_rank=s, graph_rank=g, rrf_score=rrf, rerank_score=rerank)


def test_retrieve_matches_reference_golden(engine, settings):
    """The reference's RAG2Retriever.retrieve() and ours on the same channel outputs / reranker scores:
    same contexts in the same order, bit-identical rrf, ranks, refusal flag, reason text and max score."""
    with gzip.open(Path(__file__).parent / "golden" / "retrieve_golden.json.gz", "rb") as fh_:
        cases = json.loads(fh_.read().decode())
    settings.rag2_graph_enabled = True
    n_ctx = 0
    for case in cases:
        for k, v in case["settings"].items():
            setattr(settings, k, v)
        lists, w = case["lists"], case["weights"]
        r = GpuRAG2Retriever(org_id="golden", query_planner=MagicMock(), graph_enabled=True, engine=engine)
        plan = QueryPlan(original_query="q", keywords=["k"] if lists[0] is not None else [], semantic_query_text="q",
                         requires_graph=lists[2] is not None,
                         cypher_query="MATCH (e) RETURN e" if lists[2] is not None else None,
                         weights={"lexical": w[0], "semantic": w[1], "graph": w[2]})
        scores = {k: fh(v) for k, v in case["rerank"].items()}

        async def native(query, documents):
            return [scores.get(d, 0.5) for d in documents]

        async def identity(c):
            return c

        async def go():
            with patch.object(r, "_lexical_search", new_callable=AsyncMock) as ml, \
                    patch.object(r, "_semantic_search", new_callable=AsyncMock) as ms, \
                    patch.object(r, "_graph_search", new_callable=AsyncMock) as mg, \
                    patch.object(r, "_expand_to_parents", side_effect=identity), \
                    patch.object(r, "_rerank_batch_native", side_effect=native):
                r.query_planner.plan_async = AsyncMock(return_value=plan)
                ml.return_value = [row(c) for c in (lists[0] or [])]
                ms.return_value = [row(c) for c in (lists[1] or [])]
                mg.return_value = [row(c) for c in (lists[2] or [])]
                return await r.retrieve("q", top_k=case["top_k"], skip_rerank=case["skip_rerank"])

        res = asyncio.run(go())
        o = case["out"]
        assert isinstance(res, RetrievalResult) and res.success == o["success"]
        assert (res.refused, res.refusal_reason, float(res.max_rerank_score).hex()) == (o["refused"], o["reason"], o["max"])
        assert sorted(res.timings) == o["timings"]
        got = [{"id": int(c.child_id), "rrf": c.rrf_score.hex(),
                "ranks": [c.lexical_rank or 0, c.semantic_rank or 0, c.graph_rank or 0],
                "rerank": None if c.rerank_score is None else float(c.rerank_score).hex(),
                "modality": c.modality, "page": c.page, "parent_id": c.parent_id} for c in res.contexts]
        assert got == o["contexts"]
        n_ctx += len(got)
    assert n_ctx > 100


def test_fuse_rrf_known_answers(engine):
    """tests/test_rag2_triple_hybrid.py:345-459 and tests/test_rag2_retrieval.py:121-184 of the reference."""
    r = GpuRAG2Retriever(org_id="test", engine=engine)
    w = {"lexical": 0.7, "semantic": 0.8, "graph": 1.0}
    c = cand(1, l=1, s=2, g=3)
    out = r._fuse_rrf([c], w)
    assert out[0] is c and abs(c.rrf_score - (0.7 / 61 + 0.8 / 62 + 1.0 / 63)) < 1e-3
    assert c.rrf_score.hex() == (0.0 + 0.7 / 61 + 0.8 / 62 + 1.0 / 63).hex()
    lex, sem, gr = cand("lex", l=1), cand("sem", s=1), cand("gr", g=1)
    out = r._fuse_rrf([lex, sem, gr], w)
    assert [x.child_id for x in out] == ["gr", "sem", "lex"]            # graph > semantic > lexical at rank 1
    multi, single = cand("multi", l=2, s=2, g=2), cand("single", g=1)
    assert r._fuse_rrf([single, multi], w)[0] is multi                    # three channels at 2 beat graph-only at 1
    c = cand(1, l=1, s=1)
    r._fuse_rrf([c], {"lexical": 2.0, "semantic": 0.5})
    assert abs(c.rrf_score - (2.0 / 61 + 0.5 / 61)) < 1e-9
    assert r._fuse_rrf([], w) == []
    # exact fp64 ties keep their input order (Python's stable sort): L10, S20, G40 all equal 0.01
    a, b, c3 = cand("L10", l=10), cand("S20", s=20), cand("G40", g=40)
    assert [x.child_id for x in r._fuse_rrf([b, c3, a], w)] == ["S20", "G40", "L10"]
    assert a.rrf_score == b.rrf_score == c3.rrf_score == fh("0x1.47ae147ae147bp-7")
    # 1-ulp near tie: S12 and G30 rank above L3
    out = r._fuse_rrf([cand("L3", l=3), cand("S12", s=12), cand("G30", g=30)], w)
    assert [x.child_id for x in out] == ["S12", "G30", "L3"]


def test_apply_safety_known_answers(engine, settings):
    """tests/test_rag2_triple_hybrid.py:795-896, tests/test_rag2_retrieval.py:190-238 of the reference."""
    r = GpuRAG2Retriever(org_id="test", engine=engine)
    settings.rag2_safety_threshold, settings.rag2_denoise_alpha = 0.6, 0.6
    cs = [cand(i, rrf=0.02, rerank=s) for i, s in enumerate([0.9, 0.7, 0.54, 0.5399999, 0.0, 0.3])]
    final, refused, reason, mx = r._apply_safety(cs, 5)
    assert [c.child_id for c in final] == ["0", "1", "2"] and not refused and reason is None and mx == 0.9
    final, refused, reason, mx = r._apply_safety([cand(0, rrf=0.02, rerank=0.59)], 5)
    assert (final, refused, reason, mx) == ([], True, "Max score 0.59 below threshold 0.6", 0.59)
    assert r._apply_safety([], 5) == ([], True, "No candidates after reranking", 0.0)
    settings.rag2_denoise_alpha = 0.5
    final, refused, _, _ = r._apply_safety([cand(0, rerank=0.95), cand(1, rerank=0.2)], 5)
    assert [c.child_id for c in final] == ["0"] and not refused
    settings.rag2_safety_threshold = 0.0
    final, refused, _, mx = r._apply_safety([cand(i, rrf=0.03 - 0.001 * i) for i in range(8)], 3)
    assert [c.child_id for c in final] == ["0", "1", "2"] and mx == 0.03


WORDS = [f"w{i}" for i in range(400)]


def _corpus(n, D, seed=5):
    g = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, len(WORDS) + 1)
    p /= p.sum()
    chunks = []
    for i in range(n):
        L = int(g.integers(12, 60))
        text = " ".join(g.choice(WORDS, size=L, p=p))
        chunks.append({"child_id": f"c{i}", "parent_id": f"p{i // 4}", "document_id": f"d{i // 16}", "text": text,
                       "page": 1 + i % 9, "modality": "text", "collection": "a" if i % 3 else "b"})
    emb = torch.from_numpy(g.standard_normal((n, D)).astype(np.float32))
    parents = {f"p{j}": {"text": f"parent text {j}", "section_heading": f"h{j}"} for j in range((n + 3) // 4)}
    return chunks, emb, parents


class _Embedder:
    def __init__(self, table):
        self.table = table

    def embed_query(self, text):
        return self.table[text]


def test_resident_index_end_to_end_vs_oracle(engine, settings):
    n, D, Td, d = 700, 128, 64, 128
    chunks, emb, parents = _corpus(n, D)
    g = torch.Generator().manual_seed(9)
    tok = torch.randn((n, Td, d), generator=g)
    tok = (tok / tok.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=1024, token_store=tok)
    qv = (emb[123] + 0.3 * torch.randn(D, generator=g))
    qtok = torch.randn((24, d), generator=g)
    r = GpuRAG2Retriever(org_id="t", embedder=_Embedder({"find me": qv.tolist()}), query_planner=MagicMock(),
                         index=ix, token_encoder=lambda q: qtok, graph_enabled=False, lexical_match="any")

    # semantic channel == oracle dense top-k on the bf16-rounded, normalised inputs
    rows = asyncio.run(r._semantic_search("find me", None, 50))
    qn = (qv / qv.norm()).to(torch.bfloat16).float().numpy()[None]
    wi, ws = od.dense_topk(qn, ix.X.float().cpu().numpy(), 50)
    assert [x["child_id"] for x in rows] == [f"c{i}" for i in wi[0]] and rows[0]["child_id"] == "c123"
    assert np.allclose([x["similarity"] for x in rows], ws[0], rtol=1e-3)
    assert set(rows[0]) >= {"child_id", "parent_id", "document_id", "text", "page", "modality"}

    # lexical channel == oracle BM25 over the same tokenisation
    kw = ["w3", "W17", "w150", "nosuchword"]
    rows = asyncio.run(r._lexical_search(kw, None, 50))
    d_l, t_l, f_l, lens = [], [], [], []
    for i, c in enumerate(chunks):
        toks = R.tokenize(c["text"])
        lens.append(len(toks))
        for t in set(toks):
            d_l.append(i); t_l.append(ix.vocab[t]); f_l.append(toks.count(t))
    orc = ob.CsrIndex.from_coo(np.array(d_l), np.array(t_l), np.array(f_l), np.array(lens), len(ix.vocab))
    bi, bs, bc = ob.bm25_topk(orc, [[ix.vocab["w3"], ix.vocab["w17"], ix.vocab["w150"]]], 50)
    assert [x["child_id"] for x in rows] == [f"c{i}" for i in bi[0, :bc[0]]]
    assert np.array_equal(np.array([x["rank"] for x in rows], dtype=np.float32), bs[0, :bc[0]])
    assert asyncio.run(r._lexical_search(["nosuchword"], None, 50)) == []
    # collection predicate, evaluated inside K2 / K1: the exact top-10 of collection "b"
    rows_b = asyncio.run(r._lexical_search(kw, "b", 10))
    assert rows_b and all(chunks[ix.id_of[x["child_id"]]]["collection"] == "b" for x in rows_b)
    tags = np.array([0 if c["collection"] == "a" else 1 for c in chunks])
    fi, fs, fc = ob.bm25_topk(orc, [[ix.vocab["w3"], ix.vocab["w17"], ix.vocab["w150"]]], 10, tags=tags, want=[1])
    assert [x["child_id"] for x in rows_b] == [f"c{i}" for i in fi[0, :fc[0]]]
    assert asyncio.run(r._lexical_search(kw, "no-such-collection", 10)) == []
    sem_b = asyncio.run(r._semantic_search("find me", "b", 10))
    assert len(sem_b) == 10 and all(chunks[ix.id_of[x["child_id"]]]["collection"] == "b" for x in sem_b)
    sem_all = asyncio.run(r._semantic_search("find me", None, 250))
    with pytest.raises(ValueError, match="252"):      # the reference RPC honours any p_limit; K1 says so instead of clamping
        asyncio.run(r._semantic_search("find me", None, 300))
    with pytest.raises(ValueError, match="256"):
        asyncio.run(r._lexical_search(kw, None, 300))
    assert [x["child_id"] for x in sem_b] == [x["child_id"] for x in sem_all
                                              if chunks[ix.id_of[x["child_id"]]]["collection"] == "b"][:10]

    # the reference's own predicate: every keyword must match (`tsv @@ plainto_tsquery`, 20260114_rag2_schema.sql:369);
    # a repeated keyword counts once; an unknown keyword empties the result
    r_all = GpuRAG2Retriever(org_id="t", embedder=_Embedder({"find me": qv.tolist()}), index=ix)
    assert r_all.lexical_match == "all"
    rows_all = asyncio.run(r_all._lexical_search(["w3", "w17", "W3"], None, 50))
    ai, as_, ac = ob.bm25_topk(orc, [[ix.vocab["w3"], ix.vocab["w17"]]], 50, require_all=True)
    assert ac[0] > 0 and [x["child_id"] for x in rows_all] == [f"c{i}" for i in ai[0, :ac[0]]]
    assert np.array_equal(np.array([x["rank"] for x in rows_all], dtype=np.float32), as_[0, :ac[0]])
    assert all({"w3", "w17"} <= set(R.tokenize(chunks[ix.id_of[x["child_id"]]]["text"])) for x in rows_all)
    assert asyncio.run(r_all._lexical_search(kw, None, 50)) == []

    # the reference-shaped reranker: document TEXTS in, one score in [0, 1] per text out (reranker.py:287-291)
    rr = R.GpuMaxSimReranker(ix, lambda q: qtok)
    texts = [chunks[5]["text"], "a text the index has never seen", parents["p3"]["text"]]
    sc = asyncio.run(rr._rerank_batch_native("find me", texts))
    qt_ = (qtok / qtok.norm(dim=-1, keepdim=True)).to(torch.bfloat16).float().numpy()[None]
    kids = [i for i, c in enumerate(chunks) if c["parent_id"] == "p3"]
    w5 = om.maxsim(qt_, tok.float().numpy(), np.array([[ix._rows_first(chunks[5]["text"])] + kids]))[0]
    to01 = lambda v: float(np.clip(0.5 * (v / 24 + 1.0), 0, 1))
    assert np.isclose(sc[0], to01(w5[0]), rtol=1e-3) and sc[1] == 0.5
    assert np.isclose(sc[2], max(to01(v) for v in w5[1:]), rtol=1e-3)
    assert R.Reranker is R.GpuMaxSimReranker

    # whole pipeline, MaxSim rerank included, against the oracle
    settings.rag2_rerank_top_k, settings.rag2_safety_threshold, settings.rag2_denoise_alpha = 20, 0.0, 0.0
    res = asyncio.run(r.retrieve("find me", top_k=20, skip_planning=True))
    assert res.success and not res.refused and len(res.contexts) == 20
    assert {"planning", "retrieval", "fusion", "expansion", "rerank", "safety"} <= set(res.timings)
    ids = [ix.id_of[c.child_id] for c in res.contexts]
    qt = (qtok / qtok.norm(dim=-1, keepdim=True)).to(torch.bfloat16).float().numpy()[None]
    want = om.maxsim(qt, tok.float().numpy(), np.array([ids]))[0]
    want01 = np.clip(0.5 * (want / 24 + 1.0), 0, 1)
    assert np.allclose([c.rerank_score for c in res.contexts], want01, rtol=1e-3)
    assert all(a.rerank_score >= b.rerank_score for a, b in zip(res.contexts, res.contexts[1:]))
    assert res.contexts[0].parent_text.startswith("parent text") and res.contexts[0].section_heading
    assert res.max_rerank_score == max(c.rerank_score for c in res.contexts)
    # default threshold 0.6: MaxSim of random tokens stays below it -> refusal is data, not an exception
    settings.rag2_safety_threshold = 0.99
    res = asyncio.run(r.retrieve("find me", top_k=5, skip_planning=True))
    assert res.success and res.refused and "below threshold" in res.refusal_reason and res.contexts == []


def test_retrieve_batch_matches_oracle_fusion(engine):
    n, D = 900, 64
    chunks, emb, parents = _corpus(n, D, seed=6)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=1024)
    r = GpuRAG2Retriever(org_id="t", index=ix, lexical_match="any")
    g = torch.Generator().manual_seed(3)
    Q = torch.randn((5, D), generator=g)
    kws = [["w1", "w9"], ["w2"], ["zzz"], ["w5", "w6", "w7"], []]
    graph = [[f"c{(7 * b + j) % n}" for j in range(10)] for b in range(5)]
    out = r.retrieve_batch(["q"] * 5, Q, kws, graph_ids=graph, top_k=40, k_sem=30, k_lex=20)
    Qn = (Q / Q.norm(dim=1, keepdim=True)).to(torch.bfloat16).float().numpy()
    wi, _ = od.dense_topk(Qn, ix.X.float().cpu().numpy(), 30)
    for b in range(5):
        terms = [ix.vocab[w] for w in kws[b] if w in ix.vocab]
        lex = []
        if terms:
            d = ix.bm25
            # oracle BM25 through the index's own postings is covered above; here reuse the GPU lexical channel
            lex = [ix.id_of[x["child_id"]] for x in asyncio.run(r._lexical_search(kws[b], None, 20))]
        rows = of.fuse(of.RAG2, [lex, [int(x) for x in wi[b]], [ix.id_of[c] for c in graph[b]]], top_k=40,
                       tie_mode=of.TIE_CHUNK_ID)
        got = [(ix.id_of[c.child_id], c.rrf_score.hex(), (c.lexical_rank or 0, c.semantic_rank or 0, c.graph_rank or 0))
               for c in out[b]]
        assert got == [(x["id"], x["rrf"].hex(), x["ranks"]) for x in rows]


def test_coalescing_front_end_equals_direct_batches(engine):
    """Concurrent single requests through CoalescingFrontEnd come back exactly as one direct retrieve_batch returns
    them (collections included: the predicate runs inside K1 / K2 per query)."""
    import functools
    from triple_hybrid_rag_b200.frontend import CoalescingFrontEnd
    n, D = 900, 64
    chunks, emb, parents = _corpus(n, D, seed=6)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=1024)
    r = GpuRAG2Retriever(org_id="t", index=ix, lexical_match="any")
    g = torch.Generator().manual_seed(4)
    nq = 12
    Q = torch.randn((nq, D), generator=g)
    kws = [[f"w{(3 * b) % 40}", f"w{(5 * b + 1) % 90}"] for b in range(nq)]
    colls = [None, "a", "b"] * 4
    direct = r.retrieve_batch(["q"] * nq, Q, kws, top_k=30, k_sem=25, k_lex=20, collections=colls)

    def batch_fn(queries, vectors, keywords, graph, collections):
        return r.retrieve_batch(queries, vectors, keywords, graph_ids=graph, top_k=30, k_sem=25, k_lex=20,
                                collections=collections)

    async def go():
        fe = CoalescingFrontEnd(batch_fn, max_batch=8, max_wait_ms=50)
        out = await asyncio.gather(*[fe.retrieve_candidates("q", Q[b], kws[b], collection=colls[b]) for b in range(nq)])
        return fe, out
    fe, out = asyncio.run(go())
    assert fe.batches == [8, 4]
    key = lambda lst: [(c.child_id, c.rrf_score.hex(), c.lexical_rank, c.semantic_rank) for c in lst]
    assert [key(x) for x in out] == [key(x) for x in direct]
    for b in range(nq):
        if colls[b] is not None:
            assert out[b] and all(chunks[ix.id_of[c.child_id]]["collection"] == colls[b] for c in out[b])


def test_front_end_serialises_overlapping_batches(engine):
    """overlapping batches must not share the handle's scratch.  Many small batches are in flight at once
    (max_batch 4, no waiting); every request must come back exactly as a direct call returns it."""
    from triple_hybrid_rag_b200.frontend import CoalescingFrontEnd
    n, D = 3000, 64
    chunks, emb, parents = _corpus(n, D, seed=12)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=256)
    r = GpuRAG2Retriever(org_id="t", index=ix, lexical_match="any")
    g = torch.Generator().manual_seed(5)
    nq = 64
    Q = torch.randn((nq, D), generator=g)
    kws = [[f"w{(7 * b) % 60}", f"w{(11 * b + 3) % 200}", f"w{b % 17}"] for b in range(nq)]
    direct = [r.retrieve_batch(["q"], Q[b:b + 1], [kws[b]], top_k=20, k_sem=20, k_lex=20)[0] for b in range(nq)]

    def batch_fn(queries, vectors, keywords, graph, collections):
        return r.retrieve_batch(queries, vectors, keywords, graph_ids=graph, top_k=20, k_sem=20, k_lex=20)

    async def go():
        fe = CoalescingFrontEnd(batch_fn, max_batch=4, max_wait_ms=0.0)
        out = await asyncio.gather(*[fe.retrieve_candidates("q", Q[b], kws[b]) for b in range(nq)])
        await fe.drain()
        fe.close()
        return fe, out
    fe, out = asyncio.run(go())
    assert len(fe.batches) >= nq // 4
    key = lambda lst: [(c.child_id, c.rrf_score.hex(), c.lexical_rank, c.semantic_rank) for c in lst]
    assert [key(x) for x in out] == [key(x) for x in direct]


def test_resident_index_save_load(engine, tmp_path):
    chunks, emb, parents = _corpus(300, 64, seed=8)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=256)
    r = GpuRAG2Retriever(org_id="t", embedder=_Embedder({"q": emb[5].tolist()}), index=ix, lexical_match="any")
    want_l = asyncio.run(r._lexical_search(["w2", "w11"], None, 20))
    want_s = asyncio.run(r._semantic_search("q", None, 20))
    ix.save(tmp_path / "ix.pt")
    ix2 = ResidentIndex.load(engine, tmp_path / "ix.pt")
    r2 = GpuRAG2Retriever(org_id="t", embedder=_Embedder({"q": emb[5].tolist()}), index=ix2, lexical_match="any")
    assert asyncio.run(r2._lexical_search(["w2", "w11"], None, 20)) == want_l
    assert asyncio.run(r2._semantic_search("q", None, 20)) == want_s


def test_no_engine_fails_loudly():
    r = GpuRAG2Retriever(org_id="t")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r._fuse_rrf([cand(1, l=1)], {})


def test_coalesced_tool_boundary_equals_single_calls(engine, settings):
    """concurrent `retrieve` calls share ONE K1 + K2 + K3 batch through CoalescedRetriever and each
    comes back exactly as its own RAG2Retriever.retrieve would (same contexts, bit-identical rrf / rerank scores,
    refusal and max score); the tool response carries the reference's keys (crm_knowledge.py:126-182)."""
    from triple_hybrid_rag_b200.tool import CoalescedRetriever, format_tool_response, search_knowledge_base_rag2
    n, D, Td = 1200, 64, 64
    chunks, emb, parents = _corpus(n, D, seed=21)
    g = torch.Generator().manual_seed(2)
    tok = torch.randn((n, Td, 128), generator=g)
    tok = (tok / tok.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    ix = ResidentIndex(engine, chunks, emb, parents, blk_docs=512, token_store=tok)
    queries = [f"w{3 * i % 50} w{7 * i % 120} w{i % 9}" for i in range(10)]
    table = {q: (emb[11 * i] + 0.4 * torch.randn(D, generator=g)).tolist() for i, q in enumerate(queries)}
    qtoks = {q: torch.randn((16, 128), generator=g) for q in queries}
    r = GpuRAG2Retriever(org_id="t", embedder=_Embedder(table), index=ix, token_encoder=lambda q: qtoks[q],
                         lexical_match="any")
    settings.rag2_rerank_top_k, settings.rag2_safety_threshold, settings.rag2_denoise_alpha = 20, 0.5, 0.95
    colls = [None, "a", "b", None, "a", None, None, "b", None, None]
    direct = [asyncio.run(r.retrieve(q, collection=c, top_k=5, skip_planning=True)) for q, c in zip(queries, colls)]

    async def go():
        cr = CoalescedRetriever(r, max_batch=16, max_wait_ms=30)
        out = await asyncio.gather(*[cr.retrieve(q, collection=c, top_k=5, skip_planning=True) for q, c in zip(queries, colls)])
        await cr.drain()
        cr.close()
        return cr, out
    cr, out = asyncio.run(go())
    assert cr._fe.batches == [10]                     # one launch served all ten calls
    key = lambda res: (res.refused, res.refusal_reason, float(res.max_rerank_score).hex(),
                       [(c.child_id, c.rrf_score.hex(), float(c.rerank_score).hex(), c.lexical_rank, c.semantic_rank, c.parent_text)
                        for c in res.contexts])
    assert [key(x) for x in out] == [key(x) for x in direct]
    assert any(not x.refused and x.contexts for x in out)
    d = format_tool_response(queries[0], colls[0], out[0])
    assert d["search_type"] == "rag2_triple_hybrid" and d["success"]
    if not out[0].refused:
        assert d["result_count"] == len(out[0].contexts) and set(d["results"][0]) >= {
            "chunk_id", "parent_id", "document_id", "content", "page", "modality", "relevance_rank", "similarity_score",
            "rerank_score", "is_table", "lexical_rank", "semantic_rank", "graph_rank"}
        assert d["results"][0]["content"].startswith("parent text") and {"planning", "retrieval"} <= set(d["timings_ms"])
    d2 = search_knowledge_base_rag2(queries[1], colls[1], 5, retriever=_SkipPlanning(r))
    assert d2["query"] == queries[1] and d2["category"] == "a" and ("refused" in d2 or d2["result_count"] <= 5)


class _SkipPlanning:
    """retrieve(query=, collection=, top_k=) as the tool calls it, with the fallback plan (no LLM planner here)."""
    def __init__(self, r):
        self.r = r

    def retrieve(self, query, collection=None, top_k=None):
        return self.r.retrieve(query, collection=collection, top_k=top_k, skip_planning=True)


def test_index_from_exported_table_rows(engine, settings):
    """the corpus as the reference stores it (rag_child_chunks / rag_documents / rag_parent_chunks rows,
    vectors rendered as text) -> export.from_tables -> ResidentIndex gives the same channel lists as the index built
    from the same chunks directly, and the collection predicate follows the DOCUMENT's collection."""
    from triple_hybrid_rag_b200.export import from_tables
    n, D = 600, 128
    chunks, emb, parents = _corpus(n, D, seed=11)
    doc_coll = {}
    for c in chunks:                      # one collection per document, as in rag_documents
        doc_coll.setdefault(c["document_id"], c["collection"])
        c["collection"] = doc_coll[c["document_id"]]
    child_rows = [{"id": c["child_id"], "parent_id": c["parent_id"], "document_id": c["document_id"], "org_id": "o1",
                   "index_in_parent": i % 4, "text": c["text"], "page": c["page"], "modality": c["modality"],
                   "content_hash": f"h{i}", "embedding_1024": "[" + ",".join(repr(float(x)) for x in emb[i]) + "]"}
                  for i, c in enumerate(chunks)]
    child_rows.append({**child_rows[0], "id": "foreign", "org_id": "o2"})
    doc_rows = [{"id": d, "org_id": "o1", "collection": col} for d, col in doc_coll.items()]
    parent_rows = [{"id": pid, "text": p["text"], "section_heading": p["section_heading"]} for pid, p in parents.items()]
    t_chunks, t_emb, t_parents, st = from_tables(child_rows, doc_rows, parent_rows, org_id="o1")
    assert st.children_kept == n and st.other_org == 1 and torch.equal(t_emb, emb)
    qv = emb[77] + 0.2 * torch.randn(D, generator=torch.Generator().manual_seed(1))
    out = []
    for ch, em, pa in ((chunks, emb, parents), (t_chunks, t_emb, t_parents)):
        ix = ResidentIndex(engine, ch, em, pa, blk_docs=256)
        r = GpuRAG2Retriever(org_id="o1", embedder=_Embedder({"q": qv.tolist()}), query_planner=MagicMock(), index=ix,
                             graph_enabled=False, lexical_match="any")
        sem = asyncio.run(r._semantic_search("q", "a", 40))
        lex = asyncio.run(r._lexical_search(["w2", "w9", "w40"], "b", 40))
        cands = asyncio.run(r._expand_to_parents([R.RetrievalCandidate(child_id=x["child_id"], parent_id=x["parent_id"],
                                                                       document_id=x["document_id"], text=x["text"],
                                                                       page=x["page"], modality=x["modality"])
                                                  for x in sem[:3]]))
        out.append((sem, lex, [(c.parent_text, c.section_heading) for c in cands]))
    assert out[0] == out[1] and len(out[0][0]) == 40 and out[0][1]
    assert all(doc_coll[x["document_id"]] == "a" for x in out[1][0]) and all(doc_coll[x["document_id"]] == "b" for x in out[1][1])
