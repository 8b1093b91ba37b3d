"""Full-size properties (no CPU oracle can finish at these sizes): on a multi-million-chunk corpus the
kernels must be self-consistent — returned scores equal an independent fp64 / fp32 recomputation for the
returned ids, lists are ordered by (score desc, id asc), planted queries find their chunk, and scoring the
corpus in two shards + K5 merge reproduces the unsharded answer bit for bit (the multi-GPU path's claim)."""
import os

import pytest
import torch

from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.index import BM25Index, bm25_idf, pack_queries

pytestmark = pytest.mark.gpu
N = int(os.environ.get("THR_TEST_FULL_N", 4_000_000))
D, B, K, V = 1536, 256, 100, 100_000


def _ordered(scores, ids):
    s0, s1, i0, i1 = scores[:, :-1], scores[:, 1:], ids[:, :-1], ids[:, 1:]
    return bool(((s0 > s1) | ((s0 == s1) & (i0 < i1))).all())


def test_dense_full_size_properties(engine):
    dev = engine.device
    X = synth.dense_rows(0, N, D, device=dev)
    Q = synth.dense_queries(B, D, X, n_plant=N // 8)
    engine.dense_index_set(X)
    ids, sc, cnt, gap = engine.dense_topk(Q, K)
    engine.sync()
    assert (cnt == K).all() and (ids >= 0).all() and (ids < N).all()
    assert _ordered(sc, ids)
    assert all(torch.unique(r).numel() == K for r in ids[:16])
    # scores == fp64 dot of the bf16 inputs for the returned rows (independent torch recomputation)
    ref = torch.einsum("bkd,bd->bk", X[ids.reshape(-1)].view(B, K, D).double(), Q.double())
    assert torch.allclose(sc, ref, rtol=1e-12, atol=1e-12)
    assert (gap > 0).all()                                     # exactness certificate
    # planted queries (odd rows): the planted chunk is the best match
    gj = torch.Generator().manual_seed(4322)
    j = torch.randint(0, N // 8, (B,), generator=gj).to(dev)
    assert (ids[1::2, 0] == j[1::2]).all()
    # two shards + K5 merge == unsharded, bit for bit
    h = (N // 2) // 16384 * 16384
    parts = []
    for lo, hi in ((0, h), (h, N)):
        engine.dense_index_set(X[lo:hi], id_base=lo)
        parts.append(engine.dense_topk(Q, K))
    engine.sync()
    g_sc = torch.stack([p[1] for p in parts])
    g_id = torch.stack([p[0] for p in parts])
    g_ct = torch.stack([p[2] for p in parts])
    m_sc, m_id, m_ct = engine.merge_topk(g_sc, g_id, g_ct, K)
    engine.sync()
    assert torch.equal(m_id, ids) and torch.equal(m_sc, sc) and (m_ct == K).all()


def test_bm25_full_size_properties(engine):
    dev = engine.device
    G = 262144
    n = N // G * G
    parts = []
    for gb in range(n // G):
        doc, term, tf, L = synth.bm25_block_coo(gb, G, V=V, device=dev)
        parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=200.0, idf=torch.zeros(V)))
    df = sum(p.df for p in parts)
    idf = bm25_idf(df, n)
    whole = BM25Index.concat(parts, idf=idf)
    queries = synth.bm25_queries(B, V=V)
    qt, qo = pack_queries(queries, dev)
    engine.bm25_index_set(whole.skip, whole.postings, whole.idf, whole.n_docs, whole.blk_docs, V)
    ids, sc, cnt = engine.bm25_topk(qt, qo, K)
    engine.sync()
    assert (cnt == K).all() and _ordered(sc, ids) and (ids < n).all()
    # recompute the fp32 scores of the returned docs from the postings, terms in query order
    post = whole.postings
    for q in range(0, B, 37):
        acc = torch.zeros(K, dtype=torch.float32, device=dev)
        for t in queries[q]:
            lo, hi = int(whole.skip[t * whole.n_blk]), int(whole.skip[(t + 1) * whole.n_blk])
            d = post[lo:hi, 0].to(torch.int64)
            imp = post[lo:hi, 1].view(torch.float32)
            pos = torch.searchsorted(d, ids[q])
            pos = pos.clamp(max=max(hi - lo - 1, 0))
            hit = d[pos] == ids[q]
            acc = torch.where(hit, acc + (idf[t].to(dev) * imp[pos]), acc)   # fp32 multiply, fp32 add
        assert torch.equal(acc.view(torch.int32), sc[q].view(torch.int32))
    # two shards with the GLOBAL idf + K5 merge == unsharded, bit for bit
    half = len(parts) // 2
    res = []
    for sub, base in ((parts[:half], 0), (parts[half:], half * G)):
        ix = BM25Index.concat(sub, idf=idf)
        engine.bm25_index_set(ix.skip, ix.postings, ix.idf, ix.n_docs, ix.blk_docs, V, id_base=base)
        res.append(engine.bm25_topk(qt, qo, K))
        engine.sync()
    g_sc = torch.stack([r[1].double() for r in res])
    g_id = torch.stack([r[0] for r in res])
    g_ct = torch.stack([r[2] for r in res])
    m_sc, m_id, m_ct = engine.merge_topk(g_sc, g_id, g_ct, K)
    engine.sync()
    assert torch.equal(m_id, ids) and torch.equal(m_sc.float().view(torch.int32), sc.view(torch.int32))


def test_two_stream_search_equals_single_stream(engine):
    """TripleHybridSearcher.overlap (K1's and K2's kernel chains on two streams, one scratch arena per channel) must not
    change a bit of the step's output: fused lists, both channel lists, counts and the dense certificate, over several
    steps with different query batches, on a corpus large enough for the seed pass and multi-range BM25 units."""
    from triple_hybrid_rag_b200.pipeline import TripleHybridSearcher
    dev = engine.device
    n, G = 1_048_576, 262144
    X = synth.dense_rows(0, n, D, device=dev)
    parts = []
    for gb in range(n // G):
        doc, term, tf, L = synth.bm25_block_coo(gb, G, V=V, device=dev)
        parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=200.0, idf=torch.zeros(V)))
    idf = bm25_idf(sum(p.df for p in parts), n)
    index = BM25Index.concat(parts, idf=idf)
    s = TripleHybridSearcher(engine)
    s.set_dense(X)
    s.set_bm25(index)
    outs = {False: [], True: []}
    for step in range(3):
        Q = synth.dense_queries(B, D, X, n_plant=n // (2 + step))
        qt, qo = pack_queries(synth.bm25_queries(B, V=V, seed=100 + step), dev)
        for ov in (False, True, True, False):
            s.overlap = ov
            o = s.search(Q, qt, qo, None, k_sem=K, k_lex=K, top_k=K)
            engine.sync()
            outs[ov].append([t.clone() for t in (o.ids, o.rrf, o.count, o.sem_ids, o.sem_scores, o.lex_ids, o.lex_scores,
                                                 o.lex_count, o.gap)])
    s.overlap = False
    assert len(outs[True]) == len(outs[False]) == 6
    for a, b in zip(outs[False], outs[True]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert int(outs[True][0][2].min()) == K
