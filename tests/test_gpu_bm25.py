"""K2 parity: CUDA BM25 top-k vs the oracle — ids and fp32 scores bit-exact (the summation order is
part of the definition); fp32 scores within 1e-3 relative of fp64 accumulation."""
import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.index import BM25Index, pack_queries

pytestmark = pytest.mark.gpu


def _corpus(n_docs, V, blk_docs):
    doc, term, tf, L = synth.bm25_block_coo(0, n_docs, V=V)
    idx = BM25Index.build(doc, term, tf, L, V, blk_docs=blk_docs)
    orc = ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L.numpy(), V)
    return idx, orc


def _check(engine, idx, orc, queries, k, id_base=0, require_all=False):
    d = idx.to(engine.device)
    engine.bm25_index_set(d.skip, d.postings, d.idf, d.n_docs, d.blk_docs, d.V, id_base=id_base)
    qt, qo = pack_queries(queries, engine.device)
    ids, sc, cnt = engine.bm25_topk(qt, qo, k, require_all=require_all)
    engine.sync()
    wi, ws, wc = ob.bm25_topk(orc, queries, k, id_base=id_base, require_all=require_all)
    ids, sc, cnt = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, wc)
    assert np.array_equal(ids, wi), f"{(ids != wi).sum()} of {ids.size} ids differ"
    assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("n_docs,V,blk,B,k", [(20_000, 5_000, 1024, 64, 100), (50_000, 20_000, 2048, 128, 100),
                                              (3_000, 500, 256, 17, 10)])
def test_bm25_matches_oracle(engine, n_docs, V, blk, B, k):
    idx, orc = _corpus(n_docs, V, blk)
    qs = synth.bm25_queries(B, V=V, min_rank=min(100, V // 10))
    _check(engine, idx, orc, qs, k)


def test_bm25_heavy_terms_split_stages_and_edge_queries(engine):
    """Head terms (df ~ n_docs) fill the candidate list within one round: exercises the overflow roll-back
    with its serial redo and the compactions; also empty, unknown-term and duplicate-term queries."""
    idx, orc = _corpus(40_000, 2_000, 2048)
    qs = [[0, 1, 2, 3, 4, 5, 6, 7], [0], [1999], [], [5000, -3], [10, 10, 11], [3, 700, 1500], list(range(32))]
    _check(engine, idx, orc, qs, 256)
    _check(engine, idx, orc, qs, 1, id_base=123456789)


@pytest.mark.parametrize("blk", [256, 2048])
def test_bm25_and_semantics(engine, blk):
    """`tsv @@ plainto_tsquery` (20260114_rag2_schema.sql:369): every distinct query term must match.  Frequent
    terms so that intersections are non-empty; also repeated, unknown and single terms, and k larger than the
    number of matching docs."""
    idx, orc = _corpus(30_000, 3_000, blk)
    qs = [[0, 1], [2, 5, 9], [0, 0, 3], [1, 40, 200], [7], [3, 2999], [4, 5000], [], [10, 11, 12, 13, 14, 15],
          [0, 1, 2, 3, 4, 5, 6, 7], [100, 300, 20]]
    for k in (10, 100, 256):
        _check(engine, idx, orc, qs, k, require_all=True)
    _check(engine, idx, orc, synth.bm25_queries(32, V=3_000, min_rank=5), 50, require_all=True)


def test_bm25_rejects_negative_idf(engine):
    from triple_hybrid_rag_b200._lib import ThrError
    idx, _ = _corpus(3_000, 500, 256)
    d = idx.to(engine.device)
    bad = d.idf.clone()
    bad[7] = -1.0
    with pytest.raises(ThrError, match="idf"):
        engine.bm25_index_set(d.skip, d.postings, bad, d.n_docs, d.blk_docs, d.V)
    for v in (3.0e6, 1e-13, float("nan")):     # outside {0} U [2^-40, 2^20]: the accumulator's scaled range
        bad = d.idf.clone()
        bad[3] = v
        with pytest.raises(ThrError, match="idf"):
            engine.bm25_index_set(d.skip, d.postings, bad, d.n_docs, d.blk_docs, d.V)
    ok = d.idf.clone()
    ok[3] = 0.0                                # a zero weight is fine (the term contributes nothing)
    engine.bm25_index_set(d.skip, d.postings, ok, d.n_docs, d.blk_docs, d.V)
    engine.bm25_index_set(d.skip, d.postings, d.idf, d.n_docs, d.blk_docs, d.V)


def test_bm25_fp32_within_tolerance_of_fp64(engine):
    idx, orc = _corpus(20_000, 5_000, 1024)
    qs = synth.bm25_queries(16, V=5_000)
    _, s32, cnt = ob.bm25_topk(orc, qs, 100)
    for q, s, c in zip(qs, s32, cnt):
        ref = np.sort(ob.bm25_scores_fp64(orc, q))[::-1][:c]
        assert np.allclose(np.sort(s[:c])[::-1], ref, rtol=1e-3)


def test_bm25_exact_score_ties_resolve_by_doc_id(engine):
    """Thousands of docs with identical (tf, length) for the query terms score EXACTLY the same; the k
    best must be the smallest ids even though the kernel's warps walk different doc runs concurrently
    (a tying doc with a smaller id can arrive after the threshold was raised to that very score)."""
    n_docs, V = 60_000, 64
    rng = np.random.default_rng(11)
    doc_l, term_l, tf_l = [], [], []
    L = np.full(n_docs, 50, dtype=np.int64)
    for t, (step, tfv) in enumerate([(1, 2), (3, 1), (7, 3), (2, 2)]):      # regular patterns -> massive ties
        d = np.arange(t, n_docs, step)
        doc_l.append(d); term_l.append(np.full(d.size, t)); tf_l.append(np.full(d.size, tfv))
    d = np.sort(rng.choice(n_docs, 500, replace=False))                       # a rare term breaks some ties
    doc_l.append(d); term_l.append(np.full(d.size, 9)); tf_l.append(rng.integers(1, 4, d.size))
    doc, term, tf = (torch.from_numpy(np.concatenate(x)) for x in (doc_l, term_l, tf_l))
    Lt = torch.from_numpy(L)
    idx = BM25Index.build(doc, term, tf.to(torch.int32), Lt, V, blk_docs=2048)
    orc = ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L, V)
    qs = [[0], [1], [2, 3], [0, 1, 2, 3], [9, 0], [3, 9, 1]] * 3
    for k in (1, 10, 100, 256):
        _check(engine, idx, orc, qs, k)


def test_bm25_too_many_terms_is_an_error(engine):
    from triple_hybrid_rag_b200._lib import ThrError
    idx, _ = _corpus(3_000, 500, 256)
    d = idx.to(engine.device)
    engine.bm25_index_set(d.skip, d.postings, d.idf, d.n_docs, d.blk_docs, d.V)
    qt, qo = pack_queries([list(range(33)), [1, 2]], engine.device)
    engine.bm25_topk(qt, qo, 10)
    with pytest.raises(ThrError, match="32 terms"):
        engine.sync()
    engine.sync()  # the status word is cleared: the handle stays usable
