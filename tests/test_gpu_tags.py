"""Collection predicate inside K1 / K2 (thr_dense_topk_tagged, thr_bm25_topk_tagged): the exact top-k of the
FILTERED corpus, bit-exact against the oracles restricted by the same tags — selective tags (fewer eligible
chunks than k), dominant tags, no filter (want < 0), and a corpus large enough for the dense seed pass."""
import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from oracle import dense as od
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200._lib import ThrError
from triple_hybrid_rag_b200.index import BM25Index, pack_queries

pytestmark = pytest.mark.gpu


def _tags(n, seed):
    """Skewed collections: tag 0 ~60 %, 1 ~30 %, 2 ~9 %, 3 ~1 %, tag 4 on exactly 30 rows, tag 5 on none."""
    g = np.random.default_rng(seed)
    t = g.choice(4, size=n, p=[0.6, 0.3, 0.09, 0.01]).astype(np.uint16)
    t[g.choice(n, size=30, replace=False)] = 4
    return t


def _check_dense(engine, N, D, B, k, seed):
    dev = engine.device
    X = synth.dense_block(seed, N, D)
    Q = synth.dense_queries(B, D, X[: N // 8])
    tags = _tags(N, seed)
    want = np.random.default_rng(seed + 1).integers(-1, 6, size=B).astype(np.int32)
    want[:6] = [-1, 0, 3, 4, 5, 2][: min(6, B)]
    engine.dense_index_set(X.to(dev))
    engine.dense_tags_set(torch.from_numpy(tags).to(dev))
    ids, sc, cnt, gap = engine.dense_topk(Q.to(dev), k, want=torch.from_numpy(want).to(dev))
    engine.sync()
    wi, ws = od.dense_topk(Q.float().numpy(), X.float().numpy(), k, tags=tags, want=want)
    ids, sc, cnt = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, (wi >= 0).sum(axis=1))
    assert np.array_equal(ids, wi), f"{(ids != wi).sum()} of {ids.size} ids differ"
    live = wi >= 0
    assert np.allclose(sc[live], ws[live], rtol=1e-12, atol=1e-12)
    for q in range(B):  # the predicate itself
        if want[q] >= 0:
            assert (tags[ids[q, : cnt[q]]] == want[q]).all()
    assert cnt[want == 5].sum() == 0 and (cnt[want == 4] == min(k, 30)).all()
    # the untagged call on the same handle is unaffected by the registered tags
    ids0, _, _, _ = engine.dense_topk(Q.to(dev), k)
    engine.sync()
    assert np.array_equal(ids0.cpu().numpy(), od.dense_topk(Q.float().numpy(), X.float().numpy(), k)[0])


@pytest.mark.parametrize("N,D,B,k", [(50_000, 128, 37, 50), (30_000, 64, 200, 100)])
def test_dense_tagged_matches_oracle(engine, N, D, B, k):
    _check_dense(engine, N, D, B, k, seed=3)


def test_dense_tagged_with_seed_pass(engine):
    """1M chunks: the seed pass runs (N >= 16 x prefix), so its thresholds too must come from eligible chunks."""
    _check_dense(engine, 1_000_000, 64, 130, 100, seed=5)


def test_bm25_tagged_matches_oracle(engine):
    dev = engine.device
    n_docs, V = 60_000, 3_000
    doc, term, tf, L = synth.bm25_block_coo(0, n_docs, V=V)
    idx = BM25Index.build(doc, term, tf, L, V, blk_docs=2048).to(dev)
    orc = ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L.numpy(), V)
    engine.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V, id_base=1000)
    tags = _tags(n_docs, 9)
    engine.bm25_tags_set(torch.from_numpy(tags).to(dev))
    qs = synth.bm25_queries(40, V=V, min_rank=30) + [[0, 1, 2, 3], [5]]
    want = np.random.default_rng(2).integers(-1, 6, size=len(qs)).astype(np.int32)
    want[:6] = [-1, 0, 3, 4, 5, 2]
    qt, qo = pack_queries(qs, dev)
    for k in (100, 7):
        ids, sc, cnt = engine.bm25_topk(qt, qo, k, want=torch.from_numpy(want).to(dev))
        engine.sync()
        wi, ws, wc = ob.bm25_topk(orc, qs, k, id_base=1000, tags=tags, want=want)
        assert np.array_equal(cnt.cpu().numpy(), wc)
        assert np.array_equal(ids.cpu().numpy(), wi)
        assert np.array_equal(sc.cpu().numpy().view(np.uint32), ws.view(np.uint32))
    ids0, _, _ = engine.bm25_topk(qt, qo, 100)
    engine.sync()
    assert np.array_equal(ids0.cpu().numpy(), ob.bm25_topk(orc, qs, 100, id_base=1000)[0])


def test_tagged_call_without_tags_raises(engine):
    dev = engine.device
    engine.dense_index_set(synth.dense_block(0, 2000, 64).to(dev))
    with pytest.raises(ThrError, match="thr_dense_tags_set first"):
        engine.dense_topk(torch.zeros((2, 64), dtype=torch.bfloat16, device=dev), 5,
                          want=torch.zeros((2,), dtype=torch.int32, device=dev))
