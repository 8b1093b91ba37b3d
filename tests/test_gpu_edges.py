"""C-ABI edge cases and error behaviour on the device: empty batches, calls before an index is registered,
unsupported shapes, zero-length lists, k = 1 — errors are exceptions with the library's message, never a
silent fallback."""
import numpy as np
import pytest
import torch

from oracle import fusion as of
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200._lib import ThrError
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries

pytestmark = pytest.mark.gpu


def test_calls_before_index_set_raise():
    eng = Engine(0)
    try:
        with pytest.raises(ThrError, match="thr_dense_index_set first"):
            eng.dense_topk(torch.zeros((2, 64), dtype=torch.bfloat16, device=eng.device), 5)
        qt, qo = pack_queries([[1]], eng.device)
        with pytest.raises(ThrError, match="thr_bm25_index_set first"):
            eng.bm25_topk(qt, qo, 5)
    finally:
        eng.close()


def test_unsupported_shapes_raise(engine):
    dev = engine.device
    with pytest.raises(ThrError, match="multiple of 64"):
        engine.dense_index_set(torch.zeros((8, 96), dtype=torch.bfloat16, device=dev))
    doc, term, tf, L = synth.bm25_block_coo(0, 2000, V=100)
    idx = BM25Index.build(doc, term, tf, L, 100, blk_docs=4096).to(dev)
    with pytest.raises(ThrError, match="power of two"):
        engine.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V)
    with pytest.raises(ThrError, match="d = 64"):
        engine.maxsim(torch.zeros((1, 8, 64), dtype=torch.bfloat16, device=dev),
                      torch.zeros((4, 64, 64), dtype=torch.bfloat16, device=dev),
                      torch.zeros((1, 4), dtype=torch.int64, device=dev))
    with pytest.raises(TypeError):
        engine.dense_index_set(torch.zeros((8, 64), dtype=torch.float32, device=dev))
    with pytest.raises(ValueError):
        engine.dense_index_set(torch.zeros((8, 64), dtype=torch.bfloat16))  # host tensor


def test_empty_batches_and_k1(engine):
    dev = engine.device
    X = synth.dense_block(0, 2000, 64).to(dev)
    engine.dense_index_set(X)
    ids, sc, cnt, gap = engine.dense_topk(torch.zeros((0, 64), dtype=torch.bfloat16, device=dev), 5)
    assert ids.shape == (0, 5) and cnt.numel() == 0
    Q = synth.dense_queries(3, 64, X.cpu()).to(dev)
    ids, sc, cnt, _ = engine.dense_topk(Q, 1)
    engine.sync()
    ref = (Q.double() @ X.double().T)
    assert torch.equal(ids[:, 0], ref.argmax(dim=1)) and (cnt == 1).all()
    doc, term, tf, L = synth.bm25_block_coo(0, 2000, V=100)
    idx = BM25Index.build(doc, term, tf, L, 100, blk_docs=256).to(dev)
    engine.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V)
    qt, qo = pack_queries([], dev)
    ids, sc, cnt = engine.bm25_topk(qt, qo, 7)
    assert ids.shape == (0, 7)
    qt, qo = pack_queries([[], [], []], dev)
    ids, sc, cnt = engine.bm25_topk(qt, qo, 7)
    engine.sync()
    assert (cnt == 0).all() and (ids == -1).all() and (sc == 0).all()


def test_fuse_with_empty_and_absent_channels(engine):
    dev = engine.device
    B = 3
    off0 = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    sem = torch.tensor([5, 9, 5, 7], dtype=torch.int64, device=dev)
    sem_off = torch.tensor([0, 2, 2, 4], dtype=torch.int32, device=dev)
    w = torch.tensor([[0.7, 0.8, 1.0]] * B, dtype=torch.float64, device=dev)
    ids, rrf, rk, _, cnt = engine.fuse(0, B, [(torch.zeros(1, dtype=torch.int64, device=dev), off0, None),
                                              (sem, sem_off, None), None], w, top_k=4, max_out=4)
    engine.sync()
    assert cnt.tolist() == [2, 0, 2]
    rows = of.fuse(of.RAG2, [[], [5, 9], None])
    assert [(int(ids[0, i]), float(rrf[0, i]).hex()) for i in range(2)] == [(r["id"], r["rrf"].hex()) for r in rows]
    assert (ids[1] == -1).all()
    # safety on an empty candidate set refuses with max 0 (retrieval.py:466-467)
    keep, refused, mx = engine.safety(torch.tensor([0, 0], dtype=torch.int32, device=dev),
                                      torch.zeros(0, dtype=torch.float64, device=dev), None, None, 0.6, 0.6, 5)
    engine.sync()
    assert bool(refused[0]) and float(mx[0]) == 0.0


def test_merge_with_short_lists(engine):
    dev = engine.device
    sc = torch.tensor([[[3.0, 1.0, 0.0]], [[2.0, 2.0, 0.5]]], dtype=torch.float64, device=dev)   # [G=2,B=1,k=3]
    ids = torch.tensor([[[10, 11, 12]], [[7, 3, 4]]], dtype=torch.int64, device=dev)
    cnt = torch.tensor([[2], [3]], dtype=torch.int32, device=dev)
    m_sc, m_id, m_ct = engine.merge_topk(sc, ids, cnt, 4)
    engine.sync()
    assert m_id[0].tolist() == [10, 3, 7, 11] and m_sc[0].tolist() == [3.0, 2.0, 2.0, 1.0] and int(m_ct[0]) == 4


@pytest.mark.parametrize("G,B,k_sem,k_lex", [(3, 7, 10, 6), (8, 5, 100, 100), (2, 3, 100, 37)])
def test_fused_exchange_matches_the_torch_statement(engine, G, B, k_sem, k_lex):
    """K5 around the all-gather: thr_exchange_pack / thr_exchange_merge against pipeline.exchange_topk (the
    same step stated with torch ops, which the gloo test drives) with the oracle merge standing in for K5.
    G ranks are simulated in one process: every 'rank' packs its own lists, the messages are concatenated."""
    from oracle import merge as omg
    dev = engine.device
    g = torch.Generator().manual_seed(5)
    k = max(k_sem, k_lex)
    ranks = []
    for r in range(G):
        d_cnt = torch.randint(0, k_sem + 1, (B,), generator=g, dtype=torch.int32)
        l_cnt = torch.randint(0, k_lex + 1, (B,), generator=g, dtype=torch.int32)
        # a small value set makes cross-rank score ties common: order must fall back to the id
        d_sc = torch.randint(0, 8, (B, k_sem), generator=g).double().sort(dim=1, descending=True).values
        l_sc = torch.randint(1, 6, (B, k_lex), generator=g).float().sort(dim=1, descending=True).values
        d_ids = torch.randperm(B * k_sem * G, generator=g)[: B * k_sem].view(B, k_sem) * G + r
        l_ids = torch.randperm(B * k_lex * G, generator=g)[: B * k_lex].view(B, k_lex) * G + r
        # a rank's lists arrive sorted by (score desc, id asc), as thr_dense_topk / thr_bm25_topk write them
        # (thr_exchange_merge merges by rank and relies on it): order the ids inside every run of equal scores
        def by_id_within_ties(sc_, ids_):
            o = np.stack([np.lexsort((ids_[b].numpy(), -sc_[b].numpy())) for b in range(sc_.shape[0])])
            o = torch.from_numpy(o)
            return torch.gather(sc_, 1, o), torch.gather(ids_, 1, o)
        d_sc, d_ids = by_id_within_ties(d_sc, d_ids)
        l_sc, l_ids = by_id_within_ties(l_sc, l_ids)
        ranks.append((d_ids, d_sc, d_cnt, l_ids, l_sc, l_cnt))
    nbytes = engine.exchange_msg_bytes(B, k_sem, k_lex)
    assert nbytes == 2 * (2 * B * k * 8) + 2 * B * 4
    msgs = []
    for t in ranks:
        msg = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        engine.exchange_pack(*[x.to(dev) for x in t], msg)
        msgs.append(msg)
    got = engine.exchange_merge(torch.cat(msgs), G, B, k_sem, k_lex)
    engine.sync()
    # reference: unpack-free statement of the same merge on the host
    sc = np.full((G, 2 * B, k), -np.inf)
    ids = np.full((G, 2 * B, k), -1, dtype=np.int64)
    cnt = np.zeros((G, 2 * B), dtype=np.int32)
    for r, (d_ids, d_sc, d_cnt, l_ids, l_sc, l_cnt) in enumerate(ranks):
        sc[r, :B, :k_sem], ids[r, :B, :k_sem], cnt[r, :B] = d_sc.numpy(), d_ids.numpy(), d_cnt.numpy()
        sc[r, B:, :k_lex], ids[r, B:, :k_lex], cnt[r, B:] = l_sc.double().numpy(), l_ids.numpy(), l_cnt.numpy()
    w_s, w_i, w_c = omg.merge_topk(sc, ids, cnt, k)
    d_ids, d_sc, d_cnt, l_ids, l_sc, l_cnt = [x.cpu().numpy() for x in got]
    assert np.array_equal(d_cnt, np.minimum(w_c[:B], k_sem)) and np.array_equal(l_cnt, np.minimum(w_c[B:], k_lex))
    assert np.array_equal(d_ids, w_i[:B, :k_sem]) and np.array_equal(l_ids, w_i[B:, :k_lex])
    assert np.array_equal(d_sc, w_s[:B, :k_sem])
    want_l = np.where(w_i[B:, :k_lex] < 0, 0.0, w_s[B:, :k_lex]).astype(np.float32)
    assert np.array_equal(l_sc, want_l)
