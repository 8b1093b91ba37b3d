"""Parity at the EXACT BASELINE.json shapes (VERDICT r01 "parity gaps" item 1): the CUDA path against the CPU oracles
at cfg 2 (1M x 1536 dense, batch 256, top-100), cfg 3 (BM25 over a 1M-doc Zipf index, V = 100k, batch 1024, top-100)
and cfg 4 (MaxSim 64 x 1000 x 128 x 128 at Tq = 32 and Tq = 128).  The oracles are chunked so that they fit host
memory; the corpora are generated once (on the GPU, for speed) and copied to the host for the oracle, so both sides
see the same bytes.  Tolerances: ids bit-exact; dense / MaxSim scores within 1e-3 relative of fp64 (north_star);
BM25 fp32 score BITS equal to the fp32 oracle's, and within 1e-3 relative of the independent fp64 scipy statement."""
import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from oracle import bm25_sparse as obs
from oracle import dense as od
from oracle import maxsim as om
from oracle import maxsim_einsum as ome
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
from triple_hybrid_rag_b200.retriever import dense_error_bound

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def test_cfg2_dense_1m_x_1536_batch_256_top_100(engine):
    N, D, B, k = 1_000_000, 1536, 256, 100
    dev = engine.device
    X = synth.dense_rows(0, N, D, device=dev)
    Q = synth.dense_queries(B, D, X, n_plant=N // 8)
    engine.dense_index_set(X)
    ids, sc, cnt, gap = engine.dense_topk(Q, k)
    engine.sync()
    Xh, Qh = X.cpu(), Q.cpu()
    want_i = np.full((B, k), -1, dtype=np.int64)
    want_s = np.full((B, k), -np.inf)
    # the fp64 oracle, chunked over the corpus (a [256, 65536] fp64 score block at a time), merged by (score desc, id asc)
    for s in range(0, N, 131072):
        bi, bs = od.dense_topk(Qh.float().numpy(), Xh[s:s + 131072].float().numpy(), k, id_base=s)
        ci, cs = np.concatenate([want_i, bi], 1), np.concatenate([want_s, bs], 1)
        order = np.lexsort((ci, -cs), axis=1)[:, :k]
        want_i, want_s = np.take_along_axis(ci, order, 1), np.take_along_axis(cs, order, 1)
    ids, sc, gap = ids.cpu().numpy(), sc.cpu().numpy(), gap.cpu().numpy()
    assert (cnt.cpu().numpy() == k).all()
    assert np.array_equal(ids, want_i), f"{(ids != want_i).sum()} of {ids.size} ids differ"
    assert np.allclose(sc, want_s, rtol=REL_TOL, atol=1e-9)
    # the run certifies itself: every query's gap clears the fp32 accumulation bound
    assert (gap > dense_error_bound(D, 1.01, 1.01)).all(), float(gap.min())


def test_cfg3_bm25_1m_docs_v100k_batch_1024_top_100(engine):
    N, V, B, k = 1_000_000, 100_000, 1024, 100
    dev = engine.device
    parts, coo = [], []
    G = 262144
    lens_total = sum(int(synth.bm25_doc_lens(gb, min(G, N - gb * G), device=dev).sum().item()) for gb in range((N + G - 1) // G))
    avgdl = lens_total / N
    for gb in range((N + G - 1) // G):
        rows = min(G, N - gb * G)
        doc, term, tf, L = synth.bm25_block_coo(gb, rows, V=V, device=dev)
        parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=avgdl, idf=torch.zeros(V), n_docs_global=N))
        coo.append(((doc + gb * G).cpu().numpy(), term.cpu().numpy(), tf.cpu().numpy(), L.cpu().numpy()))
    from triple_hybrid_rag_b200.index import bm25_idf
    idf = bm25_idf(sum(p.df for p in parts), N)
    idx = BM25Index.concat(parts, idf=idf)
    idx.idf = idf.to(dev)
    del parts
    engine.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V)
    qs = synth.bm25_queries(B, V=V)
    qt, qo = pack_queries(qs, dev)
    ids, sc, cnt = engine.bm25_topk(qt, qo, k)
    engine.sync()
    ids, sc, cnt = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
    doc, term, tf, L = (np.concatenate([c[i] for c in coo]) for i in range(4))
    del coo
    # (1) the fp32 posting-list oracle, built from the COO independently of the product's index builder: bit-exact
    orc = ob.CsrIndex.from_coo(doc, term, tf, L, V, avgdl=avgdl)
    assert np.array_equal(orc.idf.view(np.uint32), idf.numpy().view(np.uint32))
    wi, ws, wc = ob.bm25_topk(orc, qs, k)
    assert np.array_equal(cnt, wc)
    assert np.array_equal(ids, wi), f"{(ids != wi).sum()} of {ids.size} ids differ"
    assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
    # (2) the independent fp64 scipy.sparse statement on a slice of the batch: scores within 1e-3, same ranking up to
    # fp32 rounding of near-ties
    spm = obs.SparseBM25(doc, term, tf, L, V, avgdl=avgdl)
    sub = list(range(0, B, 16))
    S = None
    for q in sub:
        S = spm.scores([qs[q]])[0]
        c = cnt[q]
        assert np.allclose(sc[q, :c], S[ids[q, :c]], rtol=REL_TOL)
        kth = S[ids[q, c - 1]]
        better = np.nonzero(S > kth * (1 + 1e-5))[0]
        assert set(better.tolist()) <= set(ids[q, :c].tolist())


@pytest.mark.parametrize("Tq", [32, 128])
def test_cfg4_maxsim_64_x_1000_x_128_x_128(engine, Tq):
    B, C, Td, d = 64, 1000, 128, 128
    dev = engine.device
    Qt, Dt, cand = synth.maxsim_tokens(B, C, Tq=Tq, Td=Td, d=d, device=dev)
    out = engine.maxsim(Qt, Dt, cand)
    engine.sync()
    got = out.cpu().numpy()
    Qh, Dh, ch = Qt.cpu().float(), Dt.cpu().float(), cand.cpu()
    want = ome.maxsim_einsum(Qh, Dh, ch).numpy()               # fp64, one contraction per (query, candidate block)
    assert np.allclose(got, want, rtol=REL_TOL)
    # the loop oracle on a slice (its Python loop over 64000 pairs is the slow one)
    w2 = om.maxsim(Qh[:4].numpy(), Dh.numpy(), ch[:4].numpy())
    assert np.allclose(want[:4], w2, rtol=1e-10) and np.allclose(got[:4], w2, rtol=REL_TOL)
    # ranking of the candidates by the kernel's scores == by the oracle's wherever the oracle separates them
    for b in range(0, B, 8):
        order = np.argsort(-got[b], kind="stable")
        assert (np.diff(want[b][order]) <= 1e-3 * np.abs(want[b][order][1:])).all()
