"""The batched rerank stage (thr_rerank_rows -> thr_maxsim -> thr_rerank_finish) against the reference's own
`_rerank` ordering and `_apply_safety` restated on the CPU (oracle/fusion.apply_safety is pinned by reference-made
goldens; the ordering is Python's stable sorted(..., key=rerank_score or 0, reverse=True), retrieval.py:455)."""
import numpy as np
import pytest
import torch

from oracle import fusion as of
from oracle import maxsim as om
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.pipeline import SearchOutput, TripleHybridSearcher

pytestmark = pytest.mark.gpu


def _oracle_finish(ids, rrf, count, raw, Tq, thr, alpha, top_k):
    B, C = raw.shape
    res = []
    for q in range(B):
        n = min(int(count[q]), C)
        sc = [None if raw[q, j] == -np.inf else min(1.0, max(0.0, 0.5 * (float(raw[q, j]) / Tq + 1.0))) for j in range(n)]
        order = sorted(range(n), key=lambda j: sc[j] or 0, reverse=True)
        kept, refused, _, mx = of.apply_safety([float(rrf[q, j]) for j in order], [sc[j] for j in order], thr, alpha, top_k)
        res.append(([int(ids[q, j]) for j in order], [sc[j] for j in order], set(kept), refused, mx))
    return res


@pytest.mark.parametrize("thr,alpha,top_k", [(0.0, 0.0, 100), (0.6, 0.6, 5), (0.52, 0.99, 3), (0.99, 0.5, 10)])
def test_rerank_finish_matches_reference_semantics(engine, thr, alpha, top_k):
    dev = engine.device
    g = np.random.default_rng(3)
    B, C, stride, Tq = 37, 50, 64, 32
    ids = np.full((B, stride), -1, dtype=np.int64)
    rrf = np.zeros((B, stride))
    count = g.integers(0, stride + 1, B).astype(np.int32)
    count[0], count[1] = 0, 1
    raw = (g.standard_normal((B, C)) * 6 + 4).astype(np.float32)
    raw[g.random((B, C)) < 0.1] = -np.inf              # candidates nobody scored
    raw[g.random((B, C)) < 0.05] = -float(Tq)          # rerank_score exactly 0.0: falsy, falls back to rrf
    raw[3, :] = raw[3, 0]                              # exact ties: input order must survive
    raw[4, :] = -np.inf                                # no rerank score at all: safety runs on rrf
    for q in range(B):
        ids[q, :count[q]] = g.permutation(10_000)[:count[q]]
        rrf[q, :count[q]] = np.sort(g.random(count[q]) * 0.05)[::-1]
    rrf[4, :count[4]] += 0.7
    o = engine.rerank_finish(torch.from_numpy(ids).to(dev), torch.from_numpy(rrf).to(dev), torch.from_numpy(count).to(dev),
                             torch.from_numpy(raw).to(dev), Tq, thr, alpha, top_k)
    engine.sync()
    o_ids, o_rr, o_rrf, o_keep, o_n, o_ref, o_mx = (x.cpu().numpy() for x in o)
    want = _oracle_finish(ids, rrf, count, raw, Tq, thr, alpha, top_k)
    for q, (w_ids, w_sc, w_keep, w_ref, w_mx) in enumerate(want):
        n = len(w_ids)
        assert o_n[q] == n and list(o_ids[q, :n]) == w_ids and (o_ids[q, n:] == -1).all()
        assert [None if v < 0 else float(v).hex() for v in o_rr[q, :n]] == [None if v is None else float(v).hex() for v in w_sc]
        assert bool(o_ref[q]) == w_ref and float(o_mx[q]).hex() == float(w_mx).hex()
        assert set(np.nonzero(o_keep[q])[0].tolist()) == w_keep


def test_rerank_stage_sharded_rows_and_scores(engine):
    """Two 'ranks' emulated on one GPU: each scores only the candidates of its own chunk range; the MAX of the two
    score matrices equals the unsharded scores, and the stage's result equals MaxSim-oracle + finish-oracle."""
    dev = engine.device
    B, C, Tq, Td, n_chunks = 6, 40, 32, 64, 500
    Qt, Dt, _ = synth.maxsim_tokens(B, 1, Tq=Tq, Td=Td, device="cpu")
    g = torch.Generator().manual_seed(1)
    store = torch.randn((n_chunks, Td, 128), generator=g)
    store = (store / store.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    ids = torch.stack([torch.randperm(n_chunks, generator=g)[:64] for _ in range(B)]).to(torch.int64)
    count = torch.tensor([64, 40, 17, 0, 64, 5], dtype=torch.int32)
    for q in range(B):
        ids[q, count[q]:] = -1
    rrf = torch.rand((B, 64), generator=g, dtype=torch.float64).sort(dim=1, descending=True).values * 0.04
    raws = []
    for lo, hi in ((0, 256), (256, 500)):
        rows = engine.rerank_rows(ids.to(dev), count.to(dev), C, lo, hi)
        raws.append(engine.maxsim(Qt.to(dev), store[lo:hi].to(dev).contiguous(), rows))
    raw = torch.maximum(raws[0], raws[1])
    s = TripleHybridSearcher(engine)
    s.set_token_store(store.to(dev), 0, n_chunks)
    out = SearchOutput(ids.to(dev), rrf.to(dev), None, count.to(dev), None, None, None, None, None)
    out = s.rerank(out, Qt.to(dev), C, 0.5, 0.9, 10)
    engine.sync()
    cand = torch.where((torch.arange(64)[None, :] < count[:, None]), ids, torch.full_like(ids, -1))[:, :C]
    want_raw = om.maxsim(Qt.float().numpy(), store.float().numpy(), cand.numpy())
    got_raw = raw.cpu().numpy()
    m = cand.numpy() >= 0
    assert np.allclose(got_raw[m], want_raw[m], rtol=1e-3) and np.isneginf(got_raw[~m]).all()
    want = _oracle_finish(ids.numpy(), rrf.numpy(), count.numpy(), got_raw, Tq, 0.5, 0.9, 10)
    for q, (w_ids, w_sc, w_keep, w_ref, w_mx) in enumerate(want):
        n = len(w_ids)
        assert out.rr_ids[q, :n].tolist() == w_ids and bool(out.refused[q]) == w_ref
        assert set(np.nonzero(out.rr_keep[q].cpu().numpy())[0].tolist()) == w_keep
    # a repeating synthetic store: row = (id - lo) mod period
    rows = engine.rerank_rows(ids.to(dev), count.to(dev), C, 100, 400, period=7).cpu()
    exp = torch.where((cand >= 100) & (cand < 400), (cand - 100) % 7, torch.full_like(cand, -1))
    assert torch.equal(rows, exp)
    # row_off = id_lo % period: the row is a function of the global id
    rows = engine.rerank_rows(ids.to(dev), count.to(dev), C, 100, 400, period=7, row_off=100 % 7).cpu()
    assert torch.equal(rows, torch.where((cand >= 100) & (cand < 400), cand % 7, torch.full_like(cand, -1)))
