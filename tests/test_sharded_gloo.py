"""N > 1 host path on the CPU: two gloo ranks, each holding one chunk-id range of the corpus, run the
product's exchange step (pipeline.exchange_topk: pack -> all-gather -> merge) with the oracle standing in
for the kernels, and must reproduce the unsharded oracle bit for bit.  This pins the sharding claims of
DESIGN.md: contiguous id ranges aligned to BM25 doc ranges, ids rebased by the shard's id_base, GLOBAL idf /
avgdl so shard scores equal unsharded scores, ties across shards broken by chunk id."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bm25 as ob
from oracle import dense as od
from oracle import merge as omg
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.pipeline import exchange_topk, shard_bounds

N, D, V, B, K_SEM, K_LEX, ALIGN = 5000, 64, 600, 9, 40, 25, 1024


def _data():
    X = synth.dense_block(0, N, D).float().numpy()
    X[1024:1030] = X[7]                       # exact score ties straddling the shard boundary region
    Q = synth.dense_queries(B, D, torch.from_numpy(X)).float().numpy()
    doc, term, tf, L = (t.numpy() for t in synth.bm25_block_coo(0, N, V=V))
    queries = synth.bm25_queries(B, V=V, min_rank=20)
    queries[3] = []
    return X, Q, doc, term, tf, L, queries


def _merge_fn(g_sc, g_ids, g_cnt, k):
    s, i, c = omg.merge_topk(g_sc.numpy(), g_ids.numpy(), g_cnt.numpy(), k)
    return torch.from_numpy(s), torch.from_numpy(i), torch.from_numpy(c)


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Q, doc, term, tf, L, queries = _data()
        lo, hi = shard_bounds(N, world, ALIGN)[rank:rank + 2]
        # local channels on this shard: global idf / avgdl, ids rebased by id_base = lo
        di, ds = od.dense_topk(Q, X[lo:hi], K_SEM, id_base=lo)
        dcnt = np.full((B,), min(K_SEM, hi - lo), dtype=np.int32)
        df = np.bincount(term, minlength=V)
        idf = ob.idf_table(df, N)
        m = (doc >= lo) & (doc < hi)
        idx = ob.CsrIndex.from_coo(doc[m] - lo, term[m], tf[m], L[lo:hi], V, avgdl=float(L.mean()), idf=idf)
        li, ls, lc = ob.bm25_topk(idx, queries, K_LEX, id_base=lo)
        t = torch.from_numpy
        got = exchange_topk(dist.group.WORLD, world, B, K_SEM, K_LEX, t(di), t(ds), t(dcnt), t(li), t(ls), t(lc),
                            _merge_fn)
        if rank == 0:
            out.put([x.numpy() for x in got])
        else:
            out.put(None)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_rank_exchange_equals_unsharded_oracle():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=200) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m_di, m_ds, m_dc, m_li, m_ls, m_lc = next(r for r in res if r is not None)
    X, Q, doc, term, tf, L, queries = _data()
    wi, ws = od.dense_topk(Q, X, K_SEM)
    assert np.array_equal(m_di, wi) and np.array_equal(m_ds, ws) and (m_dc == K_SEM).all()
    whole = ob.CsrIndex.from_coo(doc, term, tf, L, V)
    bi, bs, bc = ob.bm25_topk(whole, queries, K_LEX)
    assert np.array_equal(m_lc, bc) and np.array_equal(m_li, bi)
    assert np.array_equal(m_ls.view(np.uint32), bs.view(np.uint32))


def test_shard_bounds():
    assert shard_bounds(10_000_000, 1) == [0, 10_000_000]
    b = shard_bounds(10_000_000, 8)
    assert b[0] == 0 and b[-1] == 10_000_000 and all(x % 16384 == 0 for x in b[1:-1])
    assert all(y > x for x, y in zip(b, b[1:]))
    assert shard_bounds(1000, 4, 1024) == [0, 0, 0, 0, 1000]   # tiny corpus: everything lands on the last rank
