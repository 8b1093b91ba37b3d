"""Cost of the collection predicate inside K1 / K2 at full size: python scripts/tags_probe.py [chunks]
Times dense_topk / bm25_topk (batch 256, k = 100) without a filter, with 8 equal collections (every query
restricted to one of them) and with a 1 % collection."""
import sys, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
D, B, k, V = 1536, 256, 100, 100_000
eng = Engine(0); dev = eng.device
X = synth.dense_rows(0, N, D, device=dev); eng.dense_index_set(X)
parts = []
G = 262144
for gb in range((N + G - 1) // G):
    rows = min(G, N - gb * G)
    doc, term, tf, L = synth.bm25_block_coo(gb, rows, V=V, device=dev)
    parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=200.0, n_docs_global=N))
idx = BM25Index.concat(parts) if len(parts) > 1 else parts[0]
del parts
eng.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V)
Q = synth.dense_queries(B, D, X, n_plant=N // 8)
qt, qo = pack_queries(synth.bm25_queries(B, V=V), dev)
g = torch.Generator(device="cpu").manual_seed(1)
t8 = torch.randint(0, 8, (N,), generator=g, dtype=torch.int32)
t1 = torch.where(torch.rand((N,), generator=g) < 0.01, 1, 0).to(torch.int32)
cases = [("no filter", None, None),
         ("8 collections, one per query", t8.to(torch.uint16).to(dev), torch.randint(0, 8, (B,), generator=g, dtype=torch.int32).to(dev)),
         ("1 % collection", t1.to(torch.uint16).to(dev), torch.ones((B,), dtype=torch.int32, device=dev))]
eng.prof_enable(True)
for name, tags, want in cases:
    eng.dense_tags_set(tags); eng.bm25_tags_set(tags)
    for _ in range(2):
        eng.dense_topk(Q, k, want=want); eng.bm25_topk(qt, qo, k, want=want)
    eng.sync(); eng.prof_reset()
    for _ in range(3):
        ids, sc, cnt, _ = eng.dense_topk(Q, k, want=want); lids, lsc, lcnt = eng.bm25_topk(qt, qo, k, want=want)
    p = eng.prof_read()
    ms = lambda s: p[s][0] / max(p[s][1], 1)
    print(f"{name:32s} dense_score {ms('dense_score'):.3f} + seed {2 * ms('dense_seed'):.3f} + finalize {ms('dense_finalize'):.3f} ms   "
          f"bm25 {ms('bm25'):.3f} ms   (dense hits/query {cnt.float().mean().item():.0f}, lexical {lcnt.float().mean().item():.0f})", flush=True)
