"""Time the BM25 kernels alone: python scripts/bm25_probe.py [docs] [batch]"""
import sys, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
V = 100_000
eng = Engine(0); dev = eng.device
parts = []
G = 262144
for gb in range((N + G - 1) // G):
    rows = min(G, N - gb * G)
    doc, term, tf, L = synth.bm25_block_coo(gb, rows, V=V, device=dev)
    parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=200.0, n_docs_global=N))
idx = BM25Index.concat(parts) if len(parts) > 1 else parts[0]
eng.bm25_index_set(idx.skip, idx.postings, idx.idf, idx.n_docs, idx.blk_docs, idx.V)
qs = synth.bm25_queries(B, V=V)
qt, qo = pack_queries(qs, dev)
import os
cold = os.environ.get("THR_PROBE_COLD")          # flush L2 (write 512 MB) before every launch, like the step's dense kernel does
junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if cold else None
eng.prof_enable(True)
for _ in range(2):
    eng.bm25_topk(qt, qo, 100)
eng.sync(); eng.prof_reset()
for _ in range(5 if cold else 3):
    if cold:
        junk.fill_(1)
    eng.bm25_topk(qt, qo, 100)
p = eng.prof_read()
ms = p["bm25"][0] / p["bm25"][1]
by = idx.algorithmic_bytes(qs)
print(f"docs={N} B={B} bm25 {ms:.3f} ms  {by/ms/1e6:.0f} GB/s ({by/1e6:.0f} MB)  prep {p['bm25_prep'][0]/p['bm25_prep'][1]*2:.3f} ms/call")
