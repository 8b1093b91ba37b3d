"""Where does a step's time go: host enqueue vs device?  python scripts/step_probe.py [chunks]"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
from triple_hybrid_rag_b200.pipeline import TripleHybridSearcher
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
D, B, k, V = 1536, 256, 100, 100_000
eng = Engine(0); dev = eng.device
s = TripleHybridSearcher(eng)
X = synth.dense_rows(0, N, D, device=dev); s.set_dense(X)
parts = []
G = 262144
for gb in range((N + G - 1) // G):
    rows = min(G, N - gb * G)
    doc, term, tf, L = synth.bm25_block_coo(gb, rows, V=V, device=dev)
    parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=2048, avgdl=200.0, n_docs_global=N))
idx = BM25Index.concat(parts) if len(parts) > 1 else parts[0]
del parts
s.set_bm25(idx)
Q = synth.dense_queries(B, D, X, n_plant=N // 8)
qt, qo = pack_queries(synth.bm25_queries(B, V=V), dev)
graph = torch.randint(0, N, (B, 50), device=dev)
def run(n, sync_each=False):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        s.search(Q, qt, qo, graph, k_sem=k, k_lex=k, top_k=k)
        if sync_each: torch.cuda.synchronize()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3
for _ in range(3): s.search(Q, qt, qo, graph, k_sem=k, k_lex=k, top_k=k)
for label, kw in (("async", {}), ("sync-each", {"sync_each": True}), ("async", {}), ("async+prof", {})):
    if label == "async+prof": eng.prof_enable(True)
    h, d, w = run(20, **kw)
    print(f"{label:10s} host-enqueue {h:7.3f} ms/step  device {d:7.3f} ms/step  wall {w:7.3f} ms/step")
