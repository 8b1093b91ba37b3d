"""Is the slow 2nd timed step a host stall or a device stall?  python scripts/hiccup_probe.py [chunks]"""
import gc, sys, time, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
from triple_hybrid_rag_b200.pipeline import TripleHybridSearcher
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
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
def step():
    return s.search(Q, qt, qo, graph, k_sem=k, k_lex=k, top_k=k)
def trial(label, n=12, pre_sync=True, keep=True, gc_off=False, prof=False, idle_ms=0):
    eng.prof_enable(prof)
    for _ in range(30): step()
    if gc_off: gc.disable()
    if pre_sync: torch.cuda.synchronize()
    if idle_ms: time.sleep(idle_ms / 1e3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    th = [0.0] * (n + 1)
    ev[0].record(); th[0] = time.perf_counter()
    out = None
    for i in range(n):
        o = step()
        if keep: out = o
        ev[i + 1].record(); th[i + 1] = time.perf_counter()
    torch.cuda.synchronize()
    if gc_off: gc.enable()
    dev_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    host_ms = [(th[i + 1] - th[i]) * 1e3 for i in range(n)]
    print(f"{label:28s} dev  " + " ".join(f"{x:6.2f}" for x in dev_ms))
    print(f"{'':28s} host " + " ".join(f"{x:6.2f}" for x in host_ms), flush=True)
trial("baseline")
trial("baseline again")
trial("no pre-sync", pre_sync=False)
trial("gc off", gc_off=True)
trial("drop outputs", keep=False)
trial("prof on", prof=True)
trial("idle 50 ms after sync", idle_ms=50)
