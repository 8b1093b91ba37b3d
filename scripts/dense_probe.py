"""Time the dense kernels alone (no BM25/fusion): python scripts/dense_probe.py [chunks] [dim] [batch]"""
import sys, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
eng = Engine(0)
X = synth.dense_rows(0, N, D, device=eng.device)
Q = synth.dense_queries(B, D, X)
eng.dense_index_set(X)
eng.prof_enable(True)
for _ in range(3):
    eng.dense_topk(Q, 100, 28)
eng.sync(); eng.prof_reset()
for _ in range(5):
    eng.dense_topk(Q, 100, 28)
p = eng.prof_read()
ms = p["dense_score"][0] / p["dense_score"][1]
seed = p.get("dense_seed", (0.0, 1))
print(f"N={N} D={D} B={B} dense_score {ms:.3f} ms  {2*B*N*D/ms/1e9:.1f} TFLOP/s  {N*D*2/ms/1e6:.0f} GB/s  "
      f"finalize {p['dense_finalize'][0]/p['dense_finalize'][1]:.3f} ms  seed(total per call) {seed[0]/5:.3f} ms")
