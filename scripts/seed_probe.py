"""Dense seed pass length, settings interleaved in ONE process (box state and clocks shared):
python scripts/seed_probe.py [chunks]"""
import os, sys, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
D, B = 1536, 256
eng = Engine(0)
X = synth.dense_rows(0, N, D, device=eng.device)
Q = synth.dense_queries(B, D, X)
eng.dense_index_set(X)
eng.prof_enable(True)
for _ in range(10):
    eng.dense_topk(Q, 100, 28)
eng.sync()
TILES = tuple(int(x) for x in os.environ.get('THR_PROBE_TILES', '2,3,4').split(','))
tot = {t: [] for t in TILES}
for rep in range(4):
    for t in TILES:
        os.environ["THR_DENSE_SEED_TILES"] = str(t)
        eng.dense_index_set(X)      # the knob is read when the index is set (no getenv on the hot path)
        eng.dense_topk(Q, 100, 28); eng.sync(); eng.prof_reset()
        for _ in range(6):
            eng.dense_topk(Q, 100, 28)
        p = eng.prof_read()
        per_call = (p["dense_score"][0] + p["dense_seed"][0] + p["dense_finalize"][0]) / 6
        tot[t].append(per_call)
for t in TILES:
    print(f"N={N} tiles {t}: score+seed+finalize per call {sorted(tot[t])[len(tot[t]) // 2]:.3f} ms (runs: " +
          " ".join(f"{x:.3f}" for x in tot[t]) + ")")
