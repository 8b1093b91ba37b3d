"""Time the MaxSim kernel alone (BASELINE config 4): python scripts/maxsim_probe.py [B] [C] [Tq] [Td]"""
import sys, json, torch
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
Tq = int(sys.argv[3]) if len(sys.argv) > 3 else 32
Td = int(sys.argv[4]) if len(sys.argv) > 4 else 128
d = 128
eng = Engine(0)
Q, D, cand = synth.maxsim_tokens(B, C, Tq=Tq, Td=Td, d=d, device=eng.device)
eng.prof_enable(True)
for _ in range(3):
    eng.maxsim(Q, D, cand)
eng.sync(); eng.prof_reset()
for _ in range(10):
    eng.maxsim(Q, D, cand)
p = eng.prof_read()
ms = p["maxsim"][0] / p["maxsim"][1]
by = B * C * Td * d * 2 + B * Tq * d * 2 + B * C * 12
fl = 2.0 * B * Tq * C * Td * d
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6544.0}
print(f"maxsim B={B} C={C} Tq={Tq} Td={Td}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peaks['hbm_gbs']*100:.1f}% of measured HBM)  "
      f"{fl/ms/1e9:.1f} TFLOP/s  {B/ms*1e3:.0f} queries/s")
