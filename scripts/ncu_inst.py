#!/usr/bin/env python
"""Rank CUDA-C lines of an ncu report by executed warp instructions: python scripts/ncu_inst.py <rep> [top]"""
import csv, io, subprocess, collections, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; agg = collections.Counter(); samp = collections.Counter(); src = {}; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    try: n = int(r[ii]); sm = int(r[isamp])
    except ValueError: continue
    k = (cur, int(r[0])); agg[k] += n; samp[k] += sm; src[k] = r[1].strip()[:90]
tot = sum(agg.values()); ts = sum(samp.values())
print("total warp instructions", tot, "samples", ts)
for k, v in agg.most_common(top):
    print(f"{v:11d} {100*v/tot:5.1f}%  samp {100*samp[k]/ts:4.1f}%  {k[0]}:{k[1]} {src[k]}")
