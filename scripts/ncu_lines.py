#!/usr/bin/env python
"""Summarise an ncu source page by CUDA-C line: python scripts/ncu_lines.py <report.ncu-rep> [kernel-regex] [top]
Reads `ncu -i <rep> --page source --print-source cuda,sass --csv` and prints, per source line, the warp-stall
samples and the dominant stall reasons (needs -lineinfo at compile time)."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["-k", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cur_file = None
agg = collections.OrderedDict()
total = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue   # SASS rows have an empty line number
    try:
        n = int(r[i_samp])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    st = {h: int(r[i] or 0) for i, h in stall_cols}
    if key in agg:
        agg[key][0] += n
        for h, v in st.items():
            agg[key][2][h] = agg[key][2].get(h, 0) + v
        agg[key][3] += int(r[i_inst] or 0)
    else:
        agg[key] = [n, r[1].strip()[:90], st, int(r[i_inst] or 0)]
    total += n
print(f"total samples {total}")
for (f, ln), (n, src, st, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ts = " ".join(f"{h[6:]}={v}" for h, v in tops if v)
    print(f"{100.0*n/max(total,1):5.1f}% {n:7d} inst={inst:9d} {f}:{ln:<4d} {src}   [{ts}]")
