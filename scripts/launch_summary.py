#!/usr/bin/env python
"""profiles/r01_launches.csv (ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`)
-> markdown table of the 3 timed steps: python scripts/launch_summary.py > profiles/r01_launches_summary.md"""
import collections
import csv
import json

rows = list(csv.reader(open('profiles/r01_launches.csv')))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i
        break
ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
L = []
for r in rows[start + 1:]:
    if len(r) > iv:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(r[iu], 1)
        L.append((r[ik].replace('<unnamed>::', '').split('(')[0][:40], v))
# launch 0 = bm25_df_kernel (index registration); then 10 launches per step; steps 3..5 are the timed ones
steps = [L[1 + 10 * s:1 + 10 * (s + 1)] for s in range(6)]
agg = collections.OrderedDict()
for st in steps[3:]:
    for k, v in st:
        agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values()) / 3
print("# ncu launch list, round 1 (command: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`, 10M x 1536, B=256)\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare SHARES.")
print("Launch 0 is `bm25_df_kernel` (index registration); then 10 launches per step; the table averages the 3 timed steps")
print("(launches 31-60).  Later launches in `r01_launches.csv` belong to the end-to-end and batch-1 latency loops.\n")
print("| kernel | launches/step | avg µs | share of step |\n|---|---|---|---|")
for k, v in agg.items():
    print(f"| `{k}` | {len(v) // 3} | {sum(v) / len(v):.1f} | {100 * sum(v) / 3 / tot:.1f}% |")
print(f"\nsum per step under ncu: {tot / 1e3:.2f} ms.  dense_score_kernel appears twice per step: the seed pass (prefix, ~0.1 ms) and the full pass.")
b = json.load(open('profiles/r01_bench_10m_profiled_cmd.json'))
st = b['stages_ms']
s = sum(st.values())
print("\nSame command without ncu (CUDA events inside bench.py, `stages_ms`): " +
      ", ".join(f"{k} {v:.3f} ms ({100 * v / s:.1f}%)" for k, v in st.items()))
