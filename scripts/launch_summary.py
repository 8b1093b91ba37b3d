#!/usr/bin/env python
"""ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` -> markdown table of the
batch-256 steps:  python scripts/launch_summary.py profiles/r01_launches.csv profiles/r01_bench_10m_profiled_cmd.json
> profiles/r01_launches_summary.md"""
import collections
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else 'profiles/r01_launches.csv'
bench = sys.argv[2] if len(sys.argv) > 2 else 'profiles/r01_bench_10m_profiled_cmd.json'
rows = list(csv.reader(open(src)))
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
# launch 0 = bm25_df_kernel (index registration); then 10 launches per step.  Every batch-256 step does the same
# work (warm-up, timed and end-to-end steps alike); the batch-1 latency steps at the end use dense_score_kernel<1>.
steps = [L[1 + 10 * s:1 + 10 * (s + 1)] for s in range((len(L) - 1) // 10)]
steps = [st for st in steps if len(st) == 10 and st[0][0].endswith('<2>') and st[-1][0].startswith('void fuse_kernel')]
n = len(steps)
agg = collections.OrderedDict()
for st in steps:
    for k, v in st:
        agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values()) / n
print("# ncu launch list, round 1 (command: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`, 10M x 1536, B=256)\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare SHARES.")
print(f"Launch 0 is `bm25_df_kernel` (index registration); then 10 launches per step; the table averages the {n} batch-256")
print(f"steps of the run (warm-up, timed and end-to-end steps do the same work).  The last launches in `{src.split('/')[-1]}`")
print("belong to the batch-1 latency loop (`dense_score_kernel<1>`).\n")
print("| kernel | launches/step | avg µs | share of step |\n|---|---|---|---|")
for k, v in agg.items():
    print(f"| `{k}` | {len(v) // n} | {sum(v) / len(v):.1f} | {100 * sum(v) / n / tot:.1f}% |")
print(f"\nsum per step under ncu: {tot / 1e3:.2f} ms.  dense_score_kernel appears twice per step: the seed pass (prefix, ~0.1 ms) and the full pass.")
b = json.load(open(bench))
st = b['stages_ms']
s = sum(st.values())
print("\nSame command without ncu (CUDA events inside bench.py, `stages_ms`): " +
      ", ".join(f"{k} {v:.3f} ms ({100 * v / s:.1f}%)" for k, v in st.items()))
