#!/usr/bin/env python
"""ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` -> markdown table of the
batch-256 steps:  python scripts/launch_summary.py profiles/r02_launches.csv profiles/r02_bench_10m_profiled_cmd.json
> profiles/r02_launches_summary.md"""
import collections
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else 'profiles/r02_launches.csv'
bench = sys.argv[2] if len(sys.argv) > 2 else 'profiles/r02_bench_10m_profiled_cmd.json'
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
# A batch-256 step = the launches from a dense_score_kernel<2> that is followed by dense_seed_select_kernel (the seed
# pass) up to and including the next fuse_kernel.  Every batch-256 step does the same work (warm-up, timed and
# end-to-end steps alike); the BM25-alone loop and the batch-1 latency steps at the end do not match the pattern.
steps, cur = [], None
for i, (k, v) in enumerate(L):
    if cur is None:
        if 'dense_score_kernel<2' in k and i + 1 < len(L) and L[i + 1][0].startswith('dense_seed_select'):
            cur = [(k, v)]
    else:
        cur.append((k, v))
        if k.startswith('void fuse_kernel'):
            steps.append(cur)
            cur = None
        elif len(cur) > 16:
            cur = None
per = collections.Counter(len(st) for st in steps).most_common(1)[0][0]
steps = [st for st in steps if len(st) == per]
n = len(steps)
agg = collections.OrderedDict()
for st in steps:
    for k, v in st:
        agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values()) / n
what = sys.argv[3] if len(sys.argv) > 3 else "`python bench.py --steps 3 --warmup 3 --no-cpu-baseline`, 10M x 1536, B=256"
print(f"# ncu launch list (command: {what})\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare SHARES.")
print(f"Launch 0 is `bm25_df_kernel` (index registration); then {per} launches per step; the table averages the {n} batch-256")
print(f"steps of the run (warm-up, timed and end-to-end steps do the same work).  The last launches in `{src.split('/')[-1]}`")
print("belong to the BM25-alone loop and the batch-1 latency loop (`dense_score_kernel<1>`).\n")
print("| kernel | launches/step | avg µs | share of step |\n|---|---|---|---|")
for k, v in agg.items():
    print(f"| `{k}` | {len(v) // n} | {sum(v) / len(v):.1f} | {100 * sum(v) / n / tot:.1f}% |")
print(f"\nsum per step under ncu: {tot / 1e3:.2f} ms.  `dense_score_kernel<2, 1>` is the seed pass (prefix), `<2, 0>` the full pass.")
b = json.load(open(bench))
st = b['stages_ms']
s = sum(st.values())
print("\nSame command without ncu (CUDA events inside bench.py, `stages_ms`): " +
      ", ".join(f"{k} {v:.3f} ms ({100 * v / s:.1f}%)" for k, v in st.items()))
