#!/bin/bash
# per-kernel durations of the small kernels at an 8-GPU-sized shard (ncu, serialised) + full sets of three of them
mkdir -p gpurun_out
CMD="python bench.py --chunks 1250000 --steps 3 --warmup 3 --no-cpu-baseline"
K='regex:dense_|bm25_|fuse_|merge_|exchange_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02o_launches.csv $CMD > gpurun_out/r02o_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fuse_kernel|dense_finalize|dense_seed_select' -s 9 -c 3 -o gpurun_out/r02o_small $CMD > gpurun_out/r02o_small.log 2>&1
THR_BENCH_NO_PROF=1 python bench.py --chunks 1250000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02o_noprof.json 2>/dev/null
python bench.py --chunks 1250000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02o_prof.json 2>/dev/null
python - <<'PY'
import json,csv,collections
for f in ('r02o_noprof','r02o_prof'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['clocks']['sm_mhz'])
rows=[r for r in csv.reader(open('gpurun_out/r02o_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,2),'us')
PY
