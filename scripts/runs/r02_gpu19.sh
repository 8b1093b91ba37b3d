#!/bin/bash
# bench accounting change (roofline slots in the region, small kernels in their own pass) + smoke, 1 GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02s_smoke.log
python bench.py --chunks 1250000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_1p25m.json 2> gpurun_out/r02s_1p25m.err; echo "rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_10m.json 2> gpurun_out/r02s_10m.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02s_1p25m','r02s_10m'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), round(d['e2e']['value']), d['two_stream_variant']['ms_per_step'], d['stages_ms'], d['clocks']['sm_mhz'], d['result_digest'], d['roofline']['frac'], d['bm25_roofline']['frac'], d['gpu_launches'])
    except Exception as e:
        print(f,'failed',e); print(open(f'gpurun_out/{f}.err').read()[-2000:])
PY
