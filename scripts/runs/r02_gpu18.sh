#!/bin/bash
# K1 / K2 chains on two streams: on vs off at an 8-GPU-sized shard and at 10M (same box, interleaved)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retriever.py tests/test_gpu_tags.py -q -m gpu -p no:cacheprovider --timeout 600 -x > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02r_tests.log
for rep in 1 2; do
for ov in 0 1; do
  THR_OVERLAP=$ov python bench.py --chunks 1250000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02r_1p25m_ov${ov}_$rep.json 2>/dev/null
done; done
for ov in 0 1; do
  THR_OVERLAP=$ov python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_10m_ov$ov.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02r_*ov*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), round(d['e2e']['value']), d['stages_ms'], d['clocks']['sm_mhz'], d['result_digest'])
    except Exception as e: print(f,'failed',e)
PY
