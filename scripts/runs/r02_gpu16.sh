#!/bin/bash
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_dense.py tests/test_gpu_tags.py tests/test_gpu_baseline_shapes.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 -x > gpurun_out/r02p_$n.log 2>&1
  r=$?; echo "== $f exit $r"; tail -4 gpurun_out/r02p_$n.log; [ $r -ne 0 ] && rc=1
done
python bench.py --chunks 1250000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02p_bench_1p25m.json 2> gpurun_out/r02p_bench_1p25m.err; echo "bench 1.25M rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_10m.json 2> gpurun_out/r02p_bench_10m.err; echo "bench 10M rc=$?"
python - <<'PY'
import json
for f in ('r02p_bench_1p25m','r02p_bench_10m'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['stages_ms'], d['clocks']['sm_mhz'], d['result_digest'], d['sem_digest'], d['dense_certificate']['min_gap'])
    except Exception as e:
        print(f, 'failed', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
exit $rc
