#!/bin/bash
mkdir -p gpurun_out
P="python scripts/bm25_probe.py 10000000 256"
{
echo "dynamic warm"; $P | tail -1
echo "dynamic cold-L2"; THR_PROBE_COLD=1 $P | tail -1
echo "static cold-L2"; THR_PROBE_COLD=1 THR_BM25_STATIC=1 $P | tail -1
echo "static+prefetch cold-L2"; THR_PROBE_COLD=1 THR_BM25_PREFETCH=1 $P | tail -1
echo "static+prefetch warm"; THR_BM25_PREFETCH=1 $P | tail -1
} > gpurun_out/r02h_sweep.log 2>&1
cat gpurun_out/r02h_sweep.log
for cfg in "dyn:" "static:THR_BM25_STATIC=1" "pf:THR_BM25_PREFETCH=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_$name.json 2> gpurun_out/r02h_bench_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02h_bench_$name.json').read().strip().splitlines()[-1])
print('$name', d['ms_per_step'], d['stages_ms'], d['bm25_roofline']['alone']['launch_ms'], d['clocks']['sm_mhz'])
PY
done
timeout 600 python -m pytest tests/test_gpu_hybrid_rag1.py tests/test_gpu_bm25.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
