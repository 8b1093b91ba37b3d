#!/bin/bash
mkdir -p gpurun_out
P="python scripts/bm25_probe.py 10000000 256"
for f in tests/test_gpu_bm25.py tests/test_gpu_tags.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r02m_$n.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed|Error" gpurun_out/r02m_$n.log | tail -3
done
{
echo "generation tags"; $P | tail -1
echo "N=1.25M"; python scripts/bm25_probe.py 1250000 256 | tail -1
echo "cfg3"; python scripts/bm25_probe.py 1000000 1024 | tail -1
} > gpurun_out/r02m_sweep.log 2>&1
cat gpurun_out/r02m_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 1 -o gpurun_out/r02m_bm25 $P > gpurun_out/r02m_ncu_bm25.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stages_ms'], d['bm25_roofline']['alone']['launch_ms'], d['clocks']['sm_mhz'], d['result_digest'], d['lex_digest'])
PY
