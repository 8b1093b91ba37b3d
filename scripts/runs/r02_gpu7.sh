#!/bin/bash
mkdir -p gpurun_out
P="python scripts/bm25_probe.py 10000000 256"
for f in tests/test_gpu_bm25.py tests/test_gpu_tags.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r02g_$n.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed" gpurun_out/r02g_$n.log | tail -2
done
{
echo "default (static grabs, exact prefetch)"; $P | tail -1
echo "prefetch off"; THR_BM25_PREFETCH=0 $P | tail -1
echo "N=1.25M"; python scripts/bm25_probe.py 1250000 256 | tail -1
echo "N=1.25M prefetch off"; THR_BM25_PREFETCH=0 python scripts/bm25_probe.py 1250000 256 | tail -1
echo "cfg3"; python scripts/bm25_probe.py 1000000 1024 | tail -1
} > gpurun_out/r02g_sweep.log 2>&1
cat gpurun_out/r02g_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 1 -o gpurun_out/r02g_bm25 $P > gpurun_out/r02g_ncu_bm25.log 2>&1; echo "ncu exit $?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','result_digest','stages_ms','bm25_roofline','clocks','cpu_baseline')})
PY
