#!/bin/bash
# round 2: BM25 trims (sentinel descriptors, acc0 addressing, cost weights): parity, time alone, ncu; new tests.
mkdir -p gpurun_out
P="python scripts/bm25_probe.py 10000000 256"
for f in tests/test_gpu_bm25.py tests/test_gpu_tags.py tests/test_gpu_rerank.py tests/test_gpu_retriever.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r02f_$n.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed" gpurun_out/r02f_$n.log | tail -2
done
{
echo "default"; $P | tail -1
echo "range_cost 64 term 24"; THR_BM25_RANGE_COST=64 THR_BM25_TERM_COST=24 $P | tail -1
echo "range_cost 300 term 15"; THR_BM25_RANGE_COST=300 $P | tail -1
echo "units 2"; THR_BM25_UNITS_PER_SM=2 $P | tail -1
echo "N=1.25M"; python scripts/bm25_probe.py 1250000 256 | tail -1
echo "cfg3"; python scripts/bm25_probe.py 1000000 1024 | tail -1
} > gpurun_out/r02f_sweep.log 2>&1
cat gpurun_out/r02f_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 1 -o gpurun_out/r02f_bm25 $P > gpurun_out/r02f_ncu_bm25.log 2>&1; echo "ncu exit $?"
