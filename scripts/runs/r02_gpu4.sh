#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_retriever.py tests/test_gpu_ref_tests.py tests/test_gpu_rerank.py tests/test_gpu_bm25.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/r02d_$n.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed|reference tests" gpurun_out/r02d_$n.log | tail -3
done
timeout 900 python bench.py --workload cfg5 --chunks 4000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_cfg5_small.json 2> gpurun_out/r02d_cfg5_small.err
echo "cfg5 small rc=$?"; tail -c 1500 gpurun_out/r02d_cfg5_small.json; tail -3 gpurun_out/r02d_cfg5_small.err
python __graft_entry__.py smoke 2>&1 | tail -2
