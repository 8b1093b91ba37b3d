#!/bin/bash
# round 2: the full GPU test suite (new: BASELINE-shape parity, the reference's own tests, AND semantics) and the bench.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/r02c_tests_summary.log 2>&1
echo "tests rc=$?"; grep -E "^== |passed|failed|error" gpurun_out/r02c_tests_summary.log | tail -40
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
