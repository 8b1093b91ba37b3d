#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_rerank.py tests/test_gpu_hybrid_rag1.py tests/test_gpu_multi.py > gpurun_out/r02i_tests_summary.log 2>&1
echo "tests rc=$?"; grep -E "^== |passed|failed|error" gpurun_out/r02i_tests_summary.log | tail -40
bash scripts/profile_round.sh r02 > gpurun_out/r02i_profile.log 2>&1; tail -15 gpurun_out/r02i_profile.log
