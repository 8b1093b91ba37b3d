#!/bin/bash
# round 2, final single-GPU pass: every GPU test, the ncu set of the final build, MaxSim probes, the bench line, the reference arm.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_rerank.py tests/test_gpu_hybrid_rag1.py tests/test_gpu_multi.py > gpurun_out/r02z_tests_summary.log 2>&1
echo "tests rc=$?"; grep -E "^== |passed|failed|error" gpurun_out/r02z_tests_summary.log | tail -40
python __graft_entry__.py smoke 2>&1 | tail -1
bash scripts/profile_round.sh r02 > gpurun_out/r02z_profile.log 2>&1; tail -4 gpurun_out/r02z_profile.log
{ python scripts/bm25_probe.py 10000000 256; python scripts/bm25_probe.py 1250000 256; python scripts/bm25_probe.py 1000000 1024; THR_PROBE_COLD=1 python scripts/bm25_probe.py 10000000 256; } 2>&1 | grep docs= > gpurun_out/r02z_bm25_probe.log; cat gpurun_out/r02z_bm25_probe.log
timeout 1200 python bench.py --steps 30 --warmup 3 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02z_bench_n1.json
[ -n "$WITH_REFERENCE" ] && { timeout 1200 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02z_reference.json 2> gpurun_out/r02z_reference.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02z_reference.json; }
