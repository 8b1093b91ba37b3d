#!/bin/bash
# round 2, first GPU call: parity of the new BM25 kernel, then its time alone and one ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_tags.py -q -m gpu -p no:cacheprovider --timeout 300 -x > gpurun_out/r02a_bm25_tests.log 2>&1
echo "== bm25 tests exit $?"; tail -15 gpurun_out/r02a_bm25_tests.log
timeout 600 python scripts/bm25_probe.py 10000000 256 > gpurun_out/r02a_probe_10m.log 2>&1; echo "probe exit $?"; tail -3 gpurun_out/r02a_probe_10m.log
timeout 300 python scripts/bm25_probe.py 1250000 256 > gpurun_out/r02a_probe_1p25m.log 2>&1; tail -2 gpurun_out/r02a_probe_1p25m.log
timeout 300 python scripts/bm25_probe.py 1000000 1024 > gpurun_out/r02a_probe_cfg3.log 2>&1; tail -2 gpurun_out/r02a_probe_cfg3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 1 -o gpurun_out/r02a_bm25 python scripts/bm25_probe.py 10000000 256 > gpurun_out/r02a_ncu_bm25.log 2>&1; echo "ncu exit $?"
for f in tests/test_gpu_fuse.py tests/test_gpu_maxsim.py tests/test_gpu_dense.py tests/test_gpu_retriever.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py tests/test_gpu_fusion_dropin.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r02a_$n.log 2>&1
  echo "== $f exit $?"; tail -3 gpurun_out/r02a_$n.log
done
