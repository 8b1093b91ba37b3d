#!/bin/bash
# Run on the GPU box (via gpurun): every GPU test file in its own process, so that a trapped
# kernel in one file cannot poison the CUDA context of the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_fuse.py tests/test_gpu_bm25.py tests/test_gpu_maxsim.py tests/test_gpu_dense.py tests/test_gpu_retriever.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py tests/test_gpu_tags.py tests/test_gpu_fusion_dropin.py tests/test_gpu_ref_tests.py tests/test_gpu_baseline_shapes.py "$@"; do
  n=$(basename $f .py)
  timeout 1500 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$n.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -25 gpurun_out/$n.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
