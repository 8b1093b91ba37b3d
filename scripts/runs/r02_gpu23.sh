#!/bin/bash
# export adapter test on the GPU + the driver's two bench invocations with default flags (sanity of the final tree)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retriever.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r02u_retriever.log 2>&1; echo "retriever tests rc=$?"; tail -3 gpurun_out/r02u_retriever.log
timeout 900 python bench.py > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err; echo "default bench rc=$?"; tail -c 300 gpurun_out/r02u_bench_default.json
