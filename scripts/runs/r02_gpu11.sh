#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02k_reference.json 2> gpurun_out/r02k_reference.err
echo "reference rc=$? wall=$(( $(date +%s) - S )) s"; tail -c 2500 gpurun_out/r02k_reference.json; tail -3 gpurun_out/r02k_reference.err
python __graft_entry__.py smoke 2>&1 | tail -2
