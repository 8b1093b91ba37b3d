#!/bin/bash
# Run on the GPU box (via gpurun): plain bench first, then the ncu passes of the SAME command.
#   bash scripts/profile_round.sh <tag>      -> gpurun_out/<tag>_*.{log,csv,ncu-rep}
tag=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 400 gpurun_out/${tag}_plain.log; echo
K='regex:dense_|bm25_|fuse_|merge_|safety_|maxsim_|rerank_|exchange_'
# every launch of our kernels with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
# full sets: dense main pass (4th dense_score launch: seed, main, seed, main), bm25, the small kernels
ncu --set full --clock-control none --import-source on -k regex:dense_score -s 3 -c 1 -o gpurun_out/${tag}_dense $CMD > gpurun_out/${tag}_ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 1 -c 1 -o gpurun_out/${tag}_bm25 $CMD > gpurun_out/${tag}_ncu_bm25.log 2>&1
ncu --set full --clock-control none -k 'regex:fuse_kernel|dense_finalize|dense_seed_select|bm25_merge|bm25_plan|bm25_order|bm25_cost' -s 14 -c 8 -o gpurun_out/${tag}_small $CMD > gpurun_out/${tag}_ncu_small.log 2>&1
python scripts/maxsim_probe.py > gpurun_out/${tag}_maxsim_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:maxsim -s 3 -c 1 -o gpurun_out/${tag}_maxsim python scripts/maxsim_probe.py > gpurun_out/${tag}_ncu_maxsim.log 2>&1
cat gpurun_out/${tag}_maxsim_plain.log
ls -la gpurun_out | grep ${tag}_
