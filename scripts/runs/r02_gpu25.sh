#!/bin/bash
# per-kernel launch list of the final build on a 1.25M-chunk shard (the size one of 8 ranks holds)
mkdir -p gpurun_out
CMD="python bench.py --chunks 1250000 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02y_plain_1p25m.log 2>&1 || { tail -5 gpurun_out/r02y_plain_1p25m.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:dense_|bm25_|fuse_|merge_|exchange_' -c 200 --csv --log-file gpurun_out/r02y_launches_1p25m.csv $CMD > gpurun_out/r02y_list.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r02y_launches_1p25m.csv
