#!/bin/bash
# rank-merge exchange + plan kernel: parity on 2 GPUs, then the 10M bench at N = 2
N=${1:-2}
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_edges.py tests/test_gpu_bm25.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 -x > gpurun_out/r02q_$n.log 2>&1
  r=$?; echo "== $f exit $r"; tail -4 gpurun_out/r02q_$n.log; [ $r -ne 0 ] && rc=1
done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $T scripts/exchange_check.py > gpurun_out/r02q_exchange_n$N.log 2>&1; echo "exchange rc=$?"; grep -E "exchange check|->" gpurun_out/r02q_exchange_n$N.log
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02q_bench_n$N.json 2> gpurun_out/r02q_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02q_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','result_digest','sem_digest','lex_digest','stages_ms')}, d['config'].get('exchange'))
for r in d.get('stages_ms_per_rank', []): print(r)
PY
exit $rc
