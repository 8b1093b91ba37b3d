#!/bin/bash
# round 2, N GPUs: exchange check, the default bench (digests must equal N = 1's), cfg5 at 50M.
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $T scripts/exchange_check.py > gpurun_out/r02e_exchange_n$N.log 2>&1; echo "exchange rc=$?"; grep -E "exchange check|->" gpurun_out/r02e_exchange_n$N.log
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','result_digest','sem_digest','lex_digest','stages_ms')}, d['config'].get('exchange'), d['dense_certificate'])
PY
timeout 1500 $T bench.py --gpus $N --workload cfg5 --steps 20 --warmup 3 > gpurun_out/r02e_cfg5_n$N.json 2> gpurun_out/r02e_cfg5_n$N.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02e_cfg5_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02e_cfg5_n$N.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','result_digest','rerank_digest','stages_ms','setup_s','maxsim_roofline')})
except Exception as e: print('no cfg5 line', e)
PY
