#!/bin/bash
# N GPUs: the default bench only (10M corpus, strong scaling point)
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','result_digest','sem_digest','lex_digest','stages_ms')}, d['two_stream_variant']['ms_per_step'], d['clocks']['sm_mhz'])
PY
