#!/bin/bash
# N GPUs: cfg5 only (50M x 1024 with rerank)
N=${1:-4}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 900 $T bench.py --gpus $N --workload cfg5 --steps 20 --warmup 3 > gpurun_out/r02e_cfg5_n$N.json 2> gpurun_out/r02e_cfg5_n$N.err; echo "cfg5 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_cfg5_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','result_digest','rerank_digest','stages_ms')}, d['two_stream_variant']['ms_per_step'], d['clocks']['sm_mhz'])
PY
