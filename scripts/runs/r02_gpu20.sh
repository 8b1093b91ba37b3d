#!/bin/bash
# where does the two-stream gain come from?  serial / two streams / serial with BM25 first, interleaved, no event pairs
mkdir -p gpurun_out
for rep in 1 2; do
  THR_BENCH_NO_PROF=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_serial_$rep.json 2>/dev/null
  THR_BENCH_NO_PROF=1 THR_ORDER=lex_first python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_lexfirst_$rep.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'serial-mode step', round(d['ms_per_step'],4), 'two-stream', round(d['two_stream_variant']['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['result_digest'])
    except Exception as e: print(f,'failed',e)
PY
