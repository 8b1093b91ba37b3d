#!/bin/bash
mkdir -p gpurun_out
L=triple_hybrid_rag_b200/lib
P="python scripts/bm25_probe.py 10000000 256"
M="dram__bytes_read.sum,gpu__time_duration.sum"
{
echo "base"; $P | tail -1
for v in pfh pfha pfa; do echo "variant $v"; THR_LIB=$L/libthr_$v.so $P | tail -1; done
} > gpurun_out/r02l_sweep.log 2>&1
cat gpurun_out/r02l_sweep.log
for v in pfh pfha; do
  THR_LIB=$L/libthr_$v.so ncu --metrics $M --clock-control none -k regex:bm25_range_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02l_dram_$v.csv $P > /dev/null 2>&1
  echo "$v"; tail -2 gpurun_out/r02l_dram_$v.csv | cut -d, -f12-
  THR_LIB=$L/libthr_$v.so timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_$v.json 2> gpurun_out/r02l_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_bench_$v.json').read().strip().splitlines()[-1])
print('$v', d['ms_per_step'], d['stages_ms']['bm25'], d['bm25_roofline']['alone']['launch_ms'], d['clocks']['sm_mhz'], d['result_digest'])
PY
done
