#!/bin/bash
# round 2: BM25 kernel sweeps (prefetch distance, grab size, warps, units per SM) + DRAM traffic of a few of them
mkdir -p gpurun_out
L=triple_hybrid_rag_b200/lib
P="python scripts/bm25_probe.py 10000000 256"
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_tags.py -q -m gpu -p no:cacheprovider --timeout 300 -x > gpurun_out/r02b_bm25_tests.log 2>&1
echo "== bm25 tests exit $?"; tail -3 gpurun_out/r02b_bm25_tests.log
{
echo "default";            $P | tail -1
for pf in 0 8 16 32 100 200; do echo "prefetch $pf"; THR_BM25_PREFETCH=$pf $P | tail -1; done
for w in 20 16; do echo "warps $w"; THR_BM25_WARPS=$w $P | tail -1; done
for u in 1 4; do echo "units_per_sm $u"; THR_BM25_UNITS_PER_SM=$u $P | tail -1; done
for v in g2 g3 g8c64; do echo "variant $v"; THR_LIB=$L/libthr_$v.so $P | tail -1; done
echo "N=1.25M default"; python scripts/bm25_probe.py 1250000 256 | tail -1
echo "N=1.25M units 1"; THR_BM25_UNITS_PER_SM=1 python scripts/bm25_probe.py 1250000 256 | tail -1
echo "N=1.25M units 4"; THR_BM25_UNITS_PER_SM=4 python scripts/bm25_probe.py 1250000 256 | tail -1
} > gpurun_out/r02b_sweep.log 2>&1
cat gpurun_out/r02b_sweep.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum"
for pf in default 0 16; do
  if [ $pf = default ]; then unset THR_BM25_PREFETCH; else export THR_BM25_PREFETCH=$pf; fi
  ncu --metrics $M --clock-control none -k regex:bm25_range_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02b_dram_pf$pf.csv $P > /dev/null 2>&1
  echo "pf=$pf"; tail -4 gpurun_out/r02b_dram_pf$pf.csv | cut -d, -f12-
done
