#!/bin/bash
# Build libthr_<name>.so with extra -D flags for bm25.cu only (kernel-shape experiments):
#   scripts/build_variant.sh g2 "-DTHR_BM25_GRAB=2"      then run with THR_LIB=triple_hybrid_rag_b200/lib/libthr_g2.so
set -e
cd "$(dirname "$0")/../triple_hybrid_rag_b200/csrc"
make -s > /dev/null
name=$1; flags=$2
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p build/var
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -Xptxas -v -c bm25.cu -o build/var/bm25_$name.o 2> build/var/bm25_$name.log
grep -A2 "range_kernelILi2048ELb0" build/var/bm25_$name.log | grep -E "registers|spill" | head -2
$NVCC $ARCH -shared -o ../lib/libthr_$name.so build/api.o build/fuse.o build/dense_topk.o build/var/bm25_$name.o build/maxsim.o build/rerank.o -lcudart
echo built libthr_$name.so
