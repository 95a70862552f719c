#!/bin/bash
# Builds shg_vqa_b200/libshgvqa.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p _obj
pids=()
for f in core lsap gemm rowops attn optim conv; do
  [ -f $f.cu ] || continue
  if [ ! -f _obj/$f.o ] || [ $f.cu -nt _obj/$f.o ] || [ shg_common.cuh -nt _obj/$f.o ] || [ ../../include/shgvqa.h -nt _obj/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o _obj/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libshgvqa.so _obj/*.o -lcudart
echo built $(realpath ../libshgvqa.so)
