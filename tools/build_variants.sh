#!/bin/bash
# Kernel-variant experiments: builds liblpvs variants with extra -D flags on gram.cu into gpurun_variants/<name>.so
# (selected at run time with LPVS_LIB=...).  usage: tools/build_variants.sh name "-DFLAG=1" [name2 "-D..."] ...
set -e
cd "$(dirname "$0")/../lpvspectral.jl_b200/csrc"
mkdir -p ../../gpurun_variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off $flags -c gram.cu -o /tmp/gram_$name.o
  nvcc $ARCH -shared -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64 -o ../../gpurun_variants/$name.so /tmp/gram_$name.o structured.o chol.o api.o admm.o lpv.o lsq.o
  echo built $name
done
