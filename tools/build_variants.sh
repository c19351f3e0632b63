#!/bin/bash
# Differently-compiled builds of the library for kernel A/B measurements (selected with SBWT_B200_LIB=...):
#   bash tools/build_variants.sh name1 "-DFLAG=.." name2 "-DFLAG=.. -DFLAG2=.." ...
set -e
cd "$(dirname "$0")/../sbwt_b200/csrc"
mkdir -p ../../.variants
[ -f host_widen.o ] || make host_widen.o
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $flags \
      -shared -o ../../.variants/$name.so sbwt_gpu.cu host_widen.o 2> ../../.variants/$name.ptxas.log && echo "built $name" ) &
done
wait
