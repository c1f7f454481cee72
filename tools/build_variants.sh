#!/bin/bash
# Builds libt3dgeo.so variants into build/variants/<name>/ for tools/bench_variants.sh.
# usage: tools/build_variants.sh name1 "-DFLAG=1 ..." name2 "..." ...
cd "$(dirname "$0")/../tetra3d_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=../../build/variants/$name; mkdir -p $d
  ( nvcc -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags -shared -o $d/libt3dgeo.so t3d_api.cu t3d_kernels.cu t3d_sort.cu t3d_pose.cu t3d_bake.cu 2> $d/ptxas.log \
      || { echo "$name: BUILD FAILED"; grep -i error $d/ptxas.log | head; } ) &
done
wait
