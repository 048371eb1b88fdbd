#!/bin/bash
# usage (here, CPU): tools/variants.sh <unit.cu> "<name>:<nvcc defines>" ...   -> build/variants/<name>.so
# Builds one library per variant of a single translation unit (the other objects come from build/).
unit=$1; shift
mkdir -p build/variants
FLAGS="-ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
case $unit in exact_geom.cu|binning.cu|composite.cu|displace.cu) FLAGS="$FLAGS --fmad=false";; esac
for v in "$@"; do
  name=${v%%:*}; defs=${v#*:}
  ( nvcc $FLAGS $defs -Xptxas -v -c omfs-4d-video-gen_b200/csrc/$unit -o build/variants/$name.o 2> build/variants/$name.ptxas
    others=$(ls build/*.o | grep -v "/${unit%.cu}.o")
    nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so build/variants/$name.o $others
    grep -h "registers" build/variants/$name.ptxas | head -3 | sed "s/^/$name: /" ) &
done
wait
