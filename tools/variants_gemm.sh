#!/bin/bash
# usage (GPU box): tools/variants_gemm.sh  -> tools/bench_gemm.py with every build/variants/*.so
cp omfs-4d-video-gen_b200/libomfs_b200.so /tmp/lib_orig.so
for so in build/variants/*.so; do
  cp $so omfs-4d-video-gen_b200/libomfs_b200.so
  echo "== $(basename $so .so)"; python tools/bench_gemm.py 2>&1 | grep tcgen05 | cut -c1-150
done
cp /tmp/lib_orig.so omfs-4d-video-gen_b200/libomfs_b200.so
