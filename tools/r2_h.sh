#!/bin/bash
# full captures (with source) of the compositing and emit kernels on the current tree
CMD="python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu"
$CMD > gpurun_out/r2h_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'composite_kernel|emit_scatter' -s 2 -c 2 -o gpurun_out/r2h_comp -f $CMD > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_ncu.log
