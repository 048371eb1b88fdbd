#!/bin/bash
# emit_scatter at 4096 tiles (configs[3]): full capture with source
CMD="python bench.py --config 3 --steps 1 --warmup 1 --frames 4 --no-cpu"
$CMD > gpurun_out/r2k_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r2k_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'emit_scatter' -s 1 -c 1 -o gpurun_out/r2k_emit12 -f $CMD > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log
