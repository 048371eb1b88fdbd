#!/bin/bash
# usage (through gpurun, one GPU): tools/profile_round.sh <tag>
# 1. plain bench (must exit 0 before anything runs under ncu)  2. ncu launch list of the same command
# 3. ncu --set full of the top kernels.  Everything lands in gpurun_out/<tag>_*.
tag=$1
set -o pipefail
CMD="python bench.py --steps 1 --warmup 1 --frames 75 --no-cpu"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_bench.err; exit 1; }
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'composite_kernel|emit_scatter|bind_preprocess|flame_blend_tc|rs_onesweep|tile_count|face_frames|flame_lbs' -s 6 -c 12 \
    -o gpurun_out/${tag}_full -f $CMD > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'])"
