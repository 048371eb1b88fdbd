#!/bin/bash
# round 2, pass c: fused front end A/B (bit-exact suite both ways), last-batch-full A/B, launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2c_tests.log; tail -3 gpurun_out/r2c_tests.log
OMFS_FUSE_FRONT=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
for v in "1 1" "0 1" "1 0"; do
  set -- $v
  OMFS_FUSE_FRONT=$1 OMFS_COMP_LAST_FULL=$2 python bench.py --no-cpu > gpurun_out/r2c_bench_f$1_l$2.json 2> gpurun_out/r2c_bench_f$1_l$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench_f$1_l$2.json").read().strip().splitlines()[-1])
print("fuse=$1 lastfull=$2 value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3), {k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
done
CMD="python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
tail -2 gpurun_out/r2c_ncu1.log
