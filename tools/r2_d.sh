#!/bin/bash
python -m pytest tests/test_gpu_torch_ops.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2d_torchops.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_torch_ops.py 2>&1 | tail -4
for v in "1" "0"; do
  OMFS_FUSE_FRONT=$v python bench.py --no-cpu > gpurun_out/r2d_bench_f$v.json 2> gpurun_out/r2d_bench_f$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_bench_f$v.json").read().strip().splitlines()[-1])
print("fuse=$v value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3), {k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
done
