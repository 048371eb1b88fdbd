#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2e_tests_full.log; tail -3 gpurun_out/r2e_tests_full.log
python bench.py --no-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
OMFS_TRACE=1 python bench.py --no-cpu --steps 3 > /dev/null 2> gpurun_out/r2e_trace.err
grep "omfs trace" gpurun_out/r2e_trace.err | tail -8
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3), {k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
CMD="python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:'png_strip_kernel|bind_preprocess_kernel|emit_scatter|rs_onesweep' -s 4 -c 6 -o gpurun_out/r2e_full -f $CMD > gpurun_out/r2e_ncu2.log 2>&1
tail -2 gpurun_out/r2e_ncu2.log; ls -la gpurun_out/r2e_full.ncu-rep
