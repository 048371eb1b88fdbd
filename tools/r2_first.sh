#!/bin/bash
# round 2, first GPU pass: parity suite, one-GPU bench, pure-copy ceiling at one GPU
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
python tools/ubench/d2h_bw.py > gpurun_out/r2a_d2h_1.json 2> gpurun_out/r2a_d2h_1.err
cat gpurun_out/r2a_d2h_1.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["d2h_bytes_per_step"],"e2e_raw",d["e2e_raw"]["value"],"ms/step",d["ms_per_step"])
for k,v in d["stages"].items(): print(k,round(v["ms_per_launch"],4),round(v["share_of_step"],3),round(v["frac_of_hbm"],3))
PY
