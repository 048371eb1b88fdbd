#!/bin/bash
for b in 38 19 13 10; do
  python bench.py --no-cpu --frames 38 --batch $b --steps 20 > gpurun_out/r2o_b$b.json 2>gpurun_out/r2o_b$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2o_b$b.json").read().strip().splitlines()[-1])
print("batch=$b value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3))
PY
done
