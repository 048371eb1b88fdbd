#!/bin/bash
for w in 12 16 20 24 28 32; do
  OMFS_COMP_PIPE_WARPS=$w python bench.py --no-cpu --steps 8 > gpurun_out/r2n_w$w.json 2>gpurun_out/r2n_w$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_w$w.json").read().strip().splitlines()[-1])
print("warps=$w value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3))
PY
done
