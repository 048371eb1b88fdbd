#!/bin/bash
# usage (GPU box): tools/sweep_r2f.sh  -> configs[3] launch-group sweep and compositing-warps-per-SM sweep at configs[2]
show() { python - "$1" "$2" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
}
for b in 16 32 48 64; do
  timeout 120 python bench.py --config 3 --no-cpu --steps 2 --warmup 1 --batch $b > gpurun_out/sw3_b$b.json 2> gpurun_out/sw3_b$b.err
  show "cfg3 batch $b" gpurun_out/sw3_b$b.json
done
for w in 16 20 24 28; do
  OMFS_COMP_PIPE_WARPS=$w timeout 60 python bench.py --no-cpu --steps 8 --warmup 3 > gpurun_out/sw2_w$w.json 2> gpurun_out/sw2_w$w.err
  show "cfg2 pipe warps $w" gpurun_out/sw2_w$w.json
done
