#!/bin/bash
# usage (GPU box): tools/quick_r2f.sh  -> GPU tests, configs[2] bench, configs[3] at launch groups of 64 / 128 images
show() { python - "$1" "$2" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3),
          " ".join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d["stages"].items()))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 60 python bench.py --no-cpu --steps 8 --warmup 3 > gpurun_out/q2.json 2> gpurun_out/q2.err; show cfg2 gpurun_out/q2.json
for b in 64 128; do
  timeout 120 python bench.py --config 3 --no-cpu --steps 2 --warmup 1 --batch $b > gpurun_out/q3_b$b.json 2> gpurun_out/q3_b$b.err
  show "cfg3 batch $b" gpurun_out/q3_b$b.json
done
