#!/bin/bash
# usage (GPU box): tools/quick2.sh [pytest]  -> (GPU parity tests,) configs[2] and configs[3] benches with stage tables
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
if [ "$1" = pytest ]; then timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
if [ "$1" = parity ]; then timeout 100 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2; fi
timeout 60 python bench.py --no-cpu --steps 8 --warmup 3 > gpurun_out/q2.json 2> gpurun_out/q2.err; show cfg2 gpurun_out/q2.json
timeout 120 python bench.py --config 3 --no-cpu --steps 2 --warmup 1 > gpurun_out/q3.json 2> gpurun_out/q3.err; show cfg3 gpurun_out/q3.json
