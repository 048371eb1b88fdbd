#!/bin/bash
bash tools/variant_ab.sh ballot
python bench.py --no-cpu --sync-steps --steps 5 > gpurun_out/r2s_sync.json 2> gpurun_out/r2s_sync.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_sync.json").read().strip().splitlines()[-1])
print("sync-steps value",round(d["value"]),{k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
