#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
  OMFS_UNIT_ORDER=$v python bench.py --no-cpu > gpurun_out/r2m_o$v.json 2>gpurun_out/r2m_o$v.err
  OMFS_UNIT_ORDER=$v python bench.py --no-cpu --frames 38 --batch 19 > gpurun_out/r2m_small_o$v.json 2>gpurun_out/r2m_small_o$v.err
  python - <<PY
import json
for f in ("gpurun_out/r2m_o$v.json","gpurun_out/r2m_small_o$v.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("order=$v", f, "value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3), {k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
done
