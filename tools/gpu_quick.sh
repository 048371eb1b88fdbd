#!/bin/bash
# usage: tools/gpu_quick.sh <tag>   (run on the GPU box through gpurun)
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 800 gpurun_out/${tag}_tests.log
tail -c 500 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
for k,v in d["stages"].items(): print(k,round(v["ms_per_launch"],4),round(v["share_of_step"],3),round(v["frac_of_hbm"],3))
PY
