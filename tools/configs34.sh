#!/bin/bash
# configs[3] and configs[4] on one GPU, stage tables
for c in 3 4; do
  timeout 600 python bench.py --config $c --no-cpu --steps 3 --warmup 1 > gpurun_out/cfg_cfg$c.json 2> gpurun_out/cfg_cfg$c.err || tail -5 gpurun_out/cfg_cfg$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/cfg_cfg$c.json").read().strip().splitlines()[-1])
print("config $c", d["metric"], "value",round(d["value"],1),d["unit"],"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],3))
print(d["config"].get("workload"), "batch", d["config"].get("batch_segments"), "pairs/img", d["config"].get("tile_pairs_per_image"))
print({k:(round(v["ms_per_launch"],4), v["images_per_launch"]) for k,v in d["stages"].items()})
PY
done
