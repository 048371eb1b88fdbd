#!/bin/bash
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
for name in base rs5 rs3 b256 b64 base; do
  if [ $name != base ]; then cp build/variants/$name.so $lib; else cp /tmp/lib_orig.so $lib; fi
  timeout 60 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/r2y_$name.json 2> gpurun_out/r2y_$name.err
  python - $name <<PY
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2y_{n}.json").read().strip().splitlines()[-1])
print(n,"value",round(d["value"]),{k:round(v["ms_per_launch"],4) for k,v in d["stages"].items() if k in ("bind_preprocess","depth_sort","emit_scatter")})
PY
done
cp /tmp/lib_orig.so $lib
