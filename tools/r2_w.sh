#!/bin/bash
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
for name in base u1 u4 g2 g8 q128 base; do
  if [ $name != base ]; then cp build/variants/$name.so $lib; else cp /tmp/lib_orig.so $lib; fi
  timeout 60 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err
  python - $name <<PY
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2w_{n}.json").read().strip().splitlines()[-1])
print(n,"value",round(d["value"]),"composite",round(d["stages"]["composite"]["ms_per_launch"],4))
PY
done
cp /tmp/lib_orig.so $lib
