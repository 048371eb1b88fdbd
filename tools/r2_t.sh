#!/bin/bash
# A/B of emit_scatter variants on one box: stage tables only
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
for name in base ballot_c100 atomic_cdef ballot_cdef; do
  if [ $name != base ]; then cp build/variants/$name.so $lib; else cp /tmp/lib_orig.so $lib; fi
  timeout 60 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err
  python - $name <<PY
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2t_{n}.json").read().strip().splitlines()[-1])
print(n,"value",round(d["value"]),{k:round(v["ms_per_launch"],4) for k,v in d["stages"].items()})
PY
done
cp /tmp/lib_orig.so $lib
