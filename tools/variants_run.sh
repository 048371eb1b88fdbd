#!/bin/bash
# usage (GPU box): tools/variants_run.sh   -> runs the bench with every build/variants/*.so in place of the library
cp omfs-4d-video-gen_b200/libomfs_b200.so /tmp/lib_orig.so
for so in build/variants/*.so; do
  name=$(basename $so .so)
  cp $so omfs-4d-video-gen_b200/libomfs_b200.so
  python bench.py --no-cpu --steps 5 --warmup 2 > /tmp/v.json 2> /tmp/v.err || { echo "$name FAILED"; tail -3 /tmp/v.err; continue; }
  python - "$name" <<PY
import json,sys
d=json.loads(open("/tmp/v.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), " ".join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d["stages"].items()))
PY
done
cp /tmp/lib_orig.so omfs-4d-video-gen_b200/libomfs_b200.so
