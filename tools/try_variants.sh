#!/bin/bash
# usage (GPU box): tools/try_variants.sh "<bench args>" name...  -> one short bench per build/variants/<name>.so, pass/fail + stages
args=$1; shift
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
for name in "$@" base; do
  if [ $name = base ]; then cp /tmp/lib_orig.so $lib; else cp build/variants/$name.so $lib; fi
  if timeout 120 python bench.py --no-cpu $args > gpurun_out/tv_$name.json 2> gpurun_out/tv_$name.err; then
    python - $name <<PY
import json, sys
d = json.loads(open(f"gpurun_out/tv_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), " ".join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d["stages"].items()))
PY
  else echo "$name FAILED: $(tail -1 gpurun_out/tv_$name.err | cut -c1-200)"; fi
done
cp /tmp/lib_orig.so $lib
