#!/bin/bash
# usage (GPU box): tools/ab_cfg3.sh <name>...  -> tools/variant_ab.sh-style A/B at configs[2] AND configs[3] for each
# build/variants/<name>.so, the committed library last.
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
run() {  # tag
  timeout 60 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  timeout 120 python bench.py --config 3 --no-cpu --steps 3 --warmup 1 > gpurun_out/ab3_$1.json 2> gpurun_out/ab3_$1.err
}
for name in "$@"; do
  cp build/variants/$name.so $lib
  run $name
  timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider > gpurun_out/ab_${name}_tests.log 2>&1
  tail -2 gpurun_out/ab_${name}_tests.log
done
cp /tmp/lib_orig.so $lib
run base
python - "$@" base <<PY
import json, sys
for n in sys.argv[1:]:
    for pre in ("ab", "ab3"):
        try:
            d = json.loads(open(f"gpurun_out/{pre}_{n}.json").read().strip().splitlines()[-1])
            print(pre, n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]),
                  " ".join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d["stages"].items()))
        except Exception as e:
            print(pre, n, "no result:", e)
PY
