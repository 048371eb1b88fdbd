#!/bin/bash
# usage (GPU box): tools/variant_ab.sh <name>   -> bench + parity tests with build/variants/<name>.so in place of the
# library, then the bench with the committed library; every step bounded and logged under gpurun_out/ as it finishes.
name=$1
lib=omfs-4d-video-gen_b200/libomfs_b200.so
cp $lib /tmp/lib_orig.so
cp build/variants/$name.so $lib
timeout 14 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
timeout 16 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider > gpurun_out/ab_${name}_tests.log 2>&1
tail -3 gpurun_out/ab_${name}_tests.log
cp /tmp/lib_orig.so $lib
timeout 14 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
python - "$name" <<PY
import json, sys
for n in (sys.argv[1], "base"):
    try:
        d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]),
              " ".join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d["stages"].items()))
    except Exception as e:
        print(n, "no result:", e)
PY
