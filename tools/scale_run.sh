#!/bin/bash
# usage (GPU box with 8 GPUs): tools/scale_run.sh <tag>  -> gpurun_out/<tag>_scale_N.json for N = 1 2 4 8
tag=$1
python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/${tag}_scale_1.json 2> gpurun_out/${tag}_scale_1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_scale_$N.json 2> gpurun_out/${tag}_scale_$N.err || tail -5 gpurun_out/${tag}_scale_$N.err
done
python - <<PY
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/${tag}_scale_{N}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(N,"failed",e); continue
    if N==1: base=d["value"]; be=d["e2e"]["value"]
    print(N, "value", round(d["value"]), "x%.2f"%(d["value"]/base), "e2e", round(d["e2e"]["value"]), "x%.2f"%(d["e2e"]["value"]/be), "blocking", round(d["e2e_blocking"]["value"]), "ms/step", round(d["ms_per_step"],3), d["config"].get("gather"), [round(x,2) for x in d["ms_per_step_by_rank"]])
PY
