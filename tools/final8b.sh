#!/bin/bash
# round 2, last 8-GPU pass (reduced: GPU-minutes are charged x8): configs[2] strong at 8 GPUs with the copy-engine
# exchange and with the fused one.  gpurun_out/<tag>_scale8_{p2p,fused}.json
tag=${1:-r2f}
for g in p2p fused; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 \
      bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-weak --gather $g > gpurun_out/${tag}_scale8_$g.json 2> gpurun_out/${tag}_scale8_$g.err || tail -5 gpurun_out/${tag}_scale8_$g.err
  python - $g <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/${tag}_scale8_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "blocking", round(d["e2e_blocking"]["value"]),
          "ms/step", round(d["ms_per_step"], 3), "verified", d.get("gathered_frames_verified") is not None,
          [round(x, 3) for x in d["ms_per_step_by_rank"]])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
