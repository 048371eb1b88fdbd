#!/bin/bash
# round 2, final 8-GPU pass: 2-GPU exchange test, configs[2] strong (+weak) at 1/2/4/8, reference arm at 8, pure-copy
# ceilings at 1/2/4/8, configs[3] and configs[4] at 8.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2x}
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) "$@"; }
python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -3
bash tools/scale_run.sh $tag
TR 8 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference > gpurun_out/${tag}_ref8.json 2> gpurun_out/${tag}_ref8.err; cut -c1-300 gpurun_out/${tag}_ref8.json
python tools/ubench/d2h_bw.py > gpurun_out/${tag}_d2h_1.json 2> gpurun_out/${tag}_d2h_1.err
for n in 2 4 8; do TR $n tools/ubench/d2h_bw.py > gpurun_out/${tag}_d2h_$n.json 2> gpurun_out/${tag}_d2h_$n.err; done
for n in 1 2 4 8; do tail -1 gpurun_out/${tag}_d2h_$n.json | cut -c1-200; done
for c in 3 4; do
  TR 8 bench.py --gpus 8 --config $c --steps 3 --warmup 1 --no-cpu --no-weak > gpurun_out/${tag}_cfg${c}_8.json 2> gpurun_out/${tag}_cfg${c}_8.err || tail -3 gpurun_out/${tag}_cfg${c}_8.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_cfg${c}_8.json").read().strip().splitlines()[-1])
    print("config $c x8:", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "verified", d.get("gathered_frames_verified") is not None)
except Exception as e:
    print("config $c x8 failed", e)
PY
done
