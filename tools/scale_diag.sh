#!/bin/bash
# usage: tools/scale_diag.sh N   (N GPUs; diagnosis of the multi-GPU overhead)
N=$1
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu $EXTRA > gpurun_out/diag_${N}_${tag}.json 2> gpurun_out/diag_${N}_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/diag_${N}_${tag}.json").read().strip().splitlines()[-1])
print("${tag}", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), [round(x,2) for x in d["ms_per_step_by_rank"]])
PY
}
EXTRA="" run default A=1
EXTRA="" run maxctas2 NCCL_MAX_CTAS=2
EXTRA="--no-gather" run nogather A=1
