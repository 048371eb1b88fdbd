#!/bin/bash
# round 2, two-GPU pass: exchange test across GPUs, strong+weak bench at N=2 (ours and the reference arm), copy ceiling at 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -5
python tools/ubench/png_device.py > gpurun_out/r2b_png_device.json 2> gpurun_out/r2b_png_device.err; cat gpurun_out/r2b_png_device.json; tail -3 gpurun_out/r2b_png_device.err
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err
tail -c 1500 gpurun_out/r2b_bench2.err
$TR bench.py --gpus 2 --steps 2 --warmup 1 --impl reference --ref-frames 30 > gpurun_out/r2b_ref2.json 2> gpurun_out/r2b_ref2.err
cat gpurun_out/r2b_ref2.json | cut -c1-400
$TR tools/ubench/d2h_bw.py > gpurun_out/r2b_d2h_2.json 2> gpurun_out/r2b_d2h_2.err; cat gpurun_out/r2b_d2h_2.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_bench2.json").read().strip().splitlines()[-1])
print("strong",d["value"],"ms",d["ms_per_step"],d["ms_per_step_by_rank"],"verified",d["gathered_frames_verified"] is not None)
print("e2e",d["e2e"]["value"],d["e2e"]["d2h_bytes_per_step"],"raw",d["e2e_raw"]["value"])
print("weak",d.get("weak",{}).get("value"))
PY
