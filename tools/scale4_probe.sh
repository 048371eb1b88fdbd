N=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29504 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/numa_4.json 2> gpurun_out/numa_4.err || tail -5 gpurun_out/numa_4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/numa_4.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"].get("host_numa"))
PY
nvidia-smi topo -m | head -14
