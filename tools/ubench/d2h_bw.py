"""Pure-copy ceiling of the end-to-end path: device->host bandwidth into pinned memory with N ranks copying at once
(one process per GPU, as bench.py runs), plus host->device for completeness.

    python tools/ubench/d2h_bw.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ubench/d2h_bw.py

Every rank copies a `--mb` MB device buffer into its own pinned host buffer `--reps` times back to back (CUDA events,
max over ranks); rank 0 prints ONE JSON line with the aggregate GB/s.  bench.py's e2e figures are to be read against
this number: frames/s ceiling = aggregate GB/s / bytes per frame.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=236)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import bench
    numa = bench.bind_to_gpu_numa_node(local) if world > 1 else None
    n = args.mb * 1_000_000
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, (dst, src) in (("d2h", (h, d)), ("h2d", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank = [float(x.item()) for x in allms]
        else:
            per_rank = [float(ms.item())]
        out[name] = {"aggregate_GBs": world * n * args.reps / (max(per_rank) * 1e-3) / 1e9,
                     "per_rank_GBs": [n * args.reps / (m * 1e-3) / 1e9 for m in per_rank]}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps, "host_numa": numa, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
