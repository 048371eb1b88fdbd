"""Frame / view / plan sharding across the GPUs of one box.

Every (plan, view, frame) triple is independent given the replicated model, so ranks never exchange
data on the path; the only exchange is the gather of finished uint8 frames on rank 0.  Two forms:

* `gather_frames` — one collective (NCCL on GPUs, gloo in the CPU tests);
* `PeerFrameGather` — the B200 form: rank 0 exports its receive buffer over CUDA IPC and every rank
  pushes its block straight into its slot with a copy-engine peer-to-peer copy over NVLink
  (`omfs_push_frames`).  No SM is taken from the rendering kernels, so the exchange of one clip
  overlaps the rendering of the next; NCCL's send/recv kernels reached 127 GB/s into rank 0 at 8 GPUs
  and did not overlap (gpurun_out/diag_8_*.json, profiles/).
"""
from __future__ import annotations

import numpy as np


def frame_block(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(n/world) frames for `rank` (config 3: 300 frames over 1/2/4/8)."""
    per = -(-n_frames // world)
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def plan_block(n_plans: int, rank: int, world: int) -> tuple[int, int]:
    """Config 5: the 64 x 120 (plan, frame) grid is sharded by plan."""
    return frame_block(n_plans, rank, world)


def gather_frames(local_u8, n_total: int, rank: int, world: int, dst: int = 0):
    """Gather per-rank frame blocks [n_local,H,W,3] (uint8 torch tensors, same device) into
    [n_total,H,W,3] on rank `dst`.  Blocks are padded to the common ceil(n/world) length so that one
    gather moves everything; returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_u8
    per = -(-n_total // world)
    pad = torch.zeros((per,) + tuple(local_u8.shape[1:]), dtype=local_u8.dtype, device=local_u8.device)
    pad[: local_u8.shape[0]] = local_u8
    if rank == dst:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, parts, dst=dst)
        return torch.cat(parts, dim=0)[:n_total]
    dist.gather(pad, None, dst=dst)
    return None


_HOST_GROUP = None


def host_ranks():
    """(rank, world, agree) of the torchrun launch this process belongs to; (0, 1, agree) in a plain process.

    `agree(error)` is the host-side meeting point of the file-based drop-in (`render_with_gaussians` under
    torchrun: every rank renders its frame block and writes its own PNGs, so no frame crosses ranks): every rank
    passes `None` or the text of the exception it caught, and gets back the list of all ranks' entries — a rank
    that failed makes every rank raise instead of leaving the others waiting at a barrier.  It runs on a gloo
    group (CPU), created once from the torchrun environment if the caller has not initialised
    torch.distributed."""
    import os
    global _HOST_GROUP
    import sys
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # a launcher's full environment, or a group the caller already set up; a stray WORLD_SIZE alone is not a launch
    # (and a plain process never pays for importing torch here)
    launched = world > 1 and "RANK" in os.environ and "MASTER_PORT" in os.environ
    dist = sys.modules.get("torch.distributed")
    grouped = dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if not launched and not grouped:
        return 0, 1, (lambda error=None: [error])
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else dist.group.WORLD
    group = _HOST_GROUP

    def agree(error=None):
        entries = [None] * dist.get_world_size(group)
        dist.all_gather_object(entries, error, group=group)
        return entries

    return dist.get_rank(), dist.get_world_size(), agree


class PeerFrameGather:
    """Finished frames of every rank -> one [world, slot_bytes] device buffer on rank `dst`.

    `exchange(objs)` must all-gather small picklable objects across ranks (torch.distributed's
    all_gather_object bound to the process group, or any other out-of-band channel); it is used once,
    to hand the root's IPC handle to the peers.  `push(src_ptr, nbytes, stream)` enqueues this rank's
    copy on `stream` (ordered after the rendering by the caller's events); the data is complete on the
    root once every rank has synchronised that stream and the ranks have met at a barrier.
    `slot_ptr()` is also a valid uint8 output pointer for `Session.render_device`: the compositing kernel then
    stores its pixels straight into the root's slot over NVLink and no local frame buffer or copy exists
    (`bench.py --gather fused`).
    """

    def __init__(self, slot_bytes: int, rank: int, world: int, exchange, dst: int = 0):
        from . import runtime
        self.rt = runtime
        self.rank, self.world, self.dst, self.slot_bytes = rank, world, dst, slot_bytes
        self.buffer = None      # DeviceArray on the root
        self.base = 0           # device pointer of the root's buffer as seen from this process
        handle = None
        if rank == dst:
            self.buffer = runtime.DeviceArray((world, slot_bytes), np.uint8)
            self.base = self.buffer.ptr
            handle = runtime.ipc_export(self.base)
        handles = exchange(handle)
        if rank != dst:
            self.base = runtime.ipc_open(handles[dst])

    def slot_ptr(self, rank: int | None = None) -> int:
        return self.base + (self.rank if rank is None else rank) * self.slot_bytes

    def push(self, src_ptr: int, nbytes: int, stream: int = 0) -> None:
        if nbytes > self.slot_bytes:
            raise ValueError("frame block larger than its slot")
        self.rt.push_frames(self.slot_ptr(), src_ptr, nbytes, stream)

    def numpy(self) -> np.ndarray:
        """Root only: the gathered buffer [world, slot_bytes] (synchronises the device)."""
        return self.buffer.numpy()

    def close(self) -> None:
        if self.rank != self.dst and self.base:
            self.rt.ipc_close(self.base)
            self.base = 0
        if self.buffer is not None:
            self.buffer.free()
            self.buffer = None
