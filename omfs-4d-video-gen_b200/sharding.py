"""Frame / view / plan sharding across the GPUs of one box.

Every (plan, view, frame) triple is independent given the replicated model, so ranks never exchange
data on the path; the only collective is the final gather of finished uint8 frames to rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).
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
