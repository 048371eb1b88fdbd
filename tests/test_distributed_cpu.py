"""The N>1 path on CPU: frame sharding and the final gather with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import omfs_b200  # noqa: F401
from omfs_b200 import sharding


def test_frame_blocks_cover_every_frame_once():
    for n in (0, 1, 7, 300, 4800):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.frame_block(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
    assert sharding.frame_block(300, 7, 8) == (266, 300)        # config 3 on 8 GPUs: 38,38,...,34
    assert sharding.plan_block(64, 3, 8) == (24, 32)            # config 5: 8 plans per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.frame_block(n_total, rank, world)
    # stand-in for the rendered block: frame t is filled with (t % 251)
    local = torch.stack([torch.full((4, 6, 3), t % 251, dtype=torch.uint8) for t in range(lo, hi)]) \
        if hi > lo else torch.zeros((0, 4, 6, 3), dtype=torch.uint8)
    full = sharding.gather_frames(local, n_total, rank, world)
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 10])
def test_gather_frames_world2_gloo(tmp_path, n_total):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (n_total, 4, 6, 3)
    for t in range(n_total):
        assert np.all(got[t] == t % 251)
