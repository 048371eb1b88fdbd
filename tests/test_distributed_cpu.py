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


def _dropin_worker(rank, world, port, data, mdl, fail_rank):
    """render_with_gaussians under a torchrun-like launch (env only; the function sets up its own gloo group).  The
    GPU renderer is replaced by a stand-in that paints each frame with a value read from ITS parameters, so the
    test sees which rank rendered which frame from the files alone."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from omfs_b200 import render_surgery as rs

    def fake_render(model, params, av, cams, plan_offset=None, device=None, want_png=False, want_u8=True, on_pngs=None):
        if rank == fail_rank:
            raise ValueError("injected renderer failure")
        assert len(cams) == params.n_frames
        out = np.zeros((params.n_frames, cams[0].height, cams[0].width, 3), np.uint8)
        out[..., 0] = np.round(params.translation[:, 0] * 1000).astype(np.uint8)[:, None, None]   # frame id
        out[..., 1] = 10 + rank
        # the stand-in for the device sink: the host encoder of the same package
        pngs = [rs.encode_png(f) for f in out]
        if want_png and on_pngs is not None:   # the streamed form: clips arrive one by one, nothing is returned
            half = len(pngs) // 2
            on_pngs(0, pngs[:half])
            on_pngs(half, pngs[half:])
            pngs = []
        return (out if want_u8 else None, pngs) if want_png else out

    rs._render_frames = fake_render
    try:
        d = rs.render_with_gaussians(mdl, data)
        open(os.path.join(mdl, f"result_{rank}.txt"), "w").write("ok " + d)
    except Exception as e:
        open(os.path.join(mdl, f"result_{rank}.txt"), "w").write(f"{type(e).__name__} {e}")
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_render_with_gaussians_under_torchrun_world2(tmp_path, fail_rank):
    """Two ranks: stale renders purged once, every frame written exactly once under its global name by the rank
    that owns its block, both ranks return the same directory; a renderer failure on one rank raises on both."""
    from PIL import Image
    from omfs_b200 import cameras, flame_io, synthetic
    T, W, H, V = 7, 32, 24, 162
    model = synthetic.make_flame_model(seed=8, n_verts=V)
    params = synthetic.make_frame_params(T, seed=9, n_verts=V)
    params.translation[:, 0] = np.arange(T, dtype=np.float32) / 1000.0     # frame id, carried by the parameters
    av = synthetic.make_avatar(50, model.n_faces, seed=10)
    c2w = cameras.look_at_c2w((0.0, 0.0, 1.0), (0.0, 0.0, 0.0))
    data, mdl = str(tmp_path / "data"), str(tmp_path / "model")
    flame_io.write_synthetic_dataset(data, mdl, model, params, av, c2w, 0.3, W, H, iteration=3000)
    n_train = len(flame_io.load_transforms(data, "train"))
    stale = os.path.join(mdl, "train", "ours_1", "renders")
    os.makedirs(stale)
    open(os.path.join(stale, "00000.png"), "wb").write(b"stale")
    mp.spawn(_dropin_worker, args=(2, _free_port(), data, mdl, fail_rank), nprocs=2, join=True)
    results = [open(os.path.join(mdl, f"result_{r}.txt")).read() for r in range(2)]
    assert not os.path.exists(stale)
    renders = os.path.join(mdl, "train", "ours_3000", "renders")
    if fail_rank >= 0:
        assert results[1].startswith("RuntimeError Rendering failed:") and "injected renderer failure" in results[1]
        assert results[0].startswith("RuntimeError Rendering failed on rank 1") and "injected" in results[0]
        return
    assert results == ["ok " + renders] * 2
    assert sorted(os.listdir(renders)) == [f"{t:05d}.png" for t in range(n_train)]
    for t in range(n_train):
        img = np.asarray(Image.open(os.path.join(renders, f"{t:05d}.png")))
        assert img.shape == (H, W, 3) and (img[..., 0] == t).all()
        lo, hi = sharding.frame_block(n_train, 0, 2)
        assert (img[..., 1] == (10 if lo <= t < hi else 11)).all()
