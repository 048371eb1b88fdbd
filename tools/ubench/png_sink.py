"""Frame sink on the host cores (no GPU): ms per 512x512 frame of render_surgery.write_frames_png against PIL at
compress_level=1 (the sink's first version) and PIL's default level 6 (what torchvision.utils.save_image, the
upstream renderer's writer, uses).  Frames: the bench scene rendered by the CPU oracle (checker only).
usage: python tools/ubench/png_sink.py [n_frames=64] [workers=all]"""
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa
from omfs_b200 import avatar, render_surgery as rs, synthetic
import oracle
from PIL import Image

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
workers = int(sys.argv[2]) if len(sys.argv) > 2 else len(os.sched_getaffinity(0))
model, params, av, cam = synthetic.make_scene(n_gauss=100_000, n_frames=8, width=512, height=512)
u8 = oracle.to_uint8(oracle.render(model, params, avatar.bake(av), [cam.pack()] * 8, 512, 512).image)
frames = np.concatenate([u8] * ((n + 7) // 8))[:n]
d = tempfile.mkdtemp()


def timed(label, fn):
    paths = [os.path.join(d, f"{label}_{i:05d}.png") for i in range(n)]
    t = time.perf_counter()
    with ThreadPoolExecutor(workers) as pool:
        list(pool.map(fn, paths, frames))
    dt = time.perf_counter() - t
    kb = sum(os.path.getsize(p) for p in paths) / n / 1024
    print(f"{label:28s} {dt / n * 1e3:7.2f} ms/frame  {kb:7.1f} KB/frame  ({workers} threads, {n} frames)")


timed("encode_png (up + rle, lvl 1)", rs._write_png)
timed("PIL compress_level=1", lambda p, f: Image.fromarray(f, mode="RGB").save(p, compress_level=1))
timed("PIL default (level 6)", lambda p, f: Image.fromarray(f, mode="RGB").save(p))
