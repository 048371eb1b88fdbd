"""Device frame sink alone: time of omfs_png_encode over a batch of rendered frames (CUDA events, input resident),
achieved GB/s of input, compression ratio, next to the host encoders on the same frames.

    python tools/ubench/png_device.py [--frames 60] > gpurun_out/png_device.json
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa: E402,F401
from omfs_b200 import avatar, render_surgery as rs, runtime, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--gauss", type=int, default=100_000)
    args = ap.parse_args()
    import torch
    T, W, H = args.frames, args.size, args.size
    model, params, av, cam = synthetic.make_scene(n_gauss=args.gauss, n_frames=T, width=W, height=H)
    with runtime.Session(model, avatar.bake(av), W, H, max_batch=min(60, T)) as sess:
        sess.set_subject(params.shape, params.static_offset)
        frames, _ = sess.render_host(params, [cam], want_u8=True)
    L = runtime.load_library()
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(frames).to(dev)
    cap = int(L.omfs_png_max_bytes(W, H))
    ws = int(L.omfs_png_workspace_bytes(T, W, H))
    d_png = torch.empty(T * cap, dtype=torch.uint8, device=dev)
    d_off = torch.empty(T + 1, dtype=torch.int64, device=dev)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def run():
        runtime.check(L.omfs_png_encode(T, W, H, ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_png.data_ptr()),
                                        T * cap, ctypes.c_void_p(d_off.data_ptr()), ctypes.c_void_p(d_ws.data_ptr()), ws,
                                        ctypes.c_void_p(st.cuda_stream)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for _ in range(10):
        flush.zero_()   # evict the frames from L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        run()
        e1.record(st)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    total = int(d_off[T].item())
    ms = float(np.median(times))
    t0 = time.perf_counter()
    host = [rs.encode_png(f) for f in frames[:8]]
    host_ms = (time.perf_counter() - t0) / 8 * 1e3
    print(json.dumps({"frames": T, "size": [W, H], "ms_per_batch": ms, "us_per_frame": ms / T * 1e3,
                      "input_GBs": frames.nbytes / (ms * 1e-3) / 1e9, "png_bytes_per_frame": total / T,
                      "ratio": frames.nbytes / total, "host_encode_png_ms_per_frame_1_thread": host_ms,
                      "host_png_bytes_per_frame": float(np.mean([len(h) for h in host]))}), flush=True)


if __name__ == "__main__":
    main()
