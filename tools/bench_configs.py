"""Throughput of the OTHER workloads BASELINE.json names, per GPU (bench.py's contract is configs[2]):

  configs[3]  1024x1024 x 16 ring views per frame, 500k Gaussians   (rank's share of the 300 frames)
  configs[4]  plan sweep: BSSO plans x 120 frames, 512x512, 100k Gaussians (rank's share of the 64 plans)

    python tools/bench_configs.py [--frames3 8] [--plans 4] > gpurun_out/configs.json

Device-resident, CUDA-event timed (second pass; the first warms up and sizes the session), one JSON line per
config with segments/s (one segment = one rendered image) and tile pairs per segment.  The plan sweep renders
every plan with the reference's scalar edit (render_surgery.py:40-42, 119-139) applied to the jaw pose; one
blendshape GEMM call covers a plan's 120 frames.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa: E402,F401
from omfs_b200 import avatar, cameras, render_surgery as rs, runtime, synthetic  # noqa: E402

KEYS = ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")


def timed(sess, calls, reps=2):
    """calls: list of (device pointer dict, n_frames, n_views, out_ptr).  Returns ms of the last repetition."""
    import torch
    stream = torch.cuda.current_stream()
    ms = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ptrs, T, nv, out in calls:
            sess.render_device(ptrs, T, nv, d_out_u8=out, stream=stream.cuda_stream)
        e1.record(stream)
        sess.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames3", type=int, default=8, help="frames of configs[3] to render (x16 views)")
    ap.add_argument("--plans", type=int, default=4, help="plans of configs[4] to render (x120 frames)")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    runtime.check(runtime.load_library().omfs_device_check(0))

    def dev_params(p):
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k), dtype=np.float32)).to(dev) for k in KEYS}
        return t, {k: v.data_ptr() for k, v in t.items()}

    # ---- configs[3]
    W = H = 1024
    N, n_views, T = 500_000, 16, args.frames3
    model, params, av, _ = synthetic.make_scene(n_gauss=N, n_frames=T, width=W, height=H)
    cams = cameras.ring_cameras(n_views, synthetic.camera_distance(W, H), (0, 0, 0), 0.3, W, H)
    sess = runtime.Session(model, avatar.bake(av), W, H, max_batch=2 * n_views)
    sess.set_subject(params.shape, params.static_offset)
    sess.reserve_pairs(int(4.5e6) * 2 * n_views)  # ~3.7M tile pairs per 1024^2 segment (device calls do not auto-grow)
    keep, ptrs = dev_params(params)
    d_cams = torch.from_numpy(np.stack([c.pack() for c in cams])).to(dev)
    ptrs["cams"] = d_cams.data_ptr()
    out = torch.empty((T * n_views, H, W, 3), dtype=torch.uint8, device=dev)
    ms = timed(sess, [(ptrs, T, n_views, out.data_ptr())])
    S = T * n_views
    pairs = sess.stats()["pairs"] / S
    # per-stage CUDA-event times of one more pass (stages serialised on one stream: shares, not the pipelined total)
    sess.set_profiling(True)
    sess.render_device(ptrs, T, n_views, d_out_u8=out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    sess.sync()
    stages = {k: round(v["ms"], 4) for k, v in sess.stage_ms().items()}
    sess.set_profiling(False)
    print(json.dumps({"config": "configs[3]: 1024x1024 x 16 views, 500k Gaussians", "frames": T, "segments": S,
                      "ms": ms, "segments_per_s": S / ms * 1e3, "frames_per_s_all_views": T / ms * 1e3,
                      "tile_pairs_per_segment": pairs, "batch_segments": 2 * n_views,
                      "stage_ms_serialised": stages}), flush=True)
    sess.close()
    del out

    # ---- configs[4]
    W = H = 512
    T = 120
    model, params, av, cam = synthetic.make_scene(n_gauss=100_000, n_frames=T, width=W, height=H)
    sess = runtime.Session(model, avatar.bake(av), W, H, max_batch=60)
    sess.set_subject(params.shape, params.static_offset)
    plans_mm = np.linspace(-15.0, 15.0, 64)[: args.plans]
    d_cam = torch.from_numpy(cam.pack()[None]).to(dev)
    calls, keep_all = [], []
    outs = torch.empty((len(plans_mm), T, H, W, 3), dtype=torch.uint8, device=dev)
    for i, mm in enumerate(plans_mm):
        rec = rs._edit_record(params.as_dict(), 0.0, rs.compute_offset(float(mm), 1.0), None)
        keep, ptrs = dev_params(synthetic.FrameParams.from_dict(rec, n_verts=model.n_verts))
        ptrs["cams"] = d_cam.data_ptr()
        keep_all.append(keep)
        calls.append((ptrs, T, 1, outs[i].data_ptr()))
    ms = timed(sess, calls)
    S = len(plans_mm) * T
    print(json.dumps({"config": "configs[4]: plan sweep x 120 frames, 512x512, 100k Gaussians", "plans": len(plans_mm),
                      "segments": S, "ms": ms, "segments_per_s": S / ms * 1e3, "plans_per_s": len(plans_mm) / ms * 1e3,
                      "tile_pairs_per_segment": sess.stats()["pairs"] / T, "batch_segments": 60}), flush=True)
    sess.close()


if __name__ == "__main__":
    main()
