"""Host side of the drop-in on the build container's cores (no GPU): `render_surgery.main()` and
`render_with_gaussians()` with the renderer replaced by a stand-in that hands over READY uint8 frames and READY PNG
streams at once — what the device path delivers (`csrc/png.cu` encodes on the GPU) — so that only the host work around
it is timed: parameter files, PLY, FLAME model, the in-memory plan edit, PNG files written, gt/ frames, the raw frames
piped to the video encoder (a `cat > /dev/null` stand-in for ffmpeg).
   python tools/ubench/host_side.py [frames] [width]  ->  one JSON line
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa: E402,F401
from omfs_b200 import cameras, flame_io, render_surgery as rs, synthetic  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    W = H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    N = 100_000
    tmp = tempfile.mkdtemp(prefix="omfs_host_")
    data, mdl = os.path.join(tmp, "data"), os.path.join(tmp, "model")
    model = synthetic.make_flame_model(seed=8)
    params = synthetic.make_frame_params(T, seed=9, n_verts=model.n_verts)
    av = synthetic.make_avatar(N, model.n_faces, seed=10)
    c2w = cameras.look_at_c2w((0.0, 0.0, 1.0), (0.0, 0.0, 0.0))
    flame_io.write_synthetic_dataset(data, mdl, model, params, av, c2w, 0.3, W, H, iteration=3000)
    n_train = len(flame_io.load_transforms(data, "train"))
    # ready-made output of the renderer: a smooth synthetic clip and its PNG streams (host encoder of the same format)
    yy, xx = np.mgrid[0:H, 0:W]
    base = np.stack([(xx * 255 // W), (yy * 255 // H), ((xx + yy) * 255 // (W + H))], -1).astype(np.uint8)
    frames = np.stack([np.roll(base, 3 * t, axis=1) for t in range(n_train)])
    pngs = [rs.encode_png(f) for f in frames]

    def fake_render(model_, params_, av_, cams, plan_offset=None, device=None, want_png=False, want_u8=True, on_pngs=None):
        got = list(pngs[: params_.n_frames])
        if want_png and on_pngs is not None:   # streamed form: clips of 64 frames arrive one by one
            for a in range(0, len(got), 64):
                on_pngs(a, got[a:a + 64])
            got = []
        return (frames[: params_.n_frames] if want_u8 else None, got) if want_png else frames[: params_.n_frames]

    fake_ffmpeg = os.path.join(tmp, "ffmpeg")
    with open(fake_ffmpeg, "w") as f:
        f.write("#!/bin/sh\ncat > /dev/null\n")
    os.chmod(fake_ffmpeg, 0o755)
    rs._render_frames = fake_render
    rs._get_ffmpeg_path = lambda: fake_ffmpeg
    os.environ.pop("WORLD_SIZE", None)
    out = {"frames": n_train, "width": W, "height": H, "gaussians": N, "host_threads": len(os.sched_getaffinity(0)),
           "png_bytes": sum(len(p) for p in pngs)}
    argv = ["--lefort_mm", "5", "--bsso_mm", "-3", "--model_path", mdl, "--data_dir", data,
            "--output", os.path.join(tmp, "video", "final.mp4"), "--fps", "24"]
    stdout = sys.stdout
    for name, fn in (("main_s", lambda: rs.main(argv)),
                     ("render_with_gaussians_s", lambda: rs.render_with_gaussians(mdl, data, iteration=3000))):
        best = []
        for _ in range(3):
            sys.stdout = open(os.devnull, "w")
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            sys.stdout = stdout
            best.append(round(dt, 3))
        out[name] = best
    # the parts, once each
    t0 = time.perf_counter(); fr = flame_io.load_transforms(data, "train"); out["load_transforms_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); flame_io.load_dataset_params(data, fr, model.n_verts); out["load_dataset_params_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); flame_io.load_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_3000", "point_cloud.ply")); out["load_avatar_ply_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); flame_io.load_flame_model(rs._find_flame_model(mdl)); out["load_flame_model_s"] = round(time.perf_counter() - t0, 3)
    rd = os.path.join(tmp, "w"); t0 = time.perf_counter(); rs.write_png_files(rd, pngs, first=0); out["write_png_files_s"] = round(time.perf_counter() - t0, 3)
    gd = os.path.join(tmp, "g"); t0 = time.perf_counter(); rs.write_gt_frames(gd, data, fr, first=0); out["write_gt_frames_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); rs.stitch_video_frames(frames, os.path.join(tmp, "v.mp4"), fps=24); out["stitch_video_frames_s"] = round(time.perf_counter() - t0, 3)
    print(json.dumps(out))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
