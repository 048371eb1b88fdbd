"""Does a concurrent device->host copy stream slow the render kernels down?  (the synchronous host call
loses ~0.5 ms per 300 frames while its frame copies run.)  Renders the bench clip device-resident, with and without an
unrelated pinned D2H copy loop of the same volume (47 MB per 60-frame batch) on another stream."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import omfs_b200  # noqa
from omfs_b200 import avatar, runtime, synthetic

T, W, H = 300, 512, 512
model, params, av, cam = synthetic.make_scene(n_gauss=100_000, n_frames=T, width=W, height=H)
sess = runtime.Session(model, avatar.bake(av), W, H, max_batch=60)
sess.set_subject(params.shape, params.static_offset)
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(np.ascontiguousarray(getattr(params, k), dtype=np.float32)).to(dev)
     for k in ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")}
ptrs = {k: v.data_ptr() for k, v in d.items()}
d_cam = torch.from_numpy(cam.pack()[None]).to(dev)
ptrs["cams"] = d_cam.data_ptr()
out = torch.empty((T, H, W, 3), dtype=torch.uint8, device=dev)
src = torch.empty(47_185_920, dtype=torch.uint8, device=dev)
dst = torch.empty(47_185_920, dtype=torch.uint8).pin_memory()
main, side = torch.cuda.Stream(), torch.cuda.Stream()


def run(copies_per_step):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(5):
        sess.render_device(ptrs, T, 1, d_out_u8=out.data_ptr(), stream=main.cuda_stream)
        with torch.cuda.stream(side):
            for _ in range(copies_per_step):
                dst.copy_(src, non_blocking=True)
    e1.record(main)
    sess.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


run(0)
for n in (0, 5, 0, 5, 10):
    print(f"{n} x 47 MB D2H per 300-frame step: {run(n):.3f} ms per step")
