"""torch.ops.omfs.render on a B200: CUDA tensors in, a CUDA tensor out, bit for bit what the ctypes path
(omfs_session_render_host) returns, on the caller's current torch stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_torch_op_equals_ctypes_path():
    import torch
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, runtime, synthetic, torch_ops
    T, W, H = 9, 160, 112
    model, params, av, cam = synthetic.make_scene(n_gauss=6000, n_frames=T, width=W, height=H, n_verts=1202)
    cam2 = synthetic.make_scene(n_gauss=10, n_frames=1, width=W, height=H, n_verts=1202, seed=3)[3]
    baked = avatar.bake(av)
    with runtime.Session(model, baked, W, H, max_batch=4, device=0) as sess:
        sess.set_subject(params.shape, params.static_offset)
        want_u8, want_f32 = sess.render_host(params, [cam, cam2], want_u8=True, want_f32=True)
    dev = torch.device("cuda", 0)
    h = torch_ops.open_session(model, baked, W, H, max_batch=4, device=0)
    try:
        torch_ops.set_subject(h, torch.from_numpy(params.shape), torch.from_numpy(params.static_offset))
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(params, k), np.float32)).to(dev)
             for k in ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")}
        cams = torch.from_numpy(np.stack([cam.pack(), cam2.pack()]).astype(np.float32)).to(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # not the default stream: the op must follow the caller's stream
            frames = torch.ops.omfs.render(h, t["expr"], t["rotation"], t["neck_pose"], t["jaw_pose"], t["eyes_pose"],
                                           t["translation"], cams, None)
            brightest = frames.to(torch.int32).amax()          # torch work ordered after the op on the same stream
            image = torch.ops.omfs.render_image(h, t["expr"], t["rotation"], t["neck_pose"], t["jaw_pose"],
                                                t["eyes_pose"], t["translation"], cams, None)
        side.synchronize()
        torch_ops.check(h)
        assert frames.shape == (2 * T, H, W, 3) and frames.dtype == torch.uint8 and frames.device == dev
        assert np.array_equal(frames.cpu().numpy(), want_u8)
        assert int(brightest.item()) == int(want_u8.max())
        assert np.array_equal(image.cpu().numpy().view(np.uint32), want_f32.view(np.uint32))
        # wrong device / dtype / shape are named, not crashed on
        with pytest.raises(runtime.OmfsError, match="float32"):
            torch.ops.omfs.render(h, t["expr"].double(), t["rotation"], t["neck_pose"], t["jaw_pose"], t["eyes_pose"],
                                  t["translation"], cams, None)
        with pytest.raises(runtime.OmfsError, match="rotation has shape"):
            torch.ops.omfs.render(h, t["expr"], t["rotation"][:-1], t["neck_pose"], t["jaw_pose"], t["eyes_pose"],
                                  t["translation"], cams, None)
    finally:
        torch_ops.close_session(h)
