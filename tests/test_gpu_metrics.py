"""R9 on device: omfs_frame_metrics against the oracle's moments and the reference's report golden."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime
    runtime.load_library()
    return runtime


def _moments(rt, a, b):
    L = rt.load_library()
    T, H, W = a.shape[:3]
    d_a, d_b = rt.DeviceArray.from_numpy(a), rt.DeviceArray.from_numpy(b)
    d_m = rt.DeviceArray((T, 6), np.float64)
    rt.check(L.omfs_frame_metrics(T, H, W, d_a.ptr, d_b.ptr, d_m.ptr, None))
    return d_m.numpy()


@pytest.mark.parametrize("shape", [(12, 20, 24), (1, 7, 9), (3, 33, 16), (5, 512, 512), (2, 1, 4)])
def test_frame_moments_match_oracle(rt, shape):
    """Squared-error sum bit-exact (integers), luma moments to 1e-12 relative (float64 sums in another order);
    ragged sizes: a pixel count that is not a multiple of 4 (single frame), one row, one group."""
    from oracle import reference_rows as rr
    rng = np.random.default_rng(sum(shape))
    T, H, W = shape
    a = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int64) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    b[0] = a[0]                                            # an identical pair
    got, want = _moments(rt, a, b), rr.frame_moments(a, b)
    assert np.array_equal(got[:, 0], want[:, 0])
    assert got[0, 0] == 0.0
    np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=1e-12, atol=0)


def test_frame_metrics_argument_errors(rt):
    L = rt.load_library()
    d = rt.DeviceArray((2, 3, 3, 3), np.uint8)
    d_m = rt.DeviceArray((2, 6), np.float64)
    assert L.omfs_frame_metrics(2, 3, 3, d.ptr, d.ptr, d_m.ptr, None) != 0     # 27 bytes per frame: unaligned frames
    assert b"multiple of 4" in L.omfs_last_error()
    assert L.omfs_frame_metrics(1, 3, 3, d.ptr, None, d_m.ptr, None) != 0
    assert L.omfs_frame_metrics(0, 3, 3, d.ptr, d.ptr, d_m.ptr, None) == 0      # empty: nothing to do


def test_device_report_matches_reference_golden(rt, golden_dir, tmp_path):
    """generate_report_device on the golden frame sets (uploaded to HBM) writes the report the reference wrote
    from the PNG files: same rows, buckets and summary; PSNR within 1e-5 dB, SSIM within 1e-12."""
    from omfs_b200 import validation_reporting as vrep
    g = np.load(os.path.join(golden_dir, "report_golden.npz"))
    renders, gt = np.ascontiguousarray(g["renders"]), np.ascontiguousarray(g["gt"])
    T, H, W = renders.shape[:3]
    d_a, d_b = rt.DeviceArray.from_numpy(renders), rt.DeviceArray.from_numpy(gt)
    rep = vrep.generate_report_device(d_a.ptr, d_b.ptr, T, H, W, [int(i) for i in g["selected"]], tmp_path / "rep")
    want = json.loads(str(g["report"]))
    assert json.load(open(tmp_path / "rep" / "strict_scores.json")) == rep
    assert (tmp_path / "rep" / "human_review_checklist.md").read_text() == str(g["checklist"])
    assert [(r["index"], r["frame"], r["bucket"], r["progress"]) for r in rep["rows"]] == \
           [(r["index"], r["frame"], r["bucket"], r["progress"]) for r in want["rows"]]
    for r, w in zip(rep["rows"], want["rows"]):
        assert abs(r["psnr"] - w["psnr"]) <= 1e-5 and abs(r["ssim"] - w["ssim"]) <= 1e-12
    for bk, w in want["summary"]["by_bucket"].items():
        r = rep["summary"]["by_bucket"][bk]
        assert r["count"] == w["count"] and abs(r["psnr"] - w["psnr"]) <= 1e-5 and abs(r["ssim"] - w["ssim"]) <= 1e-12


def test_plan_against_baseline_on_device(rt):
    """The A/B the deterministic export exists for: a 5 mm Le Fort plan against the zero-offset render of the same
    clip, both rendered into HBM, scored without leaving the device; checked against the host metrics."""
    from oracle import reference_rows as rr
    from omfs_b200 import avatar, render_surgery as rs, synthetic, validation_reporting as vrep
    T, W, H = 6, 128, 96
    model, params, av, cam = synthetic.make_scene(n_gauss=6000, n_frames=T, width=W, height=H, n_verts=642)
    baked = avatar.bake(av)
    sess = rt.Session(model, baked, W, H, max_batch=T)
    sess.set_subject(params.shape, params.static_offset)
    frames = []
    for mm in (0.0, 5.0):
        rec = rs._edit_record(params.as_dict(), rs.compute_offset(mm, 1.0), 0.0, None)
        p = synthetic.FrameParams.from_dict(rec, n_verts=model.n_verts)
        d_in = {k: rt.DeviceArray.from_numpy(np.ascontiguousarray(getattr(p, k), dtype=np.float32))
                for k in ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")}
        d_cam = rt.DeviceArray.from_numpy(cam.pack()[None])
        ptrs = {k: v.ptr for k, v in d_in.items()}
        ptrs["cams"] = d_cam.ptr
        out = rt.DeviceArray((T, H, W, 3), np.uint8)
        sess.render_device(ptrs, T, 1, d_out_u8=out.ptr)
        sess.sync()
        frames.append(out)
    p_dev, s_dev = vrep.frame_metrics_device(frames[0].ptr, frames[1].ptr, T, H, W)
    a, b = frames[0].numpy(), frames[1].numpy()
    assert (a != b).any()                                   # the plan moved the face
    for t in range(T):
        assert abs(p_dev[t] - rr.psnr(a[t].astype(np.float32), b[t].astype(np.float32))) <= 1e-5
        assert abs(s_dev[t] - rr.ssim_global(a[t].astype(np.float32), b[t].astype(np.float32))) <= 1e-12
    sess.close()
