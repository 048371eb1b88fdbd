"""The reference-facing entry points end to end on a B200: render_with_gaussians from the on-disk
formats, main()'s flow, SurgicalCutter and the plan displacement field."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dataset(tmp_path, T=5, W=96, H=80, N=2500, V=642):
    import omfs_b200  # noqa: F401
    from omfs_b200 import cameras, flame_io, synthetic
    model = synthetic.make_flame_model(seed=8, n_verts=V)
    params = synthetic.make_frame_params(T, seed=9, n_verts=V)
    av = synthetic.make_avatar(N, model.n_faces, seed=10)
    dist = synthetic.camera_distance(W, H)
    c2w = cameras.look_at_c2w((0.0, 0.0, dist), (0.0, 0.0, 0.0))
    data, mdl = str(tmp_path / "data"), str(tmp_path / "model")
    flame_io.write_synthetic_dataset(data, mdl, model, params, av, c2w, 0.3, W, H, iteration=3000)
    return data, mdl, model, params, av, cameras.camera_from_c2w(c2w, 0.3, W, H)


def test_render_with_gaussians_from_disk_matches_oracle(tmp_path):
    import oracle
    from oracle import reference_rows as rr
    from PIL import Image
    from omfs_b200 import avatar, render_surgery as rs, synthetic
    data, mdl, model, params, av, cam = _dataset(tmp_path)
    stale = os.path.join(mdl, "train", "ours_1", "renders")
    os.makedirs(stale)
    open(os.path.join(stale, "00000.png"), "wb").write(b"stale")
    lefort, bsso = rs.compute_offset(5.0, 1.0), rs.compute_offset(-4.0, 1.5)
    tmp = rs.create_modified_dataset(data, lefort, bsso)
    try:
        out_dir = rs.render_with_gaussians(mdl, tmp)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    assert out_dir == os.path.join(mdl, "train", "ours_3000", "renders")
    assert not os.path.exists(stale)                       # old renders are purged (reference :260-267)
    n_train = 5 - 5 // 10
    names = sorted(os.listdir(out_dir))
    assert names == [f"{i:05d}.png" for i in range(n_train)]
    got = np.stack([np.asarray(Image.open(os.path.join(out_dir, n))) for n in names])
    edited = synthetic.FrameParams.from_dict(rr.modify_flame_params(params.as_dict(), lefort, bsso), n_verts=model.n_verts)
    ref = oracle.render(model, edited.slice(0, n_train), avatar.bake(av), [cam.pack()] * n_train, cam.width, cam.height)
    want = oracle.to_uint8(ref.image)
    assert got.shape == want.shape
    assert rr.psnr(got.astype(np.float32), want.astype(np.float32)) > 50.0
    assert (np.abs(got.astype(int) - want.astype(int)) > 1).mean() < 2e-3
    # the in-memory path renders the same frames without the temporary dataset
    mem = rs.render_surgery_frames(model, params.slice(0, n_train), av, [cam] * n_train, 5.0, -4.0 * 1.5, 1.0)
    assert np.array_equal(mem, got)
    # the streamed path in several clips (two frames each, two in flight): the files do not change by a byte
    files = [open(os.path.join(out_dir, n), "rb").read() for n in names]
    old_chunk, rs.STREAM_CHUNK = rs.STREAM_CHUNK, 2
    tmp = rs.create_modified_dataset(data, lefort, bsso)
    try:
        out_dir = rs.render_with_gaussians(mdl, tmp)
    finally:
        rs.STREAM_CHUNK = old_chunk
        shutil.rmtree(tmp, ignore_errors=True)
    assert sorted(os.listdir(out_dir)) == names
    assert [open(os.path.join(out_dir, n), "rb").read() for n in names] == files
    # pinned iteration + deterministic export flow
    out2 = rs.render_with_gaussians(mdl, data, iteration=3000)
    exp = rs.export_deterministic_frames(out2, str(tmp_path / "ab"), None, max_frames=2)
    assert json.load(open(os.path.join(exp, "deterministic_indices_manifest.json")))["selected_indices"] == [0, n_train - 1]


def test_main_cli_flow(tmp_path, monkeypatch):
    """render_surgery.main with the reference's flags: parameter edit (in memory by default, through the reference's
    temporary dataset copy with OMFS_MATERIALISE_DATASET=1: same frames) -> in-process render -> PNGs in the upstream
    layout -> deterministic export -> encoder.  A stand-in ffmpeg records the raw frames it is handed: they are the
    frames of the PNGs, in order; the temporary dataset, when written, is removed afterwards."""
    from PIL import Image
    from omfs_b200 import render_surgery as rs
    data, mdl, model, params, av, cam = _dataset(tmp_path, T=4)
    fake = tmp_path / "ffmpeg"
    fake.write_text("#!/bin/sh\nfor a in \"$@\"; do echo \"$a\" >> %s; done\ncat > %s\n" %
                    (tmp_path / "args.txt", tmp_path / "stdin.bin"))
    fake.chmod(0o755)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(fake))
    made = []
    real_create = rs.create_modified_dataset
    monkeypatch.setattr(rs, "create_modified_dataset", lambda *a, **k: made.append(real_create(*a, **k)) or made[-1])
    out = tmp_path / "video" / "final_prediction.mp4"
    rs.main(["--lefort_mm", "5", "--bsso_mm", "-3", "--sensitivity", "1.5", "--model_path", mdl, "--data_dir", data,
             "--output", str(out), "--fps", "24", "--export_frames_dir", str(tmp_path / "ab"),
             "--deterministic_max_frames", "2"])
    renders = os.path.join(mdl, "train", "ours_3000", "renders")
    names = sorted(os.listdir(renders))
    assert names == [f"{i:05d}.png" for i in range(4)]
    frames = np.stack([np.asarray(Image.open(os.path.join(renders, n))) for n in names])
    assert (tmp_path / "stdin.bin").read_bytes() == frames.tobytes()
    args = (tmp_path / "args.txt").read_text().split("\n")
    assert args[args.index("-framerate") + 1] == "24" and args[args.index("-s") + 1] == f"{cam.width}x{cam.height}"
    assert json.load(open(tmp_path / "ab" / "deterministic_indices_manifest.json"))["selected_indices"] == [0, 3]
    assert made == []                                       # the plan was applied in memory: no dataset copy written
    # the reference's on-disk route (edited copy written, rendered, deleted: :503-539) renders the same frames
    monkeypatch.setenv("OMFS_MATERIALISE_DATASET", "1")
    rs.main(["--lefort_mm", "5", "--bsso_mm", "-3", "--sensitivity", "1.5", "--model_path", mdl, "--data_dir", data,
             "--output", str(out), "--fps", "24"])
    monkeypatch.delenv("OMFS_MATERIALISE_DATASET")
    assert made and not os.path.exists(made[0])            # the caller's temporary dataset is cleaned up (:537-539)
    again = np.stack([np.asarray(Image.open(os.path.join(renders, n))) for n in sorted(os.listdir(renders))])
    assert np.array_equal(again, frames)
    # the plan moved the face: frames differ from a zero-offset render of the same dataset
    zero = rs.render_surgery_frames(model, params, av, [cam] * 4, 0.0, 0.0)
    assert not np.array_equal(zero, frames)


def test_render_failure_is_a_runtime_error(tmp_path):
    from omfs_b200 import flame_io, render_surgery as rs
    data, mdl, model, params, av, cam = _dataset(tmp_path, T=2)
    bad = flame_io.load_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_3000", "point_cloud.ply"))
    os.environ["OMFS_RENDER_BATCH"] = "0"                  # an invalid session config -> library error
    try:
        with pytest.raises(RuntimeError, match="Rendering failed"):
            rs.render_with_gaussians(mdl, data)
    finally:
        del os.environ["OMFS_RENDER_BATCH"]
    bad.binding[:] = model.n_faces + 5
    flame_io.save_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_3000", "point_cloud.ply"), bad)
    with pytest.raises(ValueError):
        rs.render_with_gaussians(mdl, data)


def test_surgical_cutter_reference_cases(golden_dir):
    """/root/reference/test/test_surgical_sim.py:27-119 on the GPU-backed SurgicalCutter, plus the
    golden segment/move vectors produced by the reference's own move_segments."""
    from omfs_b200 import surgical_sim as ss
    g = np.load(os.path.join(golden_dir, "surgical_sim_golden.npz"))
    mk = lambda: ss.SurgicalCutter(ss.PointMesh(g["maxilla"]), ss.PointMesh(g["mandible"]))
    cutter = mk()
    with pytest.raises(RuntimeError):
        cutter.move_segments(maxilla_mm=5.0)
    res = cutter.perform_cut(lefort_z=20, bsso_l_x=-15, bsso_r_x=15)
    assert set(res) == {"upper_skull", "mobile_maxilla", "distal_mandible", "proximal_rami"}
    for k in res:
        assert res[k].n_points > 0
        assert np.array_equal(res[k].points.astype(np.float64), g[f"cut0_{k}"]), k
    c = lambda m: np.array(m.center)
    max0, mand0, skull0, rami0 = c(cutter.mobile_maxilla), c(cutter.distal_mandible), c(cutter.upper_skull), c(cutter.proximal_rami)
    moved = cutter.move_segments(maxilla_mm=10.0, mandible_mm=0.0)
    np.testing.assert_array_almost_equal(c(moved["distal_mandible"]), mand0)
    moved = cutter.move_segments(maxilla_mm=0.0, mandible_mm=10.0)
    np.testing.assert_array_almost_equal(c(moved["mobile_maxilla"]), max0)
    moved = cutter.move_segments(maxilla_mm=5.0, mandible_mm=8.0)
    np.testing.assert_almost_equal(c(moved["mobile_maxilla"])[1] - max0[1], 5.0, decimal=1)
    np.testing.assert_almost_equal(c(moved["distal_mandible"])[1] - mand0[1], 8.0, decimal=1)
    np.testing.assert_array_almost_equal(c(moved["upper_skull"]), skull0)
    np.testing.assert_array_almost_equal(c(moved["proximal_rami"]), rami0)
    d = c(cutter.move_segments(maxilla_mm=5.0, mandible_mm=0.0, advancement_direction=(1.0, 0.0, 0.0))["mobile_maxilla"]) - max0
    assert abs(d[0] - 5.0) < 0.05 and abs(d[1]) < 0.05 and abs(d[2]) < 0.05
    with pytest.raises(ValueError):
        cutter.move_segments(maxilla_mm=1.0, mandible_mm=1.0, advancement_direction=(0.0, 0.0, 0.0))
    assert abs(c(res["upper_skull"])[2] - c(res["mobile_maxilla"])[2]) > 0.1
    # the reference's own outputs for rotated planes and rotated/translated segments
    plans = json.loads(str(g["plans"]))
    for ci in (0, 1):
        cut = json.loads(str(g[f"cut{ci}_args"]))
        cutter = mk()
        cutter.perform_cut(**cut)
        for pi, plan in enumerate(plans):
            moved = cutter.move_segments(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in plan.items()})
            for k in ("mobile_maxilla", "distal_mandible"):
                np.testing.assert_allclose(moved[k].points, g[f"cut{ci}_plan{pi}_{k}"], rtol=0, atol=4e-6)
    single = ss.SurgicalCutter(ss.PointMesh(g["maxilla"]))
    assert single.perform_cut(lefort_z=20, bsso_l_x=-20, bsso_r_x=20)["upper_skull"].n_points > 0


def test_config1_lefort_advance_on_flame_mesh():
    """BASELINE.json configs[0]: a 5 mm Le Fort I advancement on a synthetic 5 023-vertex FLAME-sized
    mesh: masks bit-exact, moved vertices bit-exact against the float64 restatement, and the
    canonical-space displacement field renders through the session."""
    import omfs_b200  # noqa: F401
    import oracle
    from oracle import reference_rows as rr
    from omfs_b200 import avatar, runtime, surgical_sim as ss, synthetic
    model = synthetic.make_flame_model(seed=21, n_verts=5023)
    params = synthetic.make_frame_params(2, seed=22, n_verts=5023)
    canon = oracle.flame_forward(model, synthetic.FrameParams(
        params.shape, *[np.zeros((1, k), np.float32) for k in (100, 3, 3, 3, 6, 3)], params.static_offset,
        np.zeros((1, 5023, 3), np.float32)))[0]
    jaw_w = model.lbs_weights[:, 2].copy()
    c = (canon.min(0).astype(np.float64) + canon.max(0).astype(np.float64)) / 2
    planes = np.zeros((3, 8))
    planes[0, :3], planes[0, 3:6] = ss._angle_to_normal((0, 1, 0), 0.0, 0.0), (c[0], c[1] - 0.01, c[2])
    planes[1, :3], planes[1, 3:6] = ss._angle_to_normal((1, 0, 0), 0.0, 0.0), (c[0] - 0.04, c[1], c[2])
    planes[2, :3], planes[2, 3:6] = ss._angle_to_normal((1, 0, 0), 0.0, 0.0), (c[0] + 0.04, c[1], c[2])
    field, mask = ss.plan_displacement_field(canon, jaw_w, planes, maxilla_mm=5.0, mandible_mm=0.0,
                                             advancement_direction=(0.0, 0.0, 1.0))
    want_pts, want_mask, _ = rr.displace_points(canon, planes, rr.make_moves(5.0, 0.0, (0.0, 0.0, 1.0), unit_scale=1e-3),
                                                jaw_w > 0.5)
    assert np.array_equal(mask, want_mask)
    assert np.array_equal(field, (want_pts - canon).astype(np.float32))
    moved = (mask & 8) != 0
    assert moved.any() and (~moved).any()
    np.testing.assert_allclose(field[moved], [[0.0, 0.0, 0.005]] * int(moved.sum()), atol=1e-6)
    assert np.all(field[~moved & ((mask & 16) == 0)] == 0.0)
    av = synthetic.make_avatar(3000, model.n_faces, seed=23)
    baked = avatar.bake(av)
    cam = synthetic.make_camera(96, 96)
    with runtime.Session(model, baked, 96, 96, max_batch=2) as sess:
        sess.set_subject(params.shape, params.static_offset, field)
        u8, img = sess.render_host(params, [cam], want_f32=True)
        verts = sess.tap_array("verts", (2, 5023, 3), np.float32)
    ref = oracle.render(model, params, baked, [cam.pack()] * 2, 96, 96, plan_offset=field)
    assert np.abs(verts - ref.verts).max() <= 1e-5
    ref2 = oracle.render(model, params, baked, [cam.pack()] * 2, 96, 96, verts=verts)
    assert np.abs(img - ref2.image).max() <= 2e-4


def test_config2_single_frame_experiment_driver(tmp_path):
    """BASELINE.json configs[1] through the reference's driver shape (single_frame_experiment.py): frame 0 of a
    512x512, 100k-Gaussian dataset on disk -> one-frame dataset -> render with zero offsets -> render + GT PNGs;
    the rendered PNG against the oracle."""
    import oracle
    from oracle import reference_rows as rr
    from PIL import Image
    from omfs_b200 import avatar, cameras, flame_io, single_frame_experiment as sfe, synthetic
    W = H = 512
    model, params, av, cam = synthetic.make_scene(n_gauss=100_000, n_frames=2, width=W, height=H)
    c2w = cameras.look_at_c2w((0.0, 0.0, synthetic.camera_distance(W, H)), (0.0, 0.0, 0.0))
    data_conda, mdl = tmp_path / "data_conda", tmp_path / "model_single_frame"
    flame_io.write_synthetic_dataset(str(data_conda), str(mdl), model, params, av, c2w, 0.3, W, H, iteration=3000,
                                     write_images=True)
    single = sfe.build_single_frame_dataset(data_conda, tmp_path / "data_single_frame")
    sfe.train_single_frame(mdl)
    render_png, gt_png = sfe.render_single_frame_and_save(mdl, single, tmp_path / "out")
    assert render_png.name == "single_frame_render.png" and gt_png.name == "single_frame_gt.png"
    assert open(gt_png, "rb").read() == open(data_conda / "images" / "00000_00.png", "rb").read()
    got = np.asarray(Image.open(render_png))
    cam0 = cameras.camera_from_c2w(c2w, 0.3, W, H)
    ref = oracle.render(model, params.slice(0, 1), avatar.bake(av), [cam0.pack()], W, H)
    want = oracle.to_uint8(ref.image)[0]
    assert got.shape == want.shape == (H, W, 3)
    assert rr.psnr(got.astype(np.float32), want.astype(np.float32)) > 50.0
    assert (np.abs(got.astype(int) - want.astype(int)) > 1).mean() < 2e-3
    assert sorted(os.listdir(mdl / "train" / "ours_3000" / "renders")) == ["00000.png"]
