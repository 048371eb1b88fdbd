"""The host-side mirror of the reference interface (render_surgery.py / surgical_sim.py names,
arguments, errors, on-disk formats) — CPU only.  The cases follow the reference's own unit tests
(/root/reference/test/test_render_surgery.py) so they read the same way."""
import json
import os

import numpy as np
import pytest

import omfs_b200  # noqa: F401
from omfs_b200 import flame_io, render_surgery as rs, surgical_sim as ss, synthetic, validation_reporting


def test_compute_offset_cases():
    assert rs.compute_offset(0.0, 1.0) == 0.0
    assert rs.compute_offset(5.0, 1.0) == pytest.approx(5.0 * 1.0 * rs.SCALE_FACTOR)
    assert rs.compute_offset(-3.0, 1.0) == pytest.approx(-3.0 * rs.SCALE_FACTOR)
    assert rs.compute_offset(5.0, 2.5) == pytest.approx(5.0 * 2.5 * rs.SCALE_FACTOR)
    assert rs.compute_offset(10.0, 0.0) == 0.0


def test_compute_offset_and_modify_match_reference_goldens(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "render_surgery_golden.npz"))
    for (mm, s), want in zip(g["offset_cases"], g["offset_values"]):
        assert rs.compute_offset(float(mm), float(s)) == float(want)
    maps = json.loads(str(g["mod_maps"]))
    for kind in ("batched", "single"):
        src = tmp_path / f"{kind}.npz"
        np.savez(src, **{k: g[f"mod_{kind}_in_{k}"] for k in ("jaw_pose", "translation", "expr", "shape")})
        for i, ((lo, bo), dm) in enumerate(zip(g["mod_args"], maps)):
            dst = tmp_path / f"{kind}_{i}.npz"
            rs.modify_flame_params(str(src), str(dst), float(lo), float(bo), deformation_map=dm)
            got = np.load(dst)
            for k in ("jaw_pose", "translation", "expr", "shape"):
                want = g[f"mod_{kind}_{i}_{k}"]
                assert np.array_equal(got[k].view(np.uint32), want.view(np.uint32)), (kind, i, k)
        again = np.load(src)
        assert np.array_equal(again["translation"], g[f"mod_{kind}_in_translation"])   # source untouched


def test_modify_flame_params_reference_cases(tmp_path):
    src, dst = tmp_path / "source.npz", tmp_path / "modified.npz"
    np.savez(src, jaw_pose=np.zeros((10, 3), np.float32), translation=np.zeros((10, 3), np.float32),
             expr=np.zeros((10, 100), np.float32), shape=np.zeros(300, np.float32))
    rs.modify_flame_params(str(src), str(dst), 0.005, 0.0)
    assert float(np.load(dst)["translation"][0, 1]) == pytest.approx(0.005, abs=1e-5)
    rs.modify_flame_params(str(src), str(dst), 0.0, 0.003)
    assert float(np.load(dst)["jaw_pose"][0, 0]) == pytest.approx(0.003, abs=1e-5)
    rs.modify_flame_params(str(src), str(dst), 0.01, 0.02)
    assert float(np.load(src)["translation"][0, 1]) == 0.0 and float(np.load(src)["jaw_pose"][0, 0]) == 0.0
    dm = {"translation_axis": 2, "jaw_axis": 1, "lefort_scale": 2.0, "bsso_scale": 0.5}
    rs.modify_flame_params(str(src), str(dst), 0.01, 0.02, deformation_map=dm)
    d = np.load(dst)
    assert float(d["translation"][0, 2]) == pytest.approx(0.02, abs=1e-5)
    assert float(d["jaw_pose"][0, 1]) == pytest.approx(0.01, abs=1e-5)


def test_rig_mode_and_deformation_map(tmp_path):
    mode, reason = rs.choose_rig_mode("hybrid_full_head", "")
    assert mode == "flame_only" and "missing" in reason
    asset = tmp_path / "asset.npz"
    np.savez(asset, version=np.array([1]))
    assert rs.choose_rig_mode("hybrid_full_head", str(asset))[0] == "hybrid_full_head"
    assert rs.choose_rig_mode("flame_only", str(asset)) == ("flame_only", "explicitly requested")
    assert rs.load_deformation_map(None) == {}
    with pytest.raises(FileNotFoundError):
        rs.load_deformation_map(str(tmp_path / "nope.json"))
    bad = tmp_path / "bad.json"
    bad.write_text("[1, 2]")
    with pytest.raises(ValueError):
        rs.load_deformation_map(str(bad))


def test_deterministic_frame_export(tmp_path):
    from PIL import Image
    frames_dir, out_dir = tmp_path / "renders", tmp_path / "out"
    frames_dir.mkdir()
    for i in range(6):
        Image.fromarray(np.full((8, 8, 3), i * 20, dtype=np.uint8)).save(frames_dir / f"{i:05d}.png")
    idx = tmp_path / "idx.json"
    idx.write_text(json.dumps({"indices": [0, 3, 5]}))
    rs.export_deterministic_frames(str(frames_dir), str(out_dir), str(idx))
    manifest = json.loads((out_dir / "deterministic_indices_manifest.json").read_text())
    assert manifest["selected_indices"] == [0, 3, 5]
    for i in (0, 3, 5):
        assert (out_dir / f"idx_{i:05d}.png").exists()
    out2 = tmp_path / "out2"
    rs.export_deterministic_frames(str(frames_dir), str(out2), None, max_frames=3)
    assert json.loads((out2 / "deterministic_indices_manifest.json").read_text())["selected_indices"] == [0, 2, 5]
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        rs.export_deterministic_frames(str(empty), str(tmp_path / "o3"))
    bad = tmp_path / "bad_idx.json"
    bad.write_text(json.dumps({"indices": ["a"]}))
    with pytest.raises(ValueError):
        rs.export_deterministic_frames(str(frames_dir), str(tmp_path / "o4"), str(bad))


def test_create_modified_dataset_round_trip(tmp_path):
    model = synthetic.make_flame_model(n_verts=162)
    params = synthetic.make_frame_params(4, n_verts=162)
    av = synthetic.make_avatar(50, model.n_faces)
    cam_c2w = np.eye(4)
    cam_c2w[2, 3] = 1.0
    data, mdl = tmp_path / "data", tmp_path / "model"
    flame_io.write_synthetic_dataset(str(data), str(mdl), model, params, av, cam_c2w, 0.3, 64, 48, iteration=1234)
    tmp = rs.create_modified_dataset(str(data), rs.compute_offset(5.0, 1.0), rs.compute_offset(2.0, 1.0))
    try:
        for t in range(4):
            a = np.load(os.path.join(tmp, "flame_param", f"{t:05d}.npz"))
            np.testing.assert_allclose(a["translation"][0, 1], params.translation[t, 1] + np.float32(0.005), atol=1e-7)
            np.testing.assert_allclose(a["jaw_pose"][0, 0], params.jaw_pose[t, 0] + np.float32(0.002), atol=1e-7)
            assert np.array_equal(a["expr"], params.expr[t:t + 1])
        tr = json.load(open(os.path.join(tmp, "transforms_train.json")))
        assert all(f["flame_param_path"] == f"flame_param/{f['timestep_index']:05d}.npz" for f in tr["frames"])
        assert os.path.exists(os.path.join(tmp, "canonical_flame_param.npz"))
        frames = flame_io.load_transforms(tmp, "train")
        got = flame_io.load_dataset_params(tmp, frames, 162)
        assert got.n_frames == len(frames) == 4 - 4 // 10
    finally:
        import shutil
        shutil.rmtree(tmp)
    # PLY round trip incl. binding indices (bit-exact) and SH layout
    av2 = flame_io.load_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_1234", "point_cloud.ply"))
    assert np.array_equal(av2.binding, av.binding)
    for k in ("xyz", "scaling", "rotation", "opacity", "sh"):
        assert np.array_equal(getattr(av2, k), getattr(av, k)), k
    m2 = flame_io.load_flame_model(os.path.join(mdl, "flame_model.npz"))
    assert np.array_equal(m2.shapedirs, model.shapedirs) and np.array_equal(m2.faces, model.faces)


def test_render_with_gaussians_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        rs.render_with_gaussians(str(tmp_path / "model"), str(tmp_path / "data"))
    # no PNG frames (or no ffmpeg at all): FileNotFoundError either way, as in the reference
    with pytest.raises(FileNotFoundError):
        rs.stitch_video(str(tmp_path), str(tmp_path / "o.mp4"))


def test_surgical_sim_host_pieces(golden_dir):
    g = np.load(os.path.join(golden_dir, "surgical_sim_golden.npz"))
    for bi, base in enumerate([(0, 0, 1), (1, 0, 0)]):
        for ai, (p, y) in enumerate(g["angles"]):
            assert np.array_equal(np.array(ss._angle_to_normal(base, float(p), float(y))), g["normals"][bi, ai])
    for d_in, want in zip(g["dirs_in"], g["dirs"]):
        assert np.array_equal(ss._normalise_direction(tuple(d_in)), want)
    with pytest.raises(ValueError):
        ss._normalise_direction((0.0, 0.0, 0.0))
    cutter = ss.SurgicalCutter(ss.PointMesh(g["maxilla"]), ss.PointMesh(g["mandible"]))
    with pytest.raises(RuntimeError):
        cutter.move_segments(maxilla_mm=5.0)
    keys = cutter.preview_planes(lefort_z=20, bsso_l_x=-15, bsso_r_x=15)
    for k in ("maxilla", "mandible", "combined", "lefort", "bsso_l", "bsso_r"):
        assert k in keys
    from oracle import reference_rows as rr
    assert np.allclose(ss._rotation(5.0, -3.0, 2.0), rr.rotation_xzy(5.0, -3.0, 2.0), atol=0)


def test_psnr_mirror(golden_dir):
    g = np.load(os.path.join(golden_dir, "psnr_golden.npz"))
    # the reference averages in float32; the moments are float64: the reference's own value to within its rounding
    assert abs(validation_reporting.psnr(g["a"], g["b"]) - float(g["psnr_ab"])) <= 1e-5
    assert validation_reporting.psnr(g["a"], g["a"]) == 99.0


def test_single_frame_dataset_matches_reference(golden_dir, tmp_path):
    """single_frame_experiment.build_single_frame_dataset against what the reference's own builder wrote for the
    same seeded dataset (tests/golden/make_golden.py: golden_single_frame): file listing, one-frame transforms for
    the three splits, batched flame_param.npz; then the guards of the other two steps."""
    import sys
    sys.path.insert(0, golden_dir)
    import make_golden
    from omfs_b200 import single_frame_experiment as sfe
    g = np.load(os.path.join(golden_dir, "single_frame_golden.npz"))
    data_conda = make_golden._tiny_dataset(str(tmp_path))
    single = sfe.build_single_frame_dataset(data_conda, tmp_path / "data_single_frame")
    listing = sorted(str(p.relative_to(single)) for p in single.rglob("*") if p.is_file())
    assert listing == json.loads(str(g["listing"]))
    want = json.loads(str(g["transforms"]))
    for split in ("train", "test", "val"):
        assert json.load(open(single / f"transforms_{split}.json")) == want[split]
    batched = dict(np.load(single / "flame_param.npz", allow_pickle=True))
    keys = sorted(k[len("batched_"):] for k in g.files if k.startswith("batched_"))
    assert sorted(batched) == keys
    for k in keys:
        assert batched[k].shape == g[f"batched_{k}"].shape and np.array_equal(batched[k], g[f"batched_{k}"])
    assert open(single / "images" / "00000_00.png", "rb").read() == open(
        os.path.join(data_conda, "images", "00000_00.png"), "rb").read()
    # rebuilding replaces the directory; an incomplete transforms file is a KeyError, as upstream
    (single / "stale.txt").write_text("x")
    sfe.build_single_frame_dataset(data_conda, single)
    assert not (single / "stale.txt").exists()
    t = json.load(open(os.path.join(data_conda, "transforms_train.json")))
    del t["fl_x"]
    json.dump(t, open(os.path.join(data_conda, "transforms_train.json"), "w"))
    with pytest.raises(KeyError):
        sfe.build_single_frame_dataset(data_conda, single)
    with pytest.raises(RuntimeError, match="Training failed"):
        sfe.train_single_frame(tmp_path / "no_model")
    sfe.train_single_frame(tmp_path / "model")            # the synthetic "trained" avatar is accepted
    with pytest.raises(SystemExit):
        sfe.main(tmp_path / "missing_data_conda")


def test_frame_sink(tmp_path, monkeypatch):
    """write_frames_png writes the upstream names and round-trips the pixels; stitch_video_frames hands the raw
    frames to ffmpeg with the reference's codec flags (a stand-in ffmpeg records what it was given)."""
    from PIL import Image
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (7, 12, 16, 3), dtype=np.uint8)
    paths = rs.write_frames_png(str(tmp_path / "renders"), frames, workers=3)
    assert [os.path.basename(q) for q in paths] == [f"{i:05d}.png" for i in range(7)]
    for q, f in zip(paths, frames):
        assert np.array_equal(np.asarray(Image.open(q)), f)
    fake = tmp_path / "ffmpeg"
    fake.write_text("#!/bin/sh\nfor a in \"$@\"; do echo \"$a\" >> %s; done\ncat > %s\n" %
                    (tmp_path / "args.txt", tmp_path / "stdin.bin"))
    fake.chmod(0o755)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(fake))
    rs.stitch_video_frames(frames, str(tmp_path / "out" / "v.mp4"), fps=25)
    args = (tmp_path / "args.txt").read_text().split("\n")
    for flag, val in (("-s", "16x12"), ("-framerate", "25"), ("-c:v", "libx264"), ("-crf", "18"), ("-preset", "medium")):
        assert args[args.index(flag) + 1] == val
    assert args[args.index("-i") + 1] == "-" and args[-2] == str(tmp_path / "out" / "v.mp4")
    assert (tmp_path / "stdin.bin").read_bytes() == frames.tobytes()
    bad = tmp_path / "ffmpeg_bad"
    bad.write_text("#!/bin/sh\ncat > /dev/null\necho boom >&2\nexit 3\n")
    bad.chmod(0o755)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(bad))
    with pytest.raises(RuntimeError, match="(?s)ffmpeg failed.*boom"):
        rs.stitch_video_frames(frames, str(tmp_path / "v2.mp4"))
    # an encoder that dies before it has read its input (the pipe breaks under the one blocking write of a clip far
    # larger than a pipe buffer) and one that floods stderr while the frames are still being written: neither may
    # hang the caller, both surface as the reference's RuntimeError with the encoder's own words
    big = np.zeros((40, 128, 128, 3), np.uint8)   # 1.9 MB
    early = tmp_path / "ffmpeg_early"
    early.write_text("#!/bin/sh\necho no such codec >&2\nexit 1\n")
    early.chmod(0o755)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(early))
    with pytest.raises(RuntimeError, match="(?s)ffmpeg failed.*no such codec"):
        rs.stitch_video_frames(big, str(tmp_path / "v4.mp4"))
    chatty = tmp_path / "ffmpeg_chatty"
    chatty.write_text("#!/bin/sh\nhead -c 300000 /dev/zero | tr '\\0' 'x' >&2\ncat > %s\n" % (tmp_path / "big.bin"))
    chatty.chmod(0o755)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(chatty))
    rs.stitch_video_frames(big, str(tmp_path / "v5.mp4"))
    assert (tmp_path / "big.bin").stat().st_size == big.size
    with pytest.raises(FileNotFoundError):
        rs.stitch_video_frames(frames[:0], str(tmp_path / "v3.mp4"))


def test_validation_report_mirror(golden_dir, tmp_path):
    """generate_report on PNG files rebuilt from the golden frames writes the reference's JSON and checklist;
    error behaviour follows validation_reporting.py:48-70."""
    import json
    from PIL import Image
    g = np.load(os.path.join(golden_dir, "report_golden.npz"))
    model = tmp_path / "model"
    for it in (500, 3000):
        (model / "train" / f"ours_{it}" / "renders").mkdir(parents=True)
        (model / "train" / f"ours_{it}" / "gt").mkdir(parents=True)
    for t in range(len(g["renders"])):
        Image.fromarray(g["renders"][t]).save(model / "train" / "ours_3000" / "renders" / f"{t:05d}.png")
        Image.fromarray(g["gt"][t]).save(model / "train" / "ours_3000" / "gt" / f"{t:05d}.png")
    det = tmp_path / "det"
    rs.export_deterministic_frames(str(model / "train" / "ours_3000" / "renders"), str(det), None, 12)
    out = tmp_path / "report"
    validation_reporting.generate_report(model, det, out)
    got, want = json.load(open(out / "strict_scores.json")), json.loads(str(g["report"]))

    def same(x, y):
        """The reference's report: identical structure, names, counts and bands; scores to 1e-5 dB / 1e-12 (the
        reference averages squared errors in float32, the shared moments are float64)."""
        if isinstance(x, dict):
            assert x.keys() == y.keys()
            for k in x:
                if k in ("psnr", "ssim") and x[k] is not None:
                    assert abs(x[k] - y[k]) <= (1e-5 if k == "psnr" else 1e-12), (k, x[k], y[k])
                else:
                    same(x[k], y[k])
        elif isinstance(x, list):
            assert len(x) == len(y)
            for p, q in zip(x, y):
                same(p, q)
        else:
            assert x == y

    same(got, want)
    assert (out / "human_review_checklist.md").read_text() == str(g["checklist"])
    a, b = g["renders"][5].astype(np.float32), g["gt"][5].astype(np.float32)
    assert abs(validation_reporting.ssim_global(a, b) - float(g["ssim_ab"])) <= 1e-12
    for p, name in json.loads(str(g["buckets"])).items():
        assert validation_reporting._bucket(float(p)) == name
    with pytest.raises(FileNotFoundError):
        validation_reporting.generate_report(tmp_path / "nope", det, out)
    with pytest.raises(FileNotFoundError):
        validation_reporting.generate_report(model, tmp_path / "no_manifest", out)
    os.makedirs(tmp_path / "m2" / "train")
    with pytest.raises(FileNotFoundError):
        validation_reporting._find_latest_train_dir(tmp_path / "m2")


def test_avatar_bake_is_the_upstream_activation():
    from omfs_b200 import avatar
    av = synthetic.make_avatar(500, 100, seed=3)
    b = avatar.bake(av)
    np.testing.assert_allclose(b["scale_lo"][:, :3], np.exp(av.scaling.astype(np.float64)), rtol=1e-7)
    sig = 1.0 / (1.0 + np.exp(-av.opacity.astype(np.float64)))
    np.testing.assert_allclose(np.exp2(b["scale_lo"][:, 3].astype(np.float64)), sig, rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(b["rot"], axis=1), 1.0, atol=1e-6)
    assert np.array_equal(avatar.binding_of(b), av.binding)
    flat = av.sh.reshape(500, 48)
    assert np.array_equal(b["sh"][5, :, 2], flat[:, 22])      # flat index 22 -> plane 5, lane 2


def test_main_flow_with_a_stand_in_renderer(tmp_path, monkeypatch):
    """main()'s control flow without a GPU: the renderer and ffmpeg are stand-ins; the parameter edit reaches the
    renderer, PNGs land in the upstream layout, the encoder gets exactly those frames, the deterministic export and
    the clean-up of the temporary dataset happen, and the renderer's errors keep the reference's exception types."""
    import json
    from PIL import Image
    import omfs_b200  # noqa: F401
    from omfs_b200 import cameras, flame_io, render_surgery as rs, synthetic
    T, W, H, V = 5, 32, 24, 162
    model = synthetic.make_flame_model(seed=8, n_verts=V)
    params = synthetic.make_frame_params(T, seed=9, n_verts=V)
    av = synthetic.make_avatar(50, model.n_faces, seed=10)
    c2w = cameras.look_at_c2w((0.0, 0.0, 1.0), (0.0, 0.0, 0.0))
    data, mdl = str(tmp_path / "data"), str(tmp_path / "model")
    flame_io.write_synthetic_dataset(data, mdl, model, params, av, c2w, 0.3, W, H, iteration=3000)
    n_train = len(flame_io.load_transforms(data, "train"))
    seen = {}

    def fake_render(model_, params_, av_, cams, plan_offset=None, device=None, want_png=False, want_u8=True, on_pngs=None):
        seen["translation"] = params_.translation.copy()
        seen["jaw"] = params_.jaw_pose.copy()
        out = np.zeros((params_.n_frames, H, W, 3), np.uint8)
        out[..., 0] = np.arange(params_.n_frames, dtype=np.uint8)[:, None, None]
        # the stand-in for the device sink: the host encoder of the same package
        pngs = [rs.encode_png(f) for f in out]
        if want_png and on_pngs is not None:   # the streamed form: clips arrive one by one, nothing is returned
            half = len(pngs) // 2
            on_pngs(0, pngs[:half])
            on_pngs(half, pngs[half:])
            pngs = []
        return (out if want_u8 else None, pngs) if want_png else out

    fake = tmp_path / "ffmpeg"
    fake.write_text("#!/bin/sh\nfor a in \"$@\"; do echo \"$a\" >> %s; done\ncat > %s\n" %
                    (tmp_path / "args.txt", tmp_path / "stdin.bin"))
    fake.chmod(0o755)
    monkeypatch.setattr(rs, "_render_frames", fake_render)
    monkeypatch.setattr(rs, "_get_ffmpeg_path", lambda: str(fake))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    made = []
    real_create = rs.create_modified_dataset
    monkeypatch.setattr(rs, "create_modified_dataset", lambda *a, **k: made.append(real_create(*a, **k)) or made[-1])
    out = tmp_path / "video" / "final.mp4"
    rs.main(["--lefort_mm", "5", "--bsso_mm", "-3", "--sensitivity", "2", "--model_path", mdl, "--data_dir", data,
             "--output", str(out), "--fps", "24", "--export_frames_dir", str(tmp_path / "ab"),
             "--deterministic_max_frames", "2"])
    # R1 + R2 reached the renderer: translation y += 5 mm * 2 * 0.001, jaw x += -3 mm * 2 * 0.001
    np.testing.assert_allclose(seen["translation"][:, 1], params.translation[:n_train, 1] + np.float32(0.01), atol=1e-7)
    np.testing.assert_allclose(seen["jaw"][:, 0], params.jaw_pose[:n_train, 0] - np.float32(0.006), atol=1e-7)
    renders = os.path.join(mdl, "train", "ours_3000", "renders")
    names = sorted(os.listdir(renders))
    assert names == [f"{i:05d}.png" for i in range(n_train)]
    frames = np.stack([np.asarray(Image.open(os.path.join(renders, n))) for n in names])
    assert (frames[..., 0] == np.arange(n_train)[:, None, None]).all()
    assert (tmp_path / "stdin.bin").read_bytes() == frames.tobytes()
    args = (tmp_path / "args.txt").read_text().split("\n")
    assert args[args.index("-framerate") + 1] == "24" and args[args.index("-s") + 1] == f"{W}x{H}"
    assert json.load(open(tmp_path / "ab" / "deterministic_indices_manifest.json"))["selected_indices"] == [0, n_train - 1]
    assert made == []                                       # the edit happened in memory: no dataset copy was written
    in_memory = {k: v.copy() for k, v in seen.items()}
    # the reference's on-disk route (edited copy written, rendered, deleted) hands the renderer the same bits
    monkeypatch.setenv("OMFS_MATERIALISE_DATASET", "1")
    rs.main(["--lefort_mm", "5", "--bsso_mm", "-3", "--sensitivity", "2", "--model_path", mdl, "--data_dir", data,
             "--output", str(out)])
    monkeypatch.delenv("OMFS_MATERIALISE_DATASET")
    assert made and not os.path.exists(made[0])            # the temporary dataset is cleaned up (:537-539)
    for k in in_memory:
        assert np.array_equal(in_memory[k].view(np.uint32), seen[k].view(np.uint32)), k

    def broken(*a, **k):
        raise ValueError("device lost")
    monkeypatch.setattr(rs, "_render_frames", broken)
    with pytest.raises(RuntimeError, match="Rendering failed:\n.*device lost"):
        rs.render_with_gaussians(mdl, data)
    bad = flame_io.load_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_3000", "point_cloud.ply"))
    bad.binding[:] = model.n_faces + 5
    flame_io.save_avatar_ply(os.path.join(mdl, "point_cloud", "iteration_3000", "point_cloud.ply"), bad)
    with pytest.raises(ValueError, match="binds Gaussians to faces"):
        rs.render_with_gaussians(mdl, data)


def test_flame_pickle_with_chumpy_arrays_loads_without_chumpy(tmp_path):
    """The published FLAME pickles wrap their arrays in chumpy objects and hold a scipy-sparse joint regressor and a
    uint32 kinematic table (what flame_fitter.py:79-108 reads with chumpy installed).  The loader needs neither
    chumpy nor any conversion step, gives the same model as the .npz form, and refuses another skeleton."""
    import pickle
    import sys
    import types
    import scipy.sparse as sp
    import omfs_b200  # noqa: F401
    from omfs_b200 import flame_io, synthetic
    model = synthetic.make_flame_model(seed=3, n_verts=162)
    V = model.n_verts

    class Ch:                                   # pickled by reference as chumpy.ch.Ch, like the real files
        def __init__(self, x):
            self.x = np.asarray(x, np.float64)
            self._dirty_vars = set()
            self._itr = None
    Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
    pkg, sub = types.ModuleType("chumpy"), types.ModuleType("chumpy.ch")
    sub.Ch = Ch
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = pkg, sub
    try:
        record = {
            "v_template": Ch(model.v_template),
            "shapedirs": Ch(model.shapedirs.T.reshape(V, 3, -1)),
            "posedirs": Ch(model.posedirs.T.reshape(V, 3, -1)),
            "J_regressor": sp.csc_matrix(model.j_regressor.astype(np.float64)),
            "J": Ch(np.zeros((5, 3))),
            "weights": Ch(model.lbs_weights),
            "f": model.faces.astype(np.uint32),
            "kintree_table": np.array([[4294967295, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=np.uint32),
            "bs_style": "lbs", "bs_type": "lrotmin",
        }
        path = str(tmp_path / "flame2023.pkl")
        with open(path, "wb") as f:
            pickle.dump(record, f, protocol=2)
        record["kintree_table"] = np.array([[4294967295, 0, 1, 2, 2], [0, 1, 2, 3, 4]], dtype=np.uint32)
        with open(str(tmp_path / "other.pkl"), "wb") as f:
            pickle.dump(record, f, protocol=2)
    finally:
        del sys.modules["chumpy"], sys.modules["chumpy.ch"]
    with pytest.raises(ModuleNotFoundError):     # what a plain pickle.load does here
        pickle.load(open(path, "rb"), encoding="latin1")
    got = flame_io.load_flame_model(path)
    for k in ("v_template", "faces", "shapedirs", "posedirs", "j_regressor", "lbs_weights", "parents"):
        a, b = getattr(got, k), getattr(model, k)
        assert a.dtype == b.dtype and a.flags["C_CONTIGUOUS"] and np.array_equal(a, b), k
    with pytest.raises(ValueError, match="kinematic tree"):
        flame_io.load_flame_model(str(tmp_path / "other.pkl"))
    # model_path/flame_model.pkl is picked up by the drop-in's model search
    os.makedirs(tmp_path / "m")
    os.replace(path, tmp_path / "m" / "flame_model.pkl")
    assert rs._find_flame_model(str(tmp_path / "m")).endswith("flame_model.pkl")


def test_read_npz_equals_numpy_load(tmp_path):
    """flame_io.read_npz (headers matched, not evaluated) returns what np.load returns — for the FLAME records
    np.savez writes, and, through its fallback, for members it does not fast-path."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import flame_io, synthetic
    p = str(tmp_path / "rec.npz")
    flame_io.save_flame_params(p, synthetic.make_frame_params(3, n_verts=162))
    odd = str(tmp_path / "odd.npz")
    np.savez_compressed(odd, x=np.arange(5), s=np.array("txt"), be=np.array([1, 2], dtype=">f4"),
                        empty=np.zeros((0, 3), np.float32), scalar=np.float32(2.5), f=np.asfortranarray(np.ones((2, 3))))
    for path in (p, odd):
        got, want = flame_io.read_npz(path), dict(np.load(path))
        assert set(got) == set(want)
        for k in want:
            assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), (path, k)
            assert got[k].dtype.newbyteorder("=") == want[k].dtype.newbyteorder("="), (path, k)
    assert all(v.flags.writeable for v in flame_io.read_npz(p).values())
    # the stored-archive walk (local file headers of one read, no zipfile) is what served the first file ...
    direct = flame_io._read_stored_zip(p)
    assert set(direct) == set(np.load(p).files) and all(np.array_equal(direct[k], np.load(p)[k]) for k in direct)
    # ... it refuses what it does not cover (compressed members, a stored archive of odd members, a truncated file,
    # no archive at all), and read_npz still answers through zipfile / np.load where an answer exists
    with pytest.raises(ValueError):
        flame_io._read_stored_zip(odd)
    stored_odd = str(tmp_path / "stored_odd.npz")
    np.savez(stored_odd, s=np.array("txt"), be=np.array([1, 2], dtype=">f4"), f=np.asfortranarray(np.ones((2, 3))))
    with pytest.raises(ValueError):
        flame_io._read_stored_zip(stored_odd)
    got = flame_io.read_npz(stored_odd)
    assert str(got["s"]) == "txt" and np.array_equal(got["be"], [1, 2]) and got["f"].shape == (2, 3)
    cut = str(tmp_path / "cut.npz")
    with open(p, "rb") as f:
        blob = f.read()
    with open(cut, "wb") as f:
        f.write(blob[: len(blob) // 2])
    with pytest.raises(ValueError):
        flame_io._read_stored_zip(cut)
    with pytest.raises(Exception):
        flame_io.read_npz(cut)
    # a large member (zip64 sizes in the extra field, as np.savez always writes them) and an empty archive member
    big = str(tmp_path / "big.npz")
    np.savez(big, a=np.arange(300_000, dtype=np.float32).reshape(1000, 300), e=np.zeros((0, 3), np.float32))
    d = flame_io._read_stored_zip(big)
    assert np.array_equal(d["a"], np.arange(300_000, dtype=np.float32).reshape(1000, 300)) and d["e"].shape == (0, 3)
