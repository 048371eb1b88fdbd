"""Input validation at the drop-in boundary — CPU only: everything here is rejected BEFORE the library
touches a device, so the checks run without a GPU (the session constructor validates the mesh and the
Gaussian bindings on the host first)."""
import os

import numpy as np
import pytest

import omfs_b200  # noqa: F401
from omfs_b200 import avatar, flame_io, render_surgery as rs, runtime, synthetic


@pytest.fixture(scope="module")
def tiny():
    model, params, av, cam = synthetic.make_scene(n_gauss=500, n_frames=2, width=64, height=64, n_verts=162)
    return model, params, av, cam


def _session(model, baked):
    return runtime.Session(model, baked, 64, 64, max_batch=2, device=0)


def test_session_rejects_binding_beyond_face_count(tiny):
    model, _, av, _ = tiny
    baked = avatar.bake(av)
    bad = dict(baked)
    xyzb = baked["xyzb"].copy()
    xyzb[7, 3] = np.array([model.n_faces], np.int32).view(np.float32)[0]   # first face index that does not exist
    bad["xyzb"] = xyzb
    with pytest.raises(runtime.OmfsError, match="bound to face"):
        _session(model, bad)
    xyzb[7, 3] = np.array([-1], np.int32).view(np.float32)[0]
    with pytest.raises(runtime.OmfsError, match="bound to face"):
        _session(model, bad)


def test_session_rejects_face_index_beyond_vertex_count(tiny):
    model, _, av, _ = tiny
    baked = avatar.bake(av)
    faces = model.faces.copy()
    faces[3, 1] = model.n_verts
    broken = synthetic.FlameModel(model.v_template, faces, model.shapedirs, model.posedirs, model.j_regressor,
                                  model.lbs_weights, model.parents)
    with pytest.raises(runtime.OmfsError, match="not a vertex index"):
        _session(broken, baked)


def test_session_names_shape_mismatches(tiny):
    model, _, av, _ = tiny
    baked = avatar.bake(av)
    short = dict(baked)
    short["sh"] = baked["sh"][:, :-1]
    with pytest.raises(runtime.OmfsError, match="sh has shape"):
        _session(model, short)


def test_subject_model_mismatch_is_named(tiny):
    model, params, av, _ = tiny
    flame_io.check_subject_matches_model(model, params, av)          # consistent inputs pass
    other = synthetic.make_frame_params(2, n_verts=model.n_verts + 120)
    with pytest.raises(ValueError, match=r"records cover \d+ vertices, the model has"):
        flame_io.check_subject_matches_model(model, other, av)
    far = synthetic.Avatar(av.xyz, av.scaling, av.rotation, av.opacity, av.sh, av.binding + model.n_faces)
    with pytest.raises(ValueError, match="binds Gaussians to faces"):
        flame_io.check_subject_matches_model(model, params, far)


def test_raw_flame_mesh_gets_the_teeth_hint():
    model = synthetic.make_flame_model(n_verts=5023)
    params = synthetic.make_frame_params(1, n_verts=5143)
    av = synthetic.make_avatar(64, model.n_faces)
    with pytest.raises(ValueError, match="export_flame_with_teeth"):
        flame_io.check_subject_matches_model(model, params, av)


def test_write_gt_frames_copies_or_reencodes(tmp_path, tiny):
    from PIL import Image
    _, _, _, cam = tiny
    data = tmp_path / "data"
    (data / "images").mkdir(parents=True)
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, size=(cam.height, cam.width, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(data / "images" / "00000_00.png")                      # already in the gt/ form
    rgba = np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)])
    rgba[:8, :, 3] = 0                                                              # transparent band -> white
    Image.fromarray(rgba, "RGBA").save(data / "images" / "00001_00.png")
    big = rng.integers(0, 256, size=(2 * cam.height, 2 * cam.width, 3), dtype=np.uint8)
    Image.fromarray(big).save(data / "images" / "00002_00.jpg")                      # other size, other codec
    frames = [flame_io.DatasetFrame(f"images/{n}", "", i, 0, cam) for i, n in
              enumerate(("00000_00.png", "00001_00.png", "00002_00.jpg", "missing.png"))]
    out = rs.write_gt_frames(str(tmp_path / "gt"), str(data), frames, first=10)
    assert [os.path.basename(p) for p in out] == ["00010.png", "00011.png", "00012.png"]
    assert open(out[0], "rb").read() == open(data / "images" / "00000_00.png", "rb").read()
    flat = np.asarray(Image.open(out[1]))
    assert flat.shape == rgb.shape and (flat[:8] == 255).all() and np.array_equal(flat[8:], rgb[8:])
    assert Image.open(out[2]).size == (cam.width, cam.height) and Image.open(out[2]).mode == "RGB"
