"""The oracle's in-tree rows (R1, R2, R5, R6, R7, R9) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import reference_rows as rr


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_compute_offset_matches_reference(golden_dir):
    g = _load(golden_dir, "render_surgery_golden.npz")
    assert float(g["scale_factor"]) == rr.SCALE_FACTOR
    for (mm, s), want in zip(g["offset_cases"], g["offset_values"]):
        assert rr.compute_offset(float(mm), float(s)) == float(want)  # bit-exact float64


def test_compute_offset_reference_test_cases():
    # the literal cases of /root/reference/test/test_render_surgery.py:26-45
    assert rr.compute_offset(0.0, 1.0) == 0.0
    assert rr.compute_offset(5.0, 1.0) == pytest.approx(0.005, abs=1e-12)
    assert rr.compute_offset(-3.0, 1.0) == pytest.approx(-0.003, abs=1e-12)
    assert rr.compute_offset(5.0, 2.5) == pytest.approx(0.0125, abs=1e-12)
    assert rr.compute_offset(10.0, 0.0) == 0.0


@pytest.mark.parametrize("kind", ["batched", "single"])
def test_modify_flame_params_matches_reference(golden_dir, kind):
    g = _load(golden_dir, "render_surgery_golden.npz")
    maps = json.loads(str(g["mod_maps"]))
    rec = {k: g[f"mod_{kind}_in_{k}"] for k in ("jaw_pose", "translation", "expr", "shape")}
    for i, ((lo, bo), dm) in enumerate(zip(g["mod_args"], maps)):
        before = {k: v.copy() for k, v in rec.items()}
        got = rr.modify_flame_params(rec, float(lo), float(bo), deformation_map=dm)
        for k in rec:
            want = g[f"mod_{kind}_{i}_{k}"]
            assert got[k].dtype == want.dtype
            assert np.array_equal(got[k].view(np.uint32), want.view(np.uint32)), (kind, i, k)
            assert np.array_equal(rec[k], before[k])  # the source record is not mutated


def test_modify_flame_params_reference_test_cases():
    # /root/reference/test/test_render_surgery.py:56-88
    rec = dict(jaw_pose=np.zeros((10, 3), np.float32), translation=np.zeros((10, 3), np.float32),
               expr=np.zeros((10, 100), np.float32), shape=np.zeros(300, np.float32))
    assert float(rr.modify_flame_params(rec, 0.005, 0.0)["translation"][0, 1]) == pytest.approx(0.005, abs=1e-5)
    assert float(rr.modify_flame_params(rec, 0.0, 0.003)["jaw_pose"][0, 0]) == pytest.approx(0.003, abs=1e-5)
    dm = {"translation_axis": 2, "jaw_axis": 1, "lefort_scale": 2.0, "bsso_scale": 0.5}
    out = rr.modify_flame_params(rec, 0.01, 0.02, deformation_map=dm)
    assert float(out["translation"][0, 2]) == pytest.approx(0.02, abs=1e-5)
    assert float(out["jaw_pose"][0, 1]) == pytest.approx(0.01, abs=1e-5)


def test_plane_normals_and_directions_match_reference(golden_dir):
    g = _load(golden_dir, "surgical_sim_golden.npz")
    for bi, base in enumerate([(0, 0, 1), (1, 0, 0)]):
        for ai, (p, y) in enumerate(g["angles"]):
            got = rr.angle_to_normal(base, float(p), float(y))
            assert np.array_equal(got, g["normals"][bi, ai])  # float64 bit-exact
    for d_in, want in zip(g["dirs_in"], g["dirs"]):
        assert np.array_equal(rr.normalise_direction(tuple(d_in)), want)
    with pytest.raises(ValueError):
        rr.normalise_direction((0.0, 0.0, 0.0))


def _planes_for(cut, center):
    ln = rr.angle_to_normal((0, 0, 1), cut.get("lefort_pitch", 0.0), cut.get("lefort_yaw", 0.0))
    bl = rr.angle_to_normal((1, 0, 0), cut.get("bsso_l_pitch", 0.0), cut.get("bsso_l_yaw", 0.0))
    br = rr.angle_to_normal((1, 0, 0), cut.get("bsso_r_pitch", 0.0), cut.get("bsso_r_yaw", 0.0))
    planes = np.zeros((3, 8))
    planes[0, :3], planes[0, 3:6] = ln, (center[0], center[1], cut["lefort_z"])
    planes[1, :3], planes[1, 3:6] = bl, (cut["bsso_l_x"], center[1], center[2])
    planes[2, :3], planes[2, 3:6] = br, (cut["bsso_r_x"], center[1], center[2])
    return planes


@pytest.mark.parametrize("ci", [0, 1])
def test_segments_and_moves_match_reference(golden_dir, ci):
    """R5/R6: masks reproduce perform_cut's four segments, and the moved points reproduce
    move_segments, on the reference's own sphere fixture (test_surgical_sim.py:18-25)."""
    g = _load(golden_dir, "surgical_sim_golden.npz")
    cut = json.loads(str(g[f"cut{ci}_args"]))
    plans = json.loads(str(g["plans"]))
    maxilla, mandible = g["maxilla"], g["mandible"]
    pts = np.concatenate([maxilla, mandible])
    is_mand = np.arange(len(pts)) >= len(maxilla)
    planes = _planes_for(cut, g[f"cut{ci}_center"])
    masks = rr.segment_masks(pts, planes, is_mand)
    seg = {
        "upper_skull": pts[(~is_mand) & ((masks & 1) == 0)],
        "mobile_maxilla": pts[(masks & 8) != 0],
        "distal_mandible": pts[(masks & 16) != 0],
    }
    for k, v in seg.items():
        assert np.array_equal(v.astype(np.float64), g[f"cut{ci}_{k}"]), k
    rami = np.concatenate([pts[is_mand & ((masks & 2) == 0)], pts[is_mand & ((masks & 4) == 0)]])
    assert np.array_equal(rami.astype(np.float64), g[f"cut{ci}_proximal_rami"])
    for pi, plan in enumerate(plans):
        moves = rr.make_moves(plan.get("maxilla_mm", 0.0), plan.get("mandible_mm", 0.0),
                              tuple(plan.get("advancement_direction", (0.0, 1.0, 0.0))),
                              tuple(plan.get("maxilla_rotation", (0.0, 0.0, 0.0))),
                              tuple(plan.get("mandible_rotation", (0.0, 0.0, 0.0))))
        moved, m2, _ = rr.displace_points(pts, planes, moves, is_mand)
        assert np.array_equal(m2, masks)
        want_max = g[f"cut{ci}_plan{pi}_mobile_maxilla"]
        want_mand = g[f"cut{ci}_plan{pi}_distal_mandible"]
        # the reference keeps float64 points; the restatement rounds the result to float32
        np.testing.assert_allclose(moved[(masks & 8) != 0], want_max, rtol=0, atol=4e-6)
        np.testing.assert_allclose(moved[(masks & 16) != 0], want_mand, rtol=0, atol=4e-6)
        fixed = (masks & 24) == 0
        assert np.array_equal(moved[fixed], pts[fixed])


def test_reference_surgical_sim_test_cases(golden_dir):
    """Numeric cases of /root/reference/test/test_surgical_sim.py:46-110 restated without PyVista."""
    g = _load(golden_dir, "surgical_sim_golden.npz")
    maxilla, mandible = g["maxilla"], g["mandible"]
    assert maxilla.shape == (362, 3)  # theta*(phi-2)+2 at 20x20
    pts = np.concatenate([maxilla, mandible])
    is_mand = np.arange(len(pts)) >= len(maxilla)
    center = (pts.min(0).astype(np.float64) + pts.max(0).astype(np.float64)) / 2
    planes = _planes_for(dict(lefort_z=20, bsso_l_x=-15, bsso_r_x=15), center)

    def centre(p):
        return (p.min(0) + p.max(0)) / 2

    masks = rr.segment_masks(pts, planes, is_mand)
    mm, md = (masks & 8) != 0, (masks & 16) != 0
    assert mm.any() and md.any() and (is_mand & ~md).any()
    moved, _, _ = rr.displace_points(pts, planes, rr.make_moves(5.0, 8.0), is_mand)
    assert centre(moved[mm])[1] - centre(pts[mm])[1] == pytest.approx(5.0, abs=0.05)
    assert centre(moved[md])[1] - centre(pts[md])[1] == pytest.approx(8.0, abs=0.05)
    moved, _, _ = rr.displace_points(pts, planes, rr.make_moves(10.0, 0.0), is_mand)
    np.testing.assert_array_almost_equal(centre(moved[md]), centre(pts[md]))
    moved, _, _ = rr.displace_points(pts, planes, rr.make_moves(5.0, 0.0, (1.0, 0.0, 0.0)), is_mand)
    d = centre(moved[mm]) - centre(pts[mm])
    assert d[0] == pytest.approx(5.0, abs=0.05) and abs(d[1]) < 0.05 and abs(d[2]) < 0.05
    with pytest.raises(ValueError):
        rr.make_moves(1.0, 1.0, (0.0, 0.0, 0.0))


def test_simple_flame_matches_reference(golden_dir):
    """R7: blendshape contraction, Rodrigues (angle + 1e-8), global rotation, translation, landmarks."""
    g = _load(golden_dir, "simple_flame_golden.npz")
    sd = g["shapedirs"]
    R = rr.axis_angle_to_matrix_r7(g["rotation"])
    np.testing.assert_allclose(R, g["rotmats"], rtol=0, atol=2e-7)
    _, lm = rr.simple_flame_forward(g["v_template"], sd[:, :, :100], sd[:, :, 300:350], g["faces"],
                                    g["lmk_faces_idx"], g["lmk_bary"], g["shape"], g["expr"], g["rotation"],
                                    g["jaw"], g["translation"])
    np.testing.assert_allclose(lm, g["landmarks"], rtol=0, atol=2e-6)


def test_psnr_matches_reference(golden_dir):
    g = _load(golden_dir, "psnr_golden.npz")
    assert rr.psnr(g["a"], g["b"]) == float(g["psnr_ab"])
    assert rr.psnr(g["a"], g["a"]) == 99.0 == float(g["psnr_aa"])


def test_blendshape_gemm_formulation_matches_einsum(golden_dir):
    """The product's GEMM formulation (U1: template + coeffs . dirs, flattened) against the
    reference's einsum on the golden basis: the same contraction, different layout."""
    g = _load(golden_dir, "simple_flame_golden.npz")
    sd = g["shapedirs"]                      # (V,3,400) as the FLAME pickle stores it
    V = sd.shape[0]
    dirs = np.ascontiguousarray(sd.reshape(V * 3, 400).T)   # (400, 3V), the layout omfs_model_desc takes
    betas = np.zeros((g["shape"].shape[0], 400), np.float32)
    betas[:, :100] = g["shape"]
    betas[:, 300:350] = g["expr"]
    v_gemm = g["v_template"].reshape(1, -1) + betas @ dirs
    v_ein = g["v_template"][None] + np.einsum("ijk,bk->bij", sd[:, :, :100], g["shape"]) + \
        np.einsum("ijk,bk->bij", sd[:, :, 300:350], g["expr"])
    np.testing.assert_allclose(v_gemm.reshape(v_ein.shape), v_ein, rtol=0, atol=1e-6)


def test_ssim_bucket_and_report_match_reference(golden_dir):
    """validation_reporting.py run as shipped (tests/golden/make_golden.py: golden_report) against the
    restatement: ssim_global, _bucket, and the whole report (rows and per-bucket summary)."""
    import json
    g = _load(golden_dir, "report_golden.npz")
    a, b = g["renders"][5].astype(np.float32), g["gt"][5].astype(np.float32)
    assert rr.ssim_global(a, b) == float(g["ssim_ab"])
    assert rr.ssim_global(a, a) == float(g["ssim_aa"]) == 1.0
    assert rr.ssim_global(a[:, :, 0], b[:, :, 1]) == float(g["ssim_gray"])
    for p, name in json.loads(str(g["buckets"])).items():
        assert rr.bucket(float(p)) == name
    exports = json.loads(str(g["manifest_exports"]))
    assert [e["index"] for e in exports] == list(g["selected"])
    renders = {f"{t:05d}.png": g["renders"][t].astype(np.float32) for t in range(len(g["renders"]))}
    gts = {f"{t:05d}.png": g["gt"][t].astype(np.float32) for t in range(len(g["gt"]))}
    assert rr.report_rows(exports, renders, gts) == json.loads(str(g["report"]))


def test_frame_moments_reproduce_report_metrics(golden_dir):
    """The six moments omfs_frame_metrics accumulates give the reference's PSNR / SSIM back (PSNR to 1e-5 dB:
    the reference averages squared errors in float32; SSIM to 1e-12)."""
    import json
    import omfs_b200  # noqa: F401
    from omfs_b200 import validation_reporting as vrep
    g = _load(golden_dir, "report_golden.npz")
    m = rr.frame_moments(g["renders"], g["gt"])
    p, s = vrep.metrics_from_moments(m, g["renders"].shape[1] * g["renders"].shape[2])
    rows = {r["index"]: r for r in json.loads(str(g["report"]))["rows"]}
    for i, r in rows.items():
        assert abs(p[i] - r["psnr"]) <= 1e-5
        assert abs(s[i] - r["ssim"]) <= 1e-12
    assert p[4] == 99.0
