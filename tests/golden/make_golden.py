"""Generate golden vectors by running the REFERENCE'S OWN CODE in this container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only), writes *.npz here

/root/reference does not exist on the GPU box, so the outputs are committed next to this script and
tests/test_oracle_golden.py only ever reads the committed files.

What can be run and how:
  * 02_Visual_Engine/render_surgery.py imports with numpy alone: compute_offset, modify_flame_params
    (R1, R2), choose_rig_mode, export_deterministic_frames are called as shipped.
  * 02_Visual_Engine/flame_fitter.py needs mediapipe (absent) and the licence-gated FLAME pickle.
    `mediapipe` is replaced by an empty stub module (SimpleFLAME never touches it) and SimpleFLAME is
    constructed from a SYNTHETIC pickle with the keys it reads (flame_fitter.py:80-120); its forward
    (R7: blendshape einsum, Rodrigues, global rotation, translation, landmarks) then runs unmodified.
  * 01_Clinical_Engine/surgical_sim.py needs pyvista (absent).  `_angle_to_normal` and
    `_normalise_direction` (R6, pure numpy) are called as shipped with a stub `pyvista` module.
    `SurgicalCutter.move_segments` (R5) runs unmodified on a minimal PolyData stand-in (points + the
    handful of methods it calls: copy, center = bounding-box centre, rotate_x/y/z about a point,
    translate, clip by half-space on POINTS).  The stand-in, not VTK, supplies the rotation
    conventions, so these vectors pin the reference's orchestration (rotation order X->Z->Y about the
    bbox centre, unit direction x mm, which segments move), not VTK's arithmetic.
"""
from __future__ import annotations

import io
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def golden_render_surgery():
    sys.path.insert(0, os.path.join(REF, "02_Visual_Engine"))
    import render_surgery as rs

    cases = [(0.0, 1.0), (5.0, 1.0), (-3.0, 1.0), (5.0, 2.5), (10.0, 0.0), (7.5, 0.1), (-15.0, 3.0), (0.5, 1.7)]
    offs = np.array([rs.compute_offset(mm, s) for mm, s in cases], dtype=np.float64)
    out = {"offset_cases": np.array(cases, dtype=np.float64), "offset_values": offs,
           "scale_factor": np.float64(rs.SCALE_FACTOR)}

    rng = np.random.default_rng(11)
    T = 10
    rec = dict(jaw_pose=rng.normal(0, 0.1, (T, 3)).astype(np.float32),
               translation=rng.normal(0, 0.01, (T, 3)).astype(np.float32),
               expr=rng.normal(0, 0.5, (T, 100)).astype(np.float32),
               shape=rng.normal(0, 1, 300).astype(np.float32))
    single = dict(jaw_pose=rec["jaw_pose"][0].copy(), translation=rec["translation"][0].copy(),
                  expr=rec["expr"][:1].copy(), shape=rec["shape"].copy())
    mods = [(0.005, 0.0, None), (0.0, 0.003, None), (0.0125, -0.0075, None),
            (0.01, 0.02, {"translation_axis": 2, "jaw_axis": 1, "lefort_scale": 2.0, "bsso_scale": 0.5})]
    with tempfile.TemporaryDirectory() as d:
        for name, r in (("batched", rec), ("single", single)):
            src = os.path.join(d, f"{name}.npz")
            np.savez(src, **r)
            for k, v in r.items():
                out[f"mod_{name}_in_{k}"] = v
            for i, (lo, bo, dm) in enumerate(mods):
                dst = os.path.join(d, f"{name}_{i}.npz")
                rs.modify_flame_params(src, dst, lo, bo, deformation_map=dm)
                got = np.load(dst)
                for k in ("jaw_pose", "translation", "expr", "shape"):
                    out[f"mod_{name}_{i}_{k}"] = got[k]
    out["mod_args"] = np.array([[lo, bo] for lo, bo, _ in mods], dtype=np.float64)
    out["mod_maps"] = np.array(json.dumps([dm for _, _, dm in mods]))
    np.savez(os.path.join(HERE, "render_surgery_golden.npz"), **out)
    print("render_surgery goldens:", len(out), "arrays")


def golden_simple_flame():
    import torch

    sys.modules.setdefault("mediapipe", types.ModuleType("mediapipe"))
    sys.path.insert(0, os.path.join(REF, "02_Visual_Engine"))
    import flame_fitter as ff
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    V, F, L = 300, 500, 12
    model = {
        "v_template": rng.normal(0, 0.1, (V, 3)),
        "shapedirs": rng.normal(0, 1e-2, (V, 3, 400)),
        "J_regressor": sp.csc_matrix(np.abs(rng.normal(0, 1, (5, V))) / V),
        "weights": np.abs(rng.normal(0, 1, (V, 5))),
        "kintree_table": np.array([[4294967295, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=np.int64),
        "f": rng.integers(0, V, (F, 3)),
    }
    bary = rng.uniform(0.1, 1.0, (L, 3))
    bary /= bary.sum(axis=1, keepdims=True)
    lmk = {"full_lmk_faces_idx": rng.integers(0, F, L), "full_lmk_bary_coords": bary}
    with tempfile.TemporaryDirectory() as d:
        pkl = os.path.join(d, "flame.pkl")
        with open(pkl, "wb") as f:
            pickle.dump(model, f)
        lpath = os.path.join(d, "lmk.npy")
        np.save(lpath, lmk, allow_pickle=True)
        ff.FLAME_LMK_PATH = lpath
        m = ff.SimpleFLAME(pkl, n_shape=100, n_expr=50)
    B = 4
    shape = rng.normal(0, 1, (B, 100)).astype(np.float32)
    expr = rng.normal(0, 0.5, (B, 50)).astype(np.float32)
    rotation = rng.normal(0, 0.3, (B, 3)).astype(np.float32)
    rotation[1] = 0.0  # the angle+1e-8 corner
    jaw = rng.normal(0, 0.1, (B, 3)).astype(np.float32)
    translation = rng.normal(0, 0.05, (B, 3)).astype(np.float32)
    with torch.no_grad():
        lm = m(torch.tensor(shape), torch.tensor(expr), torch.tensor(rotation), torch.tensor(jaw),
               torch.tensor(translation)).numpy()
        Rm = m._axis_angle_to_matrix(torch.tensor(rotation)).numpy()
    np.savez(os.path.join(HERE, "simple_flame_golden.npz"),
             v_template=model["v_template"].astype(np.float32),
             shapedirs=model["shapedirs"].astype(np.float32), faces=model["f"].astype(np.int64),
             lmk_faces_idx=lmk["full_lmk_faces_idx"].astype(np.int64), lmk_bary=bary.astype(np.float32),
             shape=shape, expr=expr, rotation=rotation, jaw=jaw, translation=translation,
             landmarks=lm, rotmats=Rm)
    print("SimpleFLAME goldens: landmarks", lm.shape)


class _StubPolyData:
    """Minimal stand-in for pv.PolyData: a float64 point cloud (see the module docstring)."""

    def __init__(self, points=None):
        self.points = np.zeros((0, 3)) if points is None else np.array(points, dtype=np.float64)

    @property
    def n_points(self):
        return len(self.points)

    @property
    def center(self):
        if not len(self.points):
            return [0.0, 0.0, 0.0]
        return list((self.points.min(axis=0) + self.points.max(axis=0)) / 2.0)

    def copy(self):
        return _StubPolyData(self.points.copy())

    def merge(self, other):
        return _StubPolyData(np.concatenate([self.points, other.points]))

    def clip(self, normal, origin, invert=False):
        d = (self.points - np.asarray(origin, dtype=np.float64)) @ np.asarray(normal, dtype=np.float64)
        keep = ~(d > 0.0) if invert else (d > 0.0)
        return _StubPolyData(self.points[keep])

    def _rot(self, axis, angle, point):
        a = np.radians(angle)
        c, s = np.cos(a), np.sin(a)
        R = {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
             "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]
        p = np.asarray(point, dtype=np.float64)
        self.points = (self.points - p) @ R.T + p

    def rotate_x(self, angle, point=(0, 0, 0), inplace=True):
        self._rot("x", angle, point)
        return self

    def rotate_y(self, angle, point=(0, 0, 0), inplace=True):
        self._rot("y", angle, point)
        return self

    def rotate_z(self, angle, point=(0, 0, 0), inplace=True):
        self._rot("z", angle, point)
        return self

    def translate(self, t, inplace=True):
        self.points = self.points + np.asarray(t, dtype=np.float64)
        return self


def golden_surgical_sim():
    stub = types.ModuleType("pyvista")
    stub.PolyData = _StubPolyData
    sys.modules["pyvista"] = stub
    sys.path.insert(0, os.path.join(REF, "01_Clinical_Engine"))
    import surgical_sim as ss

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import reference_rows as rr

    angles = [(0.0, 0.0), (10.0, 0.0), (0.0, 25.0), (-7.5, 12.25), (90.0, 45.0), (33.3, -120.0)]
    bases = [(0, 0, 1), (1, 0, 0)]
    normals = np.array([[ss._angle_to_normal(b, p, y) for p, y in angles] for b in bases], dtype=np.float64)
    dirs_in = [(0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (1.0, 2.0, -0.5), (0.0, 0.0, 3.0)]
    dirs = np.array([ss._normalise_direction(d) for d in dirs_in], dtype=np.float64)

    maxilla = rr.uv_sphere(30.0, (0, 0, 20))
    mandible = rr.uv_sphere(30.0, (0, 0, -20))
    out = {"angles": np.array(angles), "normals": normals, "dirs_in": np.array(dirs_in), "dirs": dirs,
           "maxilla": maxilla, "mandible": mandible}
    plans = [
        dict(maxilla_mm=5.0, mandible_mm=8.0),
        dict(maxilla_mm=10.0, mandible_mm=0.0),
        dict(maxilla_mm=5.0, mandible_mm=0.0, advancement_direction=(1.0, 0.0, 0.0)),
        dict(maxilla_mm=3.0, mandible_mm=-4.0, advancement_direction=(0.2, 1.0, 0.1),
             maxilla_rotation=(5.0, -3.0, 2.0), mandible_rotation=(0.0, 7.5, 0.0)),
    ]
    cuts = [dict(lefort_z=20, bsso_l_x=-15, bsso_r_x=15),
            dict(lefort_z=15, bsso_l_x=-12, bsso_r_x=18, lefort_pitch=8.0, lefort_yaw=-4.0, bsso_l_yaw=6.0,
                 bsso_r_pitch=-5.0)]
    for ci, cut in enumerate(cuts):
        cutter = ss.SurgicalCutter(_StubPolyData(maxilla), _StubPolyData(mandible))
        segs = cutter.perform_cut(**cut)
        for k, v in segs.items():
            out[f"cut{ci}_{k}"] = v.points
        out[f"cut{ci}_args"] = np.array(json.dumps(cut))
        combined_center = np.array(cutter.get_combined_mesh().center)
        out[f"cut{ci}_center"] = combined_center
        for pi, plan in enumerate(plans):
            moved = cutter.move_segments(**plan)
            for k, v in moved.items():
                out[f"cut{ci}_plan{pi}_{k}"] = v.points
    out["plans"] = np.array(json.dumps(plans))
    np.savez(os.path.join(HERE, "surgical_sim_golden.npz"), **out)
    print("surgical_sim goldens:", len(out), "arrays")


def golden_psnr():
    sys.path.insert(0, os.path.join(REF, "02_Visual_Engine"))
    import validation_reporting as vr
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 255, (16, 16, 3)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 2.0, a.shape), 0, 255).astype(np.float32)
    np.savez(os.path.join(HERE, "psnr_golden.npz"), a=a, b=b, psnr_ab=np.float64(vr.psnr(a, b)),
             psnr_aa=np.float64(vr.psnr(a, a)))
    print("psnr golden:", vr.psnr(a, b), vr.psnr(a, a))


def golden_report():
    """R9 + the strict report (validation_reporting.py:16-123) and the deterministic export
    (render_surgery.py:365-409), run as shipped on a small synthetic model directory: 12 frames of 24x20
    renders / gt, exported with max_frames=12, then generate_report.  The frames, the manifest rows and the
    report JSON are stored; ssim_global / psnr of two float images as well."""
    sys.path.insert(0, os.path.join(REF, "02_Visual_Engine"))
    import render_surgery as rs
    import validation_reporting as vr
    from pathlib import Path
    from PIL import Image

    rng = np.random.default_rng(21)
    T, H, W = 12, 20, 24
    yy, xx = np.mgrid[0:H, 0:W]
    gt = np.stack([np.stack([(xx * 9 + t * 7) % 256, (yy * 11 + t * 3) % 256, ((xx + yy) * 5 + t * 13) % 256], -1)
                   for t in range(T)]).astype(np.uint8)
    noise = rng.normal(0, 6.0, gt.shape)
    noise[4] = 0.0  # one identical pair (index 4 is exported): psnr 99.0
    renders = np.clip(gt.astype(np.float64) + noise, 0, 255).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        model = Path(d) / "model"
        for it in (500, 3000):
            (model / "train" / f"ours_{it}" / "renders").mkdir(parents=True)
            (model / "train" / f"ours_{it}" / "gt").mkdir(parents=True)
        for t in range(T):
            Image.fromarray(renders[t]).save(model / "train" / "ours_3000" / "renders" / f"{t:05d}.png")
            Image.fromarray(gt[t]).save(model / "train" / "ours_3000" / "gt" / f"{t:05d}.png")
        det = Path(d) / "det"
        rs.export_deterministic_frames(str(model / "train" / "ours_3000" / "renders"), str(det), None, 12)
        manifest = json.load(open(det / "deterministic_indices_manifest.json"))
        out = Path(d) / "report"
        vr.generate_report(model, det, out)
        report = json.load(open(out / "strict_scores.json"))
        checklist = (out / "human_review_checklist.md").read_text()
    a = renders[5].astype(np.float32)
    b = gt[5].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "report_golden.npz"), renders=renders, gt=gt,
                        manifest_exports=json.dumps(manifest["exports"]),
                        selected=np.array(manifest["selected_indices"], dtype=np.int64),
                        report=json.dumps(report), checklist=checklist,
                        ssim_ab=np.float64(vr.ssim_global(a, b)), ssim_aa=np.float64(vr.ssim_global(a, a)),
                        ssim_gray=np.float64(vr.ssim_global(a[:, :, 0], b[:, :, 1])),
                        buckets=json.dumps({str(p): vr._bucket(p) for p in (0.0, 0.19, 0.2, 0.3, 0.35, 0.5, 0.65, 0.7, 0.8, 0.81, 1.0)}))
    print("report golden:", report["summary"])


def _tiny_dataset(root):
    """The seeded on-disk subject shared by golden_single_frame and tests/test_host_mirror.py."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import omfs_b200  # noqa: F401
    from omfs_b200 import cameras, flame_io, synthetic
    model = synthetic.make_flame_model(seed=31, n_verts=162)
    params = synthetic.make_frame_params(3, seed=32, n_verts=162)
    av = synthetic.make_avatar(200, model.n_faces, seed=33)
    c2w = cameras.look_at_c2w((0.0, 0.0, synthetic.camera_distance(32, 24)), (0.0, 0.0, 0.0))
    flame_io.write_synthetic_dataset(os.path.join(root, "data_conda"), os.path.join(root, "model"), model, params, av,
                                     c2w, 0.3, 32, 24, iteration=3000, write_images=True)
    return os.path.join(root, "data_conda")


def golden_single_frame():
    """single_frame_experiment.build_single_frame_dataset (:32-81) run as shipped (its module-level directory
    constants pointed at a temporary tree) on the tiny synthetic dataset: the file listing, the one-frame
    transforms and the batched flame_param.npz it writes."""
    sys.path.insert(0, os.path.join(REF, "02_Visual_Engine"))
    from pathlib import Path
    import single_frame_experiment as sfe
    with tempfile.TemporaryDirectory() as d:
        sfe.DATA_CONDA = Path(_tiny_dataset(d))
        sfe.DATA_SINGLE = Path(d) / "data_single_frame"
        sfe.build_single_frame_dataset()
        listing = sorted(str(p.relative_to(sfe.DATA_SINGLE)) for p in sfe.DATA_SINGLE.rglob("*") if p.is_file())
        transforms = {n: json.load(open(sfe.DATA_SINGLE / f"transforms_{n}.json")) for n in ("train", "test", "val")}
        batched = dict(np.load(sfe.DATA_SINGLE / "flame_param.npz", allow_pickle=True))
    np.savez_compressed(os.path.join(HERE, "single_frame_golden.npz"), listing=json.dumps(listing),
                        transforms=json.dumps(transforms), **{f"batched_{k}": v for k, v in batched.items()})
    print("single-frame golden:", listing)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("the reference tree is not available here; the committed goldens are authoritative")
    golden_render_surgery()
    golden_psnr()
    golden_report()
    golden_single_frame()
    golden_surgical_sim()
    golden_simple_flame()
