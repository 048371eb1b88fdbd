"""numpy restatement of the IN-TREE rows of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/omfs_oracle.c).  Every function cites the reference lines it
follows; tests/golden/make_golden.py runs the reference's own code in this container and commits
its outputs, and tests/test_oracle_golden.py pins these restatements against them.
"""
from __future__ import annotations

import math

import numpy as np

SCALE_FACTOR = 0.001  # 02_Visual_Engine/render_surgery.py:35


# ---------------------------------------------------------------------------- R1
def compute_offset(input_mm: float, sensitivity: float) -> float:
    """render_surgery.py:40-42 — float64 product, left to right."""
    return input_mm * sensitivity * SCALE_FACTOR


# ---------------------------------------------------------------------------- R2
def modify_flame_params(data: dict, lefort_offset: float, bsso_offset: float, deformation_map=None) -> dict:
    """render_surgery.py:111-139 on an in-memory record.  The addend is a Python float; numpy keeps
    the array dtype (float32), i.e. the float64 product is rounded to float32 and added in float32."""
    out = dict(data)
    dm = deformation_map or {}
    trans_axis = int(dm.get("translation_axis", 1))
    jaw_axis = int(dm.get("jaw_axis", 0))
    lefort_scale = float(dm.get("lefort_scale", 1.0))
    bsso_scale = float(dm.get("bsso_scale", 1.0))
    if "translation" in out:
        t = np.array(out["translation"], copy=True)
        add = t.dtype.type(lefort_offset * lefort_scale)
        if t.ndim == 1:
            t[trans_axis] = t[trans_axis] + add
        else:
            t[:, trans_axis] = t[:, trans_axis] + add
        out["translation"] = t
    if "jaw_pose" in out:
        j = np.array(out["jaw_pose"], copy=True)
        add = j.dtype.type(bsso_offset * bsso_scale)
        if j.ndim == 1:
            j[jaw_axis] = j[jaw_axis] + add
        else:
            j[:, jaw_axis] = j[:, jaw_axis] + add
        out["jaw_pose"] = j
    return out


# ---------------------------------------------------------------------------- R6
def angle_to_normal(base_normal, pitch_deg: float, yaw_deg: float) -> np.ndarray:
    """01_Clinical_Engine/surgical_sim.py:25-47 — n = unit(Rz(yaw) Rx(pitch) n0), float64."""
    n = np.array(base_normal, dtype=float)
    pitch = np.radians(pitch_deg)
    rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
    yaw = np.radians(yaw_deg)
    rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    rotated = rz @ rx @ n
    length = np.linalg.norm(rotated)
    if length < 1e-12:
        return np.array(base_normal, dtype=float)
    return rotated / length


def normalise_direction(direction) -> np.ndarray:
    """surgical_sim.py:50-56."""
    vec = np.array(direction, dtype=float)
    length = np.linalg.norm(vec)
    if length < 1e-12:
        raise ValueError("advancement_direction must be a non-zero vector.")
    return vec / length


def plane_side(points: np.ndarray, normal, origin) -> np.ndarray:
    """(p - o) . n in float64, evaluated (dx*nx + dy*ny) + dz*nz — the order displace.cu uses."""
    p = np.asarray(points, dtype=np.float64)
    n = np.asarray(normal, dtype=np.float64)
    o = np.asarray(origin, dtype=np.float64)
    return ((p[:, 0] - o[0]) * n[0] + (p[:, 1] - o[1]) * n[1]) + (p[:, 2] - o[2]) * n[2]


def segment_masks(points, planes, is_mandible) -> np.ndarray:
    """Half-space rule of perform_cut (surgical_sim.py:180-204): clip(invert=False) keeps
    (p-o).n > 0, invert=True keeps the rest.  bit0 Le Fort mobile side, bit1 inside BSSO-L,
    bit2 inside BSSO-R, bit3 mobile maxilla, bit4 distal mandible."""
    planes = np.asarray(planes, dtype=np.float64).reshape(3, 8)
    is_mandible = np.asarray(is_mandible, dtype=bool)
    m = np.zeros(len(points), dtype=np.uint8)
    m |= (~(plane_side(points, planes[0, :3], planes[0, 3:6]) > 0.0)).astype(np.uint8) * 1
    m |= (plane_side(points, planes[1, :3], planes[1, 3:6]) > 0.0).astype(np.uint8) * 2
    m |= (~(plane_side(points, planes[2, :3], planes[2, 3:6]) > 0.0)).astype(np.uint8) * 4
    m |= ((~is_mandible) & ((m & 1) != 0)).astype(np.uint8) * 8
    m |= (is_mandible & ((m & 2) != 0) & ((m & 4) != 0)).astype(np.uint8) * 16
    return m


# ---------------------------------------------------------------------------- R5
def rotation_xzy(pitch_deg: float, yaw_deg: float, roll_deg: float) -> np.ndarray:
    """Composite of rotate_x(pitch) -> rotate_z(yaw) -> rotate_y(roll) (surgical_sim.py:298-318),
    right-handed, degrees, each about the same fixed point: R = Ry Rz Rx."""
    def rx(a):
        c, s = math.cos(math.radians(a)), math.sin(math.radians(a))
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)

    def ry(a):
        c, s = math.cos(math.radians(a)), math.sin(math.radians(a))
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)

    def rz(a):
        c, s = math.cos(math.radians(a)), math.sin(math.radians(a))
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)

    R = np.eye(3)
    if pitch_deg != 0.0:
        R = rx(pitch_deg) @ R
    if yaw_deg != 0.0:
        R = rz(yaw_deg) @ R
    if roll_deg != 0.0:
        R = ry(roll_deg) @ R
    return R


def make_moves(maxilla_mm, mandible_mm, advancement_direction=(0.0, 1.0, 0.0),
               maxilla_rotation=(0.0, 0.0, 0.0), mandible_rotation=(0.0, 0.0, 0.0), unit_scale=1.0) -> np.ndarray:
    """moves[2][12] for omfs_displace_points: rotation then translation = unit(dir) * mm * unit_scale
    (surgical_sim.py:293, 321-322)."""
    d = normalise_direction(advancement_direction)
    out = np.zeros((2, 12), dtype=np.float64)
    for i, (mm, rot) in enumerate(((maxilla_mm, maxilla_rotation), (mandible_mm, mandible_rotation))):
        R = rotation_xzy(*rot) if any(r != 0.0 for r in rot) else np.eye(3)
        out[i, :9] = R.reshape(-1)
        out[i, 9:] = d * mm * unit_scale
    return out


def displace_points(points, planes, moves, is_mandible):
    """R5 on a point set: each mobile segment rotates about its bounding-box centre
    (PyVista's mesh.center, surgical_sim.py:300, 312) and translates.  float64, the operation
    order of displace.cu; returns (moved float32 points, masks, bboxes[2,6])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    moves = np.asarray(moves, dtype=np.float64).reshape(2, 12)
    masks = segment_masks(pts, planes, is_mandible)
    out = pts.copy()
    bbox = np.zeros((2, 6), dtype=np.float32)
    bbox[:, :3] = np.inf
    bbox[:, 3:] = -np.inf
    for s, bit in enumerate((8, 16)):
        sel = (masks & bit) != 0
        if s == 1:
            sel &= (masks & 8) == 0
        if not sel.any():
            continue
        seg = pts[sel]
        bbox[s, :3] = seg.min(axis=0)
        bbox[s, 3:] = seg.max(axis=0)
        c = (bbox[s, :3].astype(np.float64) + bbox[s, 3:].astype(np.float64)) * 0.5
        q = seg.astype(np.float64) - c
        R = moves[s, :9].reshape(3, 3)
        t = moves[s, 9:]
        res = np.empty_like(q)
        for r in range(3):
            rq = (R[r, 0] * q[:, 0] + R[r, 1] * q[:, 1]) + R[r, 2] * q[:, 2]
            res[:, r] = (rq + c[r]) + t[r]
        out[sel] = res.astype(np.float32)
    return out, masks, bbox


# ---------------------------------------------------------------------------- R7
def axis_angle_to_matrix_r7(axis_angle: np.ndarray) -> np.ndarray:
    """02_Visual_Engine/flame_fitter.py:122-152 — angle = |r|, axis = r / (angle + 1e-8),
    R = I + sin K + (1 - cos) K K, float32."""
    aa = np.asarray(axis_angle, dtype=np.float32)
    angle = np.linalg.norm(aa, axis=1, keepdims=True).astype(np.float32)
    axis = aa / (angle + np.float32(1e-8))
    cos_a = np.cos(angle)[..., None]
    sin_a = np.sin(angle)[..., None]
    B = aa.shape[0]
    K = np.zeros((B, 3, 3), dtype=np.float32)
    K[:, 0, 1] = -axis[:, 2]
    K[:, 0, 2] = axis[:, 1]
    K[:, 1, 0] = axis[:, 2]
    K[:, 1, 2] = -axis[:, 0]
    K[:, 2, 0] = -axis[:, 1]
    K[:, 2, 1] = axis[:, 0]
    I = np.eye(3, dtype=np.float32)[None]
    return (I + sin_a * K + (1 - cos_a) * np.matmul(K, K)).astype(np.float32)


def simple_flame_forward(v_template, shapedirs_shape, shapedirs_expr, faces, lmk_faces_idx, lmk_bary,
                         shape, expr, rotation, jaw, translation):
    """flame_fitter.py:154-197.  shapedirs_* are (V,3,K) slices (:89-92).  Returns (vertices, landmarks)."""
    v_template = np.asarray(v_template, dtype=np.float32)
    B = shape.shape[0]
    v = np.broadcast_to(v_template[None], (B,) + v_template.shape).astype(np.float32).copy()
    v = v + np.einsum("ijk,bk->bij", shapedirs_shape, shape).astype(np.float32)
    v = v + np.einsum("ijk,bk->bij", shapedirs_expr, expr).astype(np.float32)
    jaw_angle = jaw[:, 0:1]
    lower_mask = (v_template[:, 1] < v_template[:, 1].mean()).astype(np.float32)
    jaw_offset = np.zeros_like(v)
    jaw_offset[:, :, 1] = -jaw_angle * lower_mask[None] * np.float32(0.15)
    v = v + jaw_offset
    R = axis_angle_to_matrix_r7(rotation)
    v = np.matmul(v, np.transpose(R, (0, 2, 1)))
    v = v + translation[:, None, :]
    lmk_faces = faces[lmk_faces_idx]
    lmk_verts = v[:, lmk_faces]
    landmarks = (lmk_verts * lmk_bary[None, :, :, None]).sum(axis=2)
    return v.astype(np.float32), landmarks.astype(np.float32)


# ---------------------------------------------------------------------------- R9
def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """02_Visual_Engine/validation_reporting.py:16-20 (0-255 scale, 99.0 when identical)."""
    mse = float(np.mean((a - b) ** 2))
    if mse == 0.0:
        return 99.0
    return 20.0 * math.log10(255.0 / math.sqrt(mse))


def ssim_global(a: np.ndarray, b: np.ndarray) -> float:
    """02_Visual_Engine/validation_reporting.py:23-37: one global SSIM over the whole image, on the
    BT.601 luma of RGB inputs (computed in the input dtype, float32 in the report), moments in float64."""
    if a.ndim == 3:
        a = (0.299 * a[:, :, 0] + 0.587 * a[:, :, 1] + 0.114 * a[:, :, 2])
    if b.ndim == 3:
        b = (0.299 * b[:, :, 0] + 0.587 * b[:, :, 1] + 0.114 * b[:, :, 2])
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mu_x, mu_y = a.mean(), b.mean()
    sig_x = ((a - mu_x) ** 2).mean()
    sig_y = ((b - mu_y) ** 2).mean()
    sig_xy = ((a - mu_x) * (b - mu_y)).mean()
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    return float(((2 * mu_x * mu_y + c1) * (2 * sig_xy + c2)) / ((mu_x * mu_x + mu_y * mu_y + c1) * (sig_x + sig_y + c2)))


def bucket(progress: float) -> str:
    """validation_reporting.py:40-45."""
    if progress < 0.20 or progress > 0.80:
        return "front"
    if 0.35 <= progress <= 0.65:
        return "profile"
    return "rear"


def report_rows(exports: list, renders: dict, gts: dict) -> dict:
    """validation_reporting.py:58-105 without the file system: `exports` are the manifest rows, `renders` /
    `gts` map a source file name to a float32 [H,W,3] array.  Returns {"summary", "rows"}."""
    metrics = []
    max_index = max((int(r.get("index", 0)) for r in exports), default=1)
    for row in exports:
        idx, name = int(row["index"]), row["source"]
        if name not in renders or name not in gts:
            continue
        progress = idx / max(1, max_index)
        metrics.append({"index": idx, "frame": name, "progress": progress, "bucket": bucket(progress),
                        "psnr": psnr(renders[name], gts[name]), "ssim": ssim_global(renders[name], gts[name])})
    summary = {"count": len(metrics), "by_bucket": {}}
    for bk in ("front", "profile", "rear"):
        vals = [m for m in metrics if m["bucket"] == bk]
        if not vals:
            summary["by_bucket"][bk] = {"count": 0, "psnr": None, "ssim": None}
            continue
        summary["by_bucket"][bk] = {"count": len(vals), "psnr": float(np.mean([v["psnr"] for v in vals])),
                                    "ssim": float(np.mean([v["ssim"] for v in vals]))}
    return {"summary": summary, "rows": metrics}


def frame_moments(a_u8: np.ndarray, b_u8: np.ndarray) -> np.ndarray:
    """What omfs_frame_metrics accumulates per frame pair, in float64: [sum (a-b)^2 over all channel values,
    sum x, sum y, sum x^2, sum y^2, sum xy] with x, y the float32 BT.601 luma (validation_reporting.py:24-27
    order of operations).  a_u8, b_u8: [T,H,W,3] uint8."""
    out = np.empty((a_u8.shape[0], 6), dtype=np.float64)
    for t in range(a_u8.shape[0]):
        a, b = a_u8[t].astype(np.float32), b_u8[t].astype(np.float32)
        d = a_u8[t].astype(np.int64) - b_u8[t].astype(np.int64)
        x = (np.float32(0.299) * a[:, :, 0] + np.float32(0.587) * a[:, :, 1] + np.float32(0.114) * a[:, :, 2]).astype(np.float64)
        y = (np.float32(0.299) * b[:, :, 0] + np.float32(0.587) * b[:, :, 1] + np.float32(0.114) * b[:, :, 2]).astype(np.float64)
        out[t] = [float((d * d).sum()), x.sum(), y.sum(), (x * x).sum(), (y * y).sum(), (x * y).sum()]
    return out


# ---------------------------------------------------------------------------- fixture of test_surgical_sim
def uv_sphere(radius=30.0, center=(0.0, 0.0, 0.0), theta_resolution=20, phi_resolution=20) -> np.ndarray:
    """Point set with the layout of pv.Sphere / vtkSphereSource (test/test_surgical_sim.py:18-25):
    two poles, then theta_resolution meridians x (phi_resolution - 2) parallels = 362 points at the
    default 20 x 20.  Only the point cloud matters for R5/R6 (no triangle clipping here)."""
    pts = [(0.0, 0.0, radius), (0.0, 0.0, -radius)]
    n_par = phi_resolution - 2
    for i in range(theta_resolution):
        theta = 2.0 * math.pi * i / theta_resolution
        for j in range(n_par):
            phi = math.pi * (j + 1) / (phi_resolution - 1)
            pts.append((radius * math.sin(phi) * math.cos(theta), radius * math.sin(phi) * math.sin(theta),
                        radius * math.cos(phi)))
    return (np.array(pts, dtype=np.float64) + np.asarray(center, dtype=np.float64)).astype(np.float32)
