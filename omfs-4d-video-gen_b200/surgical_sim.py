"""Drop-in for the numeric part of the reference's `01_Clinical_Engine/surgical_sim.py`, on the GPU.

`SurgicalCutter` keeps the reference's method names, arguments, return keys and exceptions
(reference :59-329) but works on point sets instead of PyVista meshes: `perform_cut` classifies
points with the reference's half-space rule (:180-204) and `move_segments` applies its rigid move
(:293-322: rotate X(pitch) -> Z(yaw) -> Y(roll) in degrees about the segment's BOUNDING-BOX centre,
then translate by unit(direction) * mm).  VTK's triangle clipping — which inserts new vertices along
the cut — is UI geometry and out of scope.

The per-point work (masks, bounding boxes, moves) runs in `omfs_displace_points`
(csrc/displace.cu, float64, bit-exact against oracle/reference_rows.py); plane normals and rotation
matrices are evaluated on the host in float64 exactly as the reference does.  No CPU fallback.

`plan_displacement_field` is the bridge the reference lacks: it turns a
surgical plan into a canonical-space FLAME vertex displacement that the render session folds into
the subject (omfs_session_set_subject's plan_offset), so the triangle-bound Gaussians follow.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np


def _angle_to_normal(base_normal: tuple, pitch_deg: float, yaw_deg: float) -> tuple:
    """n = unit(Rz(yaw) Rx(pitch) n0), float64 (reference :25-47)."""
    n = np.array(base_normal, dtype=float)
    p, y = np.radians(pitch_deg), np.radians(yaw_deg)
    rx = np.array([[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])
    rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    r = rz @ rx @ n
    length = np.linalg.norm(r)
    if length < 1e-12:
        return tuple(base_normal)
    return tuple(r / length)


def _normalise_direction(direction) -> np.ndarray:
    """Unit vector; ValueError for a (near-)zero direction (reference :50-56)."""
    vec = np.array(direction, dtype=float)
    length = np.linalg.norm(vec)
    if length < 1e-12:
        raise ValueError("advancement_direction must be a non-zero vector.")
    return vec / length


def _rotation(pitch: float, yaw: float, roll: float) -> np.ndarray:
    """rotate_x(pitch) then rotate_z(yaw) then rotate_y(roll), right-handed, degrees: R = Ry Rz Rx."""
    R = np.eye(3)
    for angle, axis in ((pitch, 0), (yaw, 2), (roll, 1)):
        if angle == 0.0:
            continue
        c, s = math.cos(math.radians(angle)), math.sin(math.radians(angle))
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        M = np.eye(3)
        M[i, i], M[i, j], M[j, i], M[j, j] = c, -s, s, c
        R = M @ R
    return R


@dataclass
class PointMesh:
    """Stand-in for pv.PolyData: a float32 point set (faces are carried along untouched)."""
    points: np.ndarray
    faces: np.ndarray | None = None

    def __post_init__(self):
        self.points = np.ascontiguousarray(self.points, dtype=np.float32).reshape(-1, 3)

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def center(self) -> list[float]:
        if not self.n_points:
            return [0.0, 0.0, 0.0]
        lo, hi = self.points.min(axis=0).astype(np.float64), self.points.max(axis=0).astype(np.float64)
        return list((lo + hi) * 0.5)

    @property
    def bounds(self):
        lo, hi = self.points.min(axis=0), self.points.max(axis=0)
        return (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def copy(self) -> "PointMesh":
        return PointMesh(self.points.copy(), self.faces)

    def merge(self, other: "PointMesh") -> "PointMesh":
        return PointMesh(np.concatenate([self.points, other.points]))


def _device_displace(points: np.ndarray, planes: np.ndarray, moves: np.ndarray, jaw_weight, mandible_first: int):
    from . import runtime
    from .runtime import DeviceArray as DA
    L = runtime.load_library()
    P = len(points)
    d_pts = DA.from_numpy(np.ascontiguousarray(points, dtype=np.float32))
    d_mask, d_out, d_bbox = DA((P,), np.uint8), DA((P, 3), np.float32), DA((12,), np.float32)
    d_w = None if jaw_weight is None else DA.from_numpy(np.ascontiguousarray(jaw_weight, dtype=np.float32))
    pl = (ctypes.c_double * 24)(*np.asarray(planes, dtype=np.float64).reshape(-1))
    mv = (ctypes.c_double * 24)(*np.asarray(moves, dtype=np.float64).reshape(-1))
    runtime.check(L.omfs_displace_points(P, d_pts.ptr, pl, mv, None if d_w is None else d_w.ptr, int(mandible_first),
                                         d_mask.ptr, d_out.ptr, d_bbox.ptr, None))
    return d_out.numpy(), d_mask.numpy(), d_bbox.numpy().reshape(2, 6)


class SurgicalCutter:
    """Three planes (Le Fort I, BSSO left, BSSO right) -> four segments; two of them move."""

    def __init__(self, maxilla_mesh: PointMesh, mandible_mesh: PointMesh | None = None):
        self.maxilla = maxilla_mesh
        self.mandible = mandible_mesh
        self.has_separate = mandible_mesh is not None and mandible_mesh.n_points > 0
        self.upper_skull = self.mobile_maxilla = self.distal_mandible = self.proximal_rami = None
        self._planes = None

    def get_combined_mesh(self) -> PointMesh:
        return self.maxilla.merge(self.mandible) if self.has_separate else self.maxilla

    def _plane_table(self, lefort_z, bsso_l_x, bsso_r_x, lefort_pitch, lefort_yaw, bsso_l_pitch, bsso_l_yaw,
                     bsso_r_pitch, bsso_r_yaw) -> np.ndarray:
        c = self.get_combined_mesh().center
        planes = np.zeros((3, 8), dtype=np.float64)
        planes[0, :3], planes[0, 3:6] = _angle_to_normal((0, 0, 1), lefort_pitch, lefort_yaw), (c[0], c[1], lefort_z)
        planes[1, :3], planes[1, 3:6] = _angle_to_normal((1, 0, 0), bsso_l_pitch, bsso_l_yaw), (bsso_l_x, c[1], c[2])
        planes[2, :3], planes[2, 3:6] = _angle_to_normal((1, 0, 0), bsso_r_pitch, bsso_r_yaw), (bsso_r_x, c[1], c[2])
        return planes

    def preview_planes(self, lefort_z: float, bsso_l_x: float, bsso_r_x: float, lefort_pitch: float = 0.0,
                       lefort_yaw: float = 0.0, bsso_l_pitch: float = 0.0, bsso_l_yaw: float = 0.0,
                       bsso_r_pitch: float = 0.0, bsso_r_yaw: float = 0.0) -> dict:
        planes = self._plane_table(lefort_z, bsso_l_x, bsso_r_x, lefort_pitch, lefort_yaw, bsso_l_pitch,
                                   bsso_l_yaw, bsso_r_pitch, bsso_r_yaw)
        combined = self.get_combined_mesh()
        b = combined.bounds
        size = max(b[1] - b[0], b[3] - b[2], b[5] - b[4]) * 1.2
        mk = lambda i: {"center": tuple(planes[i, 3:6]), "direction": tuple(planes[i, :3]), "size": float(size)}
        return {"maxilla": self.maxilla, "mandible": self.mandible, "combined": combined,
                "lefort": mk(0), "bsso_l": mk(1), "bsso_r": mk(2)}

    def perform_cut(self, lefort_z: float, bsso_l_x: float, bsso_r_x: float, lefort_pitch: float = 0.0,
                    lefort_yaw: float = 0.0, bsso_l_pitch: float = 0.0, bsso_l_yaw: float = 0.0,
                    bsso_r_pitch: float = 0.0, bsso_r_yaw: float = 0.0, lefort_flip: bool = False) -> dict:
        planes = self._plane_table(lefort_z, bsso_l_x, bsso_r_x, lefort_pitch, lefort_yaw, bsso_l_pitch,
                                   bsso_l_yaw, bsso_r_pitch, bsso_r_yaw)
        if (not self.has_separate) and lefort_flip:
            planes[0, :3] = -planes[0, :3]  # single-mesh fallback: the mobile side is the other half-space
        self._planes = planes
        identity = np.zeros((2, 12))
        identity[:, [0, 4, 8]] = 1.0
        if self.has_separate:
            pts = np.concatenate([self.maxilla.points, self.mandible.points])
            n_max = self.maxilla.n_points
            _, mask, _ = _device_displace(pts, planes, identity, None, n_max)
            is_mand = np.arange(len(pts)) >= n_max
            self._points, self._mask, self._mand_first = pts, mask, n_max
            upper = pts[(~is_mand) & ((mask & 1) == 0)]
            mobile = pts[(mask & 8) != 0]
            distal = pts[(mask & 16) != 0]
            rami = np.concatenate([pts[is_mand & ((mask & 2) == 0)], pts[is_mand & ((mask & 4) == 0)]])
        else:
            # best effort on one mesh: everything on the mobile side of Le Fort is "maxilla"; the
            # BSSO slab is taken from the same points (the reference clips the same mesh twice)
            pts = self.maxilla.points
            _, mask_a, _ = _device_displace(pts, planes, identity, None, len(pts))
            _, mask_b, _ = _device_displace(pts, planes, identity, None, 0)
            self._points, self._mask, self._mand_first = pts, mask_a, len(pts)
            upper = pts[(mask_a & 1) == 0]
            mobile = pts[(mask_a & 8) != 0]
            distal = pts[(mask_b & 16) != 0]
            rami = np.concatenate([pts[(mask_b & 2) == 0], pts[(mask_b & 4) == 0]])
        self.upper_skull, self.mobile_maxilla = PointMesh(upper), PointMesh(mobile)
        self.distal_mandible, self.proximal_rami = PointMesh(distal), PointMesh(rami)
        return {"upper_skull": self.upper_skull, "mobile_maxilla": self.mobile_maxilla,
                "distal_mandible": self.distal_mandible, "proximal_rami": self.proximal_rami}

    def move_segments(self, maxilla_mm: float = 0.0, mandible_mm: float = 0.0,
                      advancement_direction: tuple[float, float, float] = (0.0, 1.0, 0.0),
                      maxilla_rotation: tuple[float, float, float] = (0.0, 0.0, 0.0),
                      mandible_rotation: tuple[float, float, float] = (0.0, 0.0, 0.0)) -> dict:
        if self.mobile_maxilla is None or self.distal_mandible is None:
            raise RuntimeError("Call perform_cut() before move_segments().")
        adv = _normalise_direction(advancement_direction)
        moves = np.zeros((2, 12), dtype=np.float64)
        for i, (mm, rot) in enumerate(((maxilla_mm, maxilla_rotation), (mandible_mm, mandible_rotation))):
            R = _rotation(*rot) if any(r != 0.0 for r in rot) else np.eye(3)
            moves[i, :9] = R.reshape(-1)
            moves[i, 9:] = adv * mm
        # each mobile segment is its own point set, moved about ITS bounding-box centre
        out = {}
        for key, seg, slot in (("mobile_maxilla", self.mobile_maxilla, 0), ("distal_mandible", self.distal_mandible, 1)):
            if seg.n_points == 0:
                out[key] = seg.copy()
                continue
            # planes that put every point of this set into segment `slot`
            far = 1e30
            planes = np.zeros((3, 8))
            planes[0, :3], planes[0, 3:6] = (0, 0, 1), (0, 0, far)      # below Le Fort: all
            planes[1, :3], planes[1, 3:6] = (1, 0, 0), (-far, 0, 0)     # right of BSSO-L: all
            planes[2, :3], planes[2, 3:6] = (1, 0, 0), (far, 0, 0)      # left of BSSO-R: all
            moved, mask, _ = _device_displace(seg.points, planes, moves, None, 0 if slot == 1 else seg.n_points)
            assert np.all((mask & (8 if slot == 0 else 16)) != 0)
            out[key] = PointMesh(moved)
        return {"upper_skull": self.upper_skull, "mobile_maxilla": out["mobile_maxilla"],
                "distal_mandible": out["distal_mandible"], "proximal_rami": self.proximal_rami}


def plan_displacement_field(canonical_verts: np.ndarray, jaw_weight: np.ndarray, planes_flame: np.ndarray,
                            maxilla_mm: float = 0.0, mandible_mm: float = 0.0,
                            advancement_direction=(0.0, 0.0, 1.0), maxilla_rotation=(0.0, 0.0, 0.0),
                            mandible_rotation=(0.0, 0.0, 0.0), sensitivity: float = 1.0,
                            scale_factor: float = 0.001):
    """Surgical plan -> canonical-space FLAME vertex displacement [V,3] (+ the per-vertex masks).

    Vertices skinned mostly to the jaw joint (lbs weight > 0.5) between the two BSSO planes follow the
    distal mandible; the other vertices below the Le Fort plane follow the mobile maxilla; everything
    else stays.  Millimetres become FLAME units with the reference's own factor
    (mm * sensitivity * 0.001, render_surgery.py:35-42).  New capability — the reference only has the
    two scalar parameter edits."""
    adv = _normalise_direction(advancement_direction)
    unit = sensitivity * scale_factor
    moves = np.zeros((2, 12), dtype=np.float64)
    for i, (mm, rot) in enumerate(((maxilla_mm, maxilla_rotation), (mandible_mm, mandible_rotation))):
        R = _rotation(*rot) if any(r != 0.0 for r in rot) else np.eye(3)
        moves[i, :9] = R.reshape(-1)
        moves[i, 9:] = adv * mm * unit
    verts = np.ascontiguousarray(canonical_verts, dtype=np.float32)
    moved, mask, _ = _device_displace(verts, planes_flame, moves, jaw_weight, len(verts))
    return (moved - verts).astype(np.float32), mask
