"""Seeded synthetic fixtures for the surgery-render path.

Nothing real is available offline: the FLAME pickle is licence-gated
(/root/reference/.gitignore:28-29, 02_Visual_Engine/flame_fitter.py:454-458), no
trained avatar or video ships with the reference.  Everything here is generated
from `numpy.random.default_rng(seed)` with the shapes the reference's on-disk
formats use (flame_fitter.py:431-441, preprocess_video.py:314-354).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import cameras as cam_mod

FLAME_V = 5143          # FLAME-with-teeth vertex count (flame_fitter.py:439-440)
N_SHAPE = 300
N_EXPR = 100
N_JOINTS = 5
N_POSE_FEAT = 36        # (5-1) joints x 9
PARENTS = np.array([-1, 0, 1, 1, 1], dtype=np.int32)


@dataclass
class FlameModel:
    """FLAME-shaped linear model.  Arrays are float32, C-contiguous."""
    v_template: np.ndarray   # (V,3)
    faces: np.ndarray        # (F,3) int32
    shapedirs: np.ndarray    # (400, 3V)   rows 0..299 shape, 300..399 expression
    posedirs: np.ndarray     # (36, 3V)
    j_regressor: np.ndarray  # (5, V)
    lbs_weights: np.ndarray  # (V, 5)
    parents: np.ndarray = field(default_factory=lambda: PARENTS.copy())

    @property
    def n_verts(self) -> int:
        return self.v_template.shape[0]

    @property
    def n_faces(self) -> int:
        return self.faces.shape[0]


@dataclass
class FrameParams:
    """The flame_param.npz record (flame_fitter.py:5-12, 431-441)."""
    shape: np.ndarray           # (300,)
    expr: np.ndarray            # (T,100)
    rotation: np.ndarray        # (T,3)
    neck_pose: np.ndarray       # (T,3)
    jaw_pose: np.ndarray        # (T,3)
    eyes_pose: np.ndarray       # (T,6)
    translation: np.ndarray     # (T,3)
    static_offset: np.ndarray   # (1,V,3)
    dynamic_offset: np.ndarray  # (T,V,3)

    @property
    def n_frames(self) -> int:
        return self.expr.shape[0]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k in (
            "shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose",
            "translation", "static_offset", "dynamic_offset")}

    @classmethod
    def from_dict(cls, d: dict, n_verts: int = FLAME_V) -> "FrameParams":
        def two_d(a, cols):
            a = np.asarray(a, dtype=np.float32)
            return a.reshape(-1, cols)
        expr = two_d(d["expr"], np.asarray(d["expr"]).shape[-1])
        T = expr.shape[0]
        so = np.asarray(d.get("static_offset", np.zeros((1, n_verts, 3))), dtype=np.float32).reshape(1, -1, 3)
        do = d.get("dynamic_offset")
        do = np.zeros((T, so.shape[1], 3), np.float32) if do is None else np.asarray(do, np.float32).reshape(T, -1, 3)
        return cls(
            shape=np.asarray(d["shape"], dtype=np.float32).reshape(-1),
            expr=expr,
            rotation=two_d(d["rotation"], 3),
            neck_pose=two_d(d["neck_pose"], 3),
            jaw_pose=two_d(d["jaw_pose"], 3),
            eyes_pose=two_d(d["eyes_pose"], 6),
            translation=two_d(d["translation"], 3),
            static_offset=so,
            dynamic_offset=do,
        )

    def slice(self, lo: int, hi: int) -> "FrameParams":
        return FrameParams(self.shape, self.expr[lo:hi], self.rotation[lo:hi], self.neck_pose[lo:hi],
                           self.jaw_pose[lo:hi], self.eyes_pose[lo:hi], self.translation[lo:hi],
                           self.static_offset, self.dynamic_offset[lo:hi])


@dataclass
class Avatar:
    """Triangle-bound Gaussian avatar, in the raw (pre-activation) form the
    GaussianAvatars PLY stores [UPSTREAM]."""
    xyz: np.ndarray       # (N,3) face-local position
    scaling: np.ndarray   # (N,3) log-scale
    rotation: np.ndarray  # (N,4) wxyz, unnormalised
    opacity: np.ndarray   # (N,)  logit
    sh: np.ndarray        # (N,16,3) SH degree-3 coefficients
    binding: np.ndarray   # (N,) int32 parent-face index

    @property
    def n(self) -> int:
        return self.xyz.shape[0]


def _fibonacci_ellipsoid(n: int, radii) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = math.pi * (1.0 + 5.0 ** 0.5) * i
    p = np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], axis=1)
    return p * np.asarray(radii, dtype=np.float64)[None, :]


def make_flame_model(seed: int = 1234, n_verts: int = FLAME_V) -> FlameModel:
    """Head-sized ellipsoid (radii ~0.08/0.11/0.09 m), convex-hull faces (2V-4 triangles)."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    pts = _fibonacci_ellipsoid(n_verts, (0.08, 0.11, 0.09))
    hull = ConvexHull(pts)
    faces = hull.simplices.astype(np.int32)
    # orient outwards so that face frames are consistent
    c = pts[faces].mean(axis=1)
    nrm = np.cross(pts[faces[:, 1]] - pts[faces[:, 0]], pts[faces[:, 2]] - pts[faces[:, 0]])
    flip = (nrm * c).sum(axis=1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    faces = np.ascontiguousarray(faces[order])

    V = n_verts
    spectrum = 1.0 / np.sqrt(1.0 + np.arange(N_SHAPE + N_EXPR, dtype=np.float64) % N_SHAPE * 0.05)
    # smooth directions: low-order functions of position plus a little noise
    basis_pts = pts / np.array([0.08, 0.11, 0.09])
    freq = rng.normal(0.0, 2.0, size=(N_SHAPE + N_EXPR, 3))
    phase = rng.uniform(0, 2 * math.pi, size=(N_SHAPE + N_EXPR, 1))
    amp = rng.normal(0.0, 1.0, size=(N_SHAPE + N_EXPR, 1, 3))
    wave = np.sin(freq @ basis_pts.T + phase)                      # (K,V)
    shapedirs = (wave[:, :, None] * amp) * (1e-3 * spectrum)[:, None, None]
    shapedirs += rng.normal(0.0, 1e-4, size=shapedirs.shape) * spectrum[:, None, None]
    shapedirs = shapedirs.reshape(N_SHAPE + N_EXPR, 3 * V).astype(np.float32)

    posedirs = rng.normal(0.0, 1e-4, size=(N_POSE_FEAT, 3 * V)).astype(np.float32)

    seeds = np.array([[0.0, -0.02, 0.0],     # root
                      [0.0, -0.09, -0.02],   # neck
                      [0.0, -0.04, 0.05],    # jaw
                      [0.03, 0.03, 0.07],    # eye L
                      [-0.03, 0.03, 0.07]])  # eye R
    d = np.linalg.norm(pts[:, None, :] - seeds[None, :, :], axis=2)  # (V,5)
    logits = -d / 0.02
    logits[:, 0] += 1.0
    w = np.exp(logits - logits.max(axis=1, keepdims=True))
    w /= w.sum(axis=1, keepdims=True)
    lbs_weights = w.astype(np.float32)

    jr = np.zeros((N_JOINTS, V), dtype=np.float64)
    for j in range(N_JOINTS):
        near = np.argsort(d[:, j])[:64]
        wj = rng.uniform(0.2, 1.0, size=near.size)
        jr[j, near] = wj / wj.sum()
    return FlameModel(
        v_template=pts.astype(np.float32),
        faces=faces,
        shapedirs=shapedirs,
        posedirs=posedirs,
        j_regressor=jr.astype(np.float32),
        lbs_weights=lbs_weights,
    )


def make_frame_params(n_frames: int, seed: int = 99, n_verts: int = FLAME_V,
                      dynamic: bool = False) -> FrameParams:
    rng = np.random.default_rng(seed)
    T = n_frames

    def ar1(cols, sigma, rho=0.9):
        x = np.zeros((T, cols))
        x[0] = rng.normal(0, sigma, cols)
        for t in range(1, T):
            x[t] = rho * x[t - 1] + math.sqrt(1 - rho * rho) * rng.normal(0, sigma, cols)
        return x

    jaw = ar1(3, 0.1)
    jaw[:, 0] = np.abs(jaw[:, 0])
    dyn = rng.normal(0, 2e-4, size=(T, n_verts, 3)) if dynamic else np.zeros((T, n_verts, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return FrameParams(
        shape=f32(rng.normal(0, 1, N_SHAPE)),
        expr=f32(ar1(N_EXPR, 0.5)),
        rotation=f32(ar1(3, 0.1)),
        neck_pose=f32(ar1(3, 0.05)),
        jaw_pose=f32(jaw),
        eyes_pose=f32(ar1(6, 0.05)),
        translation=f32(ar1(3, 0.005)),
        static_offset=f32(rng.normal(0, 1e-3, size=(1, n_verts, 3))),
        dynamic_offset=f32(dyn),
    )


def make_avatar(n_gauss: int, n_faces: int, seed: int = 7) -> Avatar:
    rng = np.random.default_rng(seed)
    N = n_gauss
    if N >= n_faces:
        binding = np.concatenate([np.arange(n_faces), rng.integers(0, n_faces, N - n_faces)])
    else:
        binding = rng.integers(0, n_faces, N)
    # keep the avatar in the order a trained one has (unordered in face index)
    binding = binding[rng.permutation(N)].astype(np.int32)
    xyz = rng.normal(0, 0.3, size=(N, 3))
    xyz[:, 2] *= 0.15  # mostly in the triangle plane, as trained avatars are
    scaling = np.log(rng.uniform(0.05, 0.5, size=(N, 3)))
    scaling[:, 2] -= 1.0
    rot = rng.normal(0, 1, size=(N, 4))
    opacity = rng.normal(2.0, 1.5, size=N)
    sh = np.zeros((N, 16, 3))
    sh[:, 0, :] = rng.uniform(-1.0, 1.0, size=(N, 3))
    sh[:, 1:, :] = rng.normal(0, 0.1, size=(N, 15, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return Avatar(f32(xyz), f32(scaling), f32(rot), f32(opacity), f32(sh), binding)


def make_camera(width: int = 512, height: int = 512, camera_angle_x: float = 0.3,
                fill: float = 0.6, head_height: float = 0.22, target=(0.0, 0.0, 0.0)) -> cam_mod.Camera:
    """Frontal pinhole camera with the head filling `fill` of the frame height."""
    fovy = cam_mod.focal2fov(cam_mod.fov2focal(camera_angle_x, width), height)
    dist = (head_height / fill) / (2.0 * math.tan(fovy / 2.0))
    eye = np.asarray(target, dtype=np.float64) + np.array([0.0, 0.0, dist])
    return cam_mod.camera_from_c2w(cam_mod.look_at_c2w(eye, target), camera_angle_x, width, height)


def camera_distance(width: int, height: int, camera_angle_x: float = 0.3, fill: float = 0.6,
                    head_height: float = 0.22) -> float:
    fovy = cam_mod.focal2fov(cam_mod.fov2focal(camera_angle_x, width), height)
    return (head_height / fill) / (2.0 * math.tan(fovy / 2.0))


def make_scene(n_gauss: int = 100_000, n_frames: int = 1, width: int = 512, height: int = 512,
               n_verts: int = FLAME_V, seed: int = 0, dynamic: bool = False):
    """Model, frame parameters, avatar and one frontal camera — config 2/3 of BASELINE.json by default."""
    model = make_flame_model(seed=1234 + seed, n_verts=n_verts)
    params = make_frame_params(n_frames, seed=99 + seed, n_verts=n_verts, dynamic=dynamic)
    avatar = make_avatar(n_gauss, model.n_faces, seed=7 + seed)
    cam = make_camera(width, height)
    return model, params, avatar, cam
