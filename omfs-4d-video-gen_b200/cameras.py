"""Host-side camera setup for the surgery-render path.

The reference never builds a camera itself: it hands `transforms_*.json`
(written by /root/reference/02_Visual_Engine/preprocess_video.py:372-401) to the
un-vendored GaussianAvatars renderer.  The conventions below restate the public
3DGS / GaussianAvatars dataset reader [UPSTREAM, unverifiable offline]:

  * `transform_matrix` is camera-to-world in the OpenGL/Blender convention
    (x right, y up, z back); the y and z camera axes are negated to get the
    COLMAP convention (y down, z forward);
  * world->view is its inverse; matrices are handed to the kernels in the
    column-major ("transposed") form `m[col*4 + row]`, so that a point maps as
    `x' = m[0]*x + m[4]*y + m[8]*z + m[12]`;
  * the vertical field of view is derived from `camera_angle_x` and the image
    aspect; znear = 0.01, zfar = 100.

Everything here runs once per camera on the host in float64 and is rounded to
float32 at the end; both the CUDA path and the oracle consume the same 39-float
record, so camera construction is not part of the parity surface.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

ZNEAR = 0.01
ZFAR = 100.0

# One camera = 40 floats (padded), the layout the C-ABI takes (include/omfs_b200.h: omfs_camera)
CAM_FLOATS = 40


@dataclass
class Camera:
    width: int
    height: int
    fovx: float
    fovy: float
    viewmatrix: np.ndarray  # (16,) float32, column-major world->view
    projmatrix: np.ndarray  # (16,) float32, column-major world->clip (view @ proj)
    campos: np.ndarray      # (3,) float32

    @property
    def tanfovx(self) -> float:
        return math.tan(self.fovx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.fovy * 0.5)

    def pack(self) -> np.ndarray:
        """Flatten to the 40-float record: view[16] proj[16] campos[3] tanfovx tanfovy focal_x focal_y pad."""
        out = np.zeros(CAM_FLOATS, dtype=np.float32)
        out[0:16] = self.viewmatrix
        out[16:32] = self.projmatrix
        out[32:35] = self.campos
        out[35] = np.float32(self.tanfovx)
        out[36] = np.float32(self.tanfovy)
        # focal = W / (2 tan) evaluated in float32 exactly as the kernels would
        out[37] = np.float32(self.width) / (np.float32(2.0) * out[35])
        out[38] = np.float32(self.height) / (np.float32(2.0) * out[36])
        return out


def fov2focal(fov: float, pixels: float) -> float:
    return pixels / (2.0 * math.tan(fov / 2.0))


def focal2fov(focal: float, pixels: float) -> float:
    return 2.0 * math.atan(pixels / (2.0 * focal))


def _projection(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    tan_y = math.tan(fovy / 2.0)
    tan_x = math.tan(fovx / 2.0)
    top = tan_y * znear
    right = tan_x * znear
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 2.0 * znear / (2.0 * right)
    P[1, 1] = 2.0 * znear / (2.0 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_c2w(c2w_gl: np.ndarray, camera_angle_x: float, width: int, height: int) -> Camera:
    """Build a camera record from one `transforms_*.json` frame entry."""
    c2w = np.array(c2w_gl, dtype=np.float64).reshape(4, 4).copy()
    c2w[:3, 1:3] *= -1.0  # OpenGL -> COLMAP camera axes
    w2c = np.linalg.inv(c2w)
    fovx = float(camera_angle_x)
    fovy = focal2fov(fov2focal(fovx, width), height)
    proj = _projection(ZNEAR, ZFAR, fovx, fovy)
    full = proj @ w2c  # row-major math: clip = P * V * p
    campos = c2w[:3, 3]
    return Camera(
        width=int(width),
        height=int(height),
        fovx=fovx,
        fovy=fovy,
        viewmatrix=np.ascontiguousarray(w2c.T.reshape(-1)).astype(np.float32),
        projmatrix=np.ascontiguousarray(full.T.reshape(-1)).astype(np.float32),
        campos=campos.astype(np.float32),
    )


def look_at_c2w(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """OpenGL-convention camera-to-world for a camera at `eye` looking at `target`."""
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    up = np.asarray(up, dtype=np.float64)
    back = eye - target
    back /= np.linalg.norm(back)
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    true_up = np.cross(back, right)
    c2w = np.eye(4, dtype=np.float64)
    c2w[:3, 0] = right
    c2w[:3, 1] = true_up
    c2w[:3, 2] = back
    c2w[:3, 3] = eye
    return c2w


def ring_cameras(n_views: int, radius: float, target, camera_angle_x: float, width: int, height: int) -> list[Camera]:
    """`n_views` cameras on a horizontal ring around `target` (config 4 of BASELINE.json)."""
    cams = []
    target = np.asarray(target, dtype=np.float64)
    for i in range(n_views):
        a = 2.0 * math.pi * i / n_views
        eye = target + radius * np.array([math.sin(a), 0.0, math.cos(a)])
        cams.append(camera_from_c2w(look_at_c2w(eye, target), camera_angle_x, width, height))
    return cams
