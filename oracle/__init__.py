"""ctypes front end of the CPU oracle (oracle/omfs_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of omfs_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64 = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "omfs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_int, c_i64, c_u64, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p
        L.orc_flame_forward.argtypes = [c_int, c_int, c_int] + [_f] * 5 + [_f] * 7 + [vp, vp, vp, _f, vp]
        L.orc_flame_forward.restype = None
        L.orc_face_frames.argtypes = [c_int, c_int, c_int, _f, _i32, _f]
        L.orc_bind_preprocess.argtypes = [c_int, c_int, c_int, c_int, _f, _f, _f, _f, _f, _f, _f, _f, _f, _u32, vp]
        L.orc_scan.argtypes = [c_i64, _u32, _u32]
        L.orc_scan.restype = c_u64
        L.orc_emit_keys.argtypes = [c_int, c_int, c_int, c_int, _f, _u32, _u32, _u64, _u32]
        L.orc_sort_pairs.argtypes = [c_u64, c_int, _u64, _u32, _u64, _u32]
        L.orc_tile_ranges.argtypes = [c_u64, c_i64, _u64, _u32]
        L.orc_composite.argtypes = [c_int, c_int, c_int, c_int, _f, _f, _f, _u32, _u32, _f, _f, vp]
        L.orc_to_uint8.argtypes = [c_int, c_int, c_int, _f, _u8]
        L.orc_num_threads.restype = c_int
        L.orc_set_num_threads.argtypes = [c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """Set the OpenMP thread count of the oracle's loops (a torchrun rank inherits OMP_NUM_THREADS=1); returns the
    count now in effect."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def _opt(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.ctypes.data_as(ctypes.c_void_p), a


def flame_forward(model, params, plan_offset=None, return_joints=False):
    """U1-U3.  `model` / `params` are duck-typed (synthetic.FlameModel / FrameParams)."""
    T = params.expr.shape[0]
    V = model.v_template.shape[0]
    n_expr = params.expr.shape[1]
    verts = np.empty((T, V, 3), dtype=np.float32)
    keep = []

    def opt(a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(ctypes.c_void_p)

    joints = np.empty((T, 5, 3), dtype=np.float32) if return_joints else None
    so = params.static_offset
    do = params.dynamic_offset
    if do is not None and not np.any(do):
        do = None
    c = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    lib().orc_flame_forward(
        T, V, n_expr, c(model.v_template), c(model.shapedirs), c(model.posedirs), c(model.j_regressor),
        c(model.lbs_weights), c(params.shape), c(params.expr), c(params.rotation), c(params.neck_pose),
        c(params.jaw_pose), c(params.eyes_pose), c(params.translation),
        opt(None if so is None else so.reshape(-1)), opt(None if do is None else do.reshape(T, -1)),
        opt(None if plan_offset is None else np.asarray(plan_offset).reshape(-1)),
        verts, None if joints is None else joints.ctypes.data_as(ctypes.c_void_p))
    return (verts, joints) if return_joints else verts


def face_frames(verts, faces):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    if verts.ndim == 2:
        verts = verts[None]
    T, V, _ = verts.shape
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    ff = np.empty((T, faces.shape[0], 20), dtype=np.float32)
    lib().orc_face_frames(T, V, faces.shape[0], verts, faces, ff)
    return ff


@dataclass
class Preprocessed:
    P0: np.ndarray            # [S,N,4] px py depth radius(int bits)
    P1: np.ndarray            # [S,N,4] ca cb cc lo
    P2: np.ndarray            # [S,N,4] r g b 0
    tiles_touched: np.ndarray  # [S,N] uint32
    mu: np.ndarray            # [S,N,3]

    @property
    def radii(self):
        return self.P0[..., 3].view(np.int32)


def bind_preprocess(ff, baked, cams, width, height, seg_frame=None):
    """U5+U6 for S segments.  ff is [T,F,20]; cams a list of packed 40-float records (one per
    segment); seg_frame[s] gives the frame index of segment s (default: s)."""
    S = len(cams)
    N = baked["xyzb"].shape[0]
    F = ff.shape[1]
    out = Preprocessed(np.zeros((S, N, 4), np.float32), np.zeros((S, N, 4), np.float32),
                       np.zeros((S, N, 4), np.float32), np.zeros((S, N), np.uint32),
                       np.zeros((S, N, 3), np.float32))
    for s in range(S):
        t = s if seg_frame is None else int(seg_frame[s])
        lib().orc_bind_preprocess(
            N, F, width, height, np.ascontiguousarray(ff[t]), baked["xyzb"], baked["scale_lo"], baked["rot"],
            baked["sh"], np.ascontiguousarray(cams[s], dtype=np.float32), out.P0[s], out.P1[s], out.P2[s],
            out.tiles_touched[s], out.mu[s].ctypes.data_as(ctypes.c_void_p))
    return out


@dataclass
class Binned:
    offsets: np.ndarray
    keys: np.ndarray
    values: np.ndarray
    sorted_keys: np.ndarray
    sorted_values: np.ndarray
    ranges: np.ndarray   # [S*tiles, 2] uint32
    n_pairs: int
    sort_bits: int


def sort_bits_for(n_segments: int, width: int, height: int) -> int:
    tiles = ((width + 15) // 16) * ((height + 15) // 16)
    total = max(1, n_segments * tiles)
    return 32 + max(1, int(total - 1).bit_length())


def binning(pre: Preprocessed, width, height):
    S, N = pre.tiles_touched.shape
    tiles = ((width + 15) // 16) * ((height + 15) // 16)
    tt = np.ascontiguousarray(pre.tiles_touched.reshape(-1))
    offsets = np.empty_like(tt)
    R = int(lib().orc_scan(tt.size, tt, offsets))
    keys = np.zeros(max(R, 1), np.uint64)
    values = np.zeros(max(R, 1), np.uint32)
    lib().orc_emit_keys(S, N, width, height, np.ascontiguousarray(pre.P0.reshape(-1)), tt, offsets, keys, values)
    bits = sort_bits_for(S, width, height)
    sk = np.zeros_like(keys)
    sv = np.zeros_like(values)
    lib().orc_sort_pairs(R, bits, keys, values, sk, sv)
    ranges = np.zeros((S * tiles, 2), np.uint32)
    lib().orc_tile_ranges(R, S * tiles, sk, ranges.reshape(-1))
    return Binned(offsets, keys[:R], values[:R], sk[:R], sv[:R], ranges, R, bits)


def composite(pre: Preprocessed, binned: Binned, width, height, bg=(1.0, 1.0, 1.0), count_evals=False):
    S, N = pre.tiles_touched.shape
    image = np.zeros((S, 3, height, width), np.float32)
    sv = binned.sorted_values if binned.n_pairs else np.zeros(1, np.uint32)
    ev = ctypes.c_uint64(0)
    lib().orc_composite(S, N, width, height, np.ascontiguousarray(pre.P0.reshape(-1)),
                        np.ascontiguousarray(pre.P1.reshape(-1)), np.ascontiguousarray(pre.P2.reshape(-1)),
                        np.ascontiguousarray(sv), np.ascontiguousarray(binned.ranges.reshape(-1)),
                        np.asarray(bg, dtype=np.float32), image.reshape(-1),
                        ctypes.cast(ctypes.pointer(ev), ctypes.c_void_p) if count_evals else None)
    return (image, int(ev.value)) if count_evals else image


def to_uint8(image):
    image = np.ascontiguousarray(image, dtype=np.float32)
    S, _, H, W = image.shape
    out = np.empty((S, H, W, 3), np.uint8)
    lib().orc_to_uint8(S, W, H, image.reshape(-1), out.reshape(-1))
    return out


@dataclass
class RenderResult:
    verts: np.ndarray
    ff: np.ndarray
    pre: Preprocessed
    binned: Binned
    image: np.ndarray


def render(model, params, baked, cams, width, height, bg=(1.0, 1.0, 1.0), seg_frame=None, plan_offset=None,
           verts=None):
    """Full chain for S = len(cams) segments.  `verts` overrides U1-U3 (the exact-domain hand-off)."""
    if verts is None:
        verts = flame_forward(model, params, plan_offset=plan_offset)
    ff = face_frames(verts, model.faces)
    pre = bind_preprocess(ff, baked, cams, width, height, seg_frame=seg_frame)
    b = binning(pre, width, height)
    img = composite(pre, b, width, height, bg=bg)
    return RenderResult(verts, ff, pre, b, img)
