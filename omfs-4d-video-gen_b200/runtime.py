"""ctypes binding of libomfs_b200.so (include/omfs_b200.h) and the Python-side render session.

This is the thin layer between the reference-facing Python entry points
(render_surgery.py in this package) and the C-ABI.  There is NO fallback: if the shared library is
missing or the device is not an sm_100 part, every call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_uint8, c_uint64, c_void_p

import numpy as np

from . import cameras as cam_mod

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libomfs_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "omfs_b200.h")

_lib = None


class OmfsError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name include/omfs_b200.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omfs_[a-z0-9_]+)\s*\(", text)))


class ModelDesc(ctypes.Structure):
    _fields_ = [("n_verts", c_int32), ("n_faces", c_int32), ("n_expr", c_int32), ("n_gauss", c_int32),
                ("v_template", c_void_p), ("shapedirs", c_void_p), ("posedirs", c_void_p),
                ("j_regressor", c_void_p), ("lbs_weights", c_void_p), ("faces", c_void_p),
                ("xyzb", c_void_p), ("scale_lo", c_void_p), ("rot", c_void_p), ("sh", c_void_p)]


class SessionConfig(ctypes.Structure):
    _fields_ = [("width", c_int32), ("height", c_int32), ("max_batch", c_int32), ("device", c_int32),
                ("gemm_impl", c_int32), ("debug_keys", c_int32), ("pair_capacity", c_uint64), ("bg", c_float * 3)]


class FramesDesc(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_views", c_int32), ("expr", c_void_p), ("rotation", c_void_p),
                ("neck_pose", c_void_p), ("jaw_pose", c_void_p), ("eyes_pose", c_void_p),
                ("translation", c_void_p), ("dynamic_offset", c_void_p), ("cams", c_void_p)]


def load_library():
    """dlopen the C-ABI library and check that it exports everything the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OmfsError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                        "there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise OmfsError(f"libomfs_b200.so does not export: {missing}")
    L.omfs_last_error.restype = c_char_p
    L.omfs_launch_count.restype = ctypes.c_ulonglong
    L.omfs_binning_workspace_bytes.restype = c_size_t
    L.omfs_binning_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_size_t]
    L.omfs_session_stream.restype = c_void_p
    L.omfs_session_stream.argtypes = [c_void_p]
    L.omfs_session_create.argtypes = [POINTER(ModelDesc), POINTER(SessionConfig), POINTER(c_void_p)]
    L.omfs_session_destroy.argtypes = [c_void_p]
    L.omfs_session_destroy.restype = None
    L.omfs_session_set_subject.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.omfs_session_render_host.argtypes = [c_void_p, POINTER(FramesDesc), c_void_p, c_void_p]
    L.omfs_session_render_device.argtypes = [c_void_p, POINTER(FramesDesc), c_void_p, c_void_p, c_void_p]
    L.omfs_session_render_host_png.argtypes = [c_void_p, POINTER(FramesDesc), c_void_p, c_size_t, c_void_p, c_void_p]
    L.omfs_png_max_bytes.restype = c_size_t
    L.omfs_png_max_bytes.argtypes = [c_int, c_int]
    L.omfs_png_workspace_bytes.restype = c_size_t
    L.omfs_png_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.omfs_png_encode.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t,
                                  c_void_p]
    L.omfs_session_sync.argtypes = [c_void_p]
    L.omfs_session_submit_host_png.argtypes = [c_void_p, POINTER(FramesDesc), c_void_p, c_size_t, c_void_p]
    L.omfs_session_collect_host_png.argtypes = [c_void_p]
    L.omfs_session_set_deferred_join.argtypes = [c_void_p, c_int]
    L.omfs_session_join.argtypes = [c_void_p, c_void_p]
    L.omfs_session_reserve_pairs.argtypes = [c_void_p, c_uint64]
    L.omfs_session_stats.argtypes = [c_void_p, POINTER(c_uint64)]
    L.omfs_session_tap.argtypes = [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_size_t)]
    L.omfs_session_dims.argtypes = [c_void_p, POINTER(c_int32)]
    L.omfs_session_set_profiling.argtypes = [c_void_p, c_int]
    L.omfs_session_stage_ms.argtypes = [c_void_p, POINTER(c_double), POINTER(c_uint64)]
    L.omfs_host_alloc.argtypes = [POINTER(c_void_p), c_size_t]
    L.omfs_host_free.argtypes = [c_void_p]
    L.omfs_device_alloc.argtypes = [POINTER(c_void_p), c_size_t]
    L.omfs_device_free.argtypes = [c_void_p]
    L.omfs_memcpy_h2d.argtypes = [c_void_p, c_void_p, c_size_t]
    L.omfs_memcpy_d2h.argtypes = [c_void_p, c_void_p, c_size_t]
    L.omfs_device_memset.argtypes = [c_void_p, c_int, c_size_t]
    vp = c_void_p
    L.omfs_flame_pose_prep.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.omfs_flame_blend_gemm.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, c_int, vp]
    L.omfs_flame_lbs.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.omfs_flame_joint_dyn.argtypes = [c_int, c_int, vp, vp, vp, vp]
    L.omfs_flame_fold_subject.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.omfs_face_frames.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp]
    L.omfs_bind_preprocess.argtypes = [c_int, c_int, c_int, c_int, c_int] + [vp] * 13
    L.omfs_binning.argtypes = [c_int, c_int, c_int, c_int, c_size_t] + [vp] * 9 + [c_size_t, vp]
    L.omfs_binning_sort_bits.argtypes = [c_int, c_int, c_int]
    L.omfs_composite.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, POINTER(c_float), vp, vp, vp, vp]
    L.omfs_to_uint8.argtypes = [c_int, c_int, c_int, vp, vp, vp]
    L.omfs_frame_metrics.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp]
    L.omfs_displace_points.argtypes = [c_int, vp, POINTER(c_double), POINTER(c_double), vp, c_int, vp, vp, vp, vp]
    L.omfs_device_check.argtypes = [c_int]
    L.omfs_set_device.argtypes = [c_int]
    L.omfs_ipc_export.argtypes = [c_void_p, c_void_p]
    L.omfs_ipc_open.argtypes = [c_void_p, POINTER(c_void_p)]
    L.omfs_ipc_close.argtypes = [c_void_p]
    L.omfs_push_frames.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().omfs_last_error().decode(errors="replace")
        raise OmfsError(f"omfs error {rc}: {msg}")


def launch_count() -> int:
    return int(load_library().omfs_launch_count())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray | None):
    # the raw address from the array interface: `a.ctypes.data_as` costs ~7 us per array, eight arrays per call — 3 % of a
    # 38-frame host call
    return None if a is None else c_void_p(a.__array_interface__["data"][0])


class PinnedArray:
    """A numpy view over page-locked host memory from the library's own allocator."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = c_void_p()
        check(load_library().omfs_host_alloc(ctypes.byref(self._p), max(nbytes, 1)))
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None and self._p.value:
            load_library().omfs_host_free(self._p)
            self._p = None
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Session:
    """Model + avatar resident on one GPU; renders batches of frames through the C-ABI.

    model : synthetic.FlameModel-like (v_template, shapedirs[400,3V], posedirs, j_regressor, lbs_weights, faces)
    baked : avatar.bake_avatar(...) dict
    """

    def __init__(self, model, baked: dict, width: int, height: int, max_batch: int = 32, device: int = 0,
                 gemm_impl: int = 0, pair_capacity: int = 0, bg=(1.0, 1.0, 1.0), n_expr: int | None = None,
                 debug_keys: bool = False):
        L = load_library()
        self._L = L
        self.width, self.height = int(width), int(height)
        self.n_verts = int(model.v_template.shape[0])
        self.n_faces = int(model.faces.shape[0])
        self.n_gauss = int(baked["xyzb"].shape[0])
        self.n_expr = int(n_expr if n_expr is not None else model.shapedirs.shape[0] - 300)
        self.max_batch = int(max_batch)
        self.device = int(device)
        keep = dict(
            v_template=_f32(model.v_template), shapedirs=_f32(model.shapedirs), posedirs=_f32(model.posedirs),
            j_regressor=_f32(model.j_regressor), lbs_weights=_f32(model.lbs_weights),
            faces=np.ascontiguousarray(model.faces, dtype=np.int32),
            xyzb=_f32(baked["xyzb"]), scale_lo=_f32(baked["scale_lo"]), rot=_f32(baked["rot"]), sh=_f32(baked["sh"]))
        for name, want in (("shapedirs", (300 + self.n_expr, 3 * self.n_verts)), ("posedirs", (36, 3 * self.n_verts)),
                           ("j_regressor", (5, self.n_verts)), ("lbs_weights", (self.n_verts, 5)),
                           ("faces", (self.n_faces, 3)), ("xyzb", (self.n_gauss, 4)), ("scale_lo", (self.n_gauss, 4)),
                           ("rot", (self.n_gauss, 4)), ("sh", (12, self.n_gauss, 4))):
            if keep[name].shape != want:
                raise OmfsError(f"Session: {name} has shape {keep[name].shape}, expected {want} for a model of "
                                f"{self.n_verts} vertices / {self.n_faces} faces and {self.n_gauss} Gaussians")
        md = ModelDesc(self.n_verts, self.n_faces, self.n_expr, self.n_gauss,
                       *[_ptr(keep[k]) for k in ("v_template", "shapedirs", "posedirs", "j_regressor",
                                                 "lbs_weights", "faces", "xyzb", "scale_lo", "rot", "sh")])
        cfg = SessionConfig(self.width, self.height, self.max_batch, self.device, int(gemm_impl), int(bool(debug_keys)),
                            int(pair_capacity), (c_float * 3)(*[float(x) for x in bg]))
        self._submitted: list = []   # clips between submit_host_png and collect_host_png (buffers kept alive)
        self._h = c_void_p()
        check(L.omfs_session_create(ctypes.byref(md), ctypes.byref(cfg), ctypes.byref(self._h)))
        del keep  # the library copied everything to the device

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.omfs_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- subject
    def set_subject(self, shape300, static_offset=None, plan_offset=None):
        sh = _f32(shape300).reshape(-1)
        if sh.size != 300:
            raise OmfsError(f"set_subject: shape has {sh.size} coefficients, the FLAME record carries 300")
        so = None if static_offset is None else _f32(static_offset).reshape(-1)
        po = None if plan_offset is None else _f32(plan_offset).reshape(-1)
        for name, a in (("static_offset", so), ("plan_offset", po)):
            if a is not None and a.size != 3 * self.n_verts:
                raise OmfsError(
                    f"set_subject: {name} covers {a.size // 3} vertices but the FLAME model has {self.n_verts}. "
                    "The reference's records are written for the 5143-vertex FLAME-with-teeth mesh "
                    "(flame_fitter.py:439, preprocess_video.py:329); a raw 5023-vertex flame2023.pkl does not match "
                    "them — export the teeth-augmented model with tools/export_flame_with_teeth.py")
        check(self._L.omfs_session_set_subject(self._h, _ptr(sh), _ptr(so), _ptr(po)))

    # -- rendering
    def _frames_desc(self, params, cams, keep: list):
        T = int(params.expr.shape[0])
        expr = _f32(params.expr)
        if expr.ndim != 2 or expr.shape[1] != self.n_expr:
            raise OmfsError(f"frames: expr has shape {expr.shape}, the session was created for {self.n_expr} "
                            "expression coefficients")
        cam_arr = _f32(np.stack([c.pack() if isinstance(c, cam_mod.Camera) else np.asarray(c) for c in cams]))
        dyn = params.dynamic_offset
        if dyn is not None and not np.any(dyn):
            dyn = None
        arrs = [expr, _f32(params.rotation), _f32(params.neck_pose), _f32(params.jaw_pose), _f32(params.eyes_pose),
                _f32(params.translation), None if dyn is None else _f32(dyn), cam_arr]
        keep.extend(arrs)
        return FramesDesc(T, len(cams), *[_ptr(a) for a in arrs]), T * len(cams)

    def render_host(self, params, cams, want_u8: bool = True, want_f32: bool = False, out_u8=None, out_f32=None):
        """Host parameters in, host frames out (uint8 [S,H,W,3] and/or float32 [S,3,H,W])."""
        keep: list = []
        fd, S = self._frames_desc(params, cams, keep)
        if want_u8 and out_u8 is None:
            out_u8 = np.empty((S, self.height, self.width, 3), np.uint8)
        if want_f32 and out_f32 is None:
            out_f32 = np.empty((S, 3, self.height, self.width), np.float32)
        check(self._L.omfs_session_render_host(self._h, ctypes.byref(fd), _ptr(out_u8) if want_u8 else None,
                                               _ptr(out_f32) if want_f32 else None))
        return (out_u8 if want_u8 else None), (out_f32 if want_f32 else None)

    def render_host_png(self, params, cams, out_png=None, out_offsets=None, out_u8=None, want_u8: bool = False):
        """Host parameters in, PNG files out: the frames are encoded on the device and only the compressed streams
        cross PCIe.  Returns (png, offsets[, u8]): frame i is png[offsets[i]:offsets[i+1]] (uint8 / uint64 arrays;
        pass pinned `out_png` / `out_offsets` to reuse buffers across calls)."""
        keep: list = []
        fd, S = self._frames_desc(params, cams, keep)
        if out_png is None:
            out_png = np.empty(S * int(self._L.omfs_png_max_bytes(self.width, self.height)), np.uint8)
        if out_offsets is None:
            out_offsets = np.zeros(S + 1, np.uint64)
        if out_offsets.size < S + 1:
            raise OmfsError(f"render_host_png: offsets array has {out_offsets.size} entries, {S + 1} needed")
        if want_u8 and out_u8 is None:
            out_u8 = np.empty((S, self.height, self.width, 3), np.uint8)
        check(self._L.omfs_session_render_host_png(self._h, ctypes.byref(fd), _ptr(out_png), out_png.nbytes,
                                                   _ptr(out_offsets), _ptr(out_u8) if want_u8 else None))
        return (out_png, out_offsets[:S + 1], out_u8) if want_u8 else (out_png, out_offsets[:S + 1])

    def submit_host_png(self, params, cams, out_png: np.ndarray, out_offsets: np.ndarray):
        """Streaming form of render_host_png: enqueue the clip and return; collect_host_png() completes the oldest
        submitted clip (at most three outstanding).  `out_png` / `out_offsets` (uint8 / uint64, S + 1 offsets) and the
        parameter arrays must stay alive and untouched until that collect; they are held here until then."""
        keep: list = [out_png, out_offsets]
        fd, S = self._frames_desc(params, cams, keep)
        if out_offsets.size < S + 1:
            raise OmfsError(f"submit_host_png: offsets array has {out_offsets.size} entries, {S + 1} needed")
        check(self._L.omfs_session_submit_host_png(self._h, ctypes.byref(fd), _ptr(out_png), out_png.nbytes,
                                                   _ptr(out_offsets)))
        self._submitted.append((keep, out_png, out_offsets, S))

    def collect_host_png(self):
        """Wait for the oldest submitted clip; returns its (png, offsets[:S + 1])."""
        if not self._submitted:
            raise OmfsError("collect_host_png: no clip outstanding")
        _, out_png, out_offsets, S = self._submitted.pop(0)
        try:
            check(self._L.omfs_session_collect_host_png(self._h))
        except OmfsError:
            self._submitted.clear()   # the library dropped every outstanding clip
            raise
        return out_png, out_offsets[:S + 1]

    def render_device(self, d_params: dict, n_frames: int, n_views: int, d_out_u8=0, d_out_f32=0, stream=0):
        """Device pointers in (ints), device pointers out.  Asynchronous; call sync()."""
        fd = FramesDesc(int(n_frames), int(n_views), d_params["expr"], d_params["rotation"],
                        d_params["neck_pose"], d_params["jaw_pose"], d_params["eyes_pose"],
                        d_params["translation"], d_params.get("dynamic_offset") or None, d_params["cams"])
        check(self._L.omfs_session_render_device(self._h, ctypes.byref(fd), d_out_u8 or None, d_out_f32 or None,
                                                 stream or None))

    def reserve_pairs(self, capacity: int):
        """Grow the per-batch tile-pair capacity (render_device callers, after an OMFS_ERR_CAPACITY)."""
        check(self._L.omfs_session_reserve_pairs(self._h, int(capacity)))

    def sync(self):
        check(self._L.omfs_session_sync(self._h))

    def set_deferred_join(self, on: bool = True):
        """render_device calls stop joining their last compositing launch into the caller's stream, so consecutive
        calls overlap; order consumers with join(stream)."""
        check(self._L.omfs_session_set_deferred_join(self._h, 1 if on else 0))

    def join(self, stream=0):
        """`stream` waits for every compositing launch enqueued so far."""
        check(self._L.omfs_session_join(self._h, stream or None))

    @property
    def stream(self) -> int:
        return int(self._L.omfs_session_stream(self._h) or 0)

    def stats(self) -> dict:
        out = (c_uint64 * 4)()
        check(self._L.omfs_session_stats(self._h, out))
        return {"pairs": int(out[0]), "launches": int(out[1]), "batches": int(out[2]), "overflow": int(out[3])}

    def dims(self) -> dict:
        out = (c_int32 * 9)()
        check(self._L.omfs_session_dims(self._h, out))
        keys = ("V", "F", "n_expr", "N", "kpad", "npad", "tiles", "last_batch_segments", "pairs_last_batch")
        return dict(zip(keys, [int(x) for x in out]))

    STAGES = ("flame", "face_frames", "bind_preprocess", "depth_sort", "tile_ranges", "emit_scatter", "composite")

    def set_profiling(self, on: bool):
        check(self._L.omfs_session_set_profiling(self._h, 1 if on else 0))

    def stage_ms(self) -> dict:
        ms = (c_double * 8)()
        calls = (c_uint64 * 8)()
        check(self._L.omfs_session_stage_ms(self._h, ms, calls))
        return {n: {"ms": float(ms[i]), "calls": int(calls[i])} for i, n in enumerate(self.STAGES)}

    def tap(self, name: str):
        p, n = c_void_p(), c_size_t()
        check(self._L.omfs_session_tap(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n)))
        return int(p.value or 0), int(n.value)

    def tap_array(self, name: str, shape, dtype) -> np.ndarray:
        """Copy a debug tap to the host (tests only)."""
        ptr, nbytes = self.tap(name)
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= nbytes, (name, out.nbytes, nbytes)
        memcpy_d2h(out, ptr)
        return out


# ---------------------------------------------------------------------------------------------
# Layout of the per-Gaussian records and of the pair list as the kernels keep them (include/omfs_b200.h), against
# the published fields the oracle restates:
#   P0 = (px, py, cull extents [two halves, a hint], radius bits)   P2 = (r, g, b, depth)
#   pair word = Gaussian index (low 28 bits) | block hints (top 4 bits)
VAL_INDEX_BITS = 28
VAL_INDEX_MASK = (1 << VAL_INDEX_BITS) - 1


def published_records(P0: np.ndarray, P2: np.ndarray):
    """(P0, P2) in the published field order — P0 = (px, py, depth, radius), P2 = (r, g, b, 0) — from the device
    layout, plus the cull extents [..., 2] in pixels (the hint that sits in the device P0.z)."""
    p0, p2 = P0.copy(), P2.copy()
    p0[..., 2] = P2[..., 3]
    p2[..., 3] = 0.0
    ext = np.ascontiguousarray(P0[..., 2]).view(np.float16).reshape(P0.shape[:-1] + (2,)).astype(np.float32)
    return p0, p2, ext


def pair_indices(vals: np.ndarray) -> np.ndarray:
    return vals & np.uint32(VAL_INDEX_MASK)


def pair_hints(vals: np.ndarray) -> np.ndarray:
    return vals >> np.uint32(VAL_INDEX_BITS)


# ---------------------------------------------------------------------------------------------
# device memory helpers through the library's own CUDA runtime (so numpy-only callers work)
def memcpy_d2h(dst: np.ndarray, src_ptr: int):
    check(load_library().omfs_memcpy_d2h(dst.ctypes.data_as(c_void_p), c_void_p(src_ptr), dst.nbytes))


def memcpy_h2d(dst_ptr: int, src: np.ndarray):
    src = np.ascontiguousarray(src)
    check(load_library().omfs_memcpy_h2d(c_void_p(dst_ptr), src.ctypes.data_as(c_void_p), src.nbytes))


class DeviceArray:
    """A device allocation with a shape/dtype, owned through the C-ABI helpers."""

    def __init__(self, shape, dtype, fill_from: np.ndarray | None = None):
        self.shape = tuple(int(x) for x in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = c_void_p()
        check(load_library().omfs_device_alloc(ctypes.byref(self._p), max(self.nbytes, 16)))
        if fill_from is not None:
            a = np.ascontiguousarray(fill_from, dtype=self.dtype)
            assert a.nbytes == self.nbytes, (a.shape, self.shape)
            memcpy_h2d(self.ptr, a)

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(a)
        return cls(a.shape, a.dtype, fill_from=a)

    @property
    def ptr(self) -> int:
        return int(self._p.value or 0)

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            memcpy_d2h(out, self.ptr)
        return out

    def zero(self):
        memcpy_h2d(self.ptr, np.zeros(self.shape, self.dtype))

    def free(self):
        if self._p is not None and self._p.value:
            load_library().omfs_device_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def png_encode_device(frames_u8: np.ndarray) -> list[bytes]:
    """uint8 [S,H,W,3] host frames -> S PNG files, encoded by the device sink (upload, omfs_png_encode, download).
    The level-1 form of the frame sink, for callers that already hold frames."""
    L = load_library()
    frames_u8 = np.ascontiguousarray(frames_u8, dtype=np.uint8)
    if frames_u8.ndim != 4 or frames_u8.shape[3] != 3:
        raise OmfsError(f"png_encode_device: expected uint8 [S,H,W,3], got {frames_u8.shape}")
    S, H, W, _ = frames_u8.shape
    if S == 0:
        return []
    cap = int(L.omfs_png_max_bytes(W, H))
    ws_bytes = int(L.omfs_png_workspace_bytes(S, W, H))
    if not cap or not ws_bytes:
        raise OmfsError(f"png_encode_device: image size {W}x{H} is not supported by the device sink")
    d_in, d_png = DeviceArray.from_numpy(frames_u8), DeviceArray((S * cap,), np.uint8)
    d_off, d_ws = DeviceArray((S + 1,), np.uint64), DeviceArray((ws_bytes,), np.uint8)
    try:
        check(L.omfs_png_encode(S, W, H, c_void_p(d_in.ptr), c_void_p(d_png.ptr), S * cap, c_void_p(d_off.ptr),
                                c_void_p(d_ws.ptr), ws_bytes, None))
        off = d_off.numpy()
        png = np.empty(int(off[S]), np.uint8)
        check(L.omfs_memcpy_d2h(png.ctypes.data_as(c_void_p), c_void_p(d_png.ptr), png.nbytes))
    finally:
        for a in (d_in, d_png, d_off, d_ws):
            a.free()
    return [png[int(off[i]):int(off[i + 1])].tobytes() for i in range(S)]


IPC_HANDLE_BYTES = 64


def ipc_export(d_ptr: int) -> bytes:
    """CUDA IPC handle of a buffer from omfs_device_alloc (DeviceArray), to be opened by another process."""
    buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
    check(load_library().omfs_ipc_export(c_void_p(d_ptr), buf))
    return buf.raw


def ipc_open(handle: bytes) -> int:
    out = c_void_p()
    buf = ctypes.create_string_buffer(handle, IPC_HANDLE_BYTES)
    check(load_library().omfs_ipc_open(buf, ctypes.byref(out)))
    return int(out.value)


def ipc_close(d_ptr: int) -> None:
    check(load_library().omfs_ipc_close(c_void_p(d_ptr)))


def push_frames(d_dst: int, d_src: int, nbytes: int, stream: int = 0) -> None:
    """Copy-engine device-to-device copy (local or to opened peer memory) on `stream`."""
    check(load_library().omfs_push_frames(c_void_p(d_dst), c_void_p(d_src), nbytes, c_void_p(stream)))
