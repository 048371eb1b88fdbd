"""Tensor-level entry: `torch.ops.omfs.*` over the C-ABI (north_star: "a drop-in behind a thin PyTorch C-ABI
extension").

Upstream's `render.py` — what the reference spawns at 02_Visual_Engine/render_surgery.py:289-315 — drives its
rasterizer with torch tensors on the GPU.  A caller of that shape (FLAME parameters as CUDA tensors, frames wanted as
a CUDA tensor) uses these operators; they pass `data_ptr()` and the CURRENT torch stream to
`omfs_session_render_device`, so the kernels are ordered with the caller's other torch work and nothing is copied.
PyTorch is plumbing here (memory, streams, dispatch); the arithmetic is libomfs_b200.so.  There is no CPU
implementation: CPU tensors are rejected.

    h = torch_ops.open_session(model, baked, width, height, max_batch=60)      # python object -> int handle
    torch_ops.set_subject(h, shape300, static_offset)
    frames = torch.ops.omfs.render(h, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, None)
    # uint8 [T * n_views, H, W, 3] on the same device, valid in stream order
    image = torch.ops.omfs.render_image(h, ...)                                # float32 [S, 3, H, W] (rasterizer layout)
    torch_ops.check(h)                                                         # waits, raises on a capacity overflow
"""
from __future__ import annotations

import torch

from . import runtime

_SESSIONS: dict[int, runtime.Session] = {}
_NEXT = [1]

_SIG = ("(int session, Tensor expr, Tensor rotation, Tensor neck_pose, Tensor jaw_pose, Tensor eyes_pose, "
        "Tensor translation, Tensor cams, Tensor? dynamic_offset) -> Tensor")
torch.library.define("omfs::render", _SIG)
torch.library.define("omfs::render_image", _SIG)


def open_session(model, baked: dict, width: int, height: int, max_batch: int = 32, device: int | None = None,
                 **kw) -> int:
    """Model + avatar to the GPU (runtime.Session); returns the integer handle the operators take."""
    if device is None:
        device = torch.cuda.current_device()
    h = _NEXT[0]
    _NEXT[0] += 1
    _SESSIONS[h] = runtime.Session(model, baked, width, height, max_batch=max_batch, device=device, **kw)
    return h


def session(handle: int) -> runtime.Session:
    try:
        return _SESSIONS[int(handle)]
    except KeyError:
        raise runtime.OmfsError(f"torch.ops.omfs: no open session with handle {handle}") from None


def set_subject(handle: int, shape300, static_offset=None, plan_offset=None) -> None:
    as_np = lambda t: None if t is None else (t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t)
    session(handle).set_subject(as_np(shape300), as_np(static_offset), as_np(plan_offset))


def check(handle: int) -> None:
    """Wait for the session's work and raise OmfsError if a batch overflowed the tile-pair capacity (then
    `session(handle).reserve_pairs(n)` and render again)."""
    session(handle).sync()


def close_session(handle: int) -> None:
    s = _SESSIONS.pop(int(handle), None)
    if s is not None:
        s.close()


def _args(sess, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset):
    named = dict(expr=expr, rotation=rotation, neck_pose=neck_pose, jaw_pose=jaw_pose, eyes_pose=eyes_pose,
                 translation=translation, cams=cams)
    if dynamic_offset is not None:
        named["dynamic_offset"] = dynamic_offset
    dev = torch.device("cuda", sess.device)
    for name, t in named.items():
        if not t.is_cuda or t.device != dev:
            raise runtime.OmfsError(f"torch.ops.omfs: {name} must live on {dev} (got {t.device}); there is no CPU path")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise runtime.OmfsError(f"torch.ops.omfs: {name} must be a contiguous float32 tensor")
    T = int(expr.shape[0])
    want = dict(expr=(T, sess.n_expr), rotation=(T, 3), neck_pose=(T, 3), jaw_pose=(T, 3), eyes_pose=(T, 6),
                translation=(T, 3))
    for name, shape in want.items():
        if tuple(named[name].shape) != shape:
            raise runtime.OmfsError(f"torch.ops.omfs: {name} has shape {tuple(named[name].shape)}, expected {shape}")
    if cams.dim() != 2 or cams.shape[1] != 40:
        raise runtime.OmfsError("torch.ops.omfs: cams must be [n_views, 40] (cameras.Camera.pack)")
    if dynamic_offset is not None and tuple(dynamic_offset.shape) != (T, sess.n_verts, 3):
        raise runtime.OmfsError(f"torch.ops.omfs: dynamic_offset must be [{T}, {sess.n_verts}, 3]")
    return T, int(cams.shape[0]), {k: v.data_ptr() for k, v in named.items()}, dev


@torch.library.impl("omfs::render", "CUDA")
def _render(session_, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset):
    sess = session(session_)
    T, nv, ptrs, dev = _args(sess, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset)
    out = torch.empty((T * nv, sess.height, sess.width, 3), dtype=torch.uint8, device=dev)
    if T:
        sess.render_device(ptrs, T, nv, d_out_u8=out.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
    return out


@torch.library.impl("omfs::render_image", "CUDA")
def _render_image(session_, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset):
    sess = session(session_)
    T, nv, ptrs, dev = _args(sess, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset)
    out = torch.empty((T * nv, 3, sess.height, sess.width), dtype=torch.float32, device=dev)
    if T:
        sess.render_device(ptrs, T, nv, d_out_f32=out.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
    return out


def _no_cpu(*args, **kw):
    raise runtime.OmfsError("torch.ops.omfs: tensors must be CUDA tensors on a B200; there is no CPU implementation")


torch.library.impl("omfs::render", "CPU")(_no_cpu)
torch.library.impl("omfs::render_image", "CPU")(_no_cpu)


@torch.library.register_fake("omfs::render")
def _render_fake(session_, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset):
    sess = session(session_)
    return expr.new_empty((expr.shape[0] * cams.shape[0], sess.height, sess.width, 3), dtype=torch.uint8)


@torch.library.register_fake("omfs::render_image")
def _render_image_fake(session_, expr, rotation, neck_pose, jaw_pose, eyes_pose, translation, cams, dynamic_offset):
    sess = session(session_)
    return expr.new_empty((expr.shape[0] * cams.shape[0], 3, sess.height, sess.width), dtype=torch.float32)
