"""The device frame sink (csrc/png.cu) on a B200: every PNG decodes to the uint8 frame bit for bit through PIL and
OpenCV, the kernel's byte stream equals the sequential host emulation of the same source (tests/emu/png_emu.cpp,
itself pinned against zlib / PIL in test_png_sink_host.py), and the session path returns the same frames as the
raw path."""
import ctypes
import io
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = ctypes.CDLL(os.path.join(HERE, "emu", "libpngemu.so"))
    L.emu_png_encode.restype = ctypes.c_longlong
    L.emu_png_encode.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                 ctypes.c_void_p]
    return L


def _emu_encode(L, img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    out = np.zeros(h * (3 * w + 1) + 64 * h + 256, np.uint8)
    n = L.emu_png_encode(w, h, img.ctypes.data, out.ctypes.data, out.size, None)
    assert n > 0, n
    return out[:n].tobytes()


def _decode_both(png):
    from PIL import Image
    import cv2
    a = np.asarray(Image.open(io.BytesIO(png)).convert("RGB"))
    b = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
    assert np.array_equal(a, b)
    return a


def _smooth(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.full((h, w, 3), 255.0)
    for _ in range(30):
        cx, cy, r = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h), rng.uniform(2, 0.2 * w)
        a = np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * r * r))[..., None]
        img = img * (1 - 0.8 * a) + 0.8 * a * rng.uniform(0, 255, 3)
    return np.clip(img + 0.5, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("shape,kind", [((512, 512), "smooth"), ((37, 53), "smooth"), ((96, 160), "noise"),
                                        ((64, 48), "flat"), ((1, 700), "low"), ((40, 1024), "smooth"),
                                        ((1, 1), "noise"), ((130, 2048), "smooth")])
def test_level1_encoder_matches_emulation_and_decodes(shape, kind):
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    frames = []
    for i in range(3):
        if kind == "smooth":
            frames.append(_smooth(h, w, i))
        elif kind == "noise":
            frames.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        elif kind == "flat":
            frames.append(np.full((h, w, 3), 17 * i, np.uint8))
        else:
            frames.append(rng.integers(0, 4, (h, w, 3), dtype=np.uint8))
    frames = np.stack(frames)
    pngs = runtime.png_encode_device(frames)
    L = _emu()
    for f, p in zip(frames, pngs):
        assert np.array_equal(_decode_both(p), f)
        assert p == _emu_encode(L, f)          # the parallel plumbing reproduces the sequential stream exactly


def test_session_png_equals_raw_frames_across_batches():
    """More frames than three batches: exercises the sink's ring, the tapered last batch and the offsets table."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, runtime, synthetic
    T, W, H = 23, 160, 112
    model, params, av, cam = synthetic.make_scene(n_gauss=6000, n_frames=T, width=W, height=H, n_verts=1202)
    with runtime.Session(model, avatar.bake(av), W, H, max_batch=4, device=0) as sess:
        sess.set_subject(params.shape, params.static_offset)
        raw, _ = sess.render_host(params, [cam], want_u8=True)
        png, off = sess.render_host_png(params, [cam])
        assert off[0] == 0 and len(off) == T + 1 and (np.diff(off.astype(np.int64)) > 0).all()
        for t in range(T):
            assert np.array_equal(_decode_both(png[int(off[t]):int(off[t + 1])].tobytes()), raw[t]), t
        # both outputs at once, two views per frame, into caller-owned (pinned) buffers
        cap = int(runtime.load_library().omfs_png_max_bytes(W, H))
        buf = runtime.PinnedArray((2 * T * cap,), np.uint8)
        offs = runtime.PinnedArray((2 * T + 1,), np.uint64)
        cam2 = synthetic.make_scene(n_gauss=10, n_frames=1, width=W, height=H, n_verts=1202, seed=3)[3]
        png2, off2, raw2 = sess.render_host_png(params, [cam, cam2], out_png=buf.array, out_offsets=offs.array,
                                                want_u8=True)
        assert raw2.shape == (2 * T, H, W, 3) and np.array_equal(raw2[0::2], raw)
        for s in (0, 1, 2 * T - 1):
            assert np.array_equal(_decode_both(png2[int(off2[s]):int(off2[s + 1])].tobytes()), raw2[s])
        # a buffer that cannot hold the streams is reported, not overrun
        with pytest.raises(runtime.OmfsError, match="too small"):
            sess.render_host_png(params, [cam], out_png=np.empty(1000, np.uint8))
        png3, off3 = sess.render_host_png(params.slice(0, 3), [cam])      # the session is usable afterwards
        assert np.array_equal(_decode_both(png3[int(off3[2]):int(off3[3])].tobytes()), raw[2])


def test_streamed_clips_equal_blocking_calls():
    """omfs_session_submit_host_png / collect: clips of different lengths (one, several and an odd number of launch
    groups) pushed two deep give, byte for byte, the PNG streams and offsets of the blocking call; a buffer that is
    too small is reported by the collect and leaves the session usable."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, runtime as rt, synthetic
    T, W, H = 11, 128, 96
    model, params, av, cam = synthetic.make_scene(n_gauss=5000, n_frames=T, width=W, height=H, n_verts=642)
    baked = avatar.bake(av)
    clips = [params.slice(0, 11), params.slice(3, 5), params.slice(1, 10), params.slice(0, 3), params.slice(2, 11)]
    with rt.Session(model, baked, W, H, max_batch=3) as sess:
        sess.set_subject(params.shape, params.static_offset)
        cap = int(rt.load_library().omfs_png_max_bytes(W, H))
        want = []
        for c in clips:
            png, off = sess.render_host_png(c, [cam])
            want.append((png[:int(off[-1])].copy(), off.copy()))
        bufs = [(rt.PinnedArray((T * cap,), np.uint8), rt.PinnedArray((T + 1,), np.uint64)) for _ in range(2)]
        got = []
        for i, c in enumerate(clips):
            if i >= 2:
                png, off = sess.collect_host_png()
                got.append((png[:int(off[-1])].copy(), off.copy()))
            sess.submit_host_png(c, [cam], bufs[i % 2][0].array, bufs[i % 2][1].array)
        with pytest.raises(rt.OmfsError, match="outstanding"):
            sess.render_host_png(clips[0], [cam])             # blocking calls wait their turn
        while len(got) < len(clips):
            png, off = sess.collect_host_png()
            got.append((png[:int(off[-1])].copy(), off.copy()))
        for (gp, go), (wp, wo) in zip(got, want):
            assert np.array_equal(go, wo) and np.array_equal(gp, wp)
        # too small a buffer: the collect says so, the session goes on
        small = rt.PinnedArray((1000,), np.uint8)
        sess.submit_host_png(clips[1], [cam], small.array, bufs[0][1].array)
        with pytest.raises(rt.OmfsError, match="too small"):
            sess.collect_host_png()
        png, off = sess.render_host_png(clips[1], [cam])
        assert np.array_equal(png[:int(off[-1])], want[1][0])
        sess.submit_host_png(clips[3], [cam], bufs[1][0].array, bufs[1][1].array)
        png, off = sess.collect_host_png()
        assert np.array_equal(png[:int(off[-1])], want[3][0]) and np.array_equal(off, want[3][1])
