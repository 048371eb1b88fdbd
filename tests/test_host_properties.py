"""Property tests (hypothesis) of the reference-facing host functions: the invariants the reference's own unit
tests spot-check (test/test_render_surgery.py), over random inputs."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import omfs_b200  # noqa: F401
from omfs_b200 import render_surgery as rs, sharding, validation_reporting as vrep
from oracle import reference_rows as rr

mm = st.floats(min_value=-15.0, max_value=15.0, allow_nan=False)
sens = st.floats(min_value=0.1, max_value=3.0, allow_nan=False)


@given(mm, sens)
def test_compute_offset_is_the_reference_product(m, s):
    assert rs.compute_offset(m, s) == m * s * 0.001 == rr.compute_offset(m, s)
    assert rs.compute_offset(0.0, s) == 0.0 and rs.compute_offset(m, 0.0) == 0.0


@settings(max_examples=40, deadline=None)
@given(mm, mm, st.integers(0, 2), st.integers(0, 2), st.floats(0.25, 4.0), st.floats(0.25, 4.0), st.integers(1, 6),
       st.booleans())
def test_parameter_edit_touches_only_the_two_axes(lefort, bsso, ax_t, ax_j, sc_l, sc_b, T, batched):
    """R2: translation[..., axis_t] += lefort*scale_l and jaw_pose[..., axis_j] += bsso*scale_b, fp32 store, every
    other key and every other column untouched, the source record not mutated; mirror == oracle restatement."""
    rng = np.random.default_rng(T * 7 + ax_t)
    shape3 = (T, 3) if batched else (3,)
    rec = {"translation": rng.normal(0, 0.01, shape3).astype(np.float32),
           "jaw_pose": rng.normal(0, 0.1, shape3).astype(np.float32),
           "expr": rng.normal(0, 0.5, (T, 100)).astype(np.float32), "shape": rng.normal(0, 1, 300).astype(np.float32)}
    before = {k: v.copy() for k, v in rec.items()}
    dmap = {"translation_axis": ax_t, "jaw_axis": ax_j, "lefort_scale": sc_l, "bsso_scale": sc_b}
    lo, bo = rs.compute_offset(lefort, 1.0), rs.compute_offset(bsso, 1.0)
    out = rs._edit_record(rec, lo, bo, dmap)
    want = rr.modify_flame_params(rec, lo, bo, dmap)
    for k in rec:
        assert np.array_equal(rec[k], before[k])                      # source untouched
        assert out[k].dtype == before[k].dtype and np.array_equal(out[k], np.asarray(want[k]))
    assert np.array_equal(out["expr"], before["expr"]) and np.array_equal(out["shape"], before["shape"])
    for key, ax, delta in (("translation", ax_t, lo * sc_l), ("jaw_pose", ax_j, bo * sc_b)):
        other = [c for c in range(3) if c != ax]
        assert np.array_equal(out[key][..., other], before[key][..., other])
        assert np.array_equal(out[key][..., ax], (before[key][..., ax] + delta).astype(np.float32))


@settings(max_examples=30, deadline=None)
@given(n_frames=st.integers(1, 40), max_frames=st.integers(1, 30))
def test_deterministic_export_selection(tmp_path_factory, n_frames, max_frames):
    """Selected indices are sorted, unique, inside the clip, include both ends when more than one frame is asked
    for, and the manifest names the source files (render_surgery.py:365-409)."""
    d = tmp_path_factory.mktemp("det")
    frames = d / "frames"
    frames.mkdir()
    for i in range(n_frames):
        (frames / f"{i:05d}.png").write_bytes(b"x")
    out = rs.export_deterministic_frames(str(frames), str(d / "out"), None, max_frames)
    man = json.load(open(os.path.join(out, "deterministic_indices_manifest.json")))
    sel = man["selected_indices"]
    assert sel == sorted(set(sel)) and 0 <= sel[0] and sel[-1] < n_frames
    assert len(sel) <= min(max_frames, n_frames)
    if min(max_frames, n_frames) > 1:
        assert sel[0] == 0 and sel[-1] == n_frames - 1
    assert [e["source"] for e in man["exports"]] == [f"{i:05d}.png" for i in sel]
    assert sorted(os.listdir(out)) == sorted([f"idx_{i:05d}.png" for i in sel] + ["deterministic_indices_manifest.json"])


@given(st.integers(0, 2000), st.integers(1, 8))
def test_frame_blocks_partition_the_clip(n, world):
    """§8e: contiguous blocks of ceil(n/world) frames (the last ones shorter or empty), every frame exactly once."""
    blocks = [sharding.frame_block(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    sizes = [hi - lo for lo, hi in blocks]
    assert min(sizes) >= 0 and max(sizes) == (-(-n // world) if n else 0) and sizes == sorted(sizes, reverse=True)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 12), st.integers(2, 12), st.integers(0, 10_000))
def test_report_metrics_from_moments(h, w, seed):
    """The device path's closed forms (metrics_from_moments on the oracle's frame moments) reproduce the
    reference metrics (psnr, ssim_global) on random uint8 frame pairs, identical pairs included."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    b = a.copy()
    b[1] = np.clip(a[1].astype(np.int64) + rng.integers(-20, 21, a[1].shape), 0, 255).astype(np.uint8)
    p, s = vrep.metrics_from_moments(rr.frame_moments(a, b), h * w)
    assert p[0] == 99.0 and abs(s[0] - 1.0) <= 1e-12
    af, bf = a[1].astype(np.float32), b[1].astype(np.float32)
    assert abs(p[1] - rr.psnr(af, bf)) <= 1e-4
    assert abs(s[1] - rr.ssim_global(af, bf)) <= 1e-9


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.integers(1, 40), st.integers(0, 2**31 - 1), st.sampled_from(["noise", "flat", "ramp"]))
def test_png_encoder_round_trips_through_independent_decoders(h, w, seed, kind):
    """render_surgery.encode_png (the frame sink's own encoder: Up filter + run-length deflate) writes standard PNGs:
    PIL and OpenCV decode every size and content back to the same pixels, and the header says 8-bit RGB."""
    import io
    import struct
    from PIL import Image
    from omfs_b200 import render_surgery as rs
    rng = np.random.default_rng(seed)
    if kind == "noise":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "flat":
        img = np.full((h, w, 3), rng.integers(0, 256), dtype=np.uint8)
    else:
        img = ((np.arange(h)[:, None, None] * 7 + np.arange(w)[None, :, None] * 3 + np.arange(3)) % 256).astype(np.uint8)
    data = rs.encode_png(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[12:16] == b"IHDR"
    assert struct.unpack(">IIBBBBB", data[16:29]) == (w, h, 8, 2, 0, 0, 0)
    pil = Image.open(io.BytesIO(data))
    pil.verify()                                                   # chunk CRCs
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))), img)
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)[..., ::-1], img)


def test_png_encoder_rejects_other_layouts():
    from omfs_b200 import render_surgery as rs
    for bad in (np.zeros((4, 4), np.uint8), np.zeros((4, 4, 4), np.uint8), np.zeros((0, 4, 3), np.uint8)):
        with pytest.raises(ValueError):
            rs.encode_png(bad)
