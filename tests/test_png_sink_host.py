"""The device frame sink's stream format, code tables and checksum algebra, without a GPU: tests/emu/png_emu.cpp
runs the product's own png_core.cuh (the source png.cu compiles) as a sequential encoder; its output must be a
PNG any decoder accepts and must decode to the input bit for bit."""
import ctypes
import io
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = ctypes.CDLL(os.path.join(HERE, "emu", "libpngemu.so"))
    L.emu_png_encode.restype = ctypes.c_longlong
    L.emu_png_encode.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                 ctypes.c_void_p]
    L.emu_png_geometry.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.emu_png_table_summary.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def encode(L, img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    geo = (ctypes.c_int * 6)()
    assert L.emu_png_geometry(w, h, geo) == 0
    out = np.zeros(h * (3 * w + 1) + 64 * geo[5] + 256, np.uint8)
    choices = (ctypes.c_int * geo[5])()
    n = L.emu_png_encode(w, h, img.ctypes.data, out.ctypes.data, out.size, choices)
    assert n > 0, n
    return out[:n].tobytes(), list(choices), list(geo)


def chunks(png):
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    at, out = 8, []
    while at < len(png):
        n, tag = struct.unpack(">I4s", png[at:at + 8])
        data = png[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", png[at + 8 + n:at + 12 + n])
        assert crc == zlib.crc32(tag + data), (tag, at)     # the piecewise CRC algebra against zlib
        out.append((tag, data))
        at += 12 + n
    assert at == len(png)
    return out


def smooth_image(h, w, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.full((h, w, 3), 255.0)
    for _ in range(40):
        cx, cy, r = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h), rng.uniform(3, 0.2 * w)
        a = np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * r * r))[..., None]
        img = img * (1 - 0.8 * a) + 0.8 * a * rng.uniform(0, 255, 3)
    return np.clip(img + 0.5, 0, 255).astype(np.uint8)


CASES = {
    "smooth_512": lambda: smooth_image(512, 512),
    "smooth_odd": lambda: smooth_image(37, 53, 1),          # row bytes not a multiple of 16, ragged last strip
    "flat_white": lambda: np.full((64, 48, 3), 255, np.uint8),
    "noise": lambda: np.random.default_rng(2).integers(0, 256, (96, 160, 3), dtype=np.uint8),
    "one_pixel": lambda: np.array([[[1, 2, 3]]], np.uint8),
    "one_row": lambda: np.random.default_rng(3).integers(0, 4, (1, 700, 3), dtype=np.uint8),
    "wide_1024": lambda: smooth_image(40, 1024, 4),
    "short_runs": lambda: np.repeat(np.random.default_rng(5).integers(0, 256, (33, 20, 3), dtype=np.uint8), 3, axis=1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_encoder_round_trips(emu, name):
    img = CASES[name]()
    png, choices, geo = encode(emu, img)
    cs = chunks(png)
    assert [c[0] for c in cs] == [b"IHDR"] + [b"IDAT"] * (geo[5] + 1) + [b"IEND"]
    h, w, _ = img.shape
    assert cs[0][1] == struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)
    raw = zlib.decompress(b"".join(d for t, d in cs if t == b"IDAT"))       # also checks Adler-32
    assert len(raw) == h * (3 * w + 1)
    lines = np.frombuffer(raw, np.uint8).reshape(h, 3 * w + 1)
    assert (lines[:, 0] == 2).all()
    back = np.cumsum(lines[:, 1:].astype(np.uint64), axis=0).astype(np.uint8).reshape(h, w, 3)   # undo Up
    assert np.array_equal(back, img)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png)).convert("RGB")), img)
    try:
        import cv2
        dec = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(dec[..., ::-1], img)
    except ImportError:
        pass
    if name == "noise":
        assert set(choices) == {8}, choices            # incompressible: every strip stored, no expansion beyond framing
        assert len(png) <= img.size + h + 19 * geo[5] + 66
    if name == "flat_white":
        assert len(png) < 600
    if name == "smooth_512":
        assert len(png) < 0.5 * img.size, len(png)


def test_every_strip_is_independent(emu):
    """Strips are separate deflate blocks ending on a byte boundary: each chunk inflates on its own."""
    img = smooth_image(128, 256, 7)
    png, _, geo = encode(emu, img)
    idat = [d for t, d in chunks(png) if t == b"IDAT"]
    rows_per_strip = geo[4]
    for i, d in enumerate(idat[:-1]):
        z = zlib.decompressobj(-15)
        raw = z.decompress(d[2:] if i == 0 else d)
        assert len(raw) == min(rows_per_strip, 128 - i * rows_per_strip) * (3 * 256 + 1)
    assert idat[-1][:5] == b"\x01\x00\x00\xff\xff" and len(idat[-1]) == 9


def test_code_tables_are_complete_prefix_codes(emu):
    for k in range(8):
        lens = (ctypes.c_int * 271)()
        hdr = ctypes.c_int()
        assert emu.emu_png_table_summary(k, lens, ctypes.byref(hdr)) == 0
        lit = list(lens)[:257]
        assert all(1 <= l <= 15 for l in lit)
        assert 0 < hdr.value <= 48 * 32
        # matches carry the distance bit (and an extra bit from length 11 up)
        sym_len = {257 + i: lens[257 + i] - 1 - (1 if i + 3 >= 11 else 0) for i in range(14)}
        codes = lit + [sym_len[257 + i] for i in range(8)] + [sym_len[265], sym_len[267], sym_len[269]]
        assert sym_len[265] == sym_len[266] and sym_len[267] == sym_len[268] and sym_len[269] == sym_len[270]
        assert sum(2.0 ** -l for l in codes) == 1.0    # Kraft equality: zlib rejects incomplete literal/length codes
