"""CPU-side checks of the exact domain (no GPU needed):

  * the PRODUCT's per-element math (csrc/exact_math.cuh, the source the CUDA kernels compile),
    built for the host by tests/emu with -ffp-contract=off, is bit-identical to the independently
    written C oracle on face frames, bind+preprocess and compositing;
  * the oracle's tiled pipeline equals a brute-force numpy renderer that knows nothing about tiles,
    keys or sorting (every pixel walks ALL Gaussians in depth order) — a check of the oracle itself.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = ctypes.CDLL(os.path.join(HERE, "emu", "libemu.so"))
    f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.emu_face_frames.argtypes = [ctypes.c_int] * 3 + [f, i32, f]
    L.emu_bind_preprocess.argtypes = [ctypes.c_int] * 4 + [f] * 9 + [u32]
    L.emu_composite.argtypes = [ctypes.c_int] * 4 + [f, f, f, u32, u32, f, f]
    return L


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_product_math_is_bit_identical_to_oracle(emu, small_scene):
    model, params, av, baked, cam = small_scene
    W, H = cam.width, cam.height
    T, N = params.n_frames, baked["n"]
    res = oracle.render(model, params, baked, [cam.pack()] * T, W, H)
    ff = np.zeros_like(res.ff)
    emu.emu_face_frames(T, model.n_verts, model.n_faces, res.verts, model.faces, ff)
    assert np.array_equal(bits(ff), bits(res.ff))
    for s in range(T):
        P = [np.zeros((N, 4), np.float32) for _ in range(3)]
        tt = np.zeros(N, np.uint32)
        emu.emu_bind_preprocess(N, model.n_faces, W, H, np.ascontiguousarray(ff[s]), baked["xyzb"], baked["scale_lo"],
                                baked["rot"], baked["sh"], cam.pack(), P[0], P[1], P[2], tt)
        assert np.array_equal(bits(P[0]), bits(res.pre.P0[s]))
        assert np.array_equal(bits(P[1]), bits(res.pre.P1[s]))
        assert np.array_equal(bits(P[2]), bits(res.pre.P2[s]))
        assert np.array_equal(tt, res.pre.tiles_touched[s])
    img = np.zeros_like(res.image)
    emu.emu_composite(T, N, W, H, res.pre.P0.reshape(-1), res.pre.P1.reshape(-1), res.pre.P2.reshape(-1),
                      res.binned.sorted_values, res.binned.ranges.reshape(-1), np.ones(3, np.float32), img.reshape(-1))
    assert np.array_equal(bits(img), bits(res.image))


def _degenerate_avatar(model, seed=41, n=600):
    """An avatar with poisoned Gaussians: NaN / infinite local positions, an exp-overflowing scale, a centre 1e30
    away, a NaN quaternion, a NaN opacity and SH coefficient."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, synthetic
    av = synthetic.make_avatar(n, model.n_faces, seed=seed)
    av.xyz[5] = np.nan
    av.xyz[6, 1] = np.inf
    av.scaling[7] = 120.0            # exp(120) overflows float32
    av.xyz[8] = 1e30
    av.rotation[9] = np.nan
    av.opacity[10] = np.nan
    av.sh[11, 3, 1] = np.nan
    av.xyz[12] = -1e30
    return av, avatar.bake(av)


def test_degenerate_gaussians_are_culled_identically(emu, small_scene):
    """Non-finite or absurd Gaussians never reach the binning: product math and oracle cull the same ones (explicit
    finiteness test in ex_bind_project, float clamps before every float->int conversion), bit for bit, and the
    rest of the frame is untouched by them."""
    model, params, _, _, cam = small_scene
    W, H = cam.width, cam.height
    with np.errstate(all="ignore"):
        av, baked = _degenerate_avatar(model)
        N = baked["n"]
        one = params.slice(0, 1)
        res = oracle.render(model, one, baked, [cam.pack()], W, H)
    assert np.isfinite(res.pre.P0).all() and np.isfinite(res.pre.P1[..., :3]).all() and np.isfinite(res.pre.P2).all()
    assert np.isnan(res.pre.P1[0, 10, 3])                    # the NaN opacity rides along, and is never blended
    assert np.isfinite(res.image).all()
    for n in (5, 6, 7, 8, 9, 12):
        assert res.pre.tiles_touched[0, n] == 0, n          # culled
    assert res.pre.tiles_touched[0, 11] > 0 and res.pre.tiles_touched[0, 10] > 0   # NaN colour / opacity do not cull
    assert (res.pre.tiles_touched[0] > 0).sum() > N // 2     # the healthy ones are still there
    P = [np.zeros((N, 4), np.float32) for _ in range(3)]
    tt = np.zeros(N, np.uint32)
    emu.emu_bind_preprocess(N, model.n_faces, W, H, np.ascontiguousarray(res.ff[0]), baked["xyzb"], baked["scale_lo"],
                            baked["rot"], baked["sh"], cam.pack(), P[0], P[1], P[2], tt)
    assert np.array_equal(tt, res.pre.tiles_touched[0])
    assert np.array_equal(bits(P[0]), bits(res.pre.P0[0])) and np.array_equal(bits(P[1]), bits(res.pre.P1[0]))
    assert np.array_equal(bits(P[2]), bits(res.pre.P2[0]))
    img = np.zeros_like(res.image)
    emu.emu_composite(1, N, W, H, res.pre.P0.reshape(-1), res.pre.P1.reshape(-1), res.pre.P2.reshape(-1),
                      res.binned.sorted_values, res.binned.ranges.reshape(-1), np.ones(3, np.float32), img.reshape(-1))
    assert np.array_equal(bits(img), bits(res.image))
    # the poisoned Gaussians change nothing but their own absence: same frame as an avatar without them
    keep = np.setdiff1d(np.arange(N), [5, 6, 7, 8, 9, 10, 12])
    from omfs_b200 import avatar as avatar_mod, synthetic
    clean = synthetic.Avatar(av.xyz[keep], av.scaling[keep], av.rotation[keep], av.opacity[keep], av.sh[keep],
                             av.binding[keep])
    ref2 = oracle.render(model, one, avatar_mod.bake(clean), [cam.pack()], W, H)
    assert np.array_equal(bits(ref2.image), bits(res.image))


def test_oracle_binning_invariants(small_scene):
    model, params, av, baked, cam = small_scene
    W, H = cam.width, cam.height
    T, N = params.n_frames, baked["n"]
    res = oracle.render(model, params, baked, [cam.pack()] * T, W, H)
    b = res.binned
    assert b.n_pairs == int(res.pre.tiles_touched.sum()) > 0
    order = np.argsort(b.keys, kind="stable")           # numpy's stable sort = the definition
    assert np.array_equal(b.sorted_keys, b.keys[order])
    assert np.array_equal(b.sorted_values, b.values[order])
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    assert b.sort_bits == 32 + int(np.ceil(np.log2(T * tiles)))
    lens = b.ranges[:, 1].astype(np.int64) - b.ranges[:, 0]
    assert int(lens.sum()) == b.n_pairs
    assert np.array_equal(np.repeat(np.arange(T * tiles), lens), (b.sorted_keys >> np.uint64(32)).astype(np.int64))
    assert np.array_equal(res.pre.radii > 0, res.pre.tiles_touched > 0)


def test_oracle_matches_untiled_brute_force():
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, synthetic
    W, H = 48, 32
    model, params, av, cam = synthetic.make_scene(n_gauss=300, n_frames=1, width=W, height=H, n_verts=162)
    baked = avatar.bake(av)
    res = oracle.render(model, params, baked, [cam.pack()], W, H)
    P0, P1, P2 = res.pre.P0[0].astype(np.float64), res.pre.P1[0].astype(np.float64), res.pre.P2[0].astype(np.float64)
    radii = res.pre.radii[0]
    vis = np.where(radii > 0)[0]
    vis = vis[np.argsort(res.pre.P0[0][vis, 2], kind="stable")]     # depth order, ties by index
    img = np.ones((3, H, W))
    thr = np.log2(1.0 / 255.0)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    for y in range(H):
        for x in range(W):
            T_, C = 1.0, np.zeros(3)
            for g in vis:
                # the tile rectangle is part of the published algorithm: a Gaussian is only seen by
                # pixels of the tiles its 3-sigma square touches
                r = float(radii[g])
                minx, maxx = int((P0[g, 0] - r) / 16), int((P0[g, 0] + r + 15) / 16)
                miny, maxy = int((P0[g, 1] - r) / 16), int((P0[g, 1] + r + 15) / 16)
                minx, maxx = min(gx, max(0, minx)), min(gx, max(0, maxx))
                miny, maxy = min(gy, max(0, miny)), min(gy, max(0, maxy))
                if not (minx <= x // 16 < maxx and miny <= y // 16 < maxy):
                    continue
                dx, dy = P0[g, 0] - x, P0[g, 1] - y
                pw = P1[g, 0] * dx * dx + P1[g, 1] * dx * dy + P1[g, 2] * dy * dy
                if pw > 0:
                    continue
                e = pw + P1[g, 3]
                if e < thr:
                    continue
                a = min(0.99, 2.0 ** e)
                if T_ * (1 - a) < 1e-4:
                    break
                C += P2[g, :3] * a * T_
                T_ *= 1 - a
            img[:, y, x] = C + T_ * 1.0
    assert np.abs(img - res.image[0]).max() < 1e-5


@pytest.mark.parametrize("case", range(8))
def test_product_math_equals_oracle_over_random_cameras(emu, case):
    """Seeded sweep of the situations one frontal camera never reaches: wide and narrow lenses, eyes inside the head
    (near-plane culls, splats far larger than the image), grazing and off-centre views, image sizes that are not
    multiples of the 16-pixel tile, big and tiny splats.  Product math (host build of exact_math.cuh) and the C
    oracle stay bit-identical in every preprocess output and in the composited frame."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, cameras, synthetic
    rng = np.random.default_rng(5000 + case)
    W, H = [(33, 47), (160, 112), (257, 65), (16, 16), (100, 100), (48, 200), (129, 127), (64, 31)][case]
    model = synthetic.make_flame_model(seed=77 + case, n_verts=642)
    params = synthetic.make_frame_params(1, seed=300 + case, n_verts=642)
    av = synthetic.make_avatar(3000, model.n_faces, seed=11 + case)
    av.scaling[:] += np.float32(rng.uniform(-2.0, 1.5))            # from sub-pixel to image-filling splats
    av.opacity[:] += np.float32(rng.uniform(-3.0, 3.0))
    baked = avatar.bake(av)
    angle = float(rng.uniform(0.05, 2.4))
    dist = float(rng.choice([0.02, 0.08, 0.3, 1.0, 5.0]))          # 0.02 / 0.08: the eye is inside the head
    d = rng.normal(size=3)
    eye = dist * d / np.linalg.norm(d)
    target = rng.normal(0, 0.05, size=3)
    cam = cameras.camera_from_c2w(cameras.look_at_c2w(eye, target), angle, W, H)
    N = baked["n"]
    res = oracle.render(model, params, baked, [cam.pack()], W, H)
    P = [np.zeros((N, 4), np.float32) for _ in range(3)]
    tt = np.zeros(N, np.uint32)
    emu.emu_bind_preprocess(N, model.n_faces, W, H, np.ascontiguousarray(res.ff[0]), baked["xyzb"], baked["scale_lo"],
                            baked["rot"], baked["sh"], cam.pack(), P[0], P[1], P[2], tt)
    assert np.array_equal(tt, res.pre.tiles_touched[0])
    for k, ref in enumerate((res.pre.P0, res.pre.P1, res.pre.P2)):
        assert np.array_equal(bits(P[k]), bits(ref[0])), k
    img = np.zeros_like(res.image)
    emu.emu_composite(1, N, W, H, res.pre.P0.reshape(-1), res.pre.P1.reshape(-1), res.pre.P2.reshape(-1),
                      res.binned.sorted_values, res.binned.ranges.reshape(-1), np.ones(3, np.float32), img.reshape(-1))
    assert np.array_equal(bits(img), bits(res.image))
    assert np.isfinite(res.image).all()
    # binning invariants at this camera: every pair's tile lies inside its Gaussian's rectangle, ranges partition
    R = res.binned.n_pairs
    assert R == int(res.pre.tiles_touched.sum())
    keys = res.binned.sorted_keys[:R]
    assert (np.diff(keys.astype(np.uint64)) >= 0).all() if R > 1 else True
