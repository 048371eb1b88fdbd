"""Bring-up / diagnostics script for a GPU box: stage-by-stage parity of the CUDA path against the
oracle, with verbose mismatch reports.  (The judged parity tests live in tests/; this prints more.)

    python tests/analysis/gpu_check.py [--n 20000] [--frames 3] [--size 256] [--gemm 0|1]
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa: E402,F401
from omfs_b200 import avatar, runtime, synthetic  # noqa: E402
from omfs_b200.runtime import DeviceArray as DA  # noqa: E402
import oracle  # noqa: E402


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def report(name, got, ref, exact=True):
    got = np.asarray(got)
    ref = np.asarray(ref)
    if exact:
        same = got.shape == ref.shape and (bits_equal(got, ref) if got.dtype == np.float32 else np.array_equal(got, ref))
        if same:
            print(f"  [OK ] {name}: bit-exact {got.shape}")
        else:
            if got.dtype == np.float32:
                neq = got.view(np.uint32) != ref.view(np.uint32)
            else:
                neq = got != ref
            idx = np.argwhere(neq)
            print(f"  [BAD] {name}: {int(neq.sum())} of {neq.size} differ; first at {idx[:3].tolist()}")
            for i in idx[:3]:
                print("        got", got[tuple(i)], "ref", ref[tuple(i)])
        return same
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
    print(f"  [{'OK ' if err <= 1e-3 else 'BAD'}] {name}: max abs err {err:.3e}")
    return err <= 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--gemm", type=int, default=0)
    ap.add_argument("--verts", type=int, default=synthetic.FLAME_V)
    args = ap.parse_args()
    L = runtime.load_library()
    runtime.check(L.omfs_device_check(0))
    W = H = args.size
    T = args.frames
    model, params, av, cam = synthetic.make_scene(n_gauss=args.n, n_frames=T, width=W, height=H, n_verts=args.verts)
    baked = avatar.bake(av)
    N, V, F = baked["n"], model.n_verts, model.n_faces
    ok = True

    # ---------------- session end to end (both GEMM implementations)
    for impl in ([args.gemm] if args.gemm else [1, 0]):
        t0 = time.time()
        sess = runtime.Session(model, baked, W, H, max_batch=2, gemm_impl=impl, debug_keys=True)
        sess.set_subject(params.shape, params.static_offset)
        u8, img = sess.render_host(params, [cam], want_f32=True)
        print(f"session impl={impl}: rendered {T} frames in {time.time()-t0:.2f}s stats={sess.stats()} dims={sess.dims()}")
        d = sess.dims()
        ref_full = oracle.render(model, params, baked, [cam.pack()] * T, W, H)
        last = T - 1 - ((T - 1) % 2)  # first frame of the last batch (max_batch=2)
        nb = T - last
        verts_chunk = sess.tap_array("verts", (T, V, 3), np.float32)
        ok &= report(f"verts (impl {impl})", verts_chunk, ref_full.verts, exact=False)
        print("     verts max err %.3e" % np.abs(verts_chunk - ref_full.verts).max())
        vp = sess.tap_array("vp", (T, d["npad"]), np.float32)
        print("     vp finite:", np.isfinite(vp).all(), "joint cols", vp[0, 3 * V:3 * V + 6])
        # exact domain, fed with the GPU's own vertices
        ref = oracle.render(model, params, baked, [cam.pack()] * T, W, H, verts=verts_chunk)
        ff = sess.tap_array("ff", (nb, F, 20), np.float32)
        ok &= report("face frames (last batch)", ff, ref.ff[last:])
        P0 = sess.tap_array("P0", (nb, N, 4), np.float32)
        P1 = sess.tap_array("P1", (nb, N, 4), np.float32)
        P2 = sess.tap_array("P2", (nb, N, 4), np.float32)
        tt = sess.tap_array("tiles_touched", (nb, N), np.uint32)
        ok &= report("P0", P0, ref.pre.P0[last:])
        ok &= report("P1", P1, ref.pre.P1[last:])
        ok &= report("P2", P2[..., :3], ref.pre.P2[last:][..., :3])
        ok &= report("tiles_touched", tt, ref.pre.tiles_touched[last:])
        # binning of the last batch vs oracle binning of the same segments
        pre_last = oracle.Preprocessed(ref.pre.P0[last:], ref.pre.P1[last:], ref.pre.P2[last:],
                                       ref.pre.tiles_touched[last:], ref.pre.mu[last:])
        bref = oracle.binning(pre_last, W, H)
        R = bref.n_pairs
        print(f"     pairs last batch: gpu {sess.dims()['pairs_last_batch']} oracle {R}")
        keys = sess.tap_array("keys", (R,), np.uint64)
        vals = sess.tap_array("vals", (R,), np.uint32)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        ranges = sess.tap_array("ranges", (nb * tiles, 2), np.uint32)
        ok &= report("sorted keys", keys, bref.sorted_keys)
        ok &= report("sorted vals", vals, bref.sorted_values)
        ok &= report("ranges", ranges, bref.ranges)
        ok &= report("image (shared verts)", img, ref.image, exact=False)
        ok &= report("image (independent)", img, ref_full.image, exact=False)
        print("     u8 identical to oracle quantisation: %.6f" % (oracle.to_uint8(ref.image) == u8).mean())
        sess.close()
    print("ALL OK" if ok else "SOME CHECKS FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
