"""Compositing cull against the bounding box of the block's LIVE pixels (CPU, oracle = checker only).

The kernel culls a list entry when the inflated box of its alpha >= 1/255 ellipse misses the warp's 8x8 block.
Pixels that saturated (T < 1e-4) can no longer change, so the test may use the bounding box of the pixels that
are still live instead of the whole block: same image, fewer evaluation iterations in blocks that are almost
done (silhouettes, gaps).  Counts per frame on the bench scene:
  A  current: ellipse box vs the 8x8 block, while any pixel of the block is live
  E  ellipse box vs the bounding box of the live pixels
  G  ellipse box vs the live pixels themselves (any live pixel inside the box; lower bound for box tests)
  C  exact footprint vs the block (any pixel with alpha >= 1/255), F the same vs live pixels only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa
from omfs_b200 import avatar, synthetic
import oracle

W = H = 512
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
model, params, av, cam = synthetic.make_scene(n_gauss=N, n_frames=1, width=W, height=H)
baked = avatar.bake(av)
res = oracle.render(model, params, baked, [cam.pack()], W, H)
P0, P1 = res.pre.P0[0], res.pre.P1[0]
vals, ranges = res.binned.sorted_values, res.binned.ranges
L2_255 = np.float32(-7.99435343685885793770)
gxt = W // 16
tot = dict(A=0, E=0, Es=0, G=0, C=0, F=0, entries=0, rounds=0, rounds_changed=0, rounds_bbox_changed=0, Es_half=0, Es_le32=0, Es_le16=0, Es_live_sum=0, live_hist=np.zeros(65, np.int64))
for tile in range(ranges.shape[0]):
    lo, hi = int(ranges[tile, 0]), int(ranges[tile, 1])
    if hi <= lo:
        continue
    g = vals[lo:hi]
    gx, gy = P0[g, 0], P0[g, 1]
    ca, cb, cc, lo_ = P1[g, 0], P1[g, 1], P1[g, 2], P1[g, 3]
    tx, ty = (tile % gxt) * 16, (tile // gxt) * 16
    xs = tx + np.arange(16, dtype=np.float32)
    ys = ty + np.arange(16, dtype=np.float32)
    dx = gx[:, None, None] - xs[None, None, :]
    dy = gy[:, None, None] - ys[None, :, None]
    pw = ca[:, None, None] * dx * dx + cb[:, None, None] * dx * dy + cc[:, None, None] * dy * dy
    e = pw + lo_[:, None, None]
    ok = (pw <= 0) & (e >= L2_255)
    alpha = np.where(ok, np.minimum(0.99, np.exp2(e.astype(np.float64))), 0.0)
    Tb = np.cumprod(1.0 - alpha, axis=0)
    stopped = Tb < 1e-4
    live = np.concatenate([np.ones((1, 16, 16), bool), ~stopped[:-1]], axis=0)
    live = np.logical_and.accumulate(live, axis=0)
    thr = (L2_255 - lo_).astype(np.float64)
    A = -ca.astype(np.float64); B = -cb.astype(np.float64) * 0.5; C = -cc.astype(np.float64)
    det = A * C - B * B
    q = -thr
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.sqrt(np.maximum(q, 0) * C / det)
        ey = np.sqrt(np.maximum(q, 0) * A / det)
    vis = thr <= 0
    inx = (xs[None, :] >= (gx - ex)[:, None]) & (xs[None, :] <= (gx + ex)[:, None])     # [n,16] column inside the box
    iny = (ys[None, :] >= (gy - ey)[:, None]) & (ys[None, :] <= (gy + ey)[:, None])
    tot["entries"] += 4 * len(g)
    for by in range(2):
        for bx in range(2):
            sx = slice(bx * 8, bx * 8 + 8); sy = slice(by * 8, by * 8 + 8)
            lv = live[:, sy, sx]                       # [n,8,8]
            lb = lv.any(axis=(1, 2))
            hitA = vis & inx[:, sx].any(axis=1) & iny[:, sy].any(axis=1)
            tot["A"] += int((hitA & lb).sum())
            # live bounding box: columns / rows that hold a live pixel
            lcol = lv.any(axis=1); lrow = lv.any(axis=2)            # [n,8]
            ar = np.arange(8)
            cmin = np.where(lcol, ar, 99).min(axis=1); cmax = np.where(lcol, ar, -1).max(axis=1)
            rmin = np.where(lrow, ar, 99).min(axis=1); rmax = np.where(lrow, ar, -1).max(axis=1)
            x0 = tx + bx * 8 + cmin; x1 = tx + bx * 8 + cmax
            y0 = ty + by * 8 + rmin; y1 = ty + by * 8 + rmax
            hitE = vis & lb & (gx + ex >= x0) & (gx - ex <= x1) & (gy + ey >= y0) & (gy - ey <= y1)
            tot["E"] += int(hitE.sum())
            # as the kernel would see it: the cull of round k runs before round k-1 is evaluated, so it uses the
            # box left by round k-2 (stale boxes are larger: still conservative)
            n = len(g)
            idx = np.arange(n)
            src = np.maximum((idx // 32 - 1) * 32, 0)
            hitEs = vis & lb & (gx + ex >= x0[src]) & (gx - ex <= x1[src]) & (gy + ey >= y0[src]) & (gy - ey <= y1[src])
            tot["Es"] += int(hitEs.sum())
            # lane occupancy of what is still evaluated: live pixels per combination, and how often one of the lane's
            # two pixel rows groups (rows 0-3 / rows 4-7) is entirely dead (a one-pixel-per-lane loop would do)
            nl = lv[hitEs].sum(axis=(1, 2))
            top = lv[hitEs][:, :4, :].any(axis=(1, 2)); bot = lv[hitEs][:, 4:, :].any(axis=(1, 2))
            tot["Es_half"] += int((~top | ~bot).sum())
            tot["Es_le32"] += int((nl <= 32).sum()); tot["Es_le16"] += int((nl <= 16).sum())
            tot["Es_live_sum"] += int(nl.sum())
            # rounds walked while live, and rounds after which the live mask / the live box changed
            last = int(lb.sum())                       # entries walked (live is monotone)
            nr = (last + 31) // 32
            tot["rounds"] += nr
            for k in range(nr):
                a = k * 32; b = min(a + 32, n - 1)
                if (lv[a] != lv[b]).any():
                    tot["rounds_changed"] += 1
                    if (x0[a], x1[a], y0[a], y1[a]) != (x0[b], x1[b], y0[b], y1[b]):
                        tot["rounds_bbox_changed"] += 1
            inbox = inx[:, None, sx] & iny[:, sy, None]
            tot["G"] += int((vis & (inbox & lv).any(axis=(1, 2))).sum())
            okb = ok[:, sy, sx]
            tot["C"] += int((okb.any(axis=(1, 2)) & lb).sum())
            tot["F"] += int((okb & lv).any(axis=(1, 2)).sum())
            ev = hitA & lb
            tot["live_hist"] += np.bincount(lv[ev].sum(axis=(1, 2)), minlength=65)
h = tot.pop("live_hist")
print(tot)
for k in ("E", "Es", "G", "C", "F"):
    print(f"{k}/A = {tot[k] / tot['A']:.3f}")
print("under Es: mean live pixels %.1f of 64; one row group dead %.3f; <=32 live %.3f; <=16 live %.3f"
      % (tot["Es_live_sum"] / tot["Es"], tot["Es_half"] / tot["Es"], tot["Es_le32"] / tot["Es"], tot["Es_le16"] / tot["Es"]))
c = np.cumsum(h) / h.sum()
print("live pixels per evaluated (entry, block) under A: share with <=8 live %.3f, <=16 %.3f, <=32 %.3f, <=48 %.3f, 64 live %.3f"
      % (c[8], c[16], c[32], c[48], h[64] / h.sum()))
