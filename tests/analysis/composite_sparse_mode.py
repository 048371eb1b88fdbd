"""Would a one-pixel-per-lane loop pay for blocks that are mostly saturated?  (CPU, oracle = checker only.)

Per 8x8 block of the bench frame, with the live-box cull as the kernel runs it: the first round boundary at which at
most `t` pixels are live, the evaluations left after it, and the net instruction count of switching there (16 of 55
issue slots saved per pair of entries, ~90 instructions to flush the parked pixels and compact the live ones).
Result on the bench frame: 957 of the 1316 blocks with work would switch at t = 32; net 5.1 % of the kernel's
instructions (3.9 % at 24, 2.7 % at 16)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import omfs_b200
from omfs_b200 import avatar, synthetic
import oracle
W = H = 512
model, params, av, cam = synthetic.make_scene(n_gauss=100_000, n_frames=1, width=W, height=H)
baked = avatar.bake(av)
res = oracle.render(model, params, baked, [cam.pack()], W, H)
P0, P1 = res.pre.P0[0], res.pre.P1[0]
vals, ranges = res.binned.sorted_values, res.binned.ranges
L2_255 = np.float32(-7.99435343685885793770)
gxt = W // 16
SAVE_PER_PAIR = 16.0
results = {thr: dict(gain=0.0, cost=0.0, switched=0) for thr in (32, 24, 16, 8)}
SWITCH_COST = 90.0
total_eval_pairs = 0
units = 0
for tile in range(ranges.shape[0]):
    lo, hi = int(ranges[tile, 0]), int(ranges[tile, 1])
    if hi <= lo: continue
    g = vals[lo:hi]
    gx, gy = P0[g, 0], P0[g, 1]
    ca, cb, cc, lo_ = P1[g, 0], P1[g, 1], P1[g, 2], P1[g, 3]
    tx, ty = (tile % gxt) * 16, (tile // gxt) * 16
    xs = tx + np.arange(16, dtype=np.float32); ys = ty + np.arange(16, dtype=np.float32)
    dx = gx[:, None, None] - xs[None, None, :]; dy = gy[:, None, None] - ys[None, :, None]
    pw = ca[:, None, None] * dx * dx + cb[:, None, None] * dx * dy + cc[:, None, None] * dy * dy
    e = pw + lo_[:, None, None]
    ok = (pw <= 0) & (e >= L2_255)
    alpha = np.where(ok, np.minimum(0.99, np.exp2(e.astype(np.float64))), 0.0)
    Tb = np.cumprod(1.0 - alpha, axis=0)
    live = np.concatenate([np.ones((1, 16, 16), bool), ~(Tb < 1e-4)[:-1]], axis=0)
    live = np.logical_and.accumulate(live, axis=0)
    thr = (L2_255 - lo_).astype(np.float64)
    A = -ca.astype(np.float64); B = -cb.astype(np.float64) * 0.5; C = -cc.astype(np.float64)
    det = A * C - B * B; q = -thr
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.sqrt(np.maximum(q, 0) * C / det); ey = np.sqrt(np.maximum(q, 0) * A / det)
    vis = thr <= 0
    n = len(g); idx = np.arange(n); src = np.maximum((idx // 32 - 1) * 32, 0)
    for by in range(2):
        for bx in range(2):
            sx = slice(bx * 8, bx * 8 + 8); sy = slice(by * 8, by * 8 + 8)
            lv = live[:, sy, sx]; lb = lv.any(axis=(1, 2))
            lcol = lv.any(axis=1); lrow = lv.any(axis=2); ar = np.arange(8)
            cmin = np.where(lcol, ar, 99).min(axis=1); cmax = np.where(lcol, ar, -1).max(axis=1)
            rmin = np.where(lrow, ar, 99).min(axis=1); rmax = np.where(lrow, ar, -1).max(axis=1)
            x0 = tx + bx * 8 + cmin; x1 = tx + bx * 8 + cmax; y0 = ty + by * 8 + rmin; y1 = ty + by * 8 + rmax
            hit = vis & lb & (gx + ex >= x0[src]) & (gx - ex <= x1[src]) & (gy + ey >= y0[src]) & (gy - ey <= y1[src])
            nl = lv.sum(axis=(1, 2))
            units += 1
            total_eval_pairs += hit.sum() / 2
            for t in results:
                # first round boundary at which live <= t (and > 0)
                rb = np.arange(0, n, 32)
                cand = rb[(nl[rb] <= t) & (nl[rb] > 0)]
                if len(cand) == 0: continue
                k = cand[0]
                after = hit[k:].sum()
                results[t]["switched"] += 1
                results[t]["gain"] += after / 2 * SAVE_PER_PAIR
                results[t]["cost"] += SWITCH_COST
print("units", units, "eval pairs", total_eval_pairs, "-> eval instr ~", total_eval_pairs * 55)
for t, r in results.items():
    print(t, "switched units", r["switched"], "gain instr", int(r["gain"]), "cost", int(r["cost"]), "net", int(r["gain"] - r["cost"]),
          "net / 13.3M = %.3f" % ((r["gain"] - r["cost"]) / 13.3e6))
