"""Compare compositing work decompositions on the bench scene (CPU, oracle = checker only).

For one frame and every 16x16 tile: how many evaluation iterations each scheme needs
  A  8x8 block per warp, one survivor list per block (bbox cull)                 [current kernel]
  B  8x8 block per warp, four 4x4 quads each with its own survivor list; trip = max over quads
  C  as A with an exact ellipse-vs-rectangle cull
  D  as B with exact cull
Liveness (early termination) is applied per block (A, C) or per quad (B, D), refreshed per entry.
"""
import sys, os
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
tot = dict(A=0, B=0, C=0, D=0, Bsum=0, Dsum=0, A_rounds=0)
for tile in range(ranges.shape[0]):
    lo, hi = int(ranges[tile, 0]), int(ranges[tile, 1])
    if hi <= lo:
        continue
    g = vals[lo:hi]
    gx, gy = P0[g, 0], P0[g, 1]
    ca, cb, cc, lo_ = P1[g, 0], P1[g, 1], P1[g, 2], P1[g, 3]
    tx, ty = (tile % gxt) * 16, (tile // gxt) * 16
    px = (tx + np.arange(16, dtype=np.float32))[None, None, :]
    py = (ty + np.arange(16, dtype=np.float32))[None, :, None]
    dx = gx[:, None, None] - px
    dy = gy[:, None, None] - py
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
    # quads: 4x4 pixel quads, index [qy(4), qx(4)]
    okq = ok.reshape(-1, 4, 4, 4, 4).any(axis=(2, 4))       # exact (discrete) footprint per quad
    liveq = live.reshape(-1, 4, 4, 4, 4).any(axis=(2, 4))
    hitq = np.zeros_like(okq)
    for qy in range(4):
        for qx in range(4):
            x0, x1 = tx + qx * 4, tx + qx * 4 + 3
            y0, y1 = ty + qy * 4, ty + qy * 4 + 3
            hitq[:, qy, qx] = vis & (gx + ex >= x0) & (gx - ex <= x1) & (gy + ey >= y0) & (gy - ey <= y1)
    for by in range(2):
        for bx in range(2):
            hq = hitq[:, by * 2:by * 2 + 2, bx * 2:bx * 2 + 2].reshape(-1, 4)
            oq = okq[:, by * 2:by * 2 + 2, bx * 2:bx * 2 + 2].reshape(-1, 4)
            lq = liveq[:, by * 2:by * 2 + 2, bx * 2:bx * 2 + 2].reshape(-1, 4)
            lb = lq.any(axis=1)
            tot["A"] += int((hq.any(axis=1) & lb).sum())
            tot["C"] += int((oq.any(axis=1) & lb).sum())
            cq = (hq & lq).sum(axis=0); tot["B"] += int(cq.max()); tot["Bsum"] += int(cq.sum())
            cq = (oq & lq).sum(axis=0); tot["D"] += int(cq.max()); tot["Dsum"] += int(cq.sum())
            tot["A_rounds"] += (int(lb.sum()) + 31) // 32
print({k: v for k, v in tot.items()})
print("B/A", tot["B"] / tot["A"], "C/A", tot["C"] / tot["A"], "D/A", tot["D"] / tot["A"], "quad balance B", tot["Bsum"] / 4 / tot["B"])
