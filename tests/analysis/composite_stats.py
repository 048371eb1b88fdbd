"""Workload statistics of the compositing stage on the bench scene (CPU, oracle = checker only).

For one frame: pairs R, (pair, 8x4 block) combinations surviving the bounding-box cull, pixel
evaluations, evaluations passing the alpha test, contributions actually blended before saturation.
Used to judge how far the compositing kernel is from its essential work.
"""
import sys
import os
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
P0, P1, P2 = res.pre.P0[0], res.pre.P1[0], res.pre.P2[0]
vals, ranges = res.binned.sorted_values, res.binned.ranges
R = res.binned.n_pairs
print("pairs R", R, "per gaussian", R / N, "visible", int((res.pre.tiles_touched[0] > 0).sum()))

L2_255 = np.float32(-7.99435343685885793770)
tot_blocks = 0
tot_eval = 0
tot_alpha = 0
tot_blend = 0
tot_blocks_live = 0     # (pair, block) combos processed before the block saturates
tot_rounds_live = 0
tot_blocks_exact = 0
tot_livebox = 0
lens = []
gxt = W // 16
for tile in range(ranges.shape[0]):
    lo, hi = int(ranges[tile, 0]), int(ranges[tile, 1])
    if hi <= lo:
        continue
    lens.append(hi - lo)
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
    ok = (pw <= 0) & (e >= L2_255)          # [n,16,16]
    alpha = np.minimum(0.99, np.exp2(e.astype(np.float64)))
    alpha = np.where(ok, alpha, 0.0)
    # transmittance before each Gaussian
    Tb = np.cumprod(1.0 - alpha, axis=0)
    Tprev = np.concatenate([np.ones((1, 16, 16)), Tb[:-1]], axis=0)
    stopped = Tb < 1e-4
    # pixel is live for gaussian i if not stopped before i
    live = np.concatenate([np.ones((1, 16, 16), bool), ~stopped[:-1]], axis=0)
    live = np.logical_and.accumulate(live, axis=0)
    tot_blend += int((ok & live).sum())
    # per 8x4 block
    okb = ok.reshape(-1, 4, 4, 2, 8).any(axis=(2, 4))           # [n, by(4), bx(2)] exact footprint per block
    liveb = live.reshape(-1, 4, 4, 2, 8).any(axis=(2, 4))
    tot_blocks_exact += int((okb & liveb).sum())
    # bbox cull of alpha>=1/255 ellipse: extents ex, ey from conic: power = -0.5*(A dx^2 + 2B dxdy + C dy^2)*log2e
    # solve: max |dx| on the level set e = L2_255
    thr = (L2_255 - lo_).astype(np.float64)   # pw >= thr (thr negative)
    A = -ca.astype(np.float64)
    B = -cb.astype(np.float64) * 0.5
    C = -cc.astype(np.float64)
    det = A * C - B * B
    q = -thr
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.sqrt(np.maximum(q, 0) * C / det)
        ey = np.sqrt(np.maximum(q, 0) * A / det)
    vis = thr <= 0
    for by in range(4):
        for bx in range(2):
            x0, x1 = tx + bx * 8, tx + bx * 8 + 7
            y0, y1 = ty + by * 4, ty + by * 4 + 3
            hit = vis & (gx + ex >= x0) & (gx - ex <= x1) & (gy + ey >= y0) & (gy - ey <= y1)
            lb = liveb[:, by, bx]
            tot_blocks += int(hit.sum())
            tot_blocks_live += int((hit & lb).sum())
            # live-bbox cull: bbox of still-live pixels, refreshed every 32 entries
            lv = live[:, by * 4:by * 4 + 4, bx * 8:bx * 8 + 8]
            for r0 in range(0, len(g), 32):
                m0 = lv[r0]
                if not m0.any():
                    break
                ys, xs = np.nonzero(m0)
                lx0, lx1, ly0, ly1 = x0 + xs.min(), x0 + xs.max(), y0 + ys.min(), y0 + ys.max()
                sl = slice(r0, r0 + 32)
                hb = vis[sl] & (gx[sl] + ex[sl] >= lx0) & (gx[sl] - ex[sl] <= lx1) & (gy[sl] + ey[sl] >= ly0) & (gy[sl] - ey[sl] <= ly1)
                tot_livebox += int((hb & lb[sl]).sum())
            n_live = int(lb.sum())
            tot_rounds_live += (n_live + 31) // 32
            sub_ok = ok[:, by * 4:by * 4 + 4, bx * 8:bx * 8 + 8]
            sub_live = live[:, by * 4:by * 4 + 4, bx * 8:bx * 8 + 8]
            m = hit & lb
            tot_eval += int(sub_live[m].sum())
            tot_alpha += int((sub_ok & sub_live)[m].sum())
lens = np.array(lens)
print("tiles with work", len(lens), "mean len", lens.mean(), "max", lens.max())
print("(pair,block) total", R * 8)
print("(pair,block) bbox-hit", tot_blocks, " live", tot_blocks_live, " exact-footprint live", tot_blocks_exact)
print("live-bbox cull hits", tot_livebox)
print("rounds live (32 entries)", tot_rounds_live)
print("pixel evals (live lanes)", tot_eval, " alpha-pass", tot_alpha, " blended(essential)", tot_blend)
print("lane efficiency of evals", tot_eval / max(1, tot_blocks_live * 32), " alpha-pass frac", tot_alpha / max(1, tot_eval))
