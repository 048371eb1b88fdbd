"""Host restatement of emit_scatter's pair walk (csrc/binning.cu: compaction into records, 32-ary cursor search,
es_starts / es_pair with the software-pipelined cursor chain), lane by lane with the kernel's integer widths, against
the plain expansion of the rectangles.  It pins the index arithmetic the GPU parity tests exercise end to end:
record packing (21-bit pair offsets, 10-bit local indices and tile coordinates, 11-bit widths), the reciprocal
multiply for k / width, a cursor that starts in the middle of a rectangle, zero-count Gaussians, warps with empty
ranges, and chunk totals up to the 2^20 maximum.  No GPU, no oracle: the expected lists come from two nested loops."""
import numpy as np
import pytest

OFF_BITS = 21
OFF_MASK = (1 << OFF_BITS) - 1
WARPS = 8
PAD = 40
U32 = 0xFFFFFFFF


def compact(rects):
    """rects: per local Gaussian (minx, rminy, rw, rows) of the band-clipped rectangle (rows == 0: no pairs).
    Returns the record array as the kernel builds it (with the sentinel pad), the pair total and the record count."""
    recs, off = [], 0
    for li, (minx, rminy, rw, rows) in enumerate(rects):
        cnt = rw * rows
        if cnt:
            assert off <= OFF_MASK and li < 1024 and minx < 1024 and rminy < 1024 and rw <= 1024
            recs.append((off | (li << OFF_BITS), minx | (rminy << 10) | (rw << 20)))
            off += cnt
    n_rec = len(recs)
    recs += [(OFF_MASK, 0)] * PAD
    return recs, off, n_rec


def es_starts(recs, J, p, lane):
    rel = ((recs[(p + 1 + lane) & U32][0] & OFF_MASK) - J) & U32
    return (1 << rel) if rel < 32 else 0


def walk(rects, gx):
    """(local Gaussian, band-local tile) of every pair, in the order the eight warps' steps produce them."""
    recs, total, n_rec = compact(rects)
    magic = [0] + [(0x80000000 // r + 1) & U32 for r in range(1, gx + 1)]
    per_warp = (total + WARPS - 1) // WARPS
    out = []
    for warp in range(WARPS):
        my_lo, my_hi = min(total, warp * per_warp), min(total, (warp + 1) * per_warp)
        p = 0
        if my_hi > my_lo:   # 32-ary search: last record that starts at or before my first pair
            c1 = sum((recs[min(lane * 32, n_rec)][0] & OFF_MASK) <= my_lo for lane in range(32))
            b = c1 - 1
            c2 = sum((recs[min(b * 32 + lane, n_rec)][0] & OFF_MASK) <= my_lo for lane in range(32))
            p = b * 32 + c2 - 1
        starts_next = 0
        for lane in range(32):
            starts_next |= es_starts(recs, my_lo, p, lane)
        J = my_lo
        while J < my_hi:
            starts, p_step = starts_next, p
            p += bin(starts).count("1")
            starts_next = 0
            for lane in range(32):   # the next step's cursor chain is issued first (probing past the end reads sentinels)
                starts_next |= es_starts(recs, J + 32, p, lane)
            for lane in range(32):
                x, y = recs[p_step + bin(starts & ((2 << lane) - 1)).count("1")]
                if J + lane >= my_hi:
                    continue
                k = (J + lane - (x & OFF_MASK)) & U32
                rw = y >> 20
                cy = ((2 * k) & U32) * magic[rw] >> 32
                tx = (y & 1023) + (k - cy * rw)
                tyl = ((y >> 10) & 1023) + cy
                out.append((x >> OFF_BITS, tyl * gx + tx))
            J += 32
    return out


def expand(rects, gx):
    return [(li, (rminy + cy) * gx + minx + cx) for li, (minx, rminy, rw, rows) in enumerate(rects)
            for cy in range(rows) for cx in range(rw)]


def random_rects(rng, n, gx, rows_max, wmax, hmax, empty):
    rects = []
    for _ in range(n):
        if rng.random() < empty:
            rects.append((0, 0, 1, 0))
            continue
        rw, rows = int(rng.integers(1, wmax + 1)), int(rng.integers(1, hmax + 1))
        rects.append((int(rng.integers(0, gx - rw + 1)), int(rng.integers(0, rows_max - rows + 1)), rw, rows))
    return rects


@pytest.mark.parametrize("seed,gx,rows_max,wmax,hmax,empty", [
    (1, 32, 32, 4, 4, 0.5),      # configs[2]-like: small rectangles, half of the chunk culled
    (2, 64, 16, 12, 6, 0.3),     # configs[3]-like band
    (3, 64, 16, 64, 16, 0.0),    # screen-filling: most steps inside one rectangle
    (4, 65, 15, 30, 15, 0.9),    # ragged band, almost everything culled: warps with empty ranges
    (5, 1, 1024, 1, 700, 0.2),   # one tile column, tall rectangles
    (6, 1024, 1, 900, 1, 0.2),   # one tile row, the widest rectangles the record format holds
])
def test_walk_equals_plain_expansion(seed, gx, rows_max, wmax, hmax, empty):
    rng = np.random.default_rng(seed)
    for n in (1024, 1000, 37):
        rects = random_rects(rng, n, gx, rows_max, wmax, hmax, empty)
        assert walk(rects, gx) == expand(rects, gx)


def test_walk_edge_cases():
    gx = 32
    assert walk([(0, 0, 1, 0)] * 1024, gx) == []                                  # nothing owns a pair
    assert walk([(3, 2, 1, 1)], gx) == [(0, 2 * gx + 3)]                          # one pair
    one = [(0, 0, 1, 0)] * 500 + [(0, 0, 32, 32)] + [(0, 0, 1, 0)] * 523          # one rectangle = the whole band
    assert walk(one, gx) == expand(one, gx)
    full = [(0, 0, 32, 32)] * 1024                                                # 2^20 pairs: the offset field's maximum
    got = walk(full, gx)
    assert len(got) == 1 << 20 and got == expand(full, gx)


def test_reciprocal_multiply_is_exact():
    """k / rw as umulhi(2k, 2^31 / rw + 1) for every width and every k the kernel can meet (k < rw * 1024)."""
    for rw in range(1, 1025):
        k = np.arange(0, rw * 1024, dtype=np.uint64)
        m = np.uint64((0x80000000 // rw + 1) & U32)
        assert np.array_equal(((2 * k) * m) >> np.uint64(32), k // np.uint64(rw))
