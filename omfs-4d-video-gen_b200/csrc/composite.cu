// composite.cu — U10: front-to-back alpha compositing, one warp per 8x8 pixel block of a 16x16 tile.
//
// Arithmetic is the canonical sequence of exact_math.cuh::ex_blend (compiled with --fmad=false,
// explicit fmaf only), so every skip / stop decision is bit-reproducible against the oracle; the
// only approximate operation is ex2.approx (the oracle uses exp2f), which moves the image by ~1e-7.
//
// Structure
//   * one warp = one 8x8 pixel block, two pixels per lane.  The tile's depth-sorted list is walked in two levels:
//     the pair WORDS are scanned 128 at a time and only the entries whose block hint names this block (set by the
//     binning from the footprint box) are queued; queued entries are then fetched 32 at a time with every lane
//     busy, tested exactly against the box of the block's live pixels, and evaluated from shared memory — the
//     kernel is FP32-issue bound, so work that is never started is the win (22 % of the (entry, block)
//     combinations pass the hint, ~19 % are evaluated);
//   * persistent one-warp CTAs that draw (segment, tile, block) work units from a ticket counter; no block
//     barriers anywhere: a finished pixel block goes straight on to the next unit;
//   * early termination per pixel (T < 1e-4) and per warp (all 64 pixels saturated);
//   * the optional uint8 HWC image (the save_image quantisation) is produced by the same kernel,
//     so the frame sink costs no extra pass over HBM.
// Roofline: 40 B per tile pair against ~256 pixel evaluations of ~7 issue slots each: the
// issue rate binds (67-76 % of the issue slots busy, DRAM 6 % of peak: profiles/r2d_ncu_summary.md), not HBM;
// bench.py reports the HBM fraction BASELINE.json asks for and says so.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "exact_math.cuh"

namespace omfs {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Ex2Dev {
    __device__ __forceinline__ float operator()(float x) const { return ex2_approx(x); }
};

// One warp = one 8x8 pixel block of a 16x16 tile; a CTA is kCompWarps such warps (default 1: at 64
// registers the 32-CTA-per-SM cap and the register file both allow 32 resident warps, and a one-warp CTA
// never waits for a slower sibling — measured 6.5 % faster than 4 warps per CTA).  The warps
// never synchronise with each other: there is no block barrier anywhere, a warp that saturates or runs out
// of Gaussians stops at once, which removes the barrier stalls (the top stall reason of the 256-thread
// tile-per-CTA version: ncu ..._issue_stalled_barrier 7.1 per issue).
//
// Per round of 32 list entries: lane l fetches entry l (index, centre, colour + packed cull extents),
// tests the Gaussian's alpha >= 1/255 footprint box against the warp's pixel block, and a ballot yields
// the entries worth evaluating, still in depth order.  Only those lanes fetch the conic and publish
// their 9 floats to the warp's shared-memory slots; every lane then evaluates the survivors with
// broadcast reads.  The next round's gathers are issued before the current round is evaluated, so the
// L2 latency of the dependent index -> record loads overlaps the arithmetic.
//
// The kernel is issue-bound (ncu: issue slots 70 % busy, FMA pipe 49 %, DRAM 3 %), so the evaluation is
// written to spend as few issue slots per (Gaussian, pixel block) as possible:
//   * survivors are stored in PAIRS, structure-of-arrays, so the operations of the exponent run as
//     packed f32x2 instructions (FADD2/FMUL2/FFMA2 of sm_100: two IEEE-rounded binary32 results per
//     issue slot — the sequence per element is exactly exact_math.cuh::ex_blend's);
//   * a skipped Gaussian gets alpha = 0 (then T*(1-0) == T and fma(c,0,C) == C bit for bit), so there is
//     no per-entry divergence (no BSSY/BSYNC); a saturated pixel parks its transmittance in Tbg and goes
//     on with T = -0; saturation is detected by one compare + vote per pair (see blend_pair).
#ifndef OMFS_COMP_WARPS
#define OMFS_COMP_WARPS 1
#endif
#ifndef OMFS_COMP_RESIDENT_WARPS
#define OMFS_COMP_RESIDENT_WARPS 32  // per SM: the 32-CTA cap with one-warp CTAs; needs <= 64 registers
#endif
#define OMFS_COMP_BOUNDS __launch_bounds__(32 * OMFS_COMP_WARPS, OMFS_COMP_RESIDENT_WARPS / OMFS_COMP_WARPS)
#ifndef OMFS_COMP_UNROLL
#define OMFS_COMP_UNROLL 2
#endif
constexpr int kCompUnroll = OMFS_COMP_UNROLL;
// 1 (default): cull list entries against the bounding box of the block's LIVE pixels instead of the whole 8x8
// block.  Saturated pixels cannot change any more, so the smaller box is just as conservative.  On the bench
// frame only 25 % of the evaluated (entry, block) combinations still have all 64 pixels live and 30 % have at most
// 16; the live box removes 13 % of the evaluations (tests/analysis/composite_live_bbox.py: 0.872x with the two-round
// staleness of the software pipeline) for ~40 instructions in the 44 % of the rounds in which a pixel stopped.
// Measured A/B on one box: 1.0125 -> 0.9634 ms per 60 frames, 33.3k -> 34.3k frames/s (profiles/r1_ab_livebox.md).
#ifndef OMFS_COMP_LIVE_BOX
#define OMFS_COMP_LIVE_BOX 1
#endif
constexpr int kCompWarps = OMFS_COMP_WARPS;  // independent pixel-block warps per CTA (the hardware caps CTAs per SM at 32)
constexpr int kPairSlots = 16;  // 32 survivors per round = 16 pairs
// one pair slot = 5 float4: [gx0 gx1 gy0 gy1] [ca0 ca1 cb0 cb1] [cc0 cc1 lo0 lo1] [r0 g0 b0 -] [r1 g1 b1 -]
constexpr int kPairFloats = 20;

// alpha of one Gaussian at one pixel: min(0.99, 2^e), or 0 when ex_blend would skip it.
// keep = (e >= log2(1/255)) && (e <= lo) is the complement of ex_blend's skip test (a NaN exponent fails the first
// compare); spelled in PTX so the two compares chain into ONE predicate and one select.
// GUARD = false drops the `e <= lo` half (the published `power > 0` rejection).  That is only allowed for Gaussians
// for which it can never fire — see well_conditioned() below — and saves one of the five issue slots.
template <bool GUARD>
__device__ __forceinline__ float alpha_of(float e, float lo) {
    float alpha = fminf(0.99f, ex2_approx(e));
    if (GUARD) {
        asm("{\n\t.reg .pred p, q;\n\t"
            "setp.le.f32 q, %1, %2;\n\t"
            "setp.ge.and.f32 p, %1, %3, q;\n\t"
            "selp.f32 %0, %0, 0f00000000, p;\n\t}"
            : "+f"(alpha)
            : "f"(e), "f"(lo), "f"(kLog2Inv255));
    } else {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.ge.f32 p, %1, %2;\n\t"
            "selp.f32 %0, %0, 0f00000000, p;\n\t}"
            : "+f"(alpha)
            : "f"(e), "f"(kLog2Inv255));
    }
    return alpha;
}

// When can `e > lo` (ex_blend's rejection of a positive power) never happen?  With A = ca dx^2, B = cb dx dy,
// C = cc dy^2 and q = A + B + C <= -lmin r^2 (lmin, lmax: the eigenvalues of the scaled conic, r^2 = dx^2 + dy^2),
// ex_blend computes c0 = RN(lo + A(1 + d1)) <= lo and e = RN(s dy + c0) with s dy = B + C + eta, |eta| <= 2 eps M,
// M = (|ca| + |cc|) r^2 <= 2 lmax r^2.  Since lo - c0 >= |A|(1 - eps) - ulp(lo)/2, the exact sum before the last
// rounding satisfies  s dy + c0 - lo <= q + ulp(lo)/2 + 3 eps M;  e > lo needs that sum to reach lo + ulp(lo)/2,
// i.e. q + 3 eps M >= 0, i.e. lmin / lmax <= 6 eps = 3.6e-7.  lmin / lmax >= D / tr^2 (D = ca cc - cb^2/4 = lmin lmax,
// tr = ca + cc, |tr| >= lmax); the kernel asks for D > 1e-4 tr^2 in float arithmetic (whose error in D is below
// 1e-7 tr^2): two and a half orders of magnitude of margin.  Every Gaussian of a trained avatar qualifies (the EWA
// dilation of 0.3 px^2 bounds the conic's condition by radius^2 / 2.7); a round that holds one that does not is
// evaluated by the guarded loop.
__device__ __forceinline__ bool well_conditioned(const float4& cn) {
    const float tr = cn.x + cn.z;
    return cn.x * cn.z - 0.25f * (cn.y * cn.y) > 1.0e-4f * (tr * tr);
}

// State of the lane's two pixels (x, y) and (x, y + 4), packed so that the blend runs as f32x2 instructions
// (one issue slot for both pixels).  A pixel that has saturated — or lies outside the image — holds
// T = -0.0f: every product with it is -0, fma(c, -0, C) == C bit for bit, and as an UNSIGNED integer -0
// never compares below 1e-4, so "still live" needs no separate flag or predicate.
struct Pixels {
    float2 T, Tbg, C0, C1, C2;
};
constexpr uint32_t kStopBits = 0x38d1b717u;  // 0.0001f; T < 0.0001f  <=>  bits(T) < kStopBits for T >= +0

__device__ __forceinline__ bool stops(float testT) { return __float_as_uint(testT) < kStopBits; }

// The saturation rule for one pixel across a PAIR of Gaussians, applied to the weights of the no-saturation
// evaluation (T1 = T(1-a0), T2 = T1(1-a1), w0 = a0 T, w1 = a1 T1 already computed).  ex_blend, twice:
//   Gaussian 0 stops the pixel (T1 < 1e-4): neither Gaussian is blended, the pixel keeps T;
//   else Gaussian 1 stops it (T2 < 1e-4): only Gaussian 0 is blended, the pixel keeps T1.
// A weight of +0 leaves the colour bit-identical (fma(c, 0, C) == C for C >= +0); a parked pixel (-0) never stops.
__device__ __forceinline__ void saturate_pixel(float T, float T1, float T2, float& w0, float& w1, float& Tnew,
                                               float& Tbg) {
    const bool s0 = stops(T1);
    const bool s1 = !s0 && stops(T2);
    w0 = s0 ? 0.0f : w0;
    w1 = (s0 || s1) ? 0.0f : w1;
    Tbg = s0 ? T : (s1 ? T1 : Tbg);
    Tnew = (s0 || s1) ? -0.0f : T2;
}

// Two consecutive Gaussians (alphas a0, a1 at the lane's two pixels) onto the two pixels.  Saturation
// happens ONCE per pixel, so the pair is evaluated as if nobody saturates (no selects: the kernel is bound by
// issue slots); one integer compare per pixel + one vote per PAIR detects the other case (19 % of the pairs:
// some pixel of the 64 saturates), and only then the weights are corrected by selects.
// T2 = T*(1-a0)*(1-a1) <= T*(1-a0), so testing T2 covers both steps.  Every packed instruction is two
// individually rounded binary32 operations: the per-pixel sequence is exactly ex_blend's.
__device__ __forceinline__ void blend_pair(const float2 a0, const float2 a1, const float4& c0, const float4& c1,
                                           Pixels& p, bool& parked) {
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 T1 = __fmul2_rn(p.T, __fadd2_rn(one, make_float2(-a0.x, -a0.y)));
    const float2 T2 = __fmul2_rn(T1, __fadd2_rn(one, make_float2(-a1.x, -a1.y)));
    float2 w0 = __fmul2_rn(a0, p.T), w1 = __fmul2_rn(a1, T1), Tn = T2;
    const bool sat = stops(T2.x) || stops(T2.y);
    if (__builtin_expect(__any_sync(0xffffffffu, sat), 0)) {
        saturate_pixel(p.T.x, T1.x, T2.x, w0.x, w1.x, Tn.x, p.Tbg.x);
        saturate_pixel(p.T.y, T1.y, T2.y, w0.y, w1.y, Tn.y, p.Tbg.y);
        parked = true;  // (warp-uniform) some pixel of the block stopped in this round
        asm volatile("" ::: "memory");  // keep this a real (warp-uniform) branch, not a chain of selects
    }
    p.C0 = __ffma2_rn(make_float2(c0.x, c0.x), w0, p.C0);
    p.C1 = __ffma2_rn(make_float2(c0.y, c0.y), w0, p.C1);
    p.C2 = __ffma2_rn(make_float2(c0.z, c0.z), w0, p.C2);
    p.C0 = __ffma2_rn(make_float2(c1.x, c1.x), w1, p.C0);
    p.C1 = __ffma2_rn(make_float2(c1.y, c1.y), w1, p.C1);
    p.C2 = __ffma2_rn(make_float2(c1.z, c1.z), w1, p.C2);
    p.T = Tn;
}

// Can the Gaussian (centre gx, gy; conic + log2 opacity in cn = (ca, cb, cc, lo), pre-scaled as in ex_blend) reach
// alpha >= 1/255 anywhere in the pixel box [x0, x1] x [y0, y1]?  The exponent e = lo + q(dx, dy) is a concave
// quadratic with its maximum at the centre, so over a box that does not contain the centre the maximum lies on one of
// the (at most two) edges facing the centre: along the segment from the centre to any point of the box q only
// decreases, and the segment enters the box through such an edge.  On an edge q is a parabola in one variable: its
// vertex, clamped to the edge, is the exact maximum.  Both candidates are points of the box, so the larger of the two
// values is the maximum (and never more).  The test is an acceleration hint outside the parity surface: the division
// is approximate and the threshold is relaxed by 0.02 (the exponent is in log2 units: 1.4 % in alpha, against
// arithmetic that is good to ~1e-5), so it can only keep an entry too many; a form that is not negative definite
// (never produced by the preprocess for finite inputs) is always kept.
__device__ __forceinline__ bool reaches_box(float gx, float gy, const float4& cn, float x0, float x1, float y0, float y1) {
    const float dx0 = gx - x1, dx1 = gx - x0, dy0 = gy - y1, dy1 = gy - y0;   // dx = gx - px over the box
    const float xs = fminf(fmaxf(0.0f, dx0), dx1), ys = fminf(fmaxf(0.0f, dy0), dy1);   // box point nearest the centre
    const float hb = -0.5f * cn.y;
    const float dyb = fminf(fmaxf(__fdividef(hb, cn.z) * xs, dy0), dy1);   // vertex of q(xs, .) on the edge dx = xs
    const float dxb = fminf(fmaxf(__fdividef(hb, cn.x) * ys, dx0), dx1);   // vertex of q(., ys) on the edge dy = ys
    const float q1 = fmaf(fmaf(cn.z, dyb, cn.y * xs), dyb, (cn.x * xs) * xs);
    const float q2 = fmaf(fmaf(cn.x, dxb, cn.y * ys), dxb, (cn.z * ys) * ys);
    const bool concave = cn.x * cn.z - 0.25f * (cn.y * cn.y) > 0.0f;
    return !concave | (cn.w + fmaxf(q1, q2) >= kLog2Inv255 - 0.02f);
}

// The published round: npairs pair slots in shared memory onto the lane's two pixels.
template <bool GUARD>
__device__ __forceinline__ void evaluate_round(const float4* rec, int npairs, const float2 npx, const float2 npy0,
                                               const float2 npy1, Pixels& px, bool& parked) {
#pragma unroll kCompUnroll
    for (int j = 0; j < npairs; j++, rec += 5) {
        const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2];
        // exact_math.cuh::ex_blend's sequence, both Gaussians of the pair per instruction.  The column
        // terms (dx, c0 = lo + ca dx dx, v = cb dx) are shared by the lane's two pixels (same x).
        const float2 lo = make_float2(q2.z, q2.w);
        const float2 dx = __fadd2_rn(make_float2(q0.x, q0.y), npx);
        const float2 u = __fmul2_rn(make_float2(q1.x, q1.y), dx);
        const float2 c0 = __ffma2_rn(u, dx, lo);
        const float2 v = __fmul2_rn(make_float2(q1.z, q1.w), dx);
        auto exponent = [&](const float2 npy) -> float2 {
            const float2 dy = __fadd2_rn(make_float2(q0.z, q0.w), npy);
            const float2 sd = __ffma2_rn(make_float2(q2.x, q2.y), dy, v);
            return __ffma2_rn(sd, dy, c0);
        };
        const float2 ea = exponent(npy0), eb = exponent(npy1);
        const float2 a0 = make_float2(alpha_of<GUARD>(ea.x, lo.x), alpha_of<GUARD>(eb.x, lo.x));  // Gaussian 0 at both pixels
        const float2 a1 = make_float2(alpha_of<GUARD>(ea.y, lo.y), alpha_of<GUARD>(eb.y, lo.y));  // Gaussian 1
        blend_pair(a0, a1, rec[3], rec[4], px, parked);
    }
}

// A warp owns an 8x8 pixel block of its tile: lane l holds the two pixels (x, y) and (x, y + 4).  Two pixels
// per lane halve the number of warps that walk a tile's list and the shared-memory broadcasts per pixel evaluated
// — the L1/shared data pipe is the unit the one-pixel-per-lane version saturated first (ncu
// l1tex__data_pipe_lsu_wavefronts 90 %) — and let the blend run packed.
constexpr int kBlocksPerTile = 4;  // 8x8 pixel blocks (warps) per 16x16 tile

// The walk over the tile's list is in two levels, because most of it is not for this block: of the (entry, 8x8
// block) combinations of the bench frame only 22 % pass the footprint test, so a walk that gathers every entry's
// records to find that out spends more issue slots and L1 wavefronts on the 78 % than on evaluating the rest
// (round-1 kernel: 43 % of its warp instructions outside the evaluation loop, 2 of 3 gathers for nothing).
//   level 1  SCAN: the list word itself says whether the entry can reach this block (block hints set by the
//            binning, common.cuh kValIndexBits).  128 entries per step: 4 coalesced loads, 4 ballots; the indices
//            that pass go, in order, into a per-warp queue in shared memory.  No record is touched.
//   level 2  ROUND: up to 32 queued indices, one per lane — every lane busy — gather centre/extents, conic, colour,
//            are tested against the box of the block's LIVE pixels, and the survivors are compacted (still in depth
//            order) into the pair slots the evaluation reads.
// The loop is software-pipelined: the list words of the next scan step and the records of the next round are
// requested before the current round is evaluated.
#ifndef OMFS_COMP_QUEUE
#define OMFS_COMP_QUEUE 256
#endif
#ifndef OMFS_COMP_SCAN_GROUPS
#define OMFS_COMP_SCAN_GROUPS 4
#endif
constexpr int kQueueCap = OMFS_COMP_QUEUE;          // queued indices per warp (power of two)
constexpr int kScanGroups = OMFS_COMP_SCAN_GROUPS;  // 32-entry groups per scan step

__global__ void OMFS_COMP_BOUNDS composite_kernel(int n_seg, int N, int width, int height, const float4* __restrict__ P0,
                                                       const float4* __restrict__ P1,
                                                       const float4* __restrict__ P2,
                                                       const uint32_t* __restrict__ vals,
                                                       const uint2* __restrict__ ranges, float bg0, float bg1,
                                                       float bg2, float* __restrict__ image,
                                                       uint8_t* __restrict__ image_u8,
                                                       unsigned long long* __restrict__ tickets) {
    // survivors of the current round, COMPACTED in depth order and stored as pairs (see kPairFloats)
    __shared__ float4 s_rec_all[kCompWarps][kPairSlots * 5];
    __shared__ uint32_t s_queue_all[kCompWarps][kQueueCap];
    float4* s_rec = s_rec_all[threadIdx.x >> 5];
    float* s_f = reinterpret_cast<float*>(s_rec);
    uint32_t* s_queue = s_queue_all[threadIdx.x >> 5];

    const int gxt = (width + kTile - 1) / kTile, gyt = (height + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    // the odd slot of a round with an odd survivor count is evaluated with whatever it holds, made
    // harmless by lo = -inf: everything else in it must be finite, so start from zeros
    for (int i = lane; i < kPairSlots * 5; i += 32) s_rec[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    // Work units = (segment, tile, pixel block).  With a ticket counter the grid is ONE resident wave of
    // persistent warps that draw units from it (the next ticket is requested while the current unit is
    // processed, so the atomic's latency is hidden); without one the grid has a CTA per unit and the
    // loop below runs once.  One CTA per unit leaves the SMs at ~72 % of their resident-warp limit (CTA
    // launch latency against ~20 us of work per unit) and pays the prologue once per unit.
    const uint32_t tiles_per_seg = (uint32_t)(gxt * gyt);
    const uint32_t units_per_seg = tiles_per_seg * kBlocksPerTile;
    // unit numbers are 32-bit: the launcher checks units_total + the grid's overshoot < 2^32
    const uint32_t units_total = (uint32_t)n_seg * units_per_seg;
    const uint32_t unit_stride = gridDim.x * kCompWarps;
    auto draw = [&]() -> uint32_t {
        uint32_t t = 0;
        if (lane == 0) t = (uint32_t)atomicAdd(tickets, 1ull);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    uint32_t unit = blockIdx.x * kCompWarps + (threadIdx.x >> 5), unit_next = 0;
    if (tickets) unit = draw();
    // a warp's units only increase, by about one wave per draw: the segment is tracked by stepping (a 32-bit
    // division costs ~25 issue slots per unit), the tile row by a float reciprocal that is exact for tiles < 2^18
    uint32_t seg = 0, seg_unit0 = 0;   // current segment and its first unit
    const float inv_gxt = 1.0f / (float)gxt;
    for (; unit < units_total; unit = unit_next) {
    unit_next = tickets ? draw() : unit + unit_stride;
    if (unit - seg_unit0 >= 8u * units_per_seg) {
        seg = unit / units_per_seg;
        seg_unit0 = seg * units_per_seg;
    }
    while (unit - seg_unit0 >= units_per_seg) {
        seg_unit0 += units_per_seg;
        seg++;
    }
    const uint32_t rem = unit - seg_unit0;
    const uint32_t tile = rem / kBlocksPerTile, sub = rem % kBlocksPerTile;
    const uint32_t tyq = (uint32_t)(((float)tile + 0.5f) * inv_gxt);
    const int bx0 = (int)(tile - tyq * (uint32_t)gxt) * kTile + (int)(sub & 1u) * 8;
    const int by0 = (int)tyq * kTile + (int)(sub >> 1) * 8;
    const int pxi = bx0 + (lane & 7);
    const int pyi = by0 + (lane >> 3);
    const float2 npx = make_float2(-(float)pxi, -(float)pxi);
    const float2 npy0 = make_float2(-(float)pyi, -(float)pyi);
    const float2 npy1 = make_float2(-(float)(pyi + 4), -(float)(pyi + 4));
    Pixels px;
    px.T = make_float2((pxi < width && pyi < height) ? 1.0f : -0.0f, (pxi < width && pyi + 4 < height) ? 1.0f : -0.0f);
    px.Tbg = px.C0 = px.C1 = px.C2 = make_float2(0.0f, 0.0f);
    // cull box = bounding box of the block's live pixels (OMFS_COMP_LIVE_BOX), else the whole block
    float wx0 = (float)bx0, wx1 = (float)(bx0 + 7), wy0 = (float)by0, wy1 = (float)(by0 + 7);
    const uint2 range = ranges[(size_t)seg * tiles_per_seg + tile];
    const uint32_t rec0 = seg * (uint32_t)N;   // first record of the segment (S*N < 2^31)

    const int len = (int)(range.y - range.x);
    const int n_groups = (len + 31) >> 5;
    const uint32_t vp = range.x + (uint32_t)lane;
    const int hint_bit = kValIndexBits + (int)sub;
    // list words of group g (entries past the end read as 0: no hint bit, never queued)
    auto fetch_group = [&](int g) -> uint32_t { return (g * 32 + lane < len) ? __ldg(vals + (vp + (uint32_t)(g * 32))) : 0u; };

    if (n_groups > 0) {
        uint32_t w[kScanGroups];
#pragma unroll
        for (int k = 0; k < kScanGroups; k++) w[k] = fetch_group(k);
        int scan_g = 0;                    // first group of the next scan step (its words are in w[])
        uint32_t q_head = 0, q_tail = 0;   // queue positions popped / pushed so far (warp-uniform)
        bool have = false, cand = false;   // a popped round's records are in (a, b, c); this lane holds a candidate
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
        while (true) {
            bool parked = false;
            // 1. publish the popped round: live-box test, ballot, compaction into the pair slots
            int cnt = 0;
            bool guard = false;   // (warp-uniform) the round holds a Gaussian that needs ex_blend's `e <= lo` test
            if (have) {
                const bool hit = cand & reaches_box(a.x, a.y, b, wx0, wx1, wy0, wy1);
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                guard = __any_sync(0xffffffffu, hit && !well_conditioned(b));
                cnt = __popc(mask);
                if (hit) {
                    const int s = __popc(mask & lanemask_lt);
                    const int h = s & 1;
                    float* d = s_f + (s >> 1) * kPairFloats + h;
                    d[0] = a.x;
                    d[2] = a.y;
                    d[4] = b.x;
                    d[6] = b.y;
                    d[8] = b.z;
                    d[10] = b.w;
                    *reinterpret_cast<float4*>(s_f + (s >> 1) * kPairFloats + 12 + 4 * h) = make_float4(c.x, c.y, c.z, 0.f);
                    if (s == cnt - 1 && h == 0) d[11] = __int_as_float(0xff800000);  // odd count: mute the partner
                }
            }
            // 2. scan step: push the hinted entries of kScanGroups groups, request the next groups' words
            if (scan_g < n_groups && q_tail - q_head <= (uint32_t)(kQueueCap - 32 * kScanGroups)) {
#pragma unroll
                for (int k = 0; k < kScanGroups; k++) {
                    const bool in = (w[k] >> hint_bit) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, in);
                    if (in) s_queue[(q_tail + __popc(m & lanemask_lt)) & (kQueueCap - 1)] = w[k] & kValIndexMask;
                    q_tail += __popc(m);
                }
                scan_g += kScanGroups;
#pragma unroll
                for (int k = 0; k < kScanGroups; k++) w[k] = fetch_group(scan_g + k);
            }
            __syncwarp();
            // 3. pop the next round and request its records
            const uint32_t take = min(q_tail - q_head, 32u);
            have = take != 0u;
            if (have) {
                cand = (uint32_t)lane < take;
                const uint32_t g = rec0 + (cand ? s_queue[(q_head + lane) & (kQueueCap - 1)] : 0u);
                q_head += take;
                a = ldg4(P0 + g);
                b = ldg4(P1 + g);
                c = ldg4(P2 + g);
            }
            // 4. evaluate the published round
            if (guard)
                evaluate_round<true>(s_rec, (cnt + 1) >> 1, npx, npy0, npy1, px, parked);
            else
                evaluate_round<false>(s_rec, (cnt + 1) >> 1, npx, npy0, npy1, px, parked);
            __syncwarp();
            if (parked) {  // only a round in which a pixel stopped can finish the block or shrink its live box
                // bit l of m0 / m1: lane l's pixel (x, y) / (x, y + 4) is still live (sign bit of T clear)
                const uint32_t m0 = __ballot_sync(0xffffffffu, (__float_as_uint(px.T.x) >> 31) == 0u);
                const uint32_t m1 = __ballot_sync(0xffffffffu, (__float_as_uint(px.T.y) >> 31) == 0u);
                if ((m0 | m1) == 0u) break;  // every pixel parked: the block is finished
#if OMFS_COMP_LIVE_BOX
                // shrink the cull box to the live pixels (a round popped before this point is tested against it at
                // its publish; a stale, larger box would be just as correct: pixels never come back)
                uint32_t cols = m0 | m1;
                cols |= cols >> 16;
                cols |= cols >> 8;
                cols &= 0xffu;
                const int xmin = __ffs(cols) - 1, xmax = 31 - __clz(cols);
                const int ymin = m0 ? ((__ffs(m0) - 1) >> 3) : (((__ffs(m1) - 1) >> 3) + 4);
                const int ymax = m1 ? (((31 - __clz(m1)) >> 3) + 4) : ((31 - __clz(m0)) >> 3);
                wx0 = (float)(bx0 + xmin);
                wx1 = (float)(bx0 + xmax);
                wy0 = (float)(by0 + ymin);
                wy1 = (float)(by0 + ymax);
#endif
            }
            if (!have && scan_g >= n_groups) break;  // list scanned, queue drained, last round evaluated
        }
    }
    const size_t hw = (size_t)width * height;
    const float Ts[2] = {px.T.x, px.T.y}, Tb[2] = {px.Tbg.x, px.Tbg.y};
    const float Cs[2][3] = {{px.C0.x, px.C1.x, px.C2.x}, {px.C0.y, px.C1.y, px.C2.y}};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int py = pyi + 4 * k;
        if (pxi < width && py < height) {
            const float Tf = (__float_as_uint(Ts[k]) >> 31) ? Tb[k] : Ts[k];  // parked: what it saturated with
            const float o0 = fmaf(Tf, bg0, Cs[k][0]), o1 = fmaf(Tf, bg1, Cs[k][1]), o2 = fmaf(Tf, bg2, Cs[k][2]);
            const size_t pix = (size_t)py * width + pxi;
            if (image) {
                float* img = image + (size_t)seg * 3 * hw;
                img[pix] = o0;
                img[hw + pix] = o1;
                img[2 * hw + pix] = o2;
            }
            if (image_u8) {
                uint8_t* o = image_u8 + ((size_t)seg * hw + pix) * 3;
                o[0] = (uint8_t)fminf(fmaxf(o0 * 255.0f + 0.5f, 0.0f), 255.0f);
                o[1] = (uint8_t)fminf(fmaxf(o1 * 255.0f + 0.5f, 0.0f), 255.0f);
                o[2] = (uint8_t)fminf(fmaxf(o2 * 255.0f + 0.5f, 0.0f), 255.0f);
            }
        }
    }
    }  // work units
    // the last warp to run out of tickets leaves the counter pair zeroed for the next launch
    if (tickets && lane == 0) {
        __threadfence();
        if (atomicAdd(tickets + 1, 1ull) == (unsigned long long)unit_stride - 1ull) {
            tickets[0] = 0ull;
            tickets[1] = 0ull;
        }
    }
}

// float [S,3,H,W] -> uint8 [S,H,W,3] as a separate pass (used when only the float image exists)
__global__ void __launch_bounds__(256) to_uint8_kernel(long long n_pix_total, long long hw,
                                                       const float* __restrict__ image,
                                                       uint8_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix_total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / hw, p = i % hw;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = image[(s * 3 + c) * hw + p] * 255.0f + 0.5f;
            out[i * 3 + c] = (uint8_t)fminf(fmaxf(v, 0.0f), 255.0f);
        }
    }
}

}  // namespace omfs

using namespace omfs;

namespace omfs {
int composite_launch(int S, int N, int width, int height, const float* d_P0, const float* d_P1, const float* d_P2,
                     const uint32_t* d_sorted_vals, const uint32_t* d_ranges, const float* bg3, float* d_image,
                     uint8_t* d_image_u8, void* d_tickets, int warps_per_sm, cudaStream_t stream) {
    OMFS_REQUIRE(S >= 0 && N > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(d_P0 && d_P1 && d_P2 && d_sorted_vals && d_ranges && bg3, "null input");
    OMFS_REQUIRE(d_image || d_image_u8, "no output requested");
    if (S == 0) return OMFS_OK;
    const int tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    const long long units = (long long)S * tiles * kBlocksPerTile;
    const long long ctas_all = (units + kCompWarps - 1) / kCompWarps;
    if (warps_per_sm <= 0 || warps_per_sm > OMFS_COMP_RESIDENT_WARPS) warps_per_sm = OMFS_COMP_RESIDENT_WARPS;
    const long long wave = (long long)kNumSMs * std::max(1, warps_per_sm / kCompWarps);  // persistent CTAs
    OMFS_REQUIRE(units < (1ll << 32) - (1ll << 20), "too many work units for one launch");
    OMFS_REQUIRE(tiles <= (1 << (kValIndexBits - 10)), "too many tiles per frame");
    OMFS_REQUIRE(d_tickets || ctas_all < (1ll << 31), "too many work units for one launch without a ticket counter");
    const int grid = (int)((d_tickets && ctas_all > wave) ? wave : ctas_all);
    composite_kernel<<<grid, 32 * kCompWarps, 0, stream>>>(
        S, N, width, height, (const float4*)d_P0, (const float4*)d_P1, (const float4*)d_P2, d_sorted_vals,
        (const uint2*)d_ranges, bg3[0], bg3[1], bg3[2], d_image, d_image_u8, (unsigned long long*)d_tickets);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}
}  // namespace omfs

extern "C" int omfs_composite(int S, int N, int width, int height, const float* d_P0, const float* d_P1,
                              const float* d_P2, const uint32_t* d_sorted_vals, const uint32_t* d_ranges,
                              const float* bg3, float* d_image, uint8_t* d_image_u8, void* d_tickets, void* stream) {
    return composite_launch(S, N, width, height, d_P0, d_P1, d_P2, d_sorted_vals, d_ranges, bg3, d_image, d_image_u8,
                            d_tickets, 0, (cudaStream_t)stream);
}

extern "C" int omfs_to_uint8(int S, int width, int height, const float* d_image, uint8_t* d_out, void* stream) {
    OMFS_REQUIRE(S >= 0 && width > 0 && height > 0 && d_image && d_out, "bad arguments");
    if (S == 0) return OMFS_OK;
    const long long hw = (long long)width * height;
    const long long total = hw * S;
    int blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    to_uint8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(total, hw, d_image, d_out);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}
