// composite.cu — U10: front-to-back alpha compositing, one 16x16 pixel tile per CTA.
//
// Arithmetic is the canonical sequence of exact_math.cuh::ex_blend (compiled with --fmad=false,
// explicit fmaf only), so every skip / stop decision is bit-reproducible against the oracle; the
// only approximate operation is ex2.approx (the oracle uses exp2f), which moves the image by ~1e-7.
//
// Structure
//   * one warp = one 8x4 pixel block (4 independent warps per CTA); the tile's depth-sorted list is consumed 32 entries
//     at a time with a lane-parallel footprint cull (extents precomputed by the preprocess kernel and
//     carried in P2.w) and a ballot, so only the ~1/3 of the tile's Gaussians that can touch the
//     block are evaluated — the kernel is FP32-issue bound, skipped (Gaussian, warp) pairs are the win;
//   * no block barriers: a finished pixel block frees its slot at once (see the kernel comment);
//   * early termination per pixel (T < 1e-4) and per warp (all 32 pixels saturated);
//   * the optional uint8 HWC image (the save_image quantisation) is produced by the same kernel,
//     so the frame sink costs no extra pass over HBM.
// Roofline: 40 B per tile pair against ~256 pixel evaluations of ~10-20
// instructions each: the FP32/SFU issue rate binds, not HBM; bench.py reports both.
#include <cuda_fp16.h>

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

// One warp = one 8x4 pixel block of a 16x16 tile; a CTA is just kCompWarps such warps of the same tile
// packed together (the hardware caps CTAs per SM at 32, packing lifts the resident warp count to the
// register limit).  The warps never synchronise with each other: there is no block barrier anywhere, a
// warp that saturates or runs out of Gaussians stops at once, which removes the barrier stalls (the top
// stall reason of the 256-thread tile-per-CTA version: ncu ..._issue_stalled_barrier 7.1 per issue).
//
// Per round of 32 list entries: lane l fetches entry l (index, centre, colour + packed cull extents),
// tests the Gaussian's alpha >= 1/255 footprint box against the warp's pixel block, and a ballot yields
// the entries worth evaluating, still in depth order.  Only those lanes fetch the conic and publish
// their 9 floats to the warp's shared-memory slots; every lane then evaluates the survivors with
// broadcast reads.  The next round's gathers are issued before the current round is evaluated, so the
// L2 latency of the dependent index -> record loads overlaps the arithmetic.
__device__ __forceinline__ void unpack_extents(float w, float& bx, float& by) {
    const uint32_t u = __float_as_uint(w);
    const __half2 h = *reinterpret_cast<const __half2*>(&u);
    bx = __low2float(h);
    by = __high2float(h);
}

constexpr int kCompWarps = 4;  // independent pixel-block warps per CTA (the hardware caps CTAs per SM at 32)

__global__ void __launch_bounds__(32 * kCompWarps) composite_kernel(int N, int width, int height, const float4* __restrict__ P0,
                                                       const float4* __restrict__ P1,
                                                       const float4* __restrict__ P2,
                                                       const uint32_t* __restrict__ vals,
                                                       const uint2* __restrict__ ranges, float bg0, float bg1,
                                                       float bg2, float* __restrict__ image,
                                                       uint8_t* __restrict__ image_u8) {
    // survivors of the current round, COMPACTED in depth order: 3 x float4 per entry
    // [gx gy ca cb | cc lo r g | b - - -], so the evaluation loop walks one pointer
    __shared__ float4 s_rec_all[kCompWarps][32 * 3];
    float4* s_rec = s_rec_all[threadIdx.x >> 5];

    const int gxt = (width + kTile - 1) / kTile, gyt = (height + kTile - 1) / kTile;
    const int unit = blockIdx.x * kCompWarps + (threadIdx.x >> 5);  // (tile, pixel block) work unit of this warp
    const int tile = unit >> 3, sub = unit & 7, seg = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const int bx0 = (tile % gxt) * kTile + (sub & 1) * 8;
    const int by0 = (tile / gxt) * kTile + (sub >> 1) * 4;
    if (bx0 >= width || by0 >= height) return;  // pixel block entirely outside the image
    const int px = bx0 + (lane & 7);
    const int py = by0 + (lane >> 3);
    const bool inside = px < width && py < height;
    const float pxf = (float)px, pyf = (float)py;
    const float wx0 = (float)bx0, wx1 = (float)(bx0 + 7), wy0 = (float)by0, wy1 = (float)(by0 + 3);
    const uint2 range = ranges[(size_t)seg * (gxt * gyt) + tile];
    const float4* p0 = P0 + (size_t)seg * N;
    const float4* p1 = P1 + (size_t)seg * N;
    const float4* p2 = P2 + (size_t)seg * N;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    bool done = !inside;

    // software pipeline: (g, a, c) of the round being evaluated, (gn, an, cn) of the next one
    uint32_t g = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (range.x + lane < range.y) {
        g = __ldg(vals + range.x + lane);
        a = ldg4(p0 + g);
        c = ldg4(p2 + g);
    }
    for (uint32_t base = range.x; base < range.y; base += 32) {
        const bool have = base + lane < range.y;
        uint32_t gn = 0;
        float4 an = make_float4(0.f, 0.f, 0.f, 0.f), cn = an;
        if (base + 32 + lane < range.y) {
            gn = __ldg(vals + base + 32 + lane);
            an = ldg4(p0 + gn);
            cn = ldg4(p2 + gn);
        }
        bool hit = false;
        if (have) {
            float ex, ey;
            unpack_extents(c.w, ex, ey);
            hit = (a.x + ex >= wx0) && (a.x - ex <= wx1) && (a.y + ey >= wy0) && (a.y - ey <= wy1);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
            if (hit) {
                const float4 b = ldg4(p1 + g);
                float4* dst = s_rec + 3 * __popc(mask & lanemask_lt);
                dst[0] = make_float4(a.x, a.y, b.x, b.y);
                dst[1] = make_float4(b.z, b.w, c.x, c.y);
                dst[2] = make_float4(c.z, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            if (!done) {
                const int cnt = __popc(mask);
                const float4* rec = s_rec;
                for (int j = 0; j < cnt; j++, rec += 3) {
                    const float4 sa = rec[0];
                    const float4 sb = rec[1];
                    const int r = ex_blend(sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w, rec[2].x, pxf, pyf, T, C0,
                                           C1, C2, Ex2Dev());
                    if (r == 2) {
                        done = true;
                        break;
                    }
                }
            }
            __syncwarp();
            if (__all_sync(0xffffffffu, done)) break;
        }
        g = gn;
        a = an;
        c = cn;
    }
    if (inside) {
        const float o0 = fmaf(T, bg0, C0), o1 = fmaf(T, bg1, C1), o2 = fmaf(T, bg2, C2);
        const size_t hw = (size_t)width * height;
        const size_t pix = (size_t)py * width + px;
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

extern "C" int omfs_composite(int S, int N, int width, int height, const float* d_P0, const float* d_P1,
                              const float* d_P2, const uint32_t* d_sorted_vals, const uint32_t* d_ranges,
                              const float* bg3, float* d_image, uint8_t* d_image_u8, void* stream) {
    OMFS_REQUIRE(S >= 0 && N > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(S <= 65535, "at most 65535 segments per call");
    OMFS_REQUIRE(d_P0 && d_P1 && d_P2 && d_sorted_vals && d_ranges && bg3, "null input");
    OMFS_REQUIRE(d_image || d_image_u8, "no output requested");
    if (S == 0) return OMFS_OK;
    const int tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    OMFS_REQUIRE((long long)tiles * 8 < (1ll << 31), "too many tiles");
    dim3 grid(tiles * 8 / kCompWarps, S);
    composite_kernel<<<grid, 32 * kCompWarps, 0, (cudaStream_t)stream>>>(N, width, height, (const float4*)d_P0,
                                                             (const float4*)d_P1, (const float4*)d_P2,
                                                             d_sorted_vals, (const uint2*)d_ranges, bg3[0], bg3[1],
                                                             bg3[2], d_image, d_image_u8);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
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
