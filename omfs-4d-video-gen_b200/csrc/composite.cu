// composite.cu — U10: front-to-back alpha compositing, one 16x16 pixel tile per CTA.
//
// Arithmetic is the canonical sequence of exact_math.cuh::ex_blend (compiled with --fmad=false,
// explicit fmaf only), so every skip / stop decision is bit-reproducible against the oracle; the
// only approximate operation is ex2.approx (the oracle uses exp2f), which moves the image by ~1e-7.
//
// Structure
//   * the tile's depth-sorted Gaussian list is consumed in chunks of 256: each thread gathers one
//     Gaussian's 9 compositing floats (centre, pre-scaled conic, log2 opacity, colour) from the
//     [S,N] float4 streams (L2-resident: every Gaussian is referenced by ~3.6 tiles) and stages
//     them in shared memory; the per-pixel loop then reads them as warp-wide broadcasts;
//   * a warp owns an 8x4 pixel block (not a 16x2 strip) and culls per block: each lane tests one
//     staged Gaussian's alpha >= 1/255 footprint (a conservative box, cull_box below) against the
//     block, a ballot gives the ~1/3 of the tile's Gaussians that can touch it, and only those are
//     evaluated — the kernel is FP32-issue bound, so skipped (Gaussian, warp) pairs are the win;
//   * early termination at three levels: per pixel (T < 1e-4), per warp (all 32 pixels done: the
//     warp stops evaluating and only helps staging), per CTA (__syncthreads_count).
//   * the optional uint8 HWC image (the save_image quantisation) is produced by the same kernel,
//     so the frame sink costs no extra pass over HBM.
// Roofline: 40 B per tile pair against ~256 pixel evaluations of ~10-20
// instructions each: the FP32/SFU issue rate binds, not HBM; bench.py reports both.
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

constexpr int kChunk = 256;

// Conservative footprint of one Gaussian for the per-warp cull.  A pixel at offset d from the centre
// can only pass the `e >= log2(1/255)` test if  q(d) = -(ca dx^2 + cb dx dy + cc dy^2) <= Lq with
// Lq = lo - log2(1/255); minimising q over dy gives dx^2 <= Lq * (-cc) / (ca*cc - cb^2/4) (and the
// symmetric bound for dy).  The half-widths are inflated (x1.001 + 0.01 px), far more than the 1e-6
// relative rounding of the kernel's own evaluation, so the cull never drops a contributing pixel:
// results stay bit-identical to evaluating every (Gaussian, pixel) pair.
__device__ __forceinline__ float4 cull_box(float gx, float gy, float ca, float cb, float cc, float lo) {
    const float Lq = lo - kLog2Inv255;
    const float D = ca * cc - 0.25f * cb * cb;
    if (!(Lq >= 0.0f)) return make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);  // can never contribute
    if (!(D > 0.0f)) return make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);    // degenerate: always test
    const float inv = Lq / D;
    const float bx = sqrtf(fmaxf(-cc * inv, 0.0f)) * 1.001f + 0.01f;
    const float by = sqrtf(fmaxf(-ca * inv, 0.0f)) * 1.001f + 0.01f;
    return make_float4(gx - bx, gx + bx, gy - by, gy + by);
}

__global__ void __launch_bounds__(256) composite_kernel(int N, int width, int height, const float4* __restrict__ P0,
                                                        const float4* __restrict__ P1,
                                                        const float4* __restrict__ P2,
                                                        const uint32_t* __restrict__ vals,
                                                        const uint2* __restrict__ ranges, float bg0, float bg1,
                                                        float bg2, float* __restrict__ image,
                                                        uint8_t* __restrict__ image_u8) {
    __shared__ float4 s_a[kChunk];     // gx gy ca cb
    __shared__ float4 s_b[kChunk];     // cc lo r g
    __shared__ float s_c[kChunk];      // b
    __shared__ float4 s_box[kChunk];   // xmin xmax ymin ymax of the alpha >= 1/255 footprint

    const int gxt = (width + kTile - 1) / kTile, gyt = (height + kTile - 1) / kTile;
    const int tile = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // a warp owns an 8x4 pixel block of the tile
    const int bx0 = (tile % gxt) * kTile + (warp & 1) * 8;
    const int by0 = (tile / gxt) * kTile + (warp >> 1) * 4;
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
    for (uint32_t base = range.x; base < range.y; base += kChunk) {
        // barrier (protects the staging buffers) + CTA-level early out
        if (__syncthreads_count(done) == 256) break;
        const uint32_t idx = base + tid;
        if (idx < range.y) {
            const uint32_t g = __ldg(vals + idx);
            const float4 a = ldg4(p0 + g), b = ldg4(p1 + g), c = ldg4(p2 + g);
            s_a[tid] = make_float4(a.x, a.y, b.x, b.y);
            s_b[tid] = make_float4(b.z, b.w, c.x, c.y);
            s_c[tid] = c.z;
            s_box[tid] = cull_box(a.x, a.y, b.x, b.y, b.z, b.w);
        }
        __syncthreads();
        const int n = (int)min((uint32_t)kChunk, range.y - base);
        if (__all_sync(0xffffffffu, done)) continue;  // warp-level: nothing left to shade here
        // 32 Gaussians at a time: each lane tests ONE Gaussian's footprint against the warp's pixel
        // block, the ballot is the list of Gaussians worth evaluating (still in depth order)
        for (int sub = 0; sub < n; sub += 32) {
            const int jj = sub + lane;
            bool hit = false;
            if (jj < n) {
                const float4 bb = s_box[jj];
                hit = (bb.y >= wx0) && (bb.x <= wx1) && (bb.w >= wy0) && (bb.z <= wy1);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            bool any_stop = false;
            while (mask) {
                const int j = sub + __ffs(mask) - 1;
                mask &= mask - 1;
                if (!done) {
                    const float4 a = s_a[j];
                    const float4 b = s_b[j];
                    const int r = ex_blend(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, s_c[j], pxf, pyf, T, C0, C1, C2,
                                           Ex2Dev());
                    if (r == 2) {
                        done = true;
                        any_stop = true;
                    }
                }
            }
            if (__any_sync(0xffffffffu, any_stop) && __all_sync(0xffffffffu, done)) break;
        }
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
    dim3 grid(tiles, S);
    composite_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(N, width, height, (const float4*)d_P0,
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
