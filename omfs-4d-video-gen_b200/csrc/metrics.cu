// metrics.cu — R9 on frames that are still in HBM: the moments behind the reference's report metrics
// (02_Visual_Engine/validation_reporting.py: psnr :16-20, ssim_global :23-37) for T pairs of uint8 frames,
// one pass over both frame sets.
//
// Per frame pair (a, b), [H,W,3] uint8 each:
//   m0 = sum (a - b)^2 over all 3HW channel values          (exact, integer)
//   m1 = sum x, m2 = sum y, m3 = sum x^2, m4 = sum y^2, m5 = sum xy   with the float32 BT.601 luma
//        x = (0.299f R + 0.587f G) + 0.114f B  (individually rounded float32 operations, as numpy evaluates
//        the reference's expression on float32 images), accumulated in float64.
// The closed forms (MSE -> PSNR, moments -> global SSIM) are finished on the host
// (validation_reporting.metrics_from_moments).
//
// Roofline: pure HBM stream, 6 B per pixel read, nothing written but 48 B per frame; 12 pixels (36 B = 9
// aligned 32-bit words) per thread and iteration from each set, grid = (CTAs per frame, T).
#include <algorithm>

#include "common.cuh"

namespace omfs {

constexpr int kFmThreads = 256;

__device__ __forceinline__ float luma_f32(float r, float g, float b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}

struct Moments {
    unsigned long long sse;
    double sx, sy, sxx, syy, sxy;
    __device__ __forceinline__ void pixel(uint32_t ar, uint32_t ag, uint32_t ab, uint32_t br, uint32_t bg, uint32_t bb) {
        const int d0 = (int)ar - (int)br, d1 = (int)ag - (int)bg, d2 = (int)ab - (int)bb;
        sse += (unsigned long long)(d0 * d0 + d1 * d1 + d2 * d2);
        const double x = (double)luma_f32((float)ar, (float)ag, (float)ab);
        const double y = (double)luma_f32((float)br, (float)bg, (float)bb);
        sx += x;
        sy += y;
        sxx += x * x;
        syy += y * y;
        sxy += x * y;
    }
};

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[3], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xffu; }

__global__ void __launch_bounds__(kFmThreads) frame_metrics_kernel(long long n_pix, const uint8_t* __restrict__ A,
                                                                   const uint8_t* __restrict__ B,
                                                                   double* __restrict__ out) {
    const int t = blockIdx.y;
    const uint8_t* a = A + (size_t)t * n_pix * 3;
    const uint8_t* b = B + (size_t)t * n_pix * 3;
    Moments m{0ull, 0.0, 0.0, 0.0, 0.0, 0.0};
    // groups of 4 pixels = 12 bytes = 3 aligned words (frames start 4-byte aligned when 3*n_pix % 4 == 0 or t == 0;
    // the launcher only takes this path when every frame start is aligned)
    const long long groups = n_pix / 4;
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(a);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(b);
    for (long long g = (long long)blockIdx.x * kFmThreads + threadIdx.x; g < groups; g += (long long)gridDim.x * kFmThreads) {
        uint32_t wa[3], wb[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            wa[k] = __ldg(aw + 3 * g + k);
            wb[k] = __ldg(bw + 3 * g + k);
        }
#pragma unroll
        for (int p = 0; p < 4; p++)
            m.pixel(byte_of(wa, 3 * p), byte_of(wa, 3 * p + 1), byte_of(wa, 3 * p + 2), byte_of(wb, 3 * p),
                    byte_of(wb, 3 * p + 1), byte_of(wb, 3 * p + 2));
    }
    // ragged tail (n_pix % 4 pixels), by the first CTA's first threads
    if (blockIdx.x == 0) {
        const long long p = groups * 4 + threadIdx.x;
        if (p < n_pix) m.pixel(a[3 * p], a[3 * p + 1], a[3 * p + 2], b[3 * p], b[3 * p + 1], b[3 * p + 2]);
    }
    // warp, then CTA reduction; one atomic per CTA and moment
    double v[6] = {(double)m.sse, m.sx, m.sy, m.sxx, m.syy, m.sxy};
    __shared__ double s_part[kFmThreads / 32][6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 6; k++) s_part[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kFmThreads / 32; w++) s += s_part[w][threadIdx.x];
        atomicAdd(out + (size_t)t * 6 + threadIdx.x, s);
    }
}

}  // namespace omfs

using namespace omfs;

extern "C" int omfs_frame_metrics(int T, int height, int width, const uint8_t* d_a_u8, const uint8_t* d_b_u8,
                                  double* d_moments, void* stream) {
    OMFS_REQUIRE(T >= 0 && height > 0 && width > 0, "bad sizes");
    OMFS_REQUIRE(T <= 65535, "at most 65535 frame pairs per call");
    OMFS_REQUIRE(d_a_u8 && d_b_u8 && d_moments, "null argument");
    if (T == 0) return OMFS_OK;
    const long long n_pix = (long long)height * width;
    // the word path needs every frame to start on a 4-byte boundary
    OMFS_REQUIRE(((uintptr_t)d_a_u8 & 3) == 0 && ((uintptr_t)d_b_u8 & 3) == 0, "frame sets must be 4-byte aligned");
    OMFS_REQUIRE(T == 1 || (n_pix * 3) % 4 == 0, "height*width*3 must be a multiple of 4 when T > 1");
    cudaStream_t st = (cudaStream_t)stream;
    OMFS_CUDA(cudaMemsetAsync(d_moments, 0, sizeof(double) * 6 * (size_t)T, st));
    // sum over at most ~4 waves of CTAs; a 512x512 frame is 65536 groups = 256 CTA-iterations
    int per_frame = ceil_div(n_pix / 4 + 1, kFmThreads * 4);
    const int cap = std::max(1, (kNumSMs * 8 + T - 1) / T);
    if (per_frame > cap) per_frame = cap;
    frame_metrics_kernel<<<dim3(per_frame, T), kFmThreads, 0, st>>>(n_pix, d_a_u8, d_b_u8, d_moments);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}
