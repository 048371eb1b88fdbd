// common.cuh — shared by every translation unit of libomfs_b200.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/omfs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libomfs_b200 is written for sm_100a (B200) only"
#endif

namespace omfs {

constexpr int kTile = OMFS_TILE;
constexpr int kFF = OMFS_FF_STRIDE;
constexpr int kCam = OMFS_CAM_FLOATS;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define OMFS_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return omfs::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define OMFS_LAUNCH_CHECK() OMFS_CUDA(cudaGetLastError())

#define OMFS_REQUIRE(cond, msg)                       \
    do {                                              \
        if (!(cond)) {                                \
            omfs::set_error("%s: %s", __func__, msg); \
            return OMFS_ERR_INVALID;                  \
        }                                             \
    } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a process that opens sessions on two
// GPUs must set it on both, and two host threads may get here at once.  One DeviceOnce per kernel (or kernel
// family) remembers what has been granted on each device.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
    std::mutex mu;
    int granted[kMaxDevices] = {0};
};
template <typename F>
inline int ensure_dyn_smem(DeviceOnce& once, F* func, int bytes) {
    int dev = 0;
    OMFS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> guard(once.mu);
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (!tracked || once.granted[dev] < bytes) {
        OMFS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (tracked) once.granted[dev] = bytes;
    }
    return OMFS_OK;
}

// launch counter (gpu_launches in bench.py is read from here)
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

// binning stages (binning.cu), called separately by the session so that it can time them
// binning_prepare zeroes the per-batch counters and returns where a fused producer (bind_preprocess_launch) is to
// accumulate the depth-digit histograms and the tile counts; depth_sort / tile_ranges are then told that this
// has happened (`prepared_and_histogrammed`, `counted`) and skip their own memset / histogram / count kernels.
int binning_prepare(int S, int N, int width, int height, size_t capacity, void* d_workspace, uint32_t** d_hist_depth,
                    uint32_t** d_tile_cnt, cudaStream_t stream);
int binning_depth_sort(int S, int N, int width, int height, size_t capacity, const uint32_t* d_depth_keys,
                       void* d_workspace, cudaStream_t stream, bool prepared_and_histogrammed = false);
int binning_tile_ranges(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                        const uint32_t* d_tiles_touched, uint32_t* d_ranges, uint32_t* d_num_pairs,
                        int* d_status_flag, unsigned long long* d_pair_accum, uint32_t* d_pair_max,
                        void* d_workspace, cudaStream_t stream, bool counted = false);
// U5+U6 (exact_geom.cu); d_hist_depth / d_tile_cnt non-null = the fused form (see binning_prepare)
int bind_preprocess_launch(int S, int N, int F, int width, int height, const float* d_ff, const int32_t* d_seg_frame,
                           const float* d_cams, const float* d_xyzb, const float* d_scale_lo, const float* d_rot,
                           const float* d_sh, float* d_P0, float* d_P1, float* d_P2, uint32_t* d_tiles_touched,
                           uint32_t* d_depth_keys, uint32_t* d_hist_depth, uint32_t* d_tile_cnt, cudaStream_t stream);
int binning_emit_scatter(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                         const uint32_t* d_tiles_touched, uint32_t* d_sorted_vals, void* d_workspace,
                         cudaStream_t stream);
int binning_rebuild_keys(int S, int N, int width, int height, size_t capacity, const uint32_t* d_ranges,
                         const uint32_t* d_vals, const uint32_t* d_depth_keys, uint64_t* d_keys64, void* d_workspace,
                         cudaStream_t stream);

// compositing (composite.cu) with an explicit number of persistent warps per SM (0 = fill the SM).  The
// session launches fewer than fit (kCompPipelinedWarps) when the next batch's front end runs beside it.
int composite_launch(int S, int N, int width, int height, const float* d_P0, const float* d_P1, const float* d_P2,
                     const uint32_t* d_sorted_vals, const uint32_t* d_ranges, const float* bg3, float* d_image,
                     uint8_t* d_image_u8, void* d_tickets, int warps_per_sm, cudaStream_t stream);

// 128-bit streaming loads / stores.  The frame-invariant avatar streams and per-frame records are
// read through the read-only path; outputs that the next kernel re-reads stay default-cached so
// they can live in the 126 MB L2 between the kernels of one batch.
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// ---- the pair list entry.  d_sorted_vals[i] = Gaussian index (low kValIndexBits bits: the parity surface) | a 4-bit
// hint (top bits): bit (kValIndexBits + 2*yhalf + xhalf) is set when the Gaussian's alpha >= 1/255 footprint box
// (centre +- the extents of P0.z) contains a pixel of that 8x8 block of the 16x16 tile.  The binning (emit_scatter) sets
// the hints, the compositing warps skip list entries whose bit for their block is clear without touching the
// Gaussian's records.  Conservative by construction of the extents (exact_geom.cu: pack_cull_extents).
constexpr int kValIndexBits = OMFS_VAL_INDEX_BITS;
constexpr uint32_t kValIndexMask = (1u << kValIndexBits) - 1u;

// the two half-precision extents packed in P0.z
__device__ __forceinline__ void unpack_extents(float w, float& bx, float& by) {
    const uint32_t u = __float_as_uint(w);
    const __half2 h = *reinterpret_cast<const __half2*>(&u);
    bx = __low2float(h);
    by = __high2float(h);
}

// 8x8-block index range [bmin, bmax] of the integer pixels p with c - e <= p <= c + e (empty: bmin > bmax).
// The float sums round like the compositing cull's own (1 ulp of a pixel coordinate against extents inflated by
// 0.02 px); the clamps keep the conversions in range for infinite extents.
__device__ __forceinline__ void block_range(float c, float e, int& bmin, int& bmax) {
    const float lo = fminf(fmaxf(ceilf(c - e), -8.0f), 1048576.0f);
    const float hi = fminf(fmaxf(floorf(c + e), -16.0f), 1048576.0f);
    bmin = (int)lo >> 3;
    bmax = (lo <= hi) ? ((int)hi >> 3) : -1048576;   // NaN or inverted: nothing
}

}  // namespace omfs
