// exact_geom.cu — U4 (face frames) and the fused U5+U6 kernel (parent-triangle transform of each
// Gaussian + cull / project / EWA covariance / SH colour).  Compiled with --fmad=false: every
// float result here is bit-reproducible against the oracle.
//
// Both kernels are pure HBM streams (80 B/face, 288 B/Gaussian).  Layout choices:
//   * all per-Gaussian inputs and outputs are float4 SoA, so one warp instruction moves 512
//     contiguous bytes;
//   * the SH block (192 of the 240 input bytes) is 12 float4 planes and is only read for
//     Gaussians that survive culling;
//   * the per-frame face records (5 float4 per face, ~0.8 MB per frame) are gathered through the
//     read-only path and stay L2-resident across the segments of a batch.
#include <cuda_fp16.h>

#include "common.cuh"
#include "exact_math.cuh"

namespace omfs {

// Conservative half-extents (pixels) of the region where a Gaussian can pass the compositing test
// `power2 + lo >= log2(1/255)`, packed as two round-UP halves into P0.z (the slot beside the centre and the radius:
// the binning reads ONE record per Gaussian for the tile rectangle and the block hints, compositing one for the
// centre and the cull box; the depth, which only the parity taps read back, rides in P2.w).  With
// q(d) = -(ca dx^2 + cb dx dy + cc dy^2) and Lq = lo - log2(1/255), minimising q over dy gives
// dx^2 <= Lq * (-cc) / (ca*cc - cb^2/4) (symmetrically for dy).  The extents are inflated (x1.002 +
// 0.02 px, then rounded up to half precision), far more than the 1e-6 relative rounding of the
// compositing arithmetic: the per-warp cull that consumes them can only skip pixels that would have
// been skipped anyway.  P0.z is an acceleration hint, NOT part of the parity surface (the oracle
// has no such field); a value of +inf means "always test", -inf "never contributes".
__device__ __forceinline__ float pack_cull_extents(float ca, float cb, float cc, float lo) {
    const float Lq = lo - kLog2Inv255;
    const float D = ca * cc - 0.25f * cb * cb;
    float bx, by;
    if (!(Lq >= 0.0f)) {
        bx = by = -INFINITY;
    } else if (!(D > 0.0f)) {
        bx = by = INFINITY;
    } else {
        const float inv = Lq / D;
        bx = sqrtf(fmaxf(-cc * inv, 0.0f)) * 1.002f + 0.02f;
        by = sqrtf(fmaxf(-ca * inv, 0.0f)) * 1.002f + 0.02f;
    }
    const __half2 h = __halves2half2(__float2half_ru(bx), __float2half_ru(by));
    return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
}

// ---------------------------------------------------------------------------------------- U4
__global__ void __launch_bounds__(256) face_frames_kernel(int T, int V, int F, const float* __restrict__ verts,
                                                          const int32_t* __restrict__ faces,
                                                          float4* __restrict__ ff) {
    const long long total = (long long)T * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / F), f = (int)(i % F);
        const float* vb = verts + (size_t)t * V * 3;
        const int i0 = __ldg(faces + f * 3 + 0), i1 = __ldg(faces + f * 3 + 1), i2 = __ldg(faces + f * 3 + 2);
        const float p0[3] = {vb[i0 * 3], vb[i0 * 3 + 1], vb[i0 * 3 + 2]};
        const float p1[3] = {vb[i1 * 3], vb[i1 * 3 + 1], vb[i1 * 3 + 2]};
        const float p2[3] = {vb[i2 * 3], vb[i2 * 3 + 1], vb[i2 * 3 + 2]};
        float o[20];
        ex_face_frame(p0, p1, p2, o);
        float4* dst = ff + (size_t)i * 5;
#pragma unroll
        for (int k = 0; k < 5; k++) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
}

// ------------------------------------------------------------------------------------- U5+U6
// One thread per (segment, Gaussian).  A block owns a range of `per_block` consecutive (segment, Gaussian) pairs:
//   unfused  grid = (blocks per segment, S): the range lies inside segment blockIdx.y;
//   fused    a 1-D grid of ONE resident wave over the flat S*N index space, so every block does the same amount
//            of work (a grid of whole segments' pieces left the last wave a quarter full: +50 % on the kernel).  A
//            range that crosses a segment boundary is walked piece by piece, counters flushed at the boundary.
//
// FUSED (the session's form): the kernel also accumulates what the binning needs from the values it has in
// registers anyway — the four digit histograms of the depth key (onesweep's upfront histogram) and the per-tile
// Gaussian counts (input of the tile-range scan) — in shared-memory counters flushed once per block.  That removes
// two kernels that re-read S*N keys and S*N (tiles_touched, P0) records (rs_histogram, tile_count: ~90 us per
// 60-frame batch) at the price of ~7.6 shared-memory atomics per Gaussian here.
#ifndef OMFS_BIND_THREADS
#define OMFS_BIND_THREADS 128  // 56 registers; small CTAs also fit beside the persistent compositing warps of the previous batch
#endif
constexpr int kBindThreads = OMFS_BIND_THREADS;
template <bool FUSED>
__global__ void __launch_bounds__(kBindThreads) bind_preprocess_kernel(
    int N, int F, int width, int height, const float4* __restrict__ ff, const int32_t* __restrict__ seg_frame,
    const float* __restrict__ cams, const float4* __restrict__ xyzb, const float4* __restrict__ scale_lo,
    const float4* __restrict__ rot, const float4* __restrict__ sh, float4* __restrict__ P0,
    float4* __restrict__ P1, float4* __restrict__ P2, uint32_t* __restrict__ tiles_touched,
    uint32_t* __restrict__ depth_keys, int S, int per_block, uint32_t* __restrict__ hist_depth /*[S][4][256]*/,
    uint32_t* __restrict__ tile_cnt /*[S][tiles]*/, int tiles) {
    __shared__ float s_cam[kCam];
    extern __shared__ uint32_t s_fused[];   // FUSED: [4][256] depth-digit counters, then the [(gy+1)][(gx+1)] corner grid of the tile counts
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31;
    long long flat_lo, flat_hi;
    if (FUSED) {
        flat_lo = (long long)blockIdx.x * per_block;
        flat_hi = min(flat_lo + per_block, (long long)S * N);
    } else {
        flat_lo = (long long)blockIdx.y * N + (long long)blockIdx.x * per_block;
        flat_hi = min(flat_lo + per_block, (long long)(blockIdx.y + 1) * N);
    }
    while (flat_lo < flat_hi) {   // one piece per segment the range touches (uniform over the block)
    const int seg = (int)(flat_lo / N);
    const int n_begin = (int)(flat_lo - (long long)seg * N);
    const int n_end = (int)min((long long)N, flat_hi - (long long)seg * N);
    flat_lo += n_end - n_begin;
    __syncthreads();   // the previous piece's counters are flushed, its camera no longer read
    if (threadIdx.x < kCam) s_cam[threadIdx.x] = __ldg(cams + (size_t)seg * kCam + threadIdx.x);
    if (FUSED)
        for (int i = threadIdx.x; i < 1024 + (gx + 1) * (gy + 1); i += blockDim.x) s_fused[i] = 0;
    __syncthreads();
    const int frame = __ldg(seg_frame + seg);

    for (int base = n_begin; base < n_end; base += blockDim.x) {
        const int n = base + threadIdx.x;
        const bool valid = n < n_end;
        bool ok = false;
        BindPre o;
        o.px = o.py = o.depth = 0.f;
        o.radius = 0;
        if (valid) {
            const float4 a = ldg4(xyzb + n);
            const float4 s = ldg4(scale_lo + n);
            const float4 q = ldg4(rot + n);
            const int b = __float_as_int(a.w);
            const float4* fr = ff + ((size_t)frame * F + b) * 5;
            float frec[20];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float4 v = ldg4(fr + k);
                frec[4 * k] = v.x;
                frec[4 * k + 1] = v.y;
                frec[4 * k + 2] = v.z;
                frec[4 * k + 3] = v.w;
            }
            ok = ex_bind_project(frec, a.x, a.y, a.z, s.x, s.y, s.z, q.x, q.y, q.z, q.w, s_cam, width, height, gx, gy, o);
            const size_t oi = (size_t)seg * N + n;
            if (!ok) {
                P0[oi] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
                P1[oi] = make_float4(0.f, 0.f, 0.f, 0.f);
                P2[oi] = make_float4(0.f, 0.f, 0.f, 0.f);
                tiles_touched[oi] = 0;
                if (depth_keys) depth_keys[oi] = 0u;
            } else {
                float dx, dy, dz;
                ex_view_dir(o.mx, o.my, o.mz, s_cam, dx, dy, dz);
                float bs[16];
                ex_sh_basis(dx, dy, dz, bs);
                // 12 float4 planes; flat index k*3+c lives in plane (flat>>2), lane (flat&3)
                float coef[48];
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    const float4 v = ldg4(sh + (size_t)j * N + n);
                    coef[4 * j] = v.x;
                    coef[4 * j + 1] = v.y;
                    coef[4 * j + 2] = v.z;
                    coef[4 * j + 3] = v.w;
                }
                float rgb[3];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float acc = bs[0] * coef[c];
#pragma unroll
                    for (int k = 1; k < 16; k++) acc = acc + bs[k] * coef[k * 3 + c];
                    acc = acc + 0.5f;
                    rgb[c] = fmaxf(acc, 0.0f);
                }
                P0[oi] = make_float4(o.px, o.py, pack_cull_extents(o.ca, o.cb, o.cc, s.w), __int_as_float(o.radius));
                P1[oi] = make_float4(o.ca, o.cb, o.cc, s.w);
                P2[oi] = make_float4(rgb[0], rgb[1], rgb[2], o.depth);
                tiles_touched[oi] = o.tiles;
                if (depth_keys) depth_keys[oi] = __float_as_uint(o.depth);  // sort key of the depth sort (binning.cu)
            }
        }
        if (FUSED) {
            // depth-key digits: bytes 0..2 straight to the counters; the exponent byte is almost always the same for a
            // warp's 32 Gaussians and is then counted with one add
            const uint32_t key = ok ? __float_as_uint(o.depth) : 0u;
            if (valid) {
                atomicAdd(&s_fused[key & 0xffu], 1u);
                atomicAdd(&s_fused[256 + ((key >> 8) & 0xffu)], 1u);
                atomicAdd(&s_fused[512 + ((key >> 16) & 0xffu)], 1u);
            }
            const uint32_t top = key >> 24;
            const uint32_t top0 = __shfl_sync(0xffffffffu, top, 0);
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            if (__all_sync(0xffffffffu, !valid || top == top0)) {
                if (lane == 0 && vmask) atomicAdd(&s_fused[768 + top0], (uint32_t)__popc(vmask));
            } else if (valid) {
                atomicAdd(&s_fused[768 + top], 1u);
            }
            if (ok) {
                // the rectangle ex_bind_project counted the tiles of, as the four corners of a 2-D difference array
                // on the (gx + 1) x (gy + 1) corner grid (counters wrap modulo 2^32; the prefix sums at the flush
                // bring them back).  Four atomics per Gaussian whatever the rectangle: the per-tile loop it replaces
                // ran as long as the largest rectangle of the warp (13 % of the kernel's issue slots).
                uint32_t* d2 = s_fused + 1024;
                const int gw = gx + 1;
                atomicAdd(d2 + o.miny * gw + o.minx, 1u);
                atomicAdd(d2 + o.miny * gw + o.maxx, 0xffffffffu);
                atomicAdd(d2 + o.maxy * gw + o.minx, 0xffffffffu);
                atomicAdd(d2 + o.maxy * gw + o.maxx, 1u);
            }
        }
    }
    if (FUSED) {
        __syncthreads();
        uint32_t* h = hist_depth + (size_t)seg * 1024;
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            const uint32_t v = s_fused[i];
            if (v) atomicAdd(h + i, v);
        }
        // difference array -> counts: inclusive prefix sums along every row, then along every column (a thread per
        // row / column; the row stride gx + 1 is odd for the usual sizes, so the row pass is free of bank conflicts)
        uint32_t* d2 = s_fused + 1024;
        const int gw = gx + 1, gh = gy + 1;
        for (int y = threadIdx.x; y < gh; y += blockDim.x) {
            uint32_t run = 0;
            for (int x = 0; x < gw; x++) {
                run += d2[y * gw + x];
                d2[y * gw + x] = run;
            }
        }
        __syncthreads();
        for (int x = threadIdx.x; x < gw; x += blockDim.x) {
            uint32_t run = 0;
            for (int y = 0; y < gh; y++) {
                run += d2[y * gw + x];
                d2[y * gw + x] = run;
            }
        }
        __syncthreads();
        uint32_t* c = tile_cnt + (size_t)seg * tiles;
        for (int i = threadIdx.x; i < tiles; i += blockDim.x) {
            const int ty = i / gx, tx = i - ty * gx;
            const uint32_t v = d2[ty * gw + tx];
            if (v) atomicAdd(c + i, v);
        }
    }
    }  // pieces
}

int bind_preprocess_launch(int S, int N, int F, int width, int height, const float* d_ff, const int32_t* d_seg_frame,
                           const float* d_cams, const float* d_xyzb, const float* d_scale_lo, const float* d_rot,
                           const float* d_sh, float* d_P0, float* d_P1, float* d_P2, uint32_t* d_tiles_touched,
                           uint32_t* d_depth_keys, uint32_t* d_hist_depth, uint32_t* d_tile_cnt, cudaStream_t stream) {
    const int tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    const bool fused = d_hist_depth && d_tile_cnt;
    if (!fused) {
        dim3 grid(ceil_div(N, kBindThreads), S);
        bind_preprocess_kernel<false><<<grid, kBindThreads, 0, stream>>>(
            N, F, width, height, (const float4*)d_ff, d_seg_frame, d_cams, (const float4*)d_xyzb, (const float4*)d_scale_lo,
            (const float4*)d_rot, (const float4*)d_sh, (float4*)d_P0, (float4*)d_P1, (float4*)d_P2, d_tiles_touched,
            d_depth_keys, S, kBindThreads, nullptr, nullptr, tiles);
    } else {
        const size_t corners = (size_t)((width + kTile - 1) / kTile + 1) * ((height + kTile - 1) / kTile + 1);
        const size_t smem = sizeof(uint32_t) * (1024 + corners);
        if (smem > 200 * 1024) {
            set_error("bind_preprocess: %d tiles per frame exceed the fused tile counters", tiles);
            return OMFS_ERR_INVALID;
        }
        static DeviceOnce once;
        int rc = ensure_dyn_smem(once, bind_preprocess_kernel<true>, (int)smem);
        if (rc) return rc;
        // one resident wave: as many blocks as fit the SMs (registers: 8 per SM; shared memory: the counters), each
        // walking an equal share of the S*N pairs, but never less than 8 rounds per block (the counter flush is
        // 1024 + tiles global atomics per piece)
        int per_sm = 8;
        const int by_smem = (int)((200 * 1024) / (smem + 1024));
        if (per_sm > by_smem) per_sm = by_smem;
        const long long total = (long long)S * N;
        long long per_block = (total + (long long)kNumSMs * per_sm - 1) / ((long long)kNumSMs * per_sm);
        per_block = (per_block + kBindThreads - 1) / kBindThreads * kBindThreads;
        if (per_block < 8 * kBindThreads) per_block = 8 * kBindThreads;
        dim3 grid(ceil_div(total, per_block), 1);
        bind_preprocess_kernel<true><<<grid, kBindThreads, smem, stream>>>(
            N, F, width, height, (const float4*)d_ff, d_seg_frame, d_cams, (const float4*)d_xyzb, (const float4*)d_scale_lo,
            (const float4*)d_rot, (const float4*)d_sh, (float4*)d_P0, (float4*)d_P1, (float4*)d_P2, d_tiles_touched,
            d_depth_keys, S, (int)per_block, d_hist_depth, d_tile_cnt, tiles);
    }
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

}  // namespace omfs

using namespace omfs;

extern "C" int omfs_face_frames(int T, int V, int F, const float* d_verts, const int32_t* d_faces, float* d_ff,
                                void* stream) {
    OMFS_REQUIRE(T >= 0 && V > 0 && F > 0, "bad sizes");
    OMFS_REQUIRE(d_verts && d_faces && d_ff, "null pointer");
    if (T == 0) return OMFS_OK;
    const long long total = (long long)T * F;
    int blocks = ceil_div(total, 256);
    const int cap = kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    face_frames_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(T, V, F, d_verts, d_faces, (float4*)d_ff);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

extern "C" int omfs_bind_preprocess(int S, int N, int F, int width, int height, const float* d_ff,
                                    const int32_t* d_seg_frame, const float* d_cams, const float* d_xyzb,
                                    const float* d_scale_lo, const float* d_rot, const float* d_sh, float* d_P0,
                                    float* d_P1, float* d_P2, uint32_t* d_tiles_touched, uint32_t* d_depth_keys,
                                    void* stream) {
    OMFS_REQUIRE(S >= 0 && N > 0 && F > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(S <= 65535, "at most 65535 segments per call");
    OMFS_REQUIRE(d_ff && d_seg_frame && d_cams && d_xyzb && d_scale_lo && d_rot && d_sh, "null input");
    OMFS_REQUIRE(d_P0 && d_P1 && d_P2 && d_tiles_touched, "null output");
    if (S == 0) return OMFS_OK;
    return bind_preprocess_launch(S, N, F, width, height, d_ff, d_seg_frame, d_cams, d_xyzb, d_scale_lo, d_rot, d_sh, d_P0,
                                  d_P1, d_P2, d_tiles_touched, d_depth_keys, nullptr, nullptr, (cudaStream_t)stream);
}
