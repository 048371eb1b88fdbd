// binning.cu — U7 (scan + key emission), U8 (onesweep LSD radix sort), U9 (tile ranges).
//
// All integer work; results are bit-exact against the oracle by construction:
//   key   = (seg*tiles + tile) << 32 | float bits of the view-space depth   (uint64)
//   value = Gaussian index inside its segment                               (uint32)
// Pairs are emitted in (segment, Gaussian) order, the sort is a stable LSD radix sort over the low
// `sort_bits` bits (8-bit digits), so equal (tile, depth) keys stay in Gaussian order.
//
// The whole batch of S segments is ONE sort problem: the segment id sits in the key's high bits,
// which costs no extra pass while S*tiles <= 2^16 (48-bit keys = 6 passes, the same 6 a single
// 512^2 frame needs for its 42 bits).  The pair count is only known on the device; every kernel
// here reads it from device memory and sizes its own work (persistent grids), so the host never
// synchronises inside a batch and the batch can be replayed as a CUDA graph.
//
// Onesweep (Adinets & Merrill 2022) per pass: one upfront kernel builds the digit histograms of ALL
// passes from a single read of the keys; then each pass is a single kernel that ranks a tile of
// keys with warp-ballot match operations, resolves its global offsets by decoupled look-back over
// the tiles before it, and scatters through shared memory so that global writes are contiguous
// per digit run.  Traffic: 8 B/key for the histogram read + 24 B/pair per pass.
#include "common.cuh"
#include "exact_math.cuh"

namespace omfs {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

constexpr int kRsThreads = 256;                   // 8 warps, one thread per digit
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                      // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;    // 4096 keys per tile
constexpr int kRadix = 256;
constexpr int kMaxPasses = 8;

constexpr uint32_t kFlagAgg = 1u << 30;     // tile aggregate available
constexpr uint32_t kFlagPrefix = 2u << 30;  // inclusive prefix available
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;

// ---------------------------------------------------------------------------------- scan
// three small kernels: per-tile sums, scan of the sums (one CTA), per-tile scan + carry.
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix and
// the block total through `total`
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wp = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t x = s_warp[w];
        if (w < warp) wp += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return wp + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(long long count,
                                                                      const uint32_t* __restrict__ in,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_warp[8];
    const long long base = (long long)blockIdx.x * kScanTile;
    uint32_t acc = 0;
    // thread t owns items [t*16, t*16+16): four 16-byte loads
    const long long first = base + (long long)threadIdx.x * kScanItems;
    if (first + kScanItems <= count) {
        const uint4* p = reinterpret_cast<const uint4*>(in + first);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 v = __ldg(p + k);
            acc += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int k = 0; k < kScanItems; k++)
            if (first + k < count) acc += __ldg(in + first + k);
    }
    uint32_t total;
    block_excl_scan_256(acc, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of the tile sums in place; writes the grand total (clamped bookkeeping
// is done by the caller kernels) to *num_pairs.
__global__ void __launch_bounds__(1024) scan_sums_kernel(int n_tiles, uint32_t* __restrict__ tile_sums,
                                                         uint32_t* __restrict__ num_pairs,
                                                         unsigned long long capacity, int* __restrict__ status_flag,
                                                         uint32_t* __restrict__ sort_count,
                                                         unsigned long long* __restrict__ pair_accum) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? tile_sums[i] : 0u;
        const uint32_t inc = warp_incl_scan(v);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            const uint32_t winc = warp_incl_scan(w);
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + inc - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t total = s_carry;
        *num_pairs = total;
        if (pair_accum) *pair_accum += total;  // running total over the batches of one render call
        if ((unsigned long long)total > capacity) {
            // overflow: flag it and sort nothing (emit_keys skips out-of-range Gaussians, so a
            // partial list would contain unwritten keys); the caller re-runs with more capacity
            *status_flag = 1;
            *sort_count = 0;
        } else {
            *sort_count = total;
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(long long count, const uint32_t* __restrict__ in,
                                                                  const uint32_t* __restrict__ tile_sums,
                                                                  uint32_t* __restrict__ out) {
    __shared__ uint32_t s_warp[8];
    const long long base = (long long)blockIdx.x * kScanTile;
    const long long first = base + (long long)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    if (first + kScanItems <= count) {
        const uint4* p = reinterpret_cast<const uint4*>(in + first);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 q = __ldg(p + k);
            v[4 * k] = q.x;
            v[4 * k + 1] = q.y;
            v[4 * k + 2] = q.z;
            v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v[k] = (first + k < count) ? __ldg(in + first + k) : 0u;
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) acc += v[k];
    uint32_t total;
    uint32_t run = block_excl_scan_256(acc, s_warp, total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        run += v[k];
        v[k] = run;  // inclusive
    }
    if (first + kScanItems <= count) {
        uint4* p = reinterpret_cast<uint4*>(out + first);
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++)
            if (first + k < count) out[first + k] = v[k];
    }
}

// ---------------------------------------------------------------------------------- key emission
// grid = (ceil(N/256), S).  Each Gaussian writes tiles_touched consecutive pairs.
__global__ void __launch_bounds__(256) emit_keys_kernel(int N, int width, int height,
                                                        const float4* __restrict__ P0,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const uint32_t* __restrict__ offsets,
                                                        unsigned long long capacity, uint64_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals) {
    const int seg = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const size_t i = (size_t)seg * N + n;
    const uint32_t tt = __ldg(tiles_touched + i);
    if (tt == 0) return;
    const uint32_t end = __ldg(offsets + i);  // inclusive scan
    uint32_t off = end - tt;
    if ((unsigned long long)end > capacity) return;  // overflow: flagged by scan_sums_kernel
    const float4 p = ldg4(P0 + i);
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
    int minx, miny, maxx, maxy;
    ex_tile_rect(p.x, p.y, __float_as_int(p.w), gx, gy, minx, miny, maxx, maxy);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(p.z);
    const uint64_t seg_base = (uint64_t)seg * (uint64_t)(gx * gy);
    for (int y = miny; y < maxy; y++)
        for (int x = minx; x < maxx; x++) {
            const uint64_t tile = seg_base + (uint64_t)(y * gx + x);
            keys[off] = (tile << 32) | depth_bits;
            vals[off] = (uint32_t)n;
            off++;
        }
}

// ---------------------------------------------------------------------------------- histograms
// hist[pass][256] for every pass from ONE read of the keys.  Warp-ballot aggregation: lanes whose
// digits match elect a leader that adds the whole group with a single shared-memory atomic, which
// collapses the heavily repeated high digits (tile id, depth exponent) to one atomic per warp.
__global__ void __launch_bounds__(256) rs_histogram_kernel(const uint64_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ sort_count, int passes,
                                                           uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t count = *sort_count;
    const int lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    // warp-granular grid-stride loop so that every lane of a warp is active for the match
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp_global * 32u; base < count; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const uint64_t k = valid ? __ldg(keys + i) : 0ull;
        for (int p = 0; p < passes; p++) {
            const uint32_t d = valid ? (uint32_t)((k >> (8 * p)) & 0xffu) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && (peers & lanemask_lt) == 0u) atomicAdd(&s_hist[p * kRadix + d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// exclusive scan of each pass's 256 bins: hist -> digit start offsets.  grid = passes, block = 256.
__global__ void __launch_bounds__(256) rs_scan_hist_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_warp[8];
    uint32_t* h = hist + blockIdx.x * kRadix;
    const uint32_t v = h[threadIdx.x];
    uint32_t total;
    const uint32_t excl = block_excl_scan_256(v, s_warp, total);
    h[threadIdx.x] = excl;
}

// ---------------------------------------------------------------------------------- onesweep pass
struct RsSmem {
    union {
        uint32_t warp_hist[kRsWarps][kRadix];  // 8 KB  (ranking phase)
        uint64_t keys[kRsTile];                // 32 KB (scatter phase)
    };
    uint32_t vals[kRsTile];          // 16 KB
    uint32_t tile_start[kRadix];     // exclusive start of each digit inside the tile
    uint32_t global_base[kRadix];    // digit d of this tile starts at global_base[d] in the output
    uint32_t scan_warp[8];
    uint32_t tile_id;
};

__global__ void __launch_bounds__(kRsThreads) rs_onesweep_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ sort_count, int shift,
    const uint32_t* __restrict__ digit_start /*[256] this pass*/, uint32_t* __restrict__ tile_counter,
    volatile uint32_t* __restrict__ status /*[max_tiles][256] this pass*/) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_smem_raw);
    const uint32_t count = *sort_count;
    const uint32_t n_tiles = (count + kRsTile - 1) / kRsTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lanemask_lt = (1u << lane) - 1u;

    while (true) {
        // tiles are handed out in increasing order, so every predecessor a tile looks back at is
        // owned by a CTA that is already running: the look-back cannot deadlock
        if (threadIdx.x == 0) sm.tile_id = atomicAdd(tile_counter, 1u);
        for (int i = threadIdx.x; i < kRsWarps * kRadix; i += kRsThreads) (&sm.warp_hist[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = sm.tile_id;
        if (tile >= n_tiles) break;
        const uint32_t tile_base = tile * kRsTile;
        // warp w owns keys [tile_base + w*512, +512); item i of lane l is element i*32 + l of that
        // chunk, so (warp, item, lane) order is memory order and the ranking below is stable
        const uint32_t warp_base = tile_base + warp * (kRsItems * 32);
        uint64_t key[kRsItems];
        uint32_t val[kRsItems];
        uint32_t rank[kRsItems];
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const bool valid = idx < count;
            key[i] = valid ? keys_in[idx] : ~0ull;
            val[i] = valid ? vals_in[idx] : 0u;
        }
        // ---- rank inside the warp with match_any; per-warp digit counters in shared memory
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const bool valid = idx < count;
            const uint32_t d = valid ? (uint32_t)((key[i] >> shift) & 0xffu) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            if (valid && lane == leader) {
                pre = sm.warp_hist[warp][d];
                sm.warp_hist[warp][d] = pre + __popc(peers);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rank[i] = pre + __popc(peers & lanemask_lt);
            __syncwarp();
        }
        __syncthreads();
        // ---- thread d: exclusive scan of digit d over the warps, tile count of digit d
        const int d = threadIdx.x;
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; w++) {
            const uint32_t c = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = cnt;  // becomes the warp's offset inside digit d
            cnt += c;
        }
        // publish the aggregate as early as possible
        volatile uint32_t* my_status = status + (size_t)tile * kRadix;
        if (tile == 0) {
            my_status[d] = kFlagPrefix | cnt;
        } else {
            my_status[d] = kFlagAgg | cnt;
        }
        // exclusive scan over digits -> where digit d starts inside the tile
        uint32_t tile_total;
        const uint32_t tstart = block_excl_scan_256(cnt, sm.scan_warp, tile_total);
        sm.tile_start[d] = tstart;
        // ---- decoupled look-back for digit d
        uint32_t excl = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            uint32_t spins = 0;
            while (true) {
                // a predecessor that never publishes would be a bug; fail loudly instead of hanging
                if (++spins > (1u << 28)) __trap();
                const uint32_t s = status[(size_t)look * kRadix + d];
                const uint32_t flag = s & kFlagMask;
                if (flag == kFlagPrefix) {
                    excl += s & kValMask;
                    break;
                }
                if (flag == kFlagAgg) {
                    excl += s & kValMask;
                    look--;
                    continue;
                }
                // not published yet: spin (the owner is running, see above)
            }
            my_status[d] = kFlagPrefix | ((excl + cnt) & kValMask);
        }
        sm.global_base[d] = digit_start[d] + excl;
        __syncthreads();
        // ---- local positions: tile_start[digit] + warp offset + rank
        uint32_t pos[kRsItems];
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const bool valid = idx < count;
            const uint32_t dg = (uint32_t)((key[i] >> shift) & 0xffu);
            pos[i] = valid ? (sm.tile_start[dg] + sm.warp_hist[warp][dg] + rank[i]) : 0xffffffffu;
        }
        __syncthreads();  // warp_hist is dead from here on: its storage becomes sm.keys
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            if (pos[i] != 0xffffffffu) {
                sm.keys[pos[i]] = key[i];
                sm.vals[pos[i]] = val[i];
            }
        }
        __syncthreads();
        // ---- scatter: consecutive threads write consecutive addresses inside each digit run
        const uint32_t in_tile = min((uint32_t)kRsTile, count - tile_base);
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t j = i * kRsThreads + threadIdx.x;
            if (j < in_tile) {
                const uint64_t k = sm.keys[j];
                const uint32_t dg = (uint32_t)((k >> shift) & 0xffu);
                const uint32_t dst = sm.global_base[dg] + (j - sm.tile_start[dg]);
                keys_out[dst] = k;
                vals_out[dst] = sm.vals[j];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- tile ranges
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint64_t* __restrict__ sorted_keys,
                                                          const uint32_t* __restrict__ sort_count,
                                                          uint32_t* __restrict__ ranges) {
    const uint32_t count = *sort_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t cur = (uint32_t)(__ldg(sorted_keys + i) >> 32);
        if (i == 0) {
            ranges[2 * cur] = 0;
        } else {
            const uint32_t prev = (uint32_t)(__ldg(sorted_keys + i - 1) >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = i;
                ranges[2 * cur] = i;
            }
        }
        if (i == count - 1) ranges[2 * cur + 1] = count;
    }
}

static inline int sort_bits_for(int S, int width, int height) {
    const long long tiles = (long long)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    long long total = (long long)S * tiles;
    if (total < 1) total = 1;
    int b = 0;
    while ((1ll << b) < total) b++;
    if (b < 1) b = 1;
    return 32 + b;
}

struct BinningWs {
    uint32_t* tile_sums;     // scan tiles
    uint32_t* hist;          // [kMaxPasses][256]
    uint32_t* tile_counter;  // [kMaxPasses]
    uint32_t* sort_count;    // [1]
    uint32_t* status;        // [passes][max_rs_tiles][256]
    size_t zero_bytes;       // prefix of the workspace that must be zeroed per call
    size_t status_stride;    // elements per pass
    size_t total;
};

static BinningWs carve(void* base, int S, int N, int width, int height, size_t capacity) {
    BinningWs w{};
    const long long count = (long long)S * N;
    const size_t n_scan_tiles = (size_t)((count + kScanTile - 1) / kScanTile) + 1;
    const int passes = (sort_bits_for(S, width, height) + 7) / 8;
    const size_t max_rs_tiles = (capacity + kRsTile - 1) / kRsTile + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? (unsigned char*)base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    // zeroed region first
    w.hist = (uint32_t*)take(sizeof(uint32_t) * kMaxPasses * kRadix);
    w.tile_counter = (uint32_t*)take(sizeof(uint32_t) * kMaxPasses);
    w.sort_count = (uint32_t*)take(sizeof(uint32_t) * 4);
    w.status_stride = max_rs_tiles * kRadix;
    w.status = (uint32_t*)take(sizeof(uint32_t) * w.status_stride * passes);
    w.zero_bytes = off;
    w.tile_sums = (uint32_t*)take(sizeof(uint32_t) * n_scan_tiles);
    w.total = off;
    return w;
}

}  // namespace omfs

using namespace omfs;

extern "C" size_t omfs_binning_workspace_bytes(int S, int N, int width, int height, size_t capacity) {
    if (S <= 0 || N <= 0 || width <= 0 || height <= 0) return 0;
    return carve(nullptr, S, N, width, height, capacity).total;
}

extern "C" int omfs_binning_sort_bits(int S, int width, int height) { return sort_bits_for(S, width, height); }

namespace omfs {

int binning_scan_emit(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                      const uint32_t* d_tiles_touched, uint32_t* d_offsets, uint64_t* d_keys, uint32_t* d_vals,
                      uint32_t* d_ranges, uint32_t* d_num_pairs, int* d_status_flag,
                      unsigned long long* d_pair_accum, void* d_workspace, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    const long long count = (long long)S * N;
    const int n_scan_tiles = ceil_div(count, kScanTile);
    const long long tiles_total = (long long)S * ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    OMFS_CUDA(cudaMemsetAsync(d_workspace, 0, w.zero_bytes, stream));
    if (d_ranges) OMFS_CUDA(cudaMemsetAsync(d_ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles_total, stream));
    scan_tile_sums_kernel<<<n_scan_tiles, kScanThreads, 0, stream>>>(count, d_tiles_touched, w.tile_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(n_scan_tiles, w.tile_sums, d_num_pairs, (unsigned long long)capacity,
                                             d_status_flag, w.sort_count, d_pair_accum);
    scan_apply_kernel<<<n_scan_tiles, kScanThreads, 0, stream>>>(count, d_tiles_touched, w.tile_sums, d_offsets);
    dim3 egrid(ceil_div(N, 256), S);
    emit_keys_kernel<<<egrid, 256, 0, stream>>>(N, width, height, (const float4*)d_P0, d_tiles_touched, d_offsets,
                                                (unsigned long long)capacity, d_keys, d_vals);
    count_launch(4);
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

// returns through *out_index which of the two buffer pairs holds the sorted result
int binning_sort(int S, int N, int width, int height, size_t capacity, uint64_t* d_keys0, uint64_t* d_keys1,
                 uint32_t* d_vals0, uint32_t* d_vals1, void* d_workspace, int* out_index, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    const int passes = (sort_bits_for(S, width, height) + 7) / 8;
    const int persistent = kNumSMs * 4;
    rs_histogram_kernel<<<persistent, 256, 0, stream>>>(d_keys0, w.sort_count, passes, w.hist);
    rs_scan_hist_kernel<<<passes, 256, 0, stream>>>(w.hist);
    count_launch(2);
    OMFS_LAUNCH_CHECK();
    static bool attr_set = false;
    if (!attr_set) {
        OMFS_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(RsSmem)));
        attr_set = true;
    }
    uint64_t* kin = d_keys0;
    uint64_t* kout = d_keys1;
    uint32_t* vin = d_vals0;
    uint32_t* vout = d_vals1;
    for (int p = 0; p < passes; p++) {
        rs_onesweep_kernel<<<persistent, kRsThreads, sizeof(RsSmem), stream>>>(
            kin, vin, kout, vout, w.sort_count, 8 * p, w.hist + p * kRadix, w.tile_counter + p,
            w.status + (size_t)p * w.status_stride);
        count_launch();
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    OMFS_LAUNCH_CHECK();
    *out_index = (kin == d_keys0) ? 0 : 1;  // after the final swap kin/vin hold the sorted pairs
    return OMFS_OK;
}

int binning_ranges(int S, int N, int width, int height, size_t capacity, const uint64_t* d_sorted_keys,
                   uint32_t* d_ranges, void* d_workspace, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    tile_ranges_kernel<<<kNumSMs * 4, 256, 0, stream>>>(d_sorted_keys, w.sort_count, d_ranges);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

}  // namespace omfs

extern "C" int omfs_binning(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                            const uint32_t* d_tiles_touched, uint32_t* d_offsets, uint64_t* d_keys0,
                            uint64_t* d_keys1, uint32_t* d_vals0, uint32_t* d_vals1, uint32_t* d_ranges,
                            uint32_t* d_num_pairs, int* d_status_flag, void* d_workspace, size_t workspace_bytes,
                            int* h_out_buffer_index, void* stream_) {
    OMFS_REQUIRE(S > 0 && N > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(S <= 65535, "at most 65535 segments per call");
    OMFS_REQUIRE(capacity > 0 && capacity < (1ull << 30), "capacity must be in (0, 2^30)");
    OMFS_REQUIRE((long long)S * N < (1ll << 31), "S*N must be below 2^31");
    OMFS_REQUIRE(d_P0 && d_tiles_touched && d_offsets && d_keys0 && d_keys1 && d_vals0 && d_vals1 && d_ranges &&
                     d_num_pairs && d_status_flag && d_workspace,
                 "null pointer");
    const int passes = (sort_bits_for(S, width, height) + 7) / 8;
    OMFS_REQUIRE(passes <= kMaxPasses, "too many sort passes");
    OMFS_REQUIRE(workspace_bytes >= carve(nullptr, S, N, width, height, capacity).total,
                 "workspace too small (omfs_binning_workspace_bytes)");
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = binning_scan_emit(S, N, width, height, capacity, d_P0, d_tiles_touched, d_offsets, d_keys0, d_vals0,
                               d_ranges, d_num_pairs, d_status_flag, nullptr, d_workspace, stream);
    if (rc) return rc;
    int idx = 0;
    rc = binning_sort(S, N, width, height, capacity, d_keys0, d_keys1, d_vals0, d_vals1, d_workspace, &idx, stream);
    if (rc) return rc;
    rc = binning_ranges(S, N, width, height, capacity, idx ? d_keys1 : d_keys0, d_ranges, d_workspace, stream);
    if (rc) return rc;
    if (h_out_buffer_index) *h_out_buffer_index = idx;
    return OMFS_OK;
}

// U7 alone (scan + emit), for the key-emission parity test: unsorted pairs land in d_keys/d_vals.
extern "C" int omfs_scan_emit(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                              const uint32_t* d_tiles_touched, uint32_t* d_offsets, uint64_t* d_keys,
                              uint32_t* d_vals, uint32_t* d_num_pairs, int* d_status_flag, void* d_workspace,
                              size_t workspace_bytes, void* stream_) {
    OMFS_REQUIRE(S > 0 && N > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(capacity > 0 && capacity < (1ull << 30), "capacity must be in (0, 2^30)");
    OMFS_REQUIRE(d_P0 && d_tiles_touched && d_offsets && d_keys && d_vals && d_num_pairs && d_status_flag &&
                     d_workspace,
                 "null pointer");
    OMFS_REQUIRE(workspace_bytes >= carve(nullptr, S, N, width, height, capacity).total,
                 "workspace too small (omfs_binning_workspace_bytes)");
    return binning_scan_emit(S, N, width, height, capacity, d_P0, d_tiles_touched, d_offsets, d_keys, d_vals, nullptr,
                             d_num_pairs, d_status_flag, nullptr, d_workspace, (cudaStream_t)stream_);
}
