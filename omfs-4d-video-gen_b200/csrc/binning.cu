// binning.cu — U7 (scan + key emission), U8 (onesweep radix sort), U9 (tile ranges).
//
// Contract (bit-exact against the oracle): the pair list of a batch, ordered by
//     key = (seg*tiles + tile) << 32 | float bits of the view-space depth,   ties by Gaussian index,
// with value = Gaussian index inside its segment, plus the per-tile [first,last) ranges.
//
// How it is produced.  A single 6-pass sort of 12-byte pairs (R ~ 3.6 N of them) is what the
// published pipeline does; on B200 that is the largest HBM consumer of the frame.  The same order
// comes out of two much smaller sorts, because an LSD radix sort is stable:
//   1. DEPTH SORT, segmented: the N Gaussians of every segment are sorted by depth bits
//      (uint32 key, value = Gaussian index; 4 passes over S*N 8-byte items).  Index order is the
//      tie-break because the first pass starts from the identity permutation.
//   2. scan of tiles_touched in that order, then EMISSION in depth order: pair = (global tile id,
//      Gaussian index).  The depth bits are never written: inside a tile, emission order already IS
//      depth order.
//   3. TILE SORT: one stable sort of the pairs by global tile id only (uint32 key, 2 passes while
//      S*tiles <= 2^16).
// Traffic per pair drops from (8 + 24*6) = 152 B to about 4 + 16*2 = 36 B, plus 4 + 16*4 B per
// Gaussian; all keys are 32-bit, which also halves the ranking registers.
// The 64-bit keys exist only on request (rebuild_keys_kernel) for the parity tests / debug taps.
//
// Both sorts are the same onesweep kernel (Adinets & Merrill 2022): one upfront histogram kernel per
// sort, then ONE kernel per pass that ranks a 4096-key tile with warp ballots (8 VOTEs per key give
// the lanes holding the same digit — MATCH.ANY retires at ~1 per 22 clk per SM on sm_100 and was the
// top stall of the first version), resolves global offsets by decoupled look-back, and scatters
// through shared memory so global writes are contiguous per digit run.  Tiles are handed out by an
// atomic counter, so a tile's predecessors are always already running (no deadlock); segments are
// independent look-back chains.  Counts live on the device: nothing here synchronises with the host.
#include "common.cuh"
#include "exact_math.cuh"

namespace omfs {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

constexpr int kRsThreads = 256;  // 8 warps, one thread per digit
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                    // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per tile
constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;

constexpr uint32_t kFlagAgg = 1u << 30;     // tile aggregate available
constexpr uint32_t kFlagPrefix = 2u << 30;  // inclusive prefix available
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;

// ---------------------------------------------------------------------------------- scan helpers
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads)
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

// per-tile sums of tiles_touched taken in DEPTH order (perm holds indices inside the segment)
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(long long count, int N,
                                                                      const uint32_t* __restrict__ tt,
                                                                      const uint32_t* __restrict__ perm,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_warp[8];
    const long long first = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (first + k < count) acc += __ldg(tt + ((first + k) / N) * N + __ldg(perm + first + k));
    uint32_t total;
    block_excl_scan_256(acc, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of the tile sums in place; pair count, overflow flag, running total
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
        const uint32_t excl = s_carry + s_warp[warp] + inc - v;
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
            // overflow: flag it and sort nothing; the caller re-runs with more capacity
            *status_flag = 1;
            *sort_count = 0;
        } else {
            *sort_count = total;
        }
    }
}

// offsets in depth order + key emission, fused: thread t owns 16 consecutive sorted Gaussians, so its
// pairs are one contiguous run of the output
__global__ void __launch_bounds__(kScanThreads) scan_emit_kernel(
    long long count, int N, int width, int height, const uint32_t* __restrict__ tt,
    const uint32_t* __restrict__ perm, const uint32_t* __restrict__ tile_sums, const float4* __restrict__ P0,
    const uint32_t* __restrict__ sort_count, uint32_t* __restrict__ tile_ids, uint32_t* __restrict__ vals) {
    __shared__ uint32_t s_warp[8];
    const long long first = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    uint32_t g[kScanItems], t[kScanItems];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        g[k] = 0;
        t[k] = 0;
        if (first + k < count) {
            g[k] = __ldg(perm + first + k);
            t[k] = __ldg(tt + ((first + k) / N) * N + g[k]);
        }
        acc += t[k];
    }
    uint32_t total;
    uint32_t off = block_excl_scan_256(acc, s_warp, total) + tile_sums[blockIdx.x];
    if (*sort_count == 0) return;  // empty batch or overflow (flagged by scan_sums_kernel)
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (t[k] == 0) continue;
        const long long seg = (first + k) / N;
        const float4 p = ldg4(P0 + seg * N + g[k]);
        int minx, miny, maxx, maxy;
        ex_tile_rect(p.x, p.y, __float_as_int(p.w), gx, gy, minx, miny, maxx, maxy);
        const uint32_t seg_base = (uint32_t)seg * (uint32_t)(gx * gy);
        for (int y = miny; y < maxy; y++)
            for (int x = minx; x < maxx; x++) {
                tile_ids[off] = seg_base + (uint32_t)(y * gx + x);
                vals[off] = g[k];
                off++;
            }
    }
}

// ---------------------------------------------------------------------------------- histograms
// hist[seg][pass][256] from ONE read of the keys.  Digits listed in uniform_mask are almost always
// identical across a warp's 32 consecutive keys (depth exponent byte, high tile byte): they are counted
// with one ballot-checked add per warp; the rest go straight to shared-memory atomics.
// grid = (blocks per segment, n_seg).  count_ptr != NULL: one segment whose length is on the device.
__global__ void __launch_bounds__(256) rs_histogram_kernel(const uint32_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ count_ptr, uint32_t seg_len,
                                                           int passes, uint32_t uniform_mask,
                                                           uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t count = count_ptr ? *count_ptr : seg_len;
    const uint32_t seg = blockIdx.y;
    const uint32_t* k_seg = keys + (size_t)seg * seg_len;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_in_grid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp_in_grid * 32u; base < count; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const uint32_t k = valid ? __ldg(k_seg + i) : 0u;
        const uint32_t n_valid = min(32u, count - base);
        for (int p = 0; p < passes; p++) {
            const uint32_t d = (k >> (8 * p)) & 0xffu;
            if ((uniform_mask >> p) & 1u) {
                const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
                if (__all_sync(0xffffffffu, !valid || d == d0)) {
                    if (lane == 0) atomicAdd(&s_hist[p * kRadix + d0], n_valid);
                    continue;
                }
            }
            if (valid) atomicAdd(&s_hist[p * kRadix + d], 1u);
        }
    }
    __syncthreads();
    uint32_t* h = hist + (size_t)seg * passes * kRadix;
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(h + i, v);
    }
}

// exclusive scan of each (segment, pass) histogram: counts -> digit start offsets.  grid = n_seg*passes
__global__ void __launch_bounds__(256) rs_scan_hist_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_warp[8];
    uint32_t* h = hist + (size_t)blockIdx.x * kRadix;
    const uint32_t v = h[threadIdx.x];
    uint32_t total;
    h[threadIdx.x] = block_excl_scan_256(v, s_warp, total);
}

// ---------------------------------------------------------------------------------- onesweep pass
struct RsSmem {
    uint32_t keys[kRsTile];                 // 16 KB  tile keys in digit order (scatter staging)
    uint32_t vals[kRsTile];                 // 16 KB
    uint32_t warp_hist[kRsWarps][kRadix];   // 8 KB   per-warp digit counters, then warp offsets
    uint32_t tile_start[kRadix];            // exclusive start of each digit inside the tile
    uint32_t global_base[kRadix];           // digit d of this tile starts here in the output
    uint32_t scan_warp[8];
    uint32_t tile_id;
};

// lanes of the warp whose 8-bit digit equals mine, from 8 ballots (one per digit bit)
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, uint32_t valid_mask) {
    uint32_t peers = valid_mask;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// One LSD pass over n_seg independent segments of seg_len keys each (count_ptr != NULL: a single
// segment whose length is read from the device).  vals_in == NULL: values are the identity
// permutation inside each segment (first pass of the depth sort).  digit_start is [n_seg][stride]
// with this pass's 256 offsets at the front of each row.
__global__ void __launch_bounds__(kRsThreads, 4) rs_onesweep_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ count_ptr, uint32_t seg_len, uint32_t n_seg,
    int shift, const uint32_t* __restrict__ digit_start, uint32_t digit_stride,
    uint32_t* __restrict__ tile_counter, volatile uint32_t* __restrict__ status /*[tiles][256] this pass*/) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_smem_raw);
    if (count_ptr) seg_len = *count_ptr;
    const uint32_t tiles_per_seg = (seg_len + kRsTile - 1) / kRsTile;
    const uint32_t n_tiles = tiles_per_seg * n_seg;
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
        const uint32_t seg = tile / tiles_per_seg, ltile = tile - seg * tiles_per_seg;
        const size_t seg_off = (size_t)seg * seg_len;
        const uint32_t tile_base = ltile * kRsTile;  // inside the segment
        // warp w owns keys [tile_base + w*512, +512); item i of lane l is element i*32 + l of that
        // chunk, so (warp, item, lane) order is memory order and the ranking below is stable
        const uint32_t warp_base = tile_base + warp * (kRsItems * 32);
        uint32_t key[kRsItems];
        uint16_t rank[kRsItems];
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            key[i] = (idx < seg_len) ? keys_in[seg_off + idx] : 0xffffffffu;
        }
        // ---- rank inside the warp: ballot peers + per-warp digit counters in shared memory
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const bool valid = idx < seg_len;
            const uint32_t d = (key[i] >> shift) & 0xffu;
            const uint32_t peers = digit_peers(d, __ballot_sync(0xffffffffu, valid));
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            if (valid && lane == leader) {
                pre = sm.warp_hist[warp][d];
                sm.warp_hist[warp][d] = pre + __popc(peers);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader & 31);
            rank[i] = (uint16_t)(pre + __popc(peers & lanemask_lt));
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
        // publish the aggregate as early as possible; the first tile of a segment has no predecessor
        volatile uint32_t* my_status = status + (size_t)tile * kRadix;
        my_status[d] = (ltile == 0 ? kFlagPrefix : kFlagAgg) | cnt;
        uint32_t tile_total;
        const uint32_t tstart = block_excl_scan_256(cnt, sm.scan_warp, tile_total);
        sm.tile_start[d] = tstart;
        // ---- decoupled look-back for digit d, inside the segment
        uint32_t excl = 0;
        if (ltile > 0) {
            uint32_t look = tile - 1;
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
                }
                // else not published yet: spin (the owner is running, see above)
            }
            my_status[d] = kFlagPrefix | ((excl + cnt) & kValMask);
        }
        sm.global_base[d] = digit_start[(size_t)seg * digit_stride + d] + excl;
        __syncthreads();
        // ---- stage keys and values in digit order (values are only touched now: they were never
        // needed for ranking, so they do not occupy registers during it)
#pragma unroll
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t idx = warp_base + i * 32 + lane;
            if (idx < seg_len) {
                const uint32_t dg = (key[i] >> shift) & 0xffu;
                const uint32_t pos = sm.tile_start[dg] + sm.warp_hist[warp][dg] + rank[i];
                sm.keys[pos] = key[i];
                sm.vals[pos] = vals_in ? vals_in[seg_off + idx] : idx;
            }
        }
        __syncthreads();
        // ---- scatter: consecutive threads write consecutive addresses inside each digit run
        const uint32_t in_tile = min((uint32_t)kRsTile, seg_len - tile_base);
#pragma unroll 4
        for (int i = 0; i < kRsItems; i++) {
            const uint32_t j = i * kRsThreads + threadIdx.x;
            if (j < in_tile) {
                const uint32_t k = sm.keys[j];
                const uint32_t dg = (k >> shift) & 0xffu;
                const size_t dst = seg_off + sm.global_base[dg] + (j - sm.tile_start[dg]);
                keys_out[dst] = k;
                vals_out[dst] = sm.vals[j];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- tile ranges
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ sorted_tiles,
                                                          const uint32_t* __restrict__ sort_count,
                                                          uint32_t* __restrict__ ranges) {
    const uint32_t count = *sort_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t cur = __ldg(sorted_tiles + i);
        if (i == 0) {
            ranges[2 * cur] = 0;
        } else {
            const uint32_t prev = __ldg(sorted_tiles + i - 1);
            if (cur != prev) {
                ranges[2 * prev + 1] = i;
                ranges[2 * cur] = i;
            }
        }
        if (i == count - 1) ranges[2 * cur + 1] = count;
    }
}

// the published 64-bit key of every pair of a list, on request (parity tests, debug taps)
__global__ void __launch_bounds__(256) rebuild_keys_kernel(int N, int tiles, const uint32_t* __restrict__ tile_ids,
                                                           const uint32_t* __restrict__ vals,
                                                           const float4* __restrict__ P0,
                                                           const uint32_t* __restrict__ sort_count,
                                                           uint64_t* __restrict__ keys64) {
    const uint32_t count = *sort_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t t = tile_ids[i], g = vals[i];
        const uint32_t seg = t / (uint32_t)tiles;
        const float4 p = ldg4(P0 + (size_t)seg * N + g);
        keys64[i] = ((uint64_t)t << 32) | (uint64_t)__float_as_uint(p.z);
    }
}

static inline int tile_bits_for(int S, int width, int height) {
    const long long tiles = (long long)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    long long total = (long long)S * tiles;
    if (total < 1) total = 1;
    int b = 0;
    while ((1ll << b) < total) b++;
    return b < 1 ? 1 : b;
}

struct BinningWs {
    // zeroed per call
    uint32_t* hist_depth;     // [S][4][256]
    uint32_t* hist_tile;      // [tile passes][256]
    uint32_t* tile_counter;   // [8]
    uint32_t* sort_count;     // [4]
    uint32_t* status_depth;   // [4][S*tiles_per_seg][256]
    uint32_t* status_tile;    // [tile passes][max tiles][256]
    size_t zero_bytes;
    // scratch
    uint32_t* tile_sums;
    uint32_t* dkeys[2];       // depth keys ping-pong  [S*N]
    uint32_t* perm[2];        // Gaussian index ping-pong [S*N]
    uint32_t* tkeys[2];       // tile ids ping-pong [capacity]
    uint32_t* tvals;          // second value buffer [capacity] (the other one is the caller's output)
    size_t status_depth_stride, status_tile_stride, total;
    int tile_passes;
};

static BinningWs carve(void* base, int S, int N, int width, int height, size_t capacity) {
    BinningWs w{};
    const size_t count = (size_t)S * N;
    const size_t n_scan_tiles = (count + kScanTile - 1) / kScanTile + 1;
    w.tile_passes = (tile_bits_for(S, width, height) + 7) / 8;
    const size_t depth_tiles = (size_t)S * (((size_t)N + kRsTile - 1) / kRsTile);
    const size_t max_pair_tiles = (capacity + kRsTile - 1) / kRsTile + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? (unsigned char*)base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    w.hist_depth = (uint32_t*)take(sizeof(uint32_t) * (size_t)S * 4 * kRadix);
    w.hist_tile = (uint32_t*)take(sizeof(uint32_t) * kMaxPasses * kRadix);
    w.tile_counter = (uint32_t*)take(sizeof(uint32_t) * 8);
    w.sort_count = (uint32_t*)take(sizeof(uint32_t) * 4);
    w.status_depth_stride = depth_tiles * kRadix;
    w.status_depth = (uint32_t*)take(sizeof(uint32_t) * w.status_depth_stride * 4);
    w.status_tile_stride = max_pair_tiles * kRadix;
    w.status_tile = (uint32_t*)take(sizeof(uint32_t) * w.status_tile_stride * w.tile_passes);
    w.zero_bytes = off;
    w.tile_sums = (uint32_t*)take(sizeof(uint32_t) * n_scan_tiles);
    for (int i = 0; i < 2; i++) w.dkeys[i] = (uint32_t*)take(sizeof(uint32_t) * count);
    for (int i = 0; i < 2; i++) w.perm[i] = (uint32_t*)take(sizeof(uint32_t) * count);
    for (int i = 0; i < 2; i++) w.tkeys[i] = (uint32_t*)take(sizeof(uint32_t) * capacity);
    w.tvals = (uint32_t*)take(sizeof(uint32_t) * capacity);
    w.total = off;
    return w;
}

// value buffer that pass p of the tile sort READS (p = tile_passes: the final result).  The two
// buffers alternate and the LAST pass must write the caller's d_sorted_vals, which fixes where the
// emission has to put the unsorted values.
static inline uint32_t* tile_vals_buffer(const BinningWs& w, uint32_t* d_sorted_vals, int p) {
    return ((w.tile_passes - p) & 1) ? w.tvals : d_sorted_vals;
}

static int set_onesweep_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        OMFS_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(RsSmem)));
        attr_set = true;
    }
    return OMFS_OK;
}

// ---- stage 1: depth sort of the Gaussians of every segment.  4 passes: the result is in perm[0].
int binning_depth_sort(int S, int N, int width, int height, size_t capacity, const uint32_t* d_depth_keys,
                       void* d_workspace, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    OMFS_CUDA(cudaMemsetAsync(d_workspace, 0, w.zero_bytes, stream));
    int rc = set_onesweep_attr();
    if (rc) return rc;
    // the depth exponent byte (bits 24..31) is nearly constant inside a warp's 32 keys
    rs_histogram_kernel<<<dim3(8, S), 256, 0, stream>>>(d_depth_keys, nullptr, (uint32_t)N, 4, 0x8u, w.hist_depth);
    rs_scan_hist_kernel<<<S * 4, 256, 0, stream>>>(w.hist_depth);
    count_launch(2);
    const uint32_t* kin = d_depth_keys;
    const uint32_t* vin = nullptr;  // identity
    for (int p = 0; p < 4; p++) {
        uint32_t* kout = w.dkeys[(p + 1) & 1];
        uint32_t* vout = w.perm[(p + 1) & 1];
        rs_onesweep_kernel<<<kNumSMs * 4, kRsThreads, sizeof(RsSmem), stream>>>(
            kin, vin, kout, vout, nullptr, (uint32_t)N, (uint32_t)S, 8 * p, w.hist_depth + p * kRadix, 4 * kRadix,
            w.tile_counter + p, w.status_depth + (size_t)p * w.status_depth_stride);
        count_launch();
        kin = kout;
        vin = vout;
    }
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

// ---- stage 2: offsets in depth order + emission of (tile id, Gaussian) pairs
int binning_scan_emit(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                      const uint32_t* d_tiles_touched, uint32_t* d_sorted_vals, uint32_t* d_ranges,
                      uint32_t* d_num_pairs, int* d_status_flag, unsigned long long* d_pair_accum,
                      void* d_workspace, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    const long long count = (long long)S * N;
    const int n_scan_tiles = ceil_div(count, kScanTile);
    const long long tiles_total = (long long)S * ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    if (d_ranges) OMFS_CUDA(cudaMemsetAsync(d_ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles_total, stream));
    scan_tile_sums_kernel<<<n_scan_tiles, kScanThreads, 0, stream>>>(count, N, d_tiles_touched, w.perm[0],
                                                                    w.tile_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(n_scan_tiles, w.tile_sums, d_num_pairs, (unsigned long long)capacity,
                                             d_status_flag, w.sort_count, d_pair_accum);
    scan_emit_kernel<<<n_scan_tiles, kScanThreads, 0, stream>>>(count, N, width, height, d_tiles_touched, w.perm[0],
                                                               w.tile_sums, (const float4*)d_P0, w.sort_count,
                                                               w.tkeys[0], tile_vals_buffer(w, d_sorted_vals, 0));
    count_launch(3);
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

// ---- stage 3: stable sort of the pairs by tile id.  Sorted values land in d_sorted_vals.
int binning_tile_sort(int S, int N, int width, int height, size_t capacity, uint32_t* d_sorted_vals,
                      void* d_workspace, const uint32_t** d_sorted_tiles_out, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    const int passes = w.tile_passes;
    int rc = set_onesweep_attr();
    if (rc) return rc;
    // every tile-id byte above the lowest is shared by long runs of consecutive pairs
    rs_histogram_kernel<<<dim3(kNumSMs * 4, 1), 256, 0, stream>>>(w.tkeys[0], w.sort_count, 0u, passes, 0xEu,
                                                                 w.hist_tile);
    rs_scan_hist_kernel<<<passes, 256, 0, stream>>>(w.hist_tile);
    count_launch(2);
    for (int p = 0; p < passes; p++) {
        rs_onesweep_kernel<<<kNumSMs * 4, kRsThreads, sizeof(RsSmem), stream>>>(
            w.tkeys[p & 1], tile_vals_buffer(w, d_sorted_vals, p), w.tkeys[(p + 1) & 1],
            tile_vals_buffer(w, d_sorted_vals, p + 1), w.sort_count, 0u, 1u, 8 * p, w.hist_tile + p * kRadix, 0u,
            w.tile_counter + 4 + p, w.status_tile + (size_t)p * w.status_tile_stride);
        count_launch();
    }
    OMFS_LAUNCH_CHECK();
    *d_sorted_tiles_out = w.tkeys[passes & 1];
    return OMFS_OK;
}

int binning_ranges(int S, int N, int width, int height, size_t capacity, const uint32_t* d_sorted_tiles,
                   uint32_t* d_ranges, void* d_workspace, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    tile_ranges_kernel<<<kNumSMs * 4, 256, 0, stream>>>(d_sorted_tiles, w.sort_count, d_ranges);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

int binning_rebuild_keys(int S, int N, int width, int height, size_t capacity, const uint32_t* d_tile_ids,
                         const uint32_t* d_vals, const float* d_P0, uint64_t* d_keys64, void* d_workspace,
                         cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    const int tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    rebuild_keys_kernel<<<kNumSMs * 4, 256, 0, stream>>>(N, tiles, d_tile_ids, d_vals, (const float4*)d_P0,
                                                         w.sort_count, d_keys64);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

}  // namespace omfs

using namespace omfs;

extern "C" size_t omfs_binning_workspace_bytes(int S, int N, int width, int height, size_t capacity) {
    if (S <= 0 || N <= 0 || width <= 0 || height <= 0) return 0;
    return carve(nullptr, S, N, width, height, capacity).total;
}

// sort key width of the published algorithm: 32 depth bits + the bits of the global tile id
extern "C" int omfs_binning_sort_bits(int S, int width, int height) { return 32 + tile_bits_for(S, width, height); }

extern "C" int omfs_binning(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                            const uint32_t* d_depth_keys, const uint32_t* d_tiles_touched,
                            uint32_t* d_sorted_vals, uint64_t* d_sorted_keys, uint64_t* d_emitted_keys,
                            uint32_t* d_emitted_vals, uint32_t* d_ranges, uint32_t* d_num_pairs, int* d_status_flag,
                            void* d_workspace, size_t workspace_bytes, void* stream_) {
    OMFS_REQUIRE(S > 0 && N > 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(S <= 65535, "at most 65535 segments per call");
    OMFS_REQUIRE(capacity > 0 && capacity < (1ull << 30), "capacity must be in (0, 2^30)");
    OMFS_REQUIRE((long long)S * N < (1ll << 31), "S*N must be below 2^31");
    OMFS_REQUIRE(tile_bits_for(S, width, height) <= 32, "too many tiles");
    OMFS_REQUIRE(d_P0 && d_depth_keys && d_tiles_touched && d_sorted_vals && d_ranges && d_num_pairs &&
                     d_status_flag && d_workspace,
                 "null pointer");
    OMFS_REQUIRE(workspace_bytes >= carve(nullptr, S, N, width, height, capacity).total,
                 "workspace too small (omfs_binning_workspace_bytes)");
    cudaStream_t stream = (cudaStream_t)stream_;
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    int rc = binning_depth_sort(S, N, width, height, capacity, d_depth_keys, d_workspace, stream);
    if (rc) return rc;
    rc = binning_scan_emit(S, N, width, height, capacity, d_P0, d_tiles_touched, d_sorted_vals, d_ranges,
                           d_num_pairs, d_status_flag, nullptr, d_workspace, stream);
    if (rc) return rc;
    if (d_emitted_keys) {
        const uint32_t* ev = tile_vals_buffer(w, d_sorted_vals, 0);
        rc = binning_rebuild_keys(S, N, width, height, capacity, w.tkeys[0], ev, d_P0, d_emitted_keys, d_workspace,
                                  stream);
        if (rc) return rc;
        if (d_emitted_vals)
            OMFS_CUDA(cudaMemcpyAsync(d_emitted_vals, ev, sizeof(uint32_t) * capacity, cudaMemcpyDeviceToDevice,
                                      stream));
    }
    const uint32_t* sorted_tiles = nullptr;
    rc = binning_tile_sort(S, N, width, height, capacity, d_sorted_vals, d_workspace, &sorted_tiles, stream);
    if (rc) return rc;
    rc = binning_ranges(S, N, width, height, capacity, sorted_tiles, d_ranges, d_workspace, stream);
    if (rc) return rc;
    if (d_sorted_keys)
        rc = binning_rebuild_keys(S, N, width, height, capacity, sorted_tiles, d_sorted_vals, d_P0, d_sorted_keys,
                                  d_workspace, stream);
    return rc;
}
