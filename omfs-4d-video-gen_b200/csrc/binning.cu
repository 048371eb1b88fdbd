// binning.cu — U7 (scan + key emission), U8 (onesweep radix sort), U9 (tile ranges).
//
// Contract (bit-exact against the oracle): the pair list of a batch, ordered by
//     key = (seg*tiles + tile) << 32 | float bits of the view-space depth,   ties by Gaussian index,
// with value = Gaussian index inside its segment, plus the per-tile [first,last) ranges.
//
// How it is produced.  A single 6-pass sort of 12-byte pairs (R ~ 3.6 N of them) is what the
// published pipeline does; on B200 that is the largest HBM consumer of the frame.  The same order
// comes out of far less traffic, because a counting / LSD radix sort is stable:
//   1. DEPTH SORT, segmented: the N Gaussians of every segment are sorted by depth bits (uint32 key,
//      value = Gaussian index; 4 onesweep passes over S*N 8-byte items).  Index order is the tie-break
//      because the first pass starts from the identity permutation.
//   2. TILE COUNTS: how many Gaussians touch each tile (shared-memory histograms), then ONE scan over
//      the S*tiles counters — which is already the tile-range table (U9) and the pair count.
//   3. EMIT + SCATTER, fused: a chunk of 1024 depth-ordered Gaussians compacts the ones that own pairs into
//      records, every warp walks its share of the chunk's pair space 32 pairs per step (every lane produces one
//      pair per step from the records: es_starts / es_pair), ranks the step's pairs per tile with one MATCH.ANY,
//      resolves the chunk's offset inside every tile by decoupled look-back over the chunks before it, and
//      writes each Gaussian index straight to its FINAL position.  It is a one-pass onesweep with one bin per
//      tile; inside a tile, emission order IS depth order, so the depth bits are never written and no pair is
//      ever moved twice.
// Traffic per pair drops from (8 + 24*6) = 152 B to 4 B written once (L2 merges the 4-byte scatters of
// neighbouring chunks: the whole value array of a batch fits in the 126 MB L2), plus 4 + 16*4 B per
// Gaussian for the depth sort.  The 64-bit keys exist only on request (rebuild_keys_kernel) for the
// parity tests / debug taps.
//
// The depth sort is onesweep (Adinets & Merrill 2022): one upfront histogram kernel, then ONE kernel
// per pass that ranks a 4096-key tile with warp ballots (8 VOTEs per key give the lanes holding the
// same digit — MATCH.ANY retires at ~1 per 22 clk per SM on sm_100 and was the top stall of the first
// version), resolves global offsets by decoupled look-back, and scatters through shared memory so
// global writes are contiguous per digit run.  Work items are handed out by an atomic counter, so an
// item's predecessors are always already running (no deadlock); segments are independent look-back
// chains.  Counts live on the device: nothing here synchronises with the host.
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

// ---------------------------------------------------------------------------------- tile counts
// cnt[seg*tiles + tile] = number of Gaussians of the segment whose rectangle covers the tile.
// grid = (blocks per segment, S); dynamic smem = tiles * 4 bytes.
__global__ void __launch_bounds__(256) tile_count_kernel(int N, int width, int height, int tiles,
                                                         const float4* __restrict__ P0,
                                                         const uint32_t* __restrict__ tt,
                                                         uint32_t* __restrict__ cnt) {
    extern __shared__ uint32_t s_cnt[];
    for (int i = threadIdx.x; i < tiles; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const int seg = blockIdx.y;
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const size_t i = (size_t)seg * N + n;
        if (__ldg(tt + i) == 0) continue;
        const float4 p = ldg4(P0 + i);
        int minx, miny, maxx, maxy;
        ex_tile_rect(p.x, p.y, __float_as_int(p.w), gx, gy, minx, miny, maxx, maxy);
        for (int y = miny; y < maxy; y++)
            for (int x = minx; x < maxx; x++) atomicAdd(&s_cnt[y * gx + x], 1u);
    }
    __syncthreads();
    uint32_t* out = cnt + (size_t)seg * tiles;
    for (int i = threadIdx.x; i < tiles; i += blockDim.x) {
        const uint32_t v = s_cnt[i];
        if (v) atomicAdd(out + i, v);
    }
}

// one CTA: exclusive scan of the S*tiles counters -> tile_start (final position of each tile's first
// pair), the tile ranges (U9; untouched tiles stay (0,0)), the pair count and the overflow flag.
// 8 counters per thread per round (8192 per round), so a 60-segment batch takes 8 rounds.
constexpr int kTsItems = 8;
__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
// The running total is carried in 64 bits: a close-up at 1024^2 (128 segments x 500k Gaussians x ~67 tiles)
// sums past 2^32, and a wrapped 32-bit total would pass the capacity test below and let emit_scatter write
// outside the value buffer.  Positions are truncated to 32 bits only when stored; they are meaningful only
// when total <= capacity < 2^30, which is exactly when nothing was truncated.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int n_tiles_total, const uint32_t* __restrict__ cnt,
                                                         uint32_t* __restrict__ tile_start,
                                                         uint32_t* __restrict__ ranges,
                                                         uint32_t* __restrict__ num_pairs,
                                                         unsigned long long capacity, int* __restrict__ status_flag,
                                                         uint32_t* __restrict__ sort_count,
                                                         unsigned long long* __restrict__ pair_accum,
                                                         uint32_t* __restrict__ pair_max) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles_total; base += 1024 * kTsItems) {
        const int first = base + threadIdx.x * kTsItems;
        const bool full = first + kTsItems <= n_tiles_total;
        uint32_t v[kTsItems];
        unsigned long long acc = 0;
        if (full) {
            // this CTA is alone on its SM: 16-byte accesses keep every request fully coalesced
            const uint4 a = *reinterpret_cast<const uint4*>(cnt + first);
            const uint4 b = *reinterpret_cast<const uint4*>(cnt + first + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < kTsItems; k++) v[k] = (first + k < n_tiles_total) ? cnt[first + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < kTsItems; k++) acc += v[k];
        const unsigned long long inc = warp_incl_scan64(acc);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = s_warp[lane];
            const unsigned long long winc = warp_incl_scan64(w);
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        unsigned long long run = s_carry + s_warp[warp] + inc - acc;
        uint32_t st[kTsItems], r0[kTsItems], r1[kTsItems];
#pragma unroll
        for (int k = 0; k < kTsItems; k++) {
            st[k] = (uint32_t)run;
            r0[k] = v[k] ? (uint32_t)run : 0u;
            r1[k] = v[k] ? (uint32_t)(run + v[k]) : 0u;
            run += v[k];
        }
        if (full) {
            uint4* ts = reinterpret_cast<uint4*>(tile_start + first);
            ts[0] = make_uint4(st[0], st[1], st[2], st[3]);
            ts[1] = make_uint4(st[4], st[5], st[6], st[7]);
            uint4* rg = reinterpret_cast<uint4*>(ranges + 2 * (size_t)first);
#pragma unroll
            for (int k = 0; k < 4; k++) rg[k] = make_uint4(r0[2 * k], r1[2 * k], r0[2 * k + 1], r1[2 * k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < kTsItems; k++)
                if (first + k < n_tiles_total) {
                    tile_start[first + k] = st[k];
                    ranges[2 * (first + k)] = r0[k];
                    ranges[2 * (first + k) + 1] = r1[k];
                }
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;
        __syncthreads();
    }
    const unsigned long long total64 = s_carry;
    const bool overflow = total64 > capacity;
    if (threadIdx.x == 0) {
        // what the caller sees is saturated, never wrapped: the auto-grow retry then asks for "as much as allowed"
        const uint32_t total = total64 > 0xffffffffull ? 0xffffffffu : (uint32_t)total64;
        *num_pairs = total;
        if (pair_accum) *pair_accum += total64;  // running total over the batches of one render call
        if (pair_max && total > *pair_max) *pair_max = total;  // largest batch: what the capacity must hold
        if (overflow) {
            // flag it and emit nothing; the caller re-runs with more capacity
            *status_flag = 1;
            *sort_count = 0;
        } else {
            *sort_count = total;
        }
    }
    if (overflow)  // no tile may point past the value buffer: the compositing pass then renders background
        for (int i = threadIdx.x; i < 2 * n_tiles_total; i += 1024) ranges[i] = 0u;
}

// ---------------------------------------------------------------------------------- fused emit + scatter
constexpr int kEsThreads = 256;
constexpr int kEsGpt = 4;                            // Gaussians per thread
constexpr int kEsChunk = kEsThreads * kEsGpt;        // 1024 depth-ordered Gaussians per work item
constexpr int kEsWarps = kEsThreads / 32;
constexpr int kEsRecPad = 40;                        // sentinel records behind the compacted list (a step probes 32 past its cursor)
constexpr uint32_t kEsOffBits = 21;                  // pair offset inside a (chunk, band): < 1024 Gaussians x 1024 tiles
constexpr uint32_t kEsOffMask = (1u << kEsOffBits) - 1u;
#ifndef OMFS_ES_MATCH
#define OMFS_ES_MATCH 1   // 1: rank a step's pairs with one MATCH.ANY instead of one ballot per tile-id bit; 2: alternate
#endif
#ifndef OMFS_ES_CTAS
#define OMFS_ES_CTAS 4    // resident CTAs per SM (41.3 KB of shared memory each at 1024 tiles, 64 registers)
#endif

// The per-tile arrays (bases, per-warp counters: 20 B per tile) never cover more than kEsBandTiles tiles: a frame with
// more tiles is processed in BANDS of whole tile rows (1024^2: 4 bands of 16 rows x 64 tiles).  A CTA loads its
// chunk's Gaussians once and runs compact / count / publish / look-back / rank / scatter once per band on the
// rectangles clipped to the band.
constexpr int kEsBandTiles = 1024;
static inline int emit_scatter_band_rows(int gx, int gy) { return std::min(gy, std::max(1, kEsBandTiles / gx)); }
// dynamic shared memory: rec[kEsChunk + pad] uint2 | gidx, bx, by [kEsChunk] u32 | base[band tiles] u32 |
// magic[gx + 1] u32 | wcnt[8][band tiles] u16.  41.3 KB at 1024 tiles: four CTAs per SM inside the 196 KB carve-out
// (60 KB of L1 left for the record gathers — the kernel is sensitive to that).
static inline size_t emit_scatter_smem(int band_tiles, int gx) {
    const int tp = (band_tiles + 1) & ~1;  // even row stride: the packed 16-bit counters are updated as 32-bit words
    return sizeof(uint2) * (kEsChunk + kEsRecPad) + sizeof(uint32_t) * (3 * kEsChunk + band_tiles + ((gx + 2) & ~1)) +
           sizeof(uint16_t) * kEsWarps * tp + 64;
}

// The pairs of a (chunk, band) are numbered Gaussian-major in depth order, row-major inside a Gaussian's clipped
// rectangle.  The Gaussians that own at least one pair are compacted into records {first pair | local index << 21,
// first column | first row << 10 | width << 20}; first pairs increase strictly, so the 32 pairs [J, J + 32) of a
// step find their owners with ONE probe of the 32 records behind the cursor `p` (the record that holds pair J or
// ends just before it), one warp-wide OR of "record starts at pair J + r" bits, and a population count: every lane
// produces one pair per step, whatever the rectangle sizes are.  (The first version let every thread walk its own
// four rectangles into a shared-memory window: 19 % of the lanes busy at 4096 tiles, 27 instructions per pair visited,
// a third of the kernel's instructions.)
struct EsPair {
    uint32_t tx, tyl, li;   // tile column, band-local tile row, local Gaussian index inside the chunk
};
// `starts` of a step: bit r set when a record's first pair is pair J + r (p = cursor before the step)
__device__ __forceinline__ uint32_t es_starts(const uint2* __restrict__ s_rec, uint32_t J, uint32_t p, int lane) {
    const uint32_t rel = (s_rec[p + 1 + lane].x & kEsOffMask) - J;   // >= 0: records behind the cursor start at J or later
    return __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
}
// the lane's pair of the step that starts at pair J (cursor p and `starts` of THAT step)
__device__ __forceinline__ EsPair es_pair(const uint2* __restrict__ s_rec, const uint32_t* __restrict__ s_magic,
                                          uint32_t J, uint32_t starts, uint32_t p, int lane, uint32_t lanemask_le) {
    const uint2 rec = s_rec[p + __popc(starts & lanemask_le)];
    const uint32_t k = J + (uint32_t)lane - (rec.x & kEsOffMask);    // my pair inside its Gaussian's rectangle
    const uint32_t rw = rec.y >> 20;
    const uint32_t cy = __umulhi(k + k, s_magic[rw]);                // k / rw, exact while k * rw < 2^31
    EsPair r;
    r.tx = (rec.y & 1023u) + (k - cy * rw);
    r.tyl = ((rec.y >> 10) & 1023u) + cy;
    r.li = rec.x >> kEsOffBits;
    return r;
}
// The walk is software-pipelined: the cursor chain (probe -> warp-wide OR -> population count) of step i + 1 is issued
// before the pairs of step i are decoded and used, so its shared-memory and reduction latency overlaps the ranking.
#define OMFS_ES_WALK_BEGIN(J0)                                   \
    uint32_t p = p0;                                             \
    uint32_t starts_next = es_starts(s_rec, (J0), p, lane);
#define OMFS_ES_WALK_STEP(J)                                     \
    const uint32_t starts = starts_next, p_step = p;             \
    p += __popc(starts);                                         \
    starts_next = es_starts(s_rec, (J) + 32u, p, lane);          \
    const EsPair g_step = es_pair(s_rec, s_magic, (J), starts, p_step, lane, lanemask_le);

__global__ void __launch_bounds__(kEsThreads, OMFS_ES_CTAS) emit_scatter_kernel(
    int S, int N, int width, int height, int tiles, int band_rows, const uint32_t* __restrict__ perm_a,
    const uint32_t* __restrict__ perm_b, const uint32_t* __restrict__ perm_select,
    const uint32_t* __restrict__ tt, const float4* __restrict__ P0, const uint32_t* __restrict__ tile_start,
    const uint32_t* __restrict__ sort_count, uint32_t* __restrict__ chunk_counter,
    volatile uint32_t* __restrict__ status /*[chunks][tiles]*/, uint32_t* __restrict__ vals_out) {
    extern __shared__ __align__(16) unsigned char es_raw[];
    constexpr int TILE_BITS = 10;   // band-local tile ids < kEsBandTiles: one ballot per bit in the ranking
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
    const int band_tiles = band_rows * gx;                             // <= kEsBandTiles (checked by the launcher)
    uint2* s_rec = reinterpret_cast<uint2*>(es_raw);                   // [kEsChunk + kEsRecPad] compacted Gaussians of the band
    uint32_t* s_gidx = reinterpret_cast<uint32_t*>(s_rec + kEsChunk + kEsRecPad);  // [kEsChunk]
    uint32_t* s_bx = s_gidx + kEsChunk;                                // [kEsChunk] 8x8-block columns the footprint reaches: min | max << 16
    uint32_t* s_by = s_bx + kEsChunk;                                  // [kEsChunk] ... rows
    uint32_t* s_base = s_by + kEsChunk;                                // [band_tiles]
    uint32_t* s_magic = s_base + band_tiles;                           // [gx + 1] 2^31 / width + 1
    uint16_t* s_wcnt = reinterpret_cast<uint16_t*>(s_magic + ((gx + 2) & ~1));  // [8][tp]
    const int tp = (band_tiles + 1) & ~1;
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_chunk;

    if (*sort_count == 0) return;  // empty batch, or overflow (flagged by tile_scan_kernel)
    // the depth sort ends in perm_a, or in perm_b when its last pass was skipped as trivial
    const uint32_t* __restrict__ perm = (*perm_select) ? perm_b : perm_a;
    const int cps = (N + kEsChunk - 1) / kEsChunk;  // chunks per segment
    const uint32_t n_chunks = (uint32_t)cps * (uint32_t)S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lanemask_lt = (1u << lane) - 1u, lanemask_le = lanemask_lt | (1u << lane);
    // two bits: do blocks b0, b0 + 1 lie in [bmin, bmax]?  (an empty range is stored as min 0xffff, max 0)
    auto half_bits = [](uint32_t b0, uint32_t bmin, uint32_t bmax) -> uint32_t {
        return (uint32_t)(b0 >= bmin && b0 <= bmax) | ((uint32_t)(b0 + 1 >= bmin && b0 + 1 <= bmax) << 1);
    };
    for (int r = threadIdx.x; r <= gx; r += kEsThreads) s_magic[r] = r ? 0x80000000u / (uint32_t)r + 1u : 0u;

    while (true) {
        // chunks are handed out in increasing order: every predecessor a chunk looks back at is owned
        // by a CTA that is already running
        if (threadIdx.x == 0) s_chunk = atomicAdd(chunk_counter, 1u);
        {
            uint32_t* z = reinterpret_cast<uint32_t*>(s_wcnt);
            for (int i = threadIdx.x; i < kEsWarps * tp / 2; i += kEsThreads) z[i] = 0;
        }
        __syncthreads();
        const uint32_t work = s_chunk;
        if (work >= n_chunks) break;
        // chunk-major order across the segments keeps the look-back short (see rs_onesweep_kernel)
        const int lc = (int)(work / (uint32_t)S), seg = (int)(work % (uint32_t)S);
        const uint32_t chunk = (uint32_t)seg * (uint32_t)cps + (uint32_t)lc;  // status slot; predecessor = chunk - 1
        const int g0 = lc * kEsChunk;
        const int n_g = min(kEsChunk, N - g0);
        const size_t seg_off = (size_t)seg * N;

        // ---- my 4 consecutive depth-ordered Gaussians: tile rectangles and pair counts.  The index loads,
        // then the record gathers, are issued as independent batches (two dependent round trips per chunk
        // instead of one chain per Gaussian); the pair count is the area of the rectangle, which is what
        // the preprocess stored in tiles_touched (a culled Gaussian has radius 0 at (0,0): area 0).
        int rminx[kEsGpt], rw[kEsGpt], fminy[kEsGpt], fmaxy[kEsGpt];   // the whole rectangle: columns, width, rows
        uint32_t gq[kEsGpt];
        float4 pq[kEsGpt];
#pragma unroll
        for (int q = 0; q < kEsGpt; q++) {
            const int li = threadIdx.x * kEsGpt + q;
            gq[q] = (li < n_g) ? __ldg(perm + seg_off + g0 + li) : 0u;
        }
#pragma unroll
        for (int q = 0; q < kEsGpt; q++) pq[q] = ldg4(P0 + seg_off + gq[q]);
#pragma unroll
        for (int q = 0; q < kEsGpt; q++) {
            const int li = threadIdx.x * kEsGpt + q;
            rminx[q] = fminy[q] = fmaxy[q] = 0;
            rw[q] = 1;
            if (li < n_g) {
                s_gidx[li] = gq[q];
                {   // block hint ranges (common.cuh: kValIndexBits): 8x8 blocks with a pixel inside centre +- extents
                    float ex, ey;
                    unpack_extents(pq[q].z, ex, ey);
                    int b0, b1;
                    block_range(pq[q].x, ex, b0, b1);
                    s_bx[li] = (b1 < 0 || b0 > b1) ? 0xffffu : ((uint32_t)max(b0, 0) | ((uint32_t)min(b1, 0xffff) << 16));
                    block_range(pq[q].y, ey, b0, b1);
                    s_by[li] = (b1 < 0 || b0 > b1) ? 0xffffu : ((uint32_t)max(b0, 0) | ((uint32_t)min(b1, 0xffff) << 16));
                }
                int minx, miny, maxx, maxy;
                ex_tile_rect(pq[q].x, pq[q].y, __float_as_int(pq[q].w), gx, gy, minx, miny, maxx, maxy);
                if ((maxx - minx) * (maxy - miny) > 0) {
                    rminx[q] = minx;
                    rw[q] = maxx - minx;
                    fminy[q] = miny;
                    fmaxy[q] = maxy;
                }
            }
        }
        // ---- one band of tile rows at a time (a frame of up to kEsBandTiles tiles is one band)
        for (int y0 = 0; y0 < gy; y0 += band_rows) {
        const int y1 = min(gy, y0 + band_rows);
        const int nt = (y1 - y0) * gx;             // tiles of this band
        const int tile_lo = y0 * gx;               // its first tile in the frame's numbering
        // ---- compact the Gaussians that own pairs in this band (one scan carries pair offsets and record slots)
        uint32_t total_pairs, n_rec;
        {
            int rminy[kEsGpt];                     // first row of the clipped rectangle, band-local
            uint32_t cnt[kEsGpt];                  // pairs of the clipped rectangle
            uint32_t acc = 0;
#pragma unroll
            for (int q = 0; q < kEsGpt; q++) {
                const int cy0 = max(fminy[q], y0), cy1 = min(fmaxy[q], y1);
                rminy[q] = cy0 - y0;
                cnt[q] = cy1 > cy0 ? (uint32_t)((cy1 - cy0) * rw[q]) : 0u;
                acc += (cnt[q] << 11) + (cnt[q] ? 1u : 0u);
            }
            uint32_t packed_total;
            uint32_t run = block_excl_scan_256(acc, s_scan, packed_total);
            total_pairs = packed_total >> 11;
            n_rec = packed_total & 2047u;
#pragma unroll
            for (int q = 0; q < kEsGpt; q++) {
                if (cnt[q]) {
                    const uint32_t li = (uint32_t)(threadIdx.x * kEsGpt + q);
                    s_rec[run & 2047u] = make_uint2((run >> 11) | (li << kEsOffBits),
                                                    (uint32_t)rminx[q] | ((uint32_t)rminy[q] << 10) | ((uint32_t)rw[q] << 20));
                    run += (cnt[q] << 11) + 1u;
                }
            }
            if (threadIdx.x < kEsRecPad) s_rec[n_rec + threadIdx.x] = make_uint2(kEsOffMask, 0u);
            __syncthreads();
        }
        // warp w owns the contiguous pair range [w*per_warp, (w+1)*per_warp) of the (chunk, band)
        const uint32_t per_warp = (total_pairs + kEsWarps - 1) / kEsWarps;
        const uint32_t my_lo = min(total_pairs, (uint32_t)warp * per_warp);
        const uint32_t my_hi = min(total_pairs, (uint32_t)(warp + 1) * per_warp);
        // cursor of my first pair: the last record that starts at or before it (32-ary search, two probes per lane;
        // slots from the record count on hold sentinels, so a clamped probe reads "starts after every pair")
        uint32_t p0 = 0;
        if (my_hi > my_lo) {
            uint32_t idx = min((uint32_t)lane * 32u, n_rec);
            const uint32_t c1 = __ballot_sync(0xffffffffu, (s_rec[idx].x & kEsOffMask) <= my_lo);
            const uint32_t b = (uint32_t)__popc(c1) - 1u;
            idx = min(b * 32u + (uint32_t)lane, n_rec);
            const uint32_t c2 = __ballot_sync(0xffffffffu, (s_rec[idx].x & kEsOffMask) <= my_lo);
            p0 = b * 32u + (uint32_t)__popc(c2) - 1u;
        }
        // ---- count: every warp counts its own pairs per tile in its own counter row (order is irrelevant here)
        {
            uint32_t* wrow = reinterpret_cast<uint32_t*>(s_wcnt + warp * tp);
            OMFS_ES_WALK_BEGIN(my_lo)
            for (uint32_t J = my_lo; J < my_hi; J += 32) {
                OMFS_ES_WALK_STEP(J)
                if (J + (uint32_t)lane < my_hi) {
                    const uint32_t t = g_step.tyl * (uint32_t)gx + g_step.tx;
                    atomicAdd(wrow + (t >> 1), 1u << (16 * (t & 1u)));
                }
            }
        }
        __syncthreads();
        // ---- per tile: offsets of the warps inside the chunk and the chunk aggregate.  EVERY aggregate of the
        // chunk is published before any look-back starts, so a successor never waits for a tile whose owner
        // thread is still busy looking back for an earlier one.
        volatile uint32_t* my_status = status + (size_t)chunk * tiles + tile_lo;
        for (int t = threadIdx.x; t < nt; t += kEsThreads) {
            uint32_t off = 0;
#pragma unroll
            for (int w = 0; w < kEsWarps; w++) {
                const uint32_t c = s_wcnt[w * tp + t];
                s_wcnt[w * tp + t] = (uint16_t)off;
                off += c;
            }
            my_status[t] = (lc == 0 ? kFlagPrefix : kFlagAgg) | off;
            s_base[t] = off;  // parked here until the look-back below turns it into the tile's base
        }
        // ---- look-back, kEsLook tiles per thread at a time: the probes of a batch are independent loads, so a
        // thread pays one round trip per batch and step, not one per tile (at 4096 tiles a thread owns 16)
        constexpr int kEsLook = 4;
        for (int t0 = threadIdx.x; t0 < nt; t0 += kEsThreads * kEsLook) {
            uint32_t excl[kEsLook], look[kEsLook];
            bool pend[kEsLook];
            bool any = false;
#pragma unroll
            for (int g = 0; g < kEsLook; g++) {
                excl[g] = 0;
                look[g] = chunk - 1;
                pend[g] = lc > 0 && t0 + g * kEsThreads < nt;
                any = any || pend[g];
            }
            uint32_t spins = 0;
            while (any) {
                if (++spins > (1u << 28)) __trap();  // a lost predecessor is a bug: fail, do not hang
                uint32_t sv[kEsLook];
#pragma unroll
                for (int g = 0; g < kEsLook; g++)
                    sv[g] = pend[g] ? status[(size_t)look[g] * tiles + tile_lo + t0 + g * kEsThreads] : 0u;
                any = false;
#pragma unroll
                for (int g = 0; g < kEsLook; g++) {
                    if (!pend[g]) continue;
                    const uint32_t flag = sv[g] & kFlagMask;
                    if (flag == kFlagPrefix) {
                        excl[g] += sv[g] & kValMask;
                        pend[g] = false;
                    } else if (flag == kFlagAgg) {
                        excl[g] += sv[g] & kValMask;
                        look[g]--;
                    }
                    any = any || pend[g];
                }
            }
#pragma unroll
            for (int g = 0; g < kEsLook; g++) {
                const int t = t0 + g * kEsThreads;
                if (t >= nt) continue;
                if (lc > 0) my_status[t] = kFlagPrefix | ((excl[g] + s_base[t]) & kValMask);
                s_base[t] = __ldg(tile_start + (size_t)seg * tiles + tile_lo + t) + excl[g];
            }
        }
        __syncthreads();  // s_base / s_wcnt are final
        // ---- rank + scatter: the same walk again, 32 pairs per step in pair order
        {
            uint16_t* wc = s_wcnt + warp * tp;
            OMFS_ES_WALK_BEGIN(my_lo)
            EsPair g;
            {   // the first step's pairs
                OMFS_ES_WALK_STEP(my_lo)
                g = g_step;
            }
            for (uint32_t J = my_lo; J < my_hi; J += 32) {
                // the NEXT step's pairs are decoded first (probing past the range's end reads sentinels: harmless), and
                // everything that depends only on the pair — tile, block hint, list word — is formed before the ranking,
                // so that the MATCH and the counter round trip overlap independent work
                EsPair g_next;
                {
                    OMFS_ES_WALK_STEP(J + 32u)
                    g_next = g_step;
                }
                const bool valid = J + (uint32_t)lane < my_hi;
                const uint32_t t = g.tyl * (uint32_t)gx + g.tx;
                uint32_t word;
                {   // block hint of the pair: which halves of tile (tx, ty) the footprint's block range reaches
                    const uint32_t ty = g.tyl + (uint32_t)y0;
                    const uint32_t rx = s_bx[g.li], ry = s_by[g.li];
                    const uint32_t hx = half_bits(2u * g.tx, rx & 0xffffu, rx >> 16), hy = half_bits(2u * ty, ry & 0xffffu, ry >> 16);
                    const uint32_t hint = ((hy & 1u) ? hx : 0u) | ((hy & 2u) ? (hx << 2) : 0u);
                    word = s_gidx[g.li] | (hint << kValIndexBits);
                }
                g = g_next;
                // lanes of the step that hold my tile
                uint32_t peers;
#if OMFS_ES_MATCH == 1
                peers = __match_any_sync(0xffffffffu, valid ? t : (0x80000000u | (uint32_t)lane));
#else
#if OMFS_ES_MATCH == 2
                if ((J >> 5) & 1u) {
                    peers = __match_any_sync(0xffffffffu, valid ? t : (0x80000000u | (uint32_t)lane));
                } else
#endif
                {
                    peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
                    for (int b = 0; b < TILE_BITS; b++) {
                        const bool bit = (t >> b) & 1u;
                        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                        peers &= bit ? bal : ~bal;
                    }
                }
#endif
                const int leader = __ffs(peers) - 1;
                uint32_t pre = 0;
                if (valid && lane == leader) {
                    pre = wc[t];
                    wc[t] = (uint16_t)(pre + __popc(peers));
                }
                pre = __shfl_sync(0xffffffffu, pre, leader & 31);
                if (valid) vals_out[s_base[t] + pre + __popc(peers & lanemask_lt)] = word;
                __syncwarp();
            }
        }
        __syncthreads();
        if (y1 < gy) {   // the next band starts from clean counters
            uint32_t* z = reinterpret_cast<uint32_t*>(s_wcnt);
            for (int i = threadIdx.x; i < kEsWarps * tp / 2; i += kEsThreads) z[i] = 0;
            __syncthreads();
        }
        }  // bands
    }
}

#undef OMFS_ES_WALK_BEGIN
#undef OMFS_ES_WALK_STEP

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

// exclusive scan of each (segment, pass) histogram: counts -> digit start offsets.  grid = n_seg*passes.
// If `nontrivial` is given, blocks of pass `watch_pass` (the top byte of the depth key) raise it when their
// segment's VISIBLE keys do not all share one digit.  Culled Gaussians carry key 0 (top byte 0; a visible depth is
// > 0.2, top byte >= 0x3e): they emit no pairs, so where they land relative to the visible ones is irrelevant, and a
// pass whose digit is constant over the visible keys of every segment only moves culled entries — it is skipped.
__global__ void __launch_bounds__(256) rs_scan_hist_kernel(uint32_t* __restrict__ hist, int passes, int watch_pass,
                                                           uint32_t seg_len, uint32_t* __restrict__ nontrivial) {
    __shared__ uint32_t s_warp[8];
    uint32_t* h = hist + (size_t)blockIdx.x * kRadix;
    const uint32_t v = h[threadIdx.x];
    (void)seg_len;
    if (nontrivial && (int)(blockIdx.x % passes) == watch_pass) {
        const int distinct = __syncthreads_count(threadIdx.x != 0 && v != 0u);
        if (threadIdx.x == 0 && distinct > 1) atomicOr(nontrivial, 1u);
    }
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
#ifndef OMFS_RS_CTAS
#define OMFS_RS_CTAS 5   // resident CTAs per SM (42 KB of shared memory each, 48 registers): 0.174 -> 0.169 ms per 60 frames against 4 (80 registers at 3: 0.187)
#endif
__global__ void __launch_bounds__(kRsThreads, OMFS_RS_CTAS) rs_onesweep_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ count_ptr, uint32_t seg_len, uint32_t n_seg,
    int shift, const uint32_t* __restrict__ digit_start, uint32_t digit_stride,
    uint32_t* __restrict__ tile_counter, volatile uint32_t* __restrict__ status /*[tiles][256] this pass*/,
    const uint32_t* __restrict__ run_if_set /*NULL: always run*/, uint32_t* __restrict__ skipped_flag) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_smem_raw);
    if (run_if_set && *run_if_set == 0u) {
        // this pass's digit is constant in every segment: the pass is the identity permutation.  Tell the
        // consumer that the result stayed in the input buffers.
        if (blockIdx.x == 0 && threadIdx.x == 0) *skipped_flag = 1u;
        return;
    }
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
        const uint32_t work = sm.tile_id;
        if (work >= n_tiles) break;
        // work items are ordered tile-major across the segments (tile 0 of every segment, then tile 1 ...):
        // the CTAs in flight spread over all segments, so a tile's predecessors have usually published their
        // inclusive prefix already and the look-back stays short
        const uint32_t ltile = work / n_seg, seg = work - ltile * n_seg;
        const uint32_t tile = seg * tiles_per_seg + ltile;  // status slot; the predecessor is tile - 1
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

// the published 64-bit key of every sorted pair, on request (parity tests, debug taps): one thread per
// (global tile, position) via the range table
__global__ void __launch_bounds__(256) rebuild_keys_kernel(int N, int tiles, long long n_tiles_total,
                                                           const uint32_t* __restrict__ ranges,
                                                           const uint32_t* __restrict__ vals,
                                                           const uint32_t* __restrict__ depth_keys,
                                                           uint64_t* __restrict__ keys64) {
    for (long long tg = blockIdx.x; tg < n_tiles_total; tg += gridDim.x) {
        const uint32_t lo = ranges[2 * tg], hi = ranges[2 * tg + 1];
        const long long seg = tg / tiles;
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            keys64[i] = ((uint64_t)tg << 32) | (uint64_t)__ldg(depth_keys + seg * N + (vals[i] & kValIndexMask));
        }
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
    uint32_t* counters;       // [8]: 0..3 depth-sort tile counters, 4 emit-scatter chunk counter
    uint32_t* sort_count;     // [4]
    uint32_t* tile_cnt;       // [S*tiles]
    uint32_t* status_depth;   // [4][S*tiles_per_seg][256]
    uint32_t* status_emit;    // [S*chunks_per_seg][tiles]
    size_t zero_bytes;
    // scratch
    uint32_t* tile_start;     // [S*tiles]
    uint32_t* dkeys[2];       // depth keys ping-pong  [S*N]
    uint32_t* perm[2];        // Gaussian index ping-pong [S*N]
    size_t status_depth_stride, total;
    int tiles, tile_bits;
};

static BinningWs carve(void* base, int S, int N, int width, int height, size_t capacity) {
    (void)capacity;
    BinningWs w{};
    const size_t count = (size_t)S * N;
    w.tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    w.tile_bits = 1;
    while ((1 << w.tile_bits) < w.tiles) w.tile_bits++;
    const size_t depth_tiles = (size_t)S * (((size_t)N + kRsTile - 1) / kRsTile);
    const size_t emit_chunks = (size_t)S * (((size_t)N + kEsChunk - 1) / kEsChunk);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? (unsigned char*)base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    w.hist_depth = (uint32_t*)take(sizeof(uint32_t) * (size_t)S * 4 * kRadix);
    w.counters = (uint32_t*)take(sizeof(uint32_t) * 8);
    w.sort_count = (uint32_t*)take(sizeof(uint32_t) * 4);
    w.tile_cnt = (uint32_t*)take(sizeof(uint32_t) * (size_t)S * w.tiles);
    w.status_depth_stride = depth_tiles * kRadix;
    w.status_depth = (uint32_t*)take(sizeof(uint32_t) * w.status_depth_stride * 4);
    w.status_emit = (uint32_t*)take(sizeof(uint32_t) * emit_chunks * w.tiles);
    w.zero_bytes = off;
    w.tile_start = (uint32_t*)take(sizeof(uint32_t) * (size_t)S * w.tiles);
    for (int i = 0; i < 2; i++) w.dkeys[i] = (uint32_t*)take(sizeof(uint32_t) * count);
    for (int i = 0; i < 2; i++) w.perm[i] = (uint32_t*)take(sizeof(uint32_t) * count);
    w.total = off;
    return w;
}

static int set_kernel_attrs(int tiles) {
    static DeviceOnce rs_once, es_once;
    int rc;
    if ((rc = ensure_dyn_smem(rs_once, rs_onesweep_kernel, (int)sizeof(RsSmem)))) return rc;
    (void)tiles;
    if ((rc = ensure_dyn_smem(es_once, emit_scatter_kernel, (int)emit_scatter_smem(kEsBandTiles, kEsBandTiles)))) return rc;
    return OMFS_OK;
}

// ---- stage 1: depth sort of the Gaussians of every segment.  4 passes: the result is in perm[0]
// (or in perm[1] with counters[6] set, when the exponent-byte pass was the identity and got skipped).
// ---- stage 0: zero the per-batch counters.  With `fused` the producer of the depth keys (bind_preprocess) fills
// the digit histograms and the tile counts itself: the pointers it accumulates into are returned.
int binning_prepare(int S, int N, int width, int height, size_t capacity, void* d_workspace, uint32_t** d_hist_depth,
                    uint32_t** d_tile_cnt, cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    OMFS_CUDA(cudaMemsetAsync(d_workspace, 0, w.zero_bytes, stream));
    if (d_hist_depth) *d_hist_depth = w.hist_depth;
    if (d_tile_cnt) *d_tile_cnt = w.tile_cnt;
    return OMFS_OK;
}

int binning_depth_sort(int S, int N, int width, int height, size_t capacity, const uint32_t* d_depth_keys,
                       void* d_workspace, cudaStream_t stream, bool prepared_and_histogrammed) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    int rc = set_kernel_attrs(w.tiles);
    if (rc) return rc;
    if (!prepared_and_histogrammed) {
        OMFS_CUDA(cudaMemsetAsync(d_workspace, 0, w.zero_bytes, stream));
        // the depth exponent byte (bits 24..31) is nearly constant inside a warp's 32 keys
        rs_histogram_kernel<<<dim3(8, S), 256, 0, stream>>>(d_depth_keys, nullptr, (uint32_t)N, 4, 0x8u, w.hist_depth);
        count_launch();
    }
    // counters[5] = "top-byte pass is not trivial", counters[6] = "top-byte pass was skipped"
    rs_scan_hist_kernel<<<S * 4, 256, 0, stream>>>(w.hist_depth, 4, 3, (uint32_t)N, w.counters + 5);
    count_launch();
    const uint32_t* kin = d_depth_keys;
    const uint32_t* vin = nullptr;  // identity
    for (int p = 0; p < 4; p++) {
        uint32_t* kout = w.dkeys[(p + 1) & 1];
        uint32_t* vout = w.perm[(p + 1) & 1];
        rs_onesweep_kernel<<<kNumSMs * OMFS_RS_CTAS, kRsThreads, sizeof(RsSmem), stream>>>(
            kin, vin, kout, vout, nullptr, (uint32_t)N, (uint32_t)S, 8 * p, w.hist_depth + p * kRadix, 4 * kRadix,
            w.counters + p, w.status_depth + (size_t)p * w.status_depth_stride, p == 3 ? w.counters + 5 : nullptr,
            w.counters + 6);
        count_launch();
        kin = kout;
        vin = vout;
    }
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

// ---- stage 2: tile counts, their scan (= tile ranges, pair count)
int binning_tile_ranges(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                        const uint32_t* d_tiles_touched, uint32_t* d_ranges, uint32_t* d_num_pairs,
                        int* d_status_flag, unsigned long long* d_pair_accum, uint32_t* d_pair_max,
                        void* d_workspace, cudaStream_t stream, bool counted) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    if (w.tiles > 12288) {
        set_error("binning: %d tiles per frame exceed the shared-memory histogram (max 12288)", w.tiles);
        return OMFS_ERR_INVALID;
    }
    if (!counted) {
        tile_count_kernel<<<dim3(32, S), 256, sizeof(uint32_t) * w.tiles, stream>>>(N, width, height, w.tiles,
                                                                                  (const float4*)d_P0, d_tiles_touched,
                                                                                  w.tile_cnt);
        count_launch();
    }
    tile_scan_kernel<<<1, 1024, 0, stream>>>(S * w.tiles, w.tile_cnt, w.tile_start, d_ranges, d_num_pairs,
                                             (unsigned long long)capacity, d_status_flag, w.sort_count,
                                             d_pair_accum, d_pair_max);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

// ---- stage 3: fused emission + counting sort by tile: Gaussian indices land at their final positions
int binning_emit_scatter(int S, int N, int width, int height, size_t capacity, const float* d_P0,
                         const uint32_t* d_tiles_touched, uint32_t* d_sorted_vals, void* d_workspace,
                         cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    if (w.tiles > (1 << (kValIndexBits - 10)) || N > (int)kValIndexMask) {   // pair word: hint | tile | local Gaussian
        set_error("binning: too many tiles per frame (%d) or Gaussians (%d)", w.tiles, N);
        return OMFS_ERR_INVALID;
    }
    int rc = set_kernel_attrs(w.tiles);
    if (rc) return rc;
    const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
    if (gx > kEsBandTiles) {
        set_error("binning: %d tile columns per frame exceed one band (%d tiles)", gx, kEsBandTiles);
        return OMFS_ERR_INVALID;
    }
    const int band_rows = emit_scatter_band_rows(gx, gy);
    emit_scatter_kernel<<<kNumSMs * OMFS_ES_CTAS, kEsThreads, emit_scatter_smem(band_rows * gx, gx), stream>>>(
        S, N, width, height, w.tiles, band_rows, w.perm[0], w.perm[1], w.counters + 6, d_tiles_touched,
        (const float4*)d_P0, w.tile_start, w.sort_count, w.counters + 4, w.status_emit, d_sorted_vals);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

int binning_rebuild_keys(int S, int N, int width, int height, size_t capacity, const uint32_t* d_ranges,
                         const uint32_t* d_vals, const uint32_t* d_depth_keys, uint64_t* d_keys64, void* d_workspace,
                         cudaStream_t stream) {
    BinningWs w = carve(d_workspace, S, N, width, height, capacity);
    rebuild_keys_kernel<<<kNumSMs * 8, 256, 0, stream>>>(N, w.tiles, (long long)S * w.tiles, d_ranges, d_vals,
                                                         d_depth_keys, d_keys64);
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
                            uint32_t* d_sorted_vals, uint64_t* d_sorted_keys, uint32_t* d_ranges,
                            uint32_t* d_num_pairs, int* d_status_flag, void* d_workspace, size_t workspace_bytes,
                            void* stream_) {
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
    int rc = binning_depth_sort(S, N, width, height, capacity, d_depth_keys, d_workspace, stream);
    if (rc) return rc;
    rc = binning_tile_ranges(S, N, width, height, capacity, d_P0, d_tiles_touched, d_ranges, d_num_pairs,
                             d_status_flag, nullptr, nullptr, d_workspace, stream);
    if (rc) return rc;
    rc = binning_emit_scatter(S, N, width, height, capacity, d_P0, d_tiles_touched, d_sorted_vals, d_workspace,
                              stream);
    if (rc) return rc;
    if (d_sorted_keys)
        rc = binning_rebuild_keys(S, N, width, height, capacity, d_ranges, d_sorted_vals, d_depth_keys, d_sorted_keys,
                                  d_workspace, stream);
    return rc;
}
