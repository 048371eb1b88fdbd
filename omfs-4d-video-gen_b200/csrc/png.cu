// png.cu — the device frame sink: uint8 frames in HBM -> complete PNG files in HBM, so that only compressed
// streams cross PCIe (replaces upstream's per-frame save_image behind
// 02_Visual_Engine/render_surgery.py:289-315 and the PNG round trip of stitch_video, :412-449).
//
// The stream format, the code tables and the tokenisation live in png_core.cuh (shared with the host
// emulation the CPU tests run).  Three kernels per batch of S frames:
//   png_strip_kernel   one CTA per strip of rows (persistent, strips drawn round-robin): Up filter, token costs
//                      under the 8 fixed code tables -> best table or stored, bit-exact parallel emission
//                      (block scan of bit counts, shared-memory OR), end-of-block + byte-aligning empty stored
//                      block, chunk framing with a parallel CRC-32 (64-byte pieces combined by polynomial
//                      multiplication), Adler-32 partial sums; the finished IDAT chunk goes to a fixed-stride slot;
//   png_layout_kernel  one CTA: per-frame chunk offsets, frame sizes, their scan (frames are packed back to back so
//                      the host copies exactly the bytes produced), Adler-32 of every frame;
//   png_pack_kernel    one CTA per strip: the chunk moves to its final byte offset; header, final IDAT and IEND.
// HBM traffic per frame: the frame is read twice by the strip kernel (the row above comes from L1/L2), the
// compressed stream is written twice (slot, final position) and read once: ~(1 + 3/ratio) x frame bytes.  The
// kernel is bound by issue slots (a few dozen integer instructions per byte), not by HBM.
#include <mutex>

#include "common.cuh"
#include "png_core.cuh"

namespace omfs {
using namespace omfs_png;

struct StripInfo {
    uint32_t bytes;            // size of the strip's IDAT chunk (framing included)
    uint32_t s1;               // sum of the strip's filtered bytes
    unsigned long long s2;     // sum of byte * (strip length - position): its share of Adler's second sum
};

constexpr int kPngThreads = 256;
constexpr int kPngWarps = kPngThreads / 32;
constexpr uint32_t kAdlerMod = 65521u;

struct PngHeader {
    uint32_t w[9];   // 33 bytes (+3 unused)
};

__device__ __forceinline__ uint4 load16_bytes(const uint8_t* p, int n, bool aligned) {
    if (aligned && n == kSeg) return __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) w[i >> 2] |= (uint32_t)__ldg(p + i) << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// OR `n` (<= 32) bits of `value` into the bit stream at bit position `pos` (shared memory words, zeroed before)
__device__ __forceinline__ void put_bits(uint32_t* out, uint32_t pos, uint32_t value, int n) {
    if (n <= 0) return;
    const unsigned long long v = (unsigned long long)(n >= 32 ? value : (value & ((1u << n) - 1u))) << (pos & 31u);
    atomicOr(out + (pos >> 5), (uint32_t)v);
    if ((uint32_t)(v >> 32)) atomicOr(out + (pos >> 5) + 1, (uint32_t)(v >> 32));
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPngWarps; w++) {
        const uint32_t x = s_warp[w];
        if (w < warp) before += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return before + inc - v;
}

// dynamic shared memory: filtered bytes [rows_per_strip][segs_per_row] uint4 | chunk words [kMaxChunk/4 + 4]
__global__ void __launch_bounds__(kPngThreads) png_strip_kernel(int S, Geometry g, const uint8_t* __restrict__ frames,
                                                                const Tables* __restrict__ tb,
                                                                uint8_t* __restrict__ slots,
                                                                StripInfo* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char png_smem[];
    uint4* s_filt = reinterpret_cast<uint4*>(png_smem);
    uint32_t* s_out = reinterpret_cast<uint32_t*>(s_filt + (size_t)g.rows_per_strip * g.segs_per_row);
    __shared__ uint32_t s_token[kNumTokens];
    __shared__ uint32_t s_lens_a[kNumTokens], s_lens_b[kNumTokens];
    __shared__ uint32_t s_crc[256];
    __shared__ uint32_t s_warp[kPngWarps];
    __shared__ unsigned long long s_red[kPngWarps][6];
    __shared__ uint32_t s_choice[4];   // table (or kNumTables = stored), data bits of the Huffman form, -, -

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kNumTokens; i += kPngThreads) {
        s_lens_a[i] = tb->lens_a[i];
        s_lens_b[i] = tb->lens_b[i];
    }
    s_crc[tid] = tb->crc_byte[tid];
    const size_t frame_bytes = (size_t)g.height * g.row_bytes;
    const bool aligned = ((reinterpret_cast<uintptr_t>(frames) | (uintptr_t)frame_bytes | (uintptr_t)g.row_bytes) & 15u) == 0;
    const int line = g.row_bytes + 1;
    const long long n_work = (long long)S * g.n_strips;

    for (long long work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int frame = (int)(work / g.n_strips), strip = (int)(work % g.n_strips);
        const int row0 = strip * g.rows_per_strip;
        const int rows = min(g.rows_per_strip, g.height - row0);
        const int nsegs = rows * g.segs_per_row;
        const int iters = (nsegs + kPngThreads - 1) / kPngThreads;   // <= 4
        const uint8_t* img = frames + (size_t)frame * frame_bytes;
        const uint32_t strip_len = (uint32_t)rows * (uint32_t)line;
        const bool zhdr = strip == 0;

        __syncthreads();   // previous strip fully written out; tables loaded
        for (int i = tid; i < kMaxChunk / 4 + 4; i += kPngThreads) s_out[i] = 0;
        // ---- 1a. Up filter into shared memory
        for (int q = tid; q < nsegs; q += kPngThreads) {
            const int r = q / g.segs_per_row, s = q - r * g.segs_per_row;
            const int n = min(kSeg, g.row_bytes - kSeg * s);
            const uint8_t* p = img + (size_t)(row0 + r) * g.row_bytes + kSeg * s;
            const uint4 cur = load16_bytes(p, n, aligned);
            uint4 up = make_uint4(0, 0, 0, 0);
            if (row0 + r > 0) up = load16_bytes(p - g.row_bytes, n, aligned);
            s_filt[q] = make_uint4(__vsub4(cur.x, up.x), __vsub4(cur.y, up.y), __vsub4(cur.z, up.z), __vsub4(cur.w, up.w));
        }
        __syncthreads();
        // ---- 1b. token costs under every table, Adler sums
        uint32_t pa[4], pb[4];   // per iteration: this thread's bit counts, one byte per table
        uint32_t cost[kNumTables];
#pragma unroll
        for (int k = 0; k < kNumTables; k++) cost[k] = 0;
        uint32_t sum_d = 0;
        unsigned long long sum_dpos = 0;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            pa[it] = pb[it] = 0;
            const int q = tid + it * kPngThreads;
            if (it < iters && q < nsegs) {
                const int r = q / g.segs_per_row, s = q - r * g.segs_per_row;
                const int n = min(kSeg, g.row_bytes - kSeg * s);
                const uint4 f = s_filt[q];
                const uint32_t w[4] = {f.x, f.y, f.z, f.w};
                const uint32_t prev = s ? (s_filt[q - 1].w >> 24) : 2u;
                uint32_t a = 0, b = 0;
                if (s == 0) {   // the row's filter-type byte, a literal 2
                    a = s_lens_a[2];
                    b = s_lens_b[2];
                }
                for_each_token(w, prev, n, [&](int tok) {
                    a += s_lens_a[tok];
                    b += s_lens_b[tok];
                });
                pa[it] = a;
                pb[it] = b;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    cost[k] += (a >> (8 * k)) & 0xffu;
                    cost[4 + k] += (b >> (8 * k)) & 0xffu;
                }
                uint32_t s1, s2;
                seg_sums(w, s1, s2);
                const uint32_t pos0 = (uint32_t)r * (uint32_t)line + 1u + (uint32_t)(kSeg * s);
                sum_d += s1;
                sum_dpos += (unsigned long long)pos0 * s1 + s2;
                if (s == 0) {
                    sum_d += 2u;
                    sum_dpos += 2ull * ((unsigned long long)r * line);
                }
            }
        }
        // block reduction of the 8 costs (two per 64-bit word), sum_d, sum_dpos
        {
            unsigned long long v[6] = {((unsigned long long)cost[1] << 32) | cost[0], ((unsigned long long)cost[3] << 32) | cost[2],
                                       ((unsigned long long)cost[5] << 32) | cost[4], ((unsigned long long)cost[7] << 32) | cost[6],
                                       (unsigned long long)sum_d, sum_dpos};
#pragma unroll
            for (int i = 0; i < 6; i++) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], d);
                if (lane == 0) s_red[warp][i] = v[i];
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long t[6] = {0, 0, 0, 0, 0, 0};
            for (int w = 0; w < kPngWarps; w++)
                for (int i = 0; i < 6; i++) t[i] += s_red[w][i];
            uint32_t best = 0, best_bits = 0xffffffffu;
            for (int k = 0; k < kNumTables; k++) {
                const uint32_t c = (uint32_t)(t[k >> 1] >> (32 * (k & 1)));
                const uint32_t bits = tb->hdr_bits[k] + c + (tb->token[k][256] >> 24);
                if (bits < best_bits) {
                    best_bits = bits;
                    best = (uint32_t)k;
                }
            }
            // Huffman form: block + 3 bits of empty stored block, padded to a byte, + 00 00 FF FF; stored form: 5 + len
            const uint32_t huff_bytes = (best_bits + 3u + 7u) / 8u + 4u;
            const uint32_t stored_bytes = 5u + strip_len;
            s_choice[0] = stored_bytes <= huff_bytes ? (uint32_t)kNumTables : best;
            s_choice[1] = best_bits;
            s_red[0][4] = t[4];
            s_red[0][5] = t[5];
        }
        __syncthreads();
        const uint32_t table = s_choice[0];
        const uint32_t data0 = 8u + (zhdr ? 2u : 0u);   // first byte after the chunk framing (and the zlib header)
        uint32_t data_bytes;                            // chunk payload
        uint8_t* out8 = reinterpret_cast<uint8_t*>(s_out);
        if (table < (uint32_t)kNumTables) {
            // ---- 2. emission with the chosen table
            for (int i = tid; i < kNumTokens; i += kPngThreads) s_token[i] = tb->token[table][i];
            const uint32_t hdr_bits = tb->hdr_bits[table];
            uint32_t cursor = data0 * 8u;
            for (uint32_t i = tid; i * 32u < hdr_bits; i += kPngThreads)
                put_bits(s_out, cursor + 32u * i, tb->hdr[table][i], (int)min(32u, hdr_bits - 32u * i));
            cursor += hdr_bits;
            __syncthreads();   // s_token ready
#pragma unroll
            for (int it = 0; it < 4; it++) {
                if (it >= iters) break;   // uniform over the CTA
                const int q = tid + it * kPngThreads;
                const uint32_t packed = table < 4 ? pa[it] : pb[it];
                const uint32_t mine = (packed >> (8 * (table & 3u))) & 0xffu;
                uint32_t total;
                uint32_t pos = cursor + block_excl_scan(mine, s_warp, total);
                cursor += total;
                if (q < nsegs) {
                    const int r = q / g.segs_per_row, s = q - r * g.segs_per_row;
                    const int n = min(kSeg, g.row_bytes - kSeg * s);
                    const uint4 f = s_filt[q];
                    const uint32_t w[4] = {f.x, f.y, f.z, f.w};
                    const uint32_t prev = s ? (s_filt[q - 1].w >> 24) : 2u;
                    // bits are gathered in a 64-bit window aligned to the stream's 32-bit words; full words go out
                    unsigned long long acc = 0;
                    uint32_t fill = pos & 31u, word = pos >> 5;
                    auto emit = [&](int tok) {
                        const uint32_t e = s_token[tok];
                        acc |= (unsigned long long)(e & 0xffffffu) << fill;
                        fill += e >> 24;
                        if (fill >= 32u) {
                            atomicOr(s_out + word, (uint32_t)acc);
                            acc >>= 32;
                            fill -= 32u;
                            word++;
                        }
                    };
                    if (s == 0) emit(2);
                    for_each_token(w, prev, n, emit);
                    if (fill) atomicOr(s_out + word, (uint32_t)acc);
                }
            }
            if (tid == 0) {
                const uint32_t e = s_token[256];
                put_bits(s_out, cursor, e & 0xffffffu, (int)(e >> 24));
                uint32_t end = cursor + (e >> 24) + 3u;   // + BFINAL=0, BTYPE=00 of the empty stored block (zeros)
                end = (end + 7u) & ~7u;
                put_bits(s_out, end + 16u, 0xffffu, 16);  // LEN = 0, NLEN = 0xffff
                s_choice[2] = end / 8u + 4u - 8u;         // payload bytes
            }
            __syncthreads();
            data_bytes = s_choice[2];
        } else {
            // ---- 2'. stored block: 00, LEN, ~LEN, then the filtered stream itself
            data_bytes = (zhdr ? 2u : 0u) + 5u + strip_len;
            if (tid == 0) {
                out8[data0 + 1] = (uint8_t)strip_len;
                out8[data0 + 2] = (uint8_t)(strip_len >> 8);
                out8[data0 + 3] = (uint8_t)~strip_len;
                out8[data0 + 4] = (uint8_t)(~strip_len >> 8);
            }
            __syncthreads();   // the header bytes share words with the first data bytes: byte stores, one writer each
            for (int q = tid; q < nsegs; q += kPngThreads) {
                const int r = q / g.segs_per_row, s = q - r * g.segs_per_row;
                const int n = min(kSeg, g.row_bytes - kSeg * s);
                const uint4 f = s_filt[q];
                const uint32_t w[4] = {f.x, f.y, f.z, f.w};
                uint8_t* d = out8 + data0 + 5u + (uint32_t)r * (uint32_t)line;
                if (s == 0) d[0] = 2;
                for (int i = 0; i < n; i++) d[1 + kSeg * s + i] = (uint8_t)seg_byte(w, i);
            }
            __syncthreads();
        }
        // ---- 3. chunk framing: length, "IDAT", zlib header, CRC-32 over type + payload
        if (tid == 0) {
            out8[0] = (uint8_t)(data_bytes >> 24);
            out8[1] = (uint8_t)(data_bytes >> 16);
            out8[2] = (uint8_t)(data_bytes >> 8);
            out8[3] = (uint8_t)data_bytes;
            out8[4] = 'I'; out8[5] = 'D'; out8[6] = 'A'; out8[7] = 'T';
            if (zhdr) {
                out8[8] = 0x78;   // deflate, 32 KB window
                out8[9] = 0x01;   // no preset dictionary, check bits
            }
        }
        __syncthreads();
        // message = bytes [4, 8 + data_bytes); thread j owns the j-th 64-byte piece counted from the END, so that every
        // piece but the first of the message is whole.  CRC-32's initial value is folded in by complementing the
        // first four message bytes, which leaves a linear (raw) remainder: pieces are computed independently, moved to
        // their place by a multiplication with x^(8 * 64 * j) and summed.
        {
            const int msg = (int)data_bytes + 4;
            const int hi = msg - tid * kPiece, lo = max(0, hi - kPiece);
            uint32_t state = 0;
            for (int i = lo; i < hi; i++) {
                const uint32_t byte = (uint32_t)out8[4 + i] ^ (i < 4 ? 0xffu : 0u);
                state = crc_step(s_crc, state, byte);
            }
            uint32_t part = (hi > 0 && state) ? crc_mulmod(state, tb->crc_shift[tid]) : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, d);
            if (lane == 0) s_warp[warp] = part;
            __syncthreads();
            if (tid == 0) {
                uint32_t crc = 0;
                for (int w = 0; w < kPngWarps; w++) crc ^= s_warp[w];
                crc = ~crc;
                uint8_t* c = out8 + 8 + data_bytes;
                c[0] = (uint8_t)(crc >> 24);
                c[1] = (uint8_t)(crc >> 16);
                c[2] = (uint8_t)(crc >> 8);
                c[3] = (uint8_t)crc;
                StripInfo si;
                si.bytes = data_bytes + 12u;
                si.s1 = (uint32_t)s_red[0][4];
                si.s2 = (unsigned long long)strip_len * s_red[0][4] - s_red[0][5];
                info[work] = si;
            }
            __syncthreads();
        }
        // ---- 4. the chunk to its slot
        {
            const uint32_t n16 = (data_bytes + 12u + 15u) / 16u;
            uint4* dst = reinterpret_cast<uint4*>(slots + (size_t)work * kMaxChunk);
            const uint4* src = reinterpret_cast<const uint4*>(s_out);
            for (uint32_t i = tid; i < n16; i += kPngThreads) dst[i] = src[i];
        }
    }
}

// One CTA.  Per frame: where each strip's chunk goes, the frame's size and Adler-32; then the scan of the sizes.
__global__ void __launch_bounds__(1024) png_layout_kernel(int S, Geometry g, const StripInfo* __restrict__ info,
                                                          uint32_t* __restrict__ strip_off,
                                                          uint32_t* __restrict__ frame_adler,
                                                          unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n_total = (unsigned long long)g.height * (g.row_bytes + 1);
    for (int base = 0; base < S; base += 1024) {
        const int f = base + threadIdx.x;
        unsigned long long size = 0;
        if (f < S) {
            uint32_t off = kPngHeaderBytes;
            unsigned long long a = 1, b = n_total % kAdlerMod, end = 0;
            for (int s = 0; s < g.n_strips; s++) {
                const StripInfo si = info[(size_t)f * g.n_strips + s];
                strip_off[(size_t)f * g.n_strips + s] = off;
                off += si.bytes;
                const int rows = min(g.rows_per_strip, g.height - s * g.rows_per_strip);
                end += (unsigned long long)rows * (g.row_bytes + 1);
                a += si.s1;
                b = (b + si.s2 % kAdlerMod + ((unsigned long long)si.s1 % kAdlerMod) * ((n_total - end) % kAdlerMod)) % kAdlerMod;
            }
            frame_adler[f] = (uint32_t)((b % kAdlerMod) << 16) | (uint32_t)(a % kAdlerMod);
            size = (unsigned long long)off + kPngTailBytes;
        }
        unsigned long long inc = size;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long n = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += n;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long start = s_carry + s_warp[warp] + inc - size;
        if (f < S) offsets[f] = start;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = start + size;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[S] = s_carry;
}

// grid (n_strips, S).  Byte-granular copy (final offsets have no alignment); the first strip's CTA also writes
// the file header, the last one the final IDAT (last deflate block 01 00 00 ff ff + Adler-32) and IEND.
__global__ void __launch_bounds__(256) png_pack_kernel(Geometry g, PngHeader header, const Tables* __restrict__ tb,
                                                       const uint8_t* __restrict__ slots,
                                                       const StripInfo* __restrict__ info,
                                                       const uint32_t* __restrict__ strip_off,
                                                       const uint32_t* __restrict__ frame_adler,
                                                       const unsigned long long* __restrict__ offsets,
                                                       uint8_t* __restrict__ png) {
    const int strip = blockIdx.x, f = blockIdx.y;
    const size_t work = (size_t)f * g.n_strips + strip;
    uint8_t* frame_out = png + offsets[f];
    const uint32_t n = info[work].bytes;
    const uint8_t* src = slots + work * kMaxChunk;
    uint8_t* dst = frame_out + strip_off[work];
    // head bytes up to 4-byte alignment of dst, then words re-aligned with a byte permute, then the tail
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
    const uint32_t head = min(n, (4u - mis) & 3u);
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint32_t n_words = (n - head) / 4u;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);   // slots are 16-byte aligned
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t i = threadIdx.x; i < n_words; i += 256) {
        const uint32_t byte0 = head + 4u * i;     // source byte offset of this destination word
        const uint32_t lo = s32[byte0 >> 2], hi = s32[(byte0 >> 2) + 1];
        d32[i] = __funnelshift_r(lo, hi, 8u * (byte0 & 3u));
    }
    const uint32_t done = head + 4u * n_words;
    if (threadIdx.x < n - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
    if (strip == 0 && threadIdx.x < kPngHeaderBytes)
        frame_out[threadIdx.x] = (uint8_t)(header.w[threadIdx.x >> 2] >> (8 * (threadIdx.x & 3)));
    if (strip == g.n_strips - 1 && threadIdx.x == 0) {
        uint8_t tail[kPngTailBytes] = {0, 0, 0, 9, 'I', 'D', 'A', 'T', 1, 0, 0, 0xff, 0xff, 0, 0, 0, 0, 0, 0, 0, 0,
                                       0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
        const uint32_t ad = frame_adler[f];
        tail[13] = (uint8_t)(ad >> 24); tail[14] = (uint8_t)(ad >> 16); tail[15] = (uint8_t)(ad >> 8); tail[16] = (uint8_t)ad;
        uint32_t c = 0xffffffffu;
        for (int i = 4; i < 17; i++) c = tb->crc_byte[(c ^ tail[i]) & 0xffu] ^ (c >> 8);
        c = ~c;
        tail[17] = (uint8_t)(c >> 24); tail[18] = (uint8_t)(c >> 16); tail[19] = (uint8_t)(c >> 8); tail[20] = (uint8_t)c;
        uint8_t* t = dst + n;
        for (int i = 0; i < kPngTailBytes; i++) t[i] = tail[i];
    }
}

// ------------------------------------------------------------------------------------------------- host side
static const Tables& host_tables() {
    static Tables* t = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        t = new Tables();
        build_tables(*t);
    });
    return *t;
}

static int device_tables(const Tables** out) {
    static std::mutex mu;
    static Tables* d_tables[kMaxDevices] = {nullptr};
    int dev = 0;
    OMFS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("png sink: device index %d out of range", dev);
        return OMFS_ERR_INVALID;
    }
    std::lock_guard<std::mutex> guard(mu);
    if (!d_tables[dev]) {
        Tables* p = nullptr;
        OMFS_CUDA(cudaMalloc(&p, sizeof(Tables)));
        OMFS_CUDA(cudaMemcpy(p, &host_tables(), sizeof(Tables), cudaMemcpyHostToDevice));
        d_tables[dev] = p;
    }
    *out = d_tables[dev];
    return OMFS_OK;
}

struct PngWs {
    uint8_t* slots;
    StripInfo* info;
    uint32_t* strip_off;
    uint32_t* frame_adler;
    size_t total;
};

static PngWs carve_png(void* base, int S, const Geometry& g) {
    PngWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? (unsigned char*)base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    const size_t n = (size_t)S * g.n_strips;
    w.slots = (uint8_t*)take(n * kMaxChunk + 16);
    w.info = (StripInfo*)take(n * sizeof(StripInfo));
    w.strip_off = (uint32_t*)take(n * sizeof(uint32_t));
    w.frame_adler = (uint32_t*)take((size_t)S * sizeof(uint32_t));
    w.total = off;
    return w;
}

static size_t png_frame_bound(const Geometry& g) {
    // every strip stored: framing 12 + zlib header 2 + stored header 5 per strip, the filtered bytes, header and tail
    return (size_t)kPngHeaderBytes + kPngTailBytes + (size_t)g.n_strips * 19 + (size_t)g.height * (g.row_bytes + 1);
}

int png_encode_launch(int S, int width, int height, const uint8_t* d_frames, uint8_t* d_png, size_t png_capacity,
                      unsigned long long* d_offsets, void* d_workspace, size_t workspace_bytes, cudaStream_t stream) {
    Geometry g;
    if (!make_geometry(width, height, g)) {
        set_error("png sink: a %d-pixel row does not fit one chunk (width <= 5450)", width);
        return OMFS_ERR_UNSUPPORTED;
    }
    const PngWs w = carve_png(d_workspace, S, g);
    OMFS_REQUIRE(workspace_bytes >= w.total, "workspace too small (omfs_png_workspace_bytes)");
    OMFS_REQUIRE(png_capacity >= (size_t)S * png_frame_bound(g), "output buffer smaller than S * omfs_png_max_bytes");
    const Tables* tb = nullptr;
    int rc = device_tables(&tb);
    if (rc) return rc;
    const size_t smem = (size_t)g.rows_per_strip * g.segs_per_row * 16 + kMaxChunk + 16;
    static DeviceOnce once;
    if ((rc = ensure_dyn_smem(once, png_strip_kernel, (int)smem))) return rc;
    const long long n_work = (long long)S * g.n_strips;
    const int grid = (int)std::min<long long>(n_work, (long long)kNumSMs * 6);
    png_strip_kernel<<<grid, kPngThreads, smem, stream>>>(S, g, d_frames, tb, w.slots, w.info);
    png_layout_kernel<<<1, 1024, 0, stream>>>(S, g, w.info, w.strip_off, w.frame_adler, d_offsets);
    PngHeader header{};
    uint8_t hb[36] = {0};
    make_png_header(host_tables(), width, height, hb);
    memcpy(header.w, hb, 36);
    png_pack_kernel<<<dim3(g.n_strips, S), 256, 0, stream>>>(g, header, tb, w.slots, w.info, w.strip_off, w.frame_adler,
                                                             d_offsets, d_png);
    count_launch(3);
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

}  // namespace omfs

using namespace omfs;

extern "C" size_t omfs_png_max_bytes(int width, int height) {
    Geometry g;
    if (width <= 0 || height <= 0 || !make_geometry(width, height, g)) return 0;
    return png_frame_bound(g);
}

extern "C" size_t omfs_png_workspace_bytes(int S, int width, int height) {
    Geometry g;
    if (S <= 0 || width <= 0 || height <= 0 || !make_geometry(width, height, g)) return 0;
    return carve_png(nullptr, S, g).total;
}

extern "C" int omfs_png_encode(int S, int width, int height, const uint8_t* d_frames_u8, uint8_t* d_png,
                               size_t png_capacity, uint64_t* d_offsets, void* d_workspace, size_t workspace_bytes,
                               void* stream) {
    OMFS_REQUIRE(S >= 0 && width > 0 && height > 0, "bad sizes");
    OMFS_REQUIRE(S <= 65535, "at most 65535 frames per call");
    OMFS_REQUIRE(d_frames_u8 && d_png && d_offsets && d_workspace, "null pointer");
    if (S == 0) return OMFS_OK;
    return png_encode_launch(S, width, height, d_frames_u8, d_png, png_capacity, (unsigned long long*)d_offsets,
                             d_workspace, workspace_bytes, (cudaStream_t)stream);
}
