// png_core.cuh — the format side of the device frame sink (png.cu): code tables, tokenisation of a
// filtered segment, CRC-32 / Adler-32 algebra.  Everything here is plain integer arithmetic; the functions
// are __host__ __device__ so that tests/emu can run the SAME source on the CPU (a sequential encoder built
// from these pieces, checked against zlib / PIL without a GPU), and the GPU test then only has to show that
// the kernel's parallel plumbing reproduces that byte stream exactly.
//
// What it replaces: the frame sink of the reference — upstream render.py's `save_image` per frame, then
// 02_Visual_Engine/render_surgery.py:412-449 (`stitch_video`: copy every PNG, ffmpeg) — SURVEY.md §8(f2).
//
// Stream produced per frame (a standard 8-bit RGB PNG, readable by any decoder):
//   signature, IHDR, one IDAT chunk per STRIP of rows, a final 21-byte IDAT (last deflate block + Adler-32), IEND.
//   Every row uses filter type 2 (Up).  A strip's filtered bytes are one deflate block, byte-aligned at its
//   end by an empty stored block (the "sync flush" form), so strips are compressed independently and their
//   chunks are simply concatenated.  A strip block is either
//     * dynamic Huffman with one of kNumTables FIXED code tables (a family of codes for Laplacian residuals of
//       increasing scale, built once on the host by an ordinary length-limited Huffman construction), picked
//       per strip by exact cost — so the kernel needs no histogram and no tree construction — plus run-length
//       matches (distance 1, length 3..16 inside a 16-byte segment), or
//     * stored, when no table beats the raw bytes.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PNG_HD __host__ __device__ __forceinline__
#else
#define PNG_HD static inline
#endif

namespace omfs_png {

constexpr int kSeg = 16;            // raw bytes per segment (one thread's unit of work)
constexpr int kNumLitLen = 268;     // deflate literal/length alphabet in use: 0..255, 256 = end of block, 257..267 = lengths 3..16
constexpr int kNumTokens = 271;     // table entries: 0..255 literal, 256 end of block, 257 + (len - 3) for a match of len 3..16
constexpr int kNumTables = 8;
constexpr int kMaxHdrWords = 48;    // dynamic-block header (3 + 14 + 57 + code lengths) <= 1536 bits
constexpr int kPiece = 64;          // bytes of a chunk one thread checksums (CRC-32)
constexpr int kMaxPieces = 256;
constexpr int kMaxChunk = kPiece * kMaxPieces;   // 16 KB: upper bound of one strip's IDAT chunk
constexpr uint32_t kCrcPoly = 0xedb88320u;

struct Tables {
    uint32_t token[kNumTables][kNumTokens];   // bits 0..23: code (+ extra bit + distance bit), LSB first; bits 24..31: length
    uint32_t lens_a[kNumTokens];              // token lengths of tables 0..3, one byte each
    uint32_t lens_b[kNumTokens];              // tables 4..7
    uint32_t hdr[kNumTables][kMaxHdrWords];   // block header: BFINAL=0, BTYPE=2, HLIT, HDIST, HCLEN, code lengths
    uint32_t hdr_bits[kNumTables];
    uint32_t crc_byte[256];                   // the byte-wise CRC-32 table
    uint32_t crc_shift[kMaxPieces];           // x^(8 * kPiece * j) mod P: moves a piece's remainder j pieces towards the front
};

// ----------------------------------------------------------------------------------------- CRC-32 algebra
// Remainders are in the reflected representation zlib uses (x^0 is bit 31).  a(x) * b(x) mod P.
PNG_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 1
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}

// register state after one more message byte (state 0, no final xor: the "raw" remainder, which is linear)
PNG_HD uint32_t crc_step(const uint32_t* table, uint32_t state, uint32_t byte) {
    return table[(state ^ byte) & 0xffu] ^ (state >> 8);
}

// ----------------------------------------------------------------------------------------- tokenisation
// 16 filtered bytes as four little-endian words.  eq bit i: byte i equals the byte before it in the stream
// (`prev` for byte 0).  Only the first n bytes are valid.
PNG_HD uint32_t eq_mask16(const uint32_t w[4], uint32_t prev, int n) {
    uint32_t mask = 0;
    uint32_t carry = prev & 0xffu;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t shifted = (w[k] << 8) | carry;   // each byte's predecessor
        carry = w[k] >> 24;
        const uint32_t x = w[k] ^ shifted;
        // per byte: 0x80 where the byte of x is zero
        const uint32_t z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
        const uint32_t t = z >> 7;
        mask |= (((t * 0x01020408u) >> 24) & 0xfu) << (4 * k);
    }
    return n >= 16 ? mask : (mask & ((1u << n) - 1u));
}

// Tokens of a segment from its eq mask: a maximal run of >= 3 bytes equal to their predecessor is ONE match
// (distance 1) starting at the run's first byte; every other byte is a literal.
//   tokens  bit i: a token starts at byte i        in_run  bit i: byte i belongs to a match
PNG_HD void token_masks(uint32_t eq, int n, uint32_t& tokens, uint32_t& in_run) {
    const uint32_t valid = n >= 16 ? 0xffffu : ((1u << n) - 1u);
    const uint32_t head = eq & (eq >> 1) & (eq >> 2);          // three in a row start here
    in_run = (head | (head << 1) | (head << 2)) & valid;
    const uint32_t start = in_run & ~(in_run << 1);
    tokens = ((~in_run) & valid) | start;
}

PNG_HD int ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

PNG_HD uint32_t seg_byte(const uint32_t w[4], int i) {
    const uint32_t lo = (i & 8) ? w[2] : w[0], hi = (i & 8) ? w[3] : w[1];
    const uint32_t v = (i & 4) ? hi : lo;
    return (v >> (8 * (i & 3))) & 0xffu;
}

// Visit the tokens of a segment in stream order: f(token index) with token index as in Tables::token.
template <typename F>
PNG_HD void for_each_token(const uint32_t w[4], uint32_t prev, int n, F&& f) {
    uint32_t tokens, in_run;
    token_masks(eq_mask16(w, prev, n), n, tokens, in_run);
    while (tokens) {
        const int i = ctz32(tokens);
        tokens &= tokens - 1u;
        if ((in_run >> i) & 1u) {
            const int len = ctz32(~(in_run >> i));   // run length from its first byte (>= 3, <= 16)
            f(257 + (len - 3));
        } else {
            f((int)seg_byte(w, i));
        }
    }
}

// byte sums of a segment for Adler-32: s1 = sum d_i, s2 = sum i * d_i (i = 0..15 inside the segment)
PNG_HD void seg_sums(const uint32_t w[4], uint32_t& s1, uint32_t& s2) {
    s1 = 0;
    s2 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t b0 = w[k] & 0xffu, b1 = (w[k] >> 8) & 0xffu, b2 = (w[k] >> 16) & 0xffu, b3 = w[k] >> 24;
        s1 += b0 + b1 + b2 + b3;
        s2 += (uint32_t)(4 * k) * (b0 + b1 + b2 + b3) + b1 + 2u * b2 + 3u * b3;
    }
}

// strip geometry shared by the kernel, the host packer and the emulation
struct Geometry {
    int width, height, row_bytes, segs_per_row, rows_per_strip, n_strips;
};

PNG_HD bool make_geometry(int width, int height, Geometry& g) {
    g.width = width;
    g.height = height;
    g.row_bytes = 3 * width;
    g.segs_per_row = (g.row_bytes + kSeg - 1) / kSeg;
    // about 12 KB of filtered bytes per strip, bounded by (a) 1024 segments per strip (4 per thread of the
    // kernel's 256) and (b) the chunk buffer, which must hold the strip even when it is stored:
    // 12 (chunk framing) + 2 (zlib header) + 5 (stored block header) + rows * (row_bytes + 1) <= kMaxChunk
    const int line = g.row_bytes + 1;
    int rows = 12288 / line;
    if (rows < 1) rows = 1;
    if (rows > 1024 / g.segs_per_row) rows = 1024 / g.segs_per_row;
    if (rows > (kMaxChunk - 32) / line) rows = (kMaxChunk - 32) / line;
    if (rows > height) rows = height;
    if (rows < 1) return false;   // a row wider than one chunk: not supported (width <= 5450)
    g.rows_per_strip = rows;
    g.n_strips = (height + rows - 1) / rows;
    return true;
}

constexpr int kPngHeaderBytes = 33;   // signature + IHDR chunk
constexpr int kPngTailBytes = 21 + 12;  // final IDAT (5-byte last block + Adler-32) + IEND

}  // namespace omfs_png

// ------------------------------------------------------------------------------------------- host only
// (plain host functions: visible to both nvcc passes, never called from device code)
#include <algorithm>
#include <cmath>
#include <vector>

namespace omfs_png {

// Huffman code lengths (<= limit) for freq[0..n): ordinary two-queue construction; if the tree is deeper than
// the limit the frequencies are halved (rounding up, so nothing reaches zero) and it is rebuilt.
inline std::vector<int> huffman_lengths(std::vector<uint64_t> freq, int limit) {
    const int n = (int)freq.size();
    std::vector<int> len(n, 0);
    for (;;) {
        struct Node { uint64_t w; int left, right; };
        std::vector<Node> nodes;
        std::vector<int> order;
        for (int i = 0; i < n; i++)
            if (freq[i]) order.push_back(i);
        if (order.size() == 1) { len[order[0]] = 1; return len; }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return freq[a] < freq[b]; });
        for (int i : order) nodes.push_back({freq[i], -1 - i, 0});
        size_t leaf = 0, inner = order.size(), n_leaves = order.size();
        auto take = [&]() -> int {
            const bool has_leaf = leaf < n_leaves, has_inner = inner < nodes.size();
            if (has_leaf && (!has_inner || nodes[leaf].w <= nodes[inner].w)) return (int)leaf++;
            return (int)inner++;
        };
        while ((n_leaves - leaf) + (nodes.size() - inner) > 1) {
            const int a = take(), b = take();
            nodes.push_back({nodes[a].w + nodes[b].w, a, b});
        }
        std::vector<int> depth(nodes.size(), 0);
        int maxd = 0;
        for (int i = (int)nodes.size() - 1; i >= (int)n_leaves; i--) {
            depth[nodes[i].left] = depth[nodes[i].right] = depth[i] + 1;
        }
        for (size_t i = 0; i < n_leaves; i++) {
            len[-1 - nodes[i].left] = depth[i];
            maxd = std::max(maxd, depth[i]);
        }
        if (maxd <= limit) return len;
        for (auto& f : freq)
            if (f) f = (f + 1) / 2;
    }
}

// canonical codes of RFC 1951 section 3.2.2, returned bit-reversed (deflate packs Huffman codes MSB first into an
// LSB-first bit stream)
inline std::vector<uint32_t> canonical_codes_reversed(const std::vector<int>& len) {
    int bl_count[16] = {0};
    for (int l : len) bl_count[l]++;
    bl_count[0] = 0;
    uint32_t next[16] = {0}, code = 0;
    for (int b = 1; b < 16; b++) {
        code = (code + bl_count[b - 1]) << 1;
        next[b] = code;
    }
    std::vector<uint32_t> out(len.size(), 0);
    for (size_t i = 0; i < len.size(); i++) {
        if (!len[i]) continue;
        uint32_t c = next[len[i]]++, r = 0;
        for (int b = 0; b < len[i]; b++) r |= ((c >> b) & 1u) << (len[i] - 1 - b);
        out[i] = r;
    }
    return out;
}

struct BitWriter {
    std::vector<uint32_t> words;
    uint32_t bits = 0;
    void put(uint32_t value, int n) {
        for (int i = 0; i < n; i++, bits++) {
            if ((bits >> 5) >= words.size()) words.push_back(0);
            words[bits >> 5] |= ((value >> i) & 1u) << (bits & 31);
        }
    }
};

inline void build_tables(Tables& t) {
    memset(&t, 0, sizeof(t));
    // residual scale of each table (mean absolute Up-filter residual it is tuned for) and the share of run matches
    static const double scale[kNumTables] = {0.35, 0.8, 1.6, 3.2, 6.4, 12.8, 25.6, 51.2};
    static const double run_share[kNumTables] = {1.0, 0.3, 0.1, 0.05, 0.03, 0.02, 0.01, 0.01};
    for (int k = 0; k < kNumTables; k++) {
        std::vector<uint64_t> freq(kNumLitLen, 0);
        double lit_total = 0;
        for (int v = 0; v < 256; v++) {
            const int s = v < 128 ? v : 256 - v;   // |signed residual|
            const double p = std::exp(-(double)s / scale[k]);
            freq[v] = (uint64_t)std::llround(p * 1e7) + 1;
            lit_total += (double)freq[v];
        }
        freq[256] = 1 + (uint64_t)(lit_total / 12000.0);   // one end-of-block per ~12k symbols
        for (int s = 257; s < kNumLitLen; s++) freq[s] = 1 + (uint64_t)(lit_total * run_share[k] / 20.0);
        freq[267] = 1 + (uint64_t)(lit_total * run_share[k]);   // a whole segment of one value: lengths 15-16
        const std::vector<int> len = huffman_lengths(freq, 15);
        const std::vector<uint32_t> code = canonical_codes_reversed(len);
        for (int v = 0; v <= 256; v++) t.token[k][v] = code[v] | ((uint32_t)len[v] << 24);
        for (int L = 3; L <= 16; L++) {
            // lengths 3..10: codes 257..264, no extra bits; 11..16: codes 265..267 with one extra bit
            const int sym = L <= 10 ? 257 + (L - 3) : 265 + (L - 11) / 2;
            const int extra_n = L <= 10 ? 0 : 1, extra_v = L <= 10 ? 0 : (L - 11) & 1;
            uint32_t bits = code[sym];
            int n = len[sym];
            bits |= (uint32_t)extra_v << n;
            n += extra_n;
            n += 1;   // the single distance code (distance 1), one bit, value 0
            t.token[k][257 + (L - 3)] = bits | ((uint32_t)n << 24);
        }
        for (int i = 0; i < kNumTokens; i++) {
            const uint32_t n = t.token[k][i] >> 24;
            if (k < 4) t.lens_a[i] |= n << (8 * k);
            else t.lens_b[i] |= n << (8 * (k - 4));
        }
        // ---- block header
        std::vector<int> seq(len.begin(), len.end());
        seq.push_back(1);   // HDIST = 0: one distance code, of length 1
        struct Cl { int sym, extra_v, extra_n; };
        std::vector<Cl> cl;
        for (size_t i = 0; i < seq.size();) {
            size_t j = i;
            while (j < seq.size() && seq[j] == seq[i]) j++;
            size_t run = j - i;
            cl.push_back({seq[i], 0, 0});
            run--;
            while (run >= 3) {
                const int r = (int)std::min<size_t>(6, run);
                cl.push_back({16, r - 3, 2});
                run -= r;
            }
            while (run--) cl.push_back({seq[i], 0, 0});
            i = j;
        }
        std::vector<uint64_t> clfreq(19, 0);
        for (const Cl& c : cl) clfreq[c.sym]++;
        const std::vector<int> cllen = huffman_lengths(clfreq, 7);
        const std::vector<uint32_t> clcode = canonical_codes_reversed(cllen);
        BitWriter bw;
        bw.put(0, 1);                   // BFINAL = 0
        bw.put(2, 2);                   // BTYPE = 10: dynamic Huffman
        bw.put(kNumLitLen - 257, 5);    // HLIT
        bw.put(0, 5);                   // HDIST
        bw.put(19 - 4, 4);              // HCLEN: all 19
        static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 19; i++) bw.put((uint32_t)cllen[order[i]], 3);
        for (const Cl& c : cl) {
            bw.put(clcode[c.sym], cllen[c.sym]);
            if (c.extra_n) bw.put((uint32_t)c.extra_v, c.extra_n);
        }
        t.hdr_bits[k] = bw.bits;
        for (size_t i = 0; i < bw.words.size() && i < (size_t)kMaxHdrWords; i++) t.hdr[k][i] = bw.words[i];
        if (bw.words.size() > (size_t)kMaxHdrWords) t.hdr_bits[k] = 0xffffffffu;   // cannot happen; checked by the tests
    }
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int b = 0; b < 8; b++) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
        t.crc_byte[i] = c;
    }
    // x^(8 * kPiece) by squaring x^8 = x^(2^3) up to x^(2^(3 + log2 kPiece))
    uint32_t step = 0x80000000u >> 1;   // x^1
    for (int i = 0; i < 3; i++) step = crc_mulmod(step, step);              // x^8
    for (int p = kPiece; p > 1; p >>= 1) step = crc_mulmod(step, step);     // x^(8 * kPiece)
    t.crc_shift[0] = 0x80000000u;   // x^0
    for (int j = 1; j < kMaxPieces; j++) t.crc_shift[j] = crc_mulmod(t.crc_shift[j - 1], step);
}

inline void put_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

inline uint32_t crc32_bytes(const Tables& t, const uint8_t* p, size_t n) {
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; i++) c = t.crc_byte[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return ~c;
}

// the 33 bytes in front of the first IDAT
inline void make_png_header(const Tables& t, int width, int height, uint8_t out[kPngHeaderBytes]) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    memcpy(out, sig, 8);
    put_be32(out + 8, 13);
    memcpy(out + 12, "IHDR", 4);
    put_be32(out + 16, (uint32_t)width);
    put_be32(out + 20, (uint32_t)height);
    out[24] = 8; out[25] = 2; out[26] = 0; out[27] = 0; out[28] = 0;   // 8-bit, RGB, deflate, adaptive, no interlace
    put_be32(out + 29, crc32_bytes(t, out + 12, 17));
}

}  // namespace omfs_png
