// png_emu.cpp — TEST INFRASTRUCTURE.  A sequential host encoder assembled from the product's own
// png_core.cuh (code tables, tokenisation, CRC/Adler algebra: the same source png.cu compiles), following
// the kernel's steps one "thread" at a time.  `-m "not gpu"` tests decode its output with zlib / PIL, which
// pins the stream format, the tables and the checksum algebra without a GPU; the GPU test then asserts that
// the kernel produces the very same bytes.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../omfs-4d-video-gen_b200/csrc/png_core.cuh"

using namespace omfs_png;

static const Tables& tables() {
    static Tables* t = nullptr;
    if (!t) {
        t = new Tables();
        build_tables(*t);
    }
    return *t;
}

static void put_bits(std::vector<uint8_t>& out, uint32_t& pos, uint32_t value, int n) {
    for (int i = 0; i < n; i++, pos++) {
        if ((pos >> 3) >= out.size()) out.resize((pos >> 3) + 1, 0);
        out[pos >> 3] |= (uint8_t)(((value >> i) & 1u) << (pos & 7));
    }
}

// CRC-32 of msg the way the kernel computes it: 64-byte pieces counted from the end, raw remainders, moved by
// polynomial multiplication, initial value folded into the first four bytes
static uint32_t crc_by_pieces(const Tables& t, const uint8_t* msg, int n) {
    uint32_t total = 0;
    for (int j = 0; j * kPiece < n; j++) {
        const int hi = n - j * kPiece, lo = hi - kPiece > 0 ? hi - kPiece : 0;
        uint32_t state = 0;
        for (int i = lo; i < hi; i++) state = crc_step(t.crc_byte, state, (uint32_t)msg[i] ^ (i < 4 ? 0xffu : 0u));
        total ^= crc_mulmod(state, t.crc_shift[j]);
    }
    return ~total;
}

extern "C" int emu_png_geometry(int width, int height, int* out6) {
    Geometry g;
    if (!make_geometry(width, height, g)) return -1;
    out6[0] = g.width; out6[1] = g.height; out6[2] = g.row_bytes; out6[3] = g.segs_per_row;
    out6[4] = g.rows_per_strip; out6[5] = g.n_strips;
    return 0;
}

// table choice per strip is reported through `choices` (n_strips entries, kNumTables = stored) when not NULL
extern "C" long long emu_png_encode(int width, int height, const uint8_t* frame, uint8_t* out, long long capacity,
                                    int* choices) {
    const Tables& t = tables();
    Geometry g;
    if (!make_geometry(width, height, g)) return -1;
    std::vector<uint8_t> png(kPngHeaderBytes);
    make_png_header(t, width, height, png.data());
    const int line = g.row_bytes + 1;
    const unsigned long long n_total = (unsigned long long)height * line;
    unsigned long long A = 1, B = n_total % 65521u, end = 0;
    for (int strip = 0; strip < g.n_strips; strip++) {
        const int row0 = strip * g.rows_per_strip;
        const int rows = g.rows_per_strip < height - row0 ? g.rows_per_strip : height - row0;
        const int nsegs = rows * g.segs_per_row;
        const uint32_t strip_len = (uint32_t)rows * line;
        // 1a. filter
        std::vector<uint32_t> filt((size_t)nsegs * 4, 0);
        for (int q = 0; q < nsegs; q++) {
            const int r = q / g.segs_per_row, s = q % g.segs_per_row;
            const int n = g.row_bytes - kSeg * s < kSeg ? g.row_bytes - kSeg * s : kSeg;
            for (int i = 0; i < n; i++) {
                const size_t at = (size_t)(row0 + r) * g.row_bytes + kSeg * s + i;
                const uint8_t cur = frame[at], up = (row0 + r) ? frame[at - g.row_bytes] : 0;
                filt[(size_t)q * 4 + (i >> 2)] |= (uint32_t)(uint8_t)(cur - up) << (8 * (i & 3));
            }
        }
        // 1b. costs and Adler sums
        unsigned long long cost[kNumTables] = {0}, sum_d = 0, sum_dpos = 0;
        std::vector<uint32_t> pa(nsegs), pb(nsegs);
        for (int q = 0; q < nsegs; q++) {
            const int r = q / g.segs_per_row, s = q % g.segs_per_row;
            const int n = g.row_bytes - kSeg * s < kSeg ? g.row_bytes - kSeg * s : kSeg;
            const uint32_t* w = &filt[(size_t)q * 4];
            const uint32_t prev = s ? (filt[(size_t)(q - 1) * 4 + 3] >> 24) : 2u;
            uint32_t a = 0, b = 0;
            if (s == 0) { a = t.lens_a[2]; b = t.lens_b[2]; }
            for_each_token(w, prev, n, [&](int tok) { a += t.lens_a[tok]; b += t.lens_b[tok]; });
            pa[q] = a; pb[q] = b;
            for (int k = 0; k < 4; k++) { cost[k] += (a >> (8 * k)) & 0xffu; cost[4 + k] += (b >> (8 * k)) & 0xffu; }
            uint32_t s1, s2;
            seg_sums(w, s1, s2);
            const uint32_t pos0 = (uint32_t)r * line + 1u + kSeg * s;
            sum_d += s1;
            sum_dpos += (unsigned long long)pos0 * s1 + s2;
            if (s == 0) { sum_d += 2; sum_dpos += 2ull * ((unsigned long long)r * line); }
        }
        uint32_t best = 0, best_bits = 0xffffffffu;
        for (int k = 0; k < kNumTables; k++) {
            const uint32_t bits = t.hdr_bits[k] + (uint32_t)cost[k] + (t.token[k][256] >> 24);
            if (bits < best_bits) { best_bits = bits; best = k; }
        }
        const uint32_t huff_bytes = (best_bits + 3u + 7u) / 8u + 4u, stored_bytes = 5u + strip_len;
        const uint32_t table = stored_bytes <= huff_bytes ? (uint32_t)kNumTables : best;
        if (choices) choices[strip] = (int)table;
        // 2. payload
        std::vector<uint8_t> chunk(8, 0);
        const bool zhdr = strip == 0;
        if (zhdr) { chunk.push_back(0x78); chunk.push_back(0x01); }
        if (table < (uint32_t)kNumTables) {
            uint32_t pos = (uint32_t)chunk.size() * 8u;
            const uint32_t hb = t.hdr_bits[table];
            for (uint32_t i = 0; i * 32u < hb; i++) put_bits(chunk, pos, t.hdr[table][i], (int)(hb - 32u * i < 32u ? hb - 32u * i : 32u));
            for (int q = 0; q < nsegs; q++) {
                const int s = q % g.segs_per_row;
                const int n = g.row_bytes - kSeg * s < kSeg ? g.row_bytes - kSeg * s : kSeg;
                const uint32_t* w = &filt[(size_t)q * 4];
                const uint32_t prev = s ? (filt[(size_t)(q - 1) * 4 + 3] >> 24) : 2u;
                const uint32_t before = pos;
                auto emit = [&](int tok) { const uint32_t e = t.token[table][tok]; put_bits(chunk, pos, e & 0xffffffu, (int)(e >> 24)); };
                if (s == 0) emit(2);
                for_each_token(w, prev, n, emit);
                const uint32_t packed = table < 4 ? pa[q] : pb[q];
                if (pos - before != ((packed >> (8 * (table & 3u))) & 0xffu)) return -2;   // the scan input must be exact
            }
            const uint32_t e = t.token[table][256];
            put_bits(chunk, pos, e & 0xffffffu, (int)(e >> 24));
            pos += 3;                       // empty stored block: BFINAL=0, BTYPE=00
            pos = (pos + 7u) & ~7u;
            chunk.resize(pos / 8, 0);
            chunk.push_back(0); chunk.push_back(0); chunk.push_back(0xff); chunk.push_back(0xff);
        } else {
            chunk.push_back(0);
            chunk.push_back((uint8_t)strip_len); chunk.push_back((uint8_t)(strip_len >> 8));
            chunk.push_back((uint8_t)~strip_len); chunk.push_back((uint8_t)(~strip_len >> 8));
            for (int r = 0; r < rows; r++) {
                chunk.push_back(2);
                for (int c = 0; c < g.row_bytes; c++) {
                    const int q = r * g.segs_per_row + c / kSeg;
                    chunk.push_back((uint8_t)seg_byte(&filt[(size_t)q * 4], c % kSeg));
                }
            }
        }
        const uint32_t data_bytes = (uint32_t)chunk.size() - 8u;
        if (data_bytes + 12u > (uint32_t)kMaxChunk) return -3;
        put_be32(chunk.data(), data_bytes);
        memcpy(chunk.data() + 4, "IDAT", 4);
        const uint32_t crc = crc_by_pieces(t, chunk.data() + 4, (int)data_bytes + 4);
        uint8_t c4[4];
        put_be32(c4, crc);
        chunk.insert(chunk.end(), c4, c4 + 4);
        png.insert(png.end(), chunk.begin(), chunk.end());
        // Adler combine, as png_layout_kernel does
        end += strip_len;
        const unsigned long long s2 = (unsigned long long)strip_len * sum_d - sum_dpos;
        A += sum_d;
        B = (B + s2 % 65521u + (sum_d % 65521u) * ((n_total - end) % 65521u)) % 65521u;
    }
    uint8_t tail[kPngTailBytes] = {0, 0, 0, 9, 'I', 'D', 'A', 'T', 1, 0, 0, 0xff, 0xff, 0, 0, 0, 0, 0, 0, 0, 0,
                                   0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
    put_be32(tail + 13, (uint32_t)((B % 65521u) << 16) | (uint32_t)(A % 65521u));
    put_be32(tail + 17, crc32_bytes(t, tail + 4, 13));
    png.insert(png.end(), tail, tail + kPngTailBytes);
    if ((long long)png.size() > capacity) return -4;
    memcpy(out, png.data(), png.size());
    return (long long)png.size();
}

extern "C" int emu_png_table_summary(int k, int* lens271, int* hdr_bits) {
    const Tables& t = tables();
    if (k < 0 || k >= kNumTables) return -1;
    for (int i = 0; i < kNumTokens; i++) lens271[i] = (int)(t.token[k][i] >> 24);
    *hdr_bits = (int)t.hdr_bits[k];
    return 0;
}
