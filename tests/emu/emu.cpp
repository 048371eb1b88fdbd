// emu.cpp — TEST INFRASTRUCTURE.  Runs the product's exact-domain per-element math
// (omfs-4d-video-gen_b200/csrc/exact_math.cuh, the same source the CUDA kernels compile) on the host,
// so that `-m "not gpu"` tests can compare it bit-for-bit with the independent C oracle without a
// GPU.  Built with -ffp-contract=off, the host equivalent of nvcc's --fmad=false.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../omfs-4d-video-gen_b200/csrc/exact_math.cuh"

using namespace omfs;

extern "C" void emu_face_frames(int T, int V, int F, const float* verts, const int32_t* faces, float* ff) {
    for (int t = 0; t < T; t++)
        for (int f = 0; f < F; f++) {
            const float* vb = verts + (size_t)t * V * 3;
            ex_face_frame(vb + 3 * faces[f * 3], vb + 3 * faces[f * 3 + 1], vb + 3 * faces[f * 3 + 2],
                          ff + ((size_t)t * F + f) * 20);
        }
}

extern "C" void emu_bind_preprocess(int N, int F, int width, int height, const float* ff, const float* xyzb,
                                    const float* scale_lo, const float* rot, const float* sh, const float* cam,
                                    float* P0, float* P1, float* P2, uint32_t* tt) {
    const int gx = (width + 15) / 16, gy = (height + 15) / 16;
    (void)F;
    for (int n = 0; n < N; n++) {
        int32_t b;
        memcpy(&b, &xyzb[n * 4 + 3], 4);
        BindPre o;
        const bool ok = ex_bind_project(ff + (size_t)b * 20, xyzb[n * 4], xyzb[n * 4 + 1], xyzb[n * 4 + 2],
                                        scale_lo[n * 4], scale_lo[n * 4 + 1], scale_lo[n * 4 + 2], rot[n * 4],
                                        rot[n * 4 + 1], rot[n * 4 + 2], rot[n * 4 + 3], cam, width, height, gx, gy, o);
        float* o0 = P0 + (size_t)n * 4;
        float* o1 = P1 + (size_t)n * 4;
        float* o2 = P2 + (size_t)n * 4;
        for (int k = 0; k < 4; k++) o0[k] = o1[k] = o2[k] = 0.f;
        tt[n] = 0;
        if (!ok) continue;
        float dx, dy, dz, bs[16];
        ex_view_dir(o.mx, o.my, o.mz, cam, dx, dy, dz);
        ex_sh_basis(dx, dy, dz, bs);
        for (int c = 0; c < 3; c++) {
            float acc = 0.f;
            for (int k = 0; k < 16; k++) {
                const int flat = k * 3 + c;
                const float coef = sh[((size_t)(flat >> 2) * N + n) * 4 + (flat & 3)];
                acc = (k == 0) ? bs[0] * coef : acc + bs[k] * coef;
            }
            acc = acc + 0.5f;
            o2[c] = fmaxf(acc, 0.0f);
        }
        o0[0] = o.px; o0[1] = o.py; o0[2] = o.depth; o0[3] = i32_as_float(o.radius);
        o1[0] = o.ca; o1[1] = o.cb; o1[2] = o.cc; o1[3] = scale_lo[n * 4 + 3];
        tt[n] = o.tiles;
    }
}

struct Exp2Host {
    float operator()(float x) const { return exp2f(x); }
};

extern "C" void emu_composite(int S, int N, int width, int height, const float* P0, const float* P1,
                              const float* P2, const uint32_t* vals, const uint32_t* ranges, const float* bg,
                              float* image) {
    const int gx = (width + 15) / 16, gy = (height + 15) / 16, tiles = gx * gy;
    for (long long gt = 0; gt < (long long)S * tiles; gt++) {
        const int seg = (int)(gt / tiles), tile = (int)(gt % tiles);
        const float* p0 = P0 + (size_t)seg * N * 4;
        const float* p1 = P1 + (size_t)seg * N * 4;
        const float* p2 = P2 + (size_t)seg * N * 4;
        for (int ly = 0; ly < 16; ly++)
            for (int lx = 0; lx < 16; lx++) {
                const int px = (tile % gx) * 16 + lx, py = (tile / gx) * 16 + ly;
                if (px >= width || py >= height) continue;
                float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
                for (uint32_t i = ranges[2 * gt]; i < ranges[2 * gt + 1]; i++) {
                    const uint32_t g = vals[i];
                    const int r = ex_blend(p0[g * 4], p0[g * 4 + 1], p1[g * 4], p1[g * 4 + 1], p1[g * 4 + 2],
                                           p1[g * 4 + 3], p2[g * 4], p2[g * 4 + 1], p2[g * 4 + 2], (float)px,
                                           (float)py, T, C0, C1, C2, Exp2Host());
                    if (r == 2) break;
                }
                const size_t hw = (size_t)width * height, pix = (size_t)py * width + px;
                float* img = image + (size_t)seg * 3 * hw;
                img[pix] = fmaf(T, bg[0], C0);
                img[hw + pix] = fmaf(T, bg[1], C1);
                img[2 * hw + pix] = fmaf(T, bg[2], C2);
            }
    }
}
