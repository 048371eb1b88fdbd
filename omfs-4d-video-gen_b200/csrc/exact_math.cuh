// exact_math.cuh — per-element arithmetic of the exact domain.
//
// Every expression here is evaluated as individually rounded IEEE binary32 operations, left to
// right as written.  The translation units that include this file are compiled with
// --fmad=false, so nvcc never contracts a*b+c; fused multiply-adds appear only where fmaf() is
// spelled out.  Division and square root are the IEEE-rounded forms (nvcc defaults
// -prec-div=true -prec-sqrt=true, no -use_fast_math anywhere in this library).
//
// The functions are __host__ __device__ so that tests/emu (a host harness, test-only) can run the
// very same source on the CPU and compare it with the independent C oracle without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define OMFS_HD __host__ __device__ __forceinline__
#else
#define OMFS_HD static inline
#endif

namespace omfs {

struct Float4 {
    float x, y, z, w;
};

OMFS_HD float ex_dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ax * bx + ay * by + az * bz;
}

OMFS_HD int32_t float_as_i32(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int32_t i;
    memcpy(&i, &f, 4);
    return i;
#endif
}

OMFS_HD float i32_as_float(int32_t i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

OMFS_HD void ex_safe_normalize(float x, float y, float z, float& ox, float& oy, float& oz, float& len) {
    const float n2 = fmaxf(ex_dot3(x, y, z, x, y, z), 1e-20f);
    const float l = sqrtf(n2);
    ox = x / l;
    oy = y / l;
    oz = z / l;
    len = l;
}

// rotation matrix (row-major) -> unit quaternion wxyz.  Branch on the largest of
// {m00, m11, m22, trace}; ties go to the lowest index.
OMFS_HD void ex_rotmat_to_quat(const float m[9], float q[4]) {
    const float trace = (m[0] + m[4]) + m[8];
    int ch = 0;
    float best = m[0];
    if (m[4] > best) { best = m[4]; ch = 1; }
    if (m[8] > best) { best = m[8]; ch = 2; }
    if (trace > best) { best = trace; ch = 3; }
    float vx, vy, vz, vw;
    if (ch == 3) {
        vx = m[7] - m[5];
        vy = m[2] - m[6];
        vz = m[3] - m[1];
        vw = 1.0f + trace;
    } else if (ch == 0) {  // i=0 j=1 k=2
        vx = (1.0f - trace) + 2.0f * m[0];
        vy = m[3] + m[1];
        vz = m[6] + m[2];
        vw = m[7] - m[5];
    } else if (ch == 1) {  // i=1 j=2 k=0
        vy = (1.0f - trace) + 2.0f * m[4];
        vz = m[7] + m[5];
        vx = m[1] + m[3];
        vw = m[2] - m[6];
    } else {  // i=2 j=0 k=1
        vz = (1.0f - trace) + 2.0f * m[8];
        vx = m[2] + m[6];
        vy = m[5] + m[7];
        vw = m[3] - m[1];
    }
    const float n = sqrtf(((vx * vx + vy * vy) + vz * vz) + vw * vw);
    q[0] = vw / n;
    q[1] = vx / n;
    q[2] = vy / n;
    q[3] = vz / n;
}

// U4.  out[20] = [cx cy cz s | qw qx qy qz | R00 R01 R02 0 | R10 R11 R12 0 | R20 R21 R22 0]
OMFS_HD void ex_face_frame(const float p0[3], const float p1[3], const float p2[3], float out[20]) {
    const float e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
    const float e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
    float a0x, a0y, a0z, l1, a1x, a1y, a1z, a2x, a2y, a2z, l;
    ex_safe_normalize(e1x, e1y, e1z, a0x, a0y, a0z, l1);
    float cx = a0y * e2z - a0z * e2y, cy = a0z * e2x - a0x * e2z, cz = a0x * e2y - a0y * e2x;
    ex_safe_normalize(cx, cy, cz, a1x, a1y, a1z, l);
    cx = a1y * a0z - a1z * a0y;
    cy = a1z * a0x - a1x * a0z;
    cz = a1x * a0y - a1y * a0x;
    ex_safe_normalize(cx, cy, cz, a2x, a2y, a2z, l);
    a2x = -a2x;
    a2y = -a2y;
    a2z = -a2z;
    const float s1 = fabsf(ex_dot3(a2x, a2y, a2z, e2x, e2y, e2z));
    const float scale = (l1 + s1) * 0.5f;
    const float R[9] = {a0x, a1x, a2x, a0y, a1y, a2y, a0z, a1z, a2z};
    float q[4];
    ex_rotmat_to_quat(R, q);
    out[0] = ((p0[0] + p1[0]) + p2[0]) / 3.0f;
    out[1] = ((p0[1] + p1[1]) + p2[1]) / 3.0f;
    out[2] = ((p0[2] + p1[2]) + p2[2]) / 3.0f;
    out[3] = scale;
    out[4] = q[0];
    out[5] = q[1];
    out[6] = q[2];
    out[7] = q[3];
    out[8] = R[0];  out[9] = R[1];  out[10] = R[2]; out[11] = 0.0f;
    out[12] = R[3]; out[13] = R[4]; out[14] = R[5]; out[15] = 0.0f;
    out[16] = R[6]; out[17] = R[7]; out[18] = R[8]; out[19] = 0.0f;
}

// tile rectangle of a splat centred at (px,py) with integer radius r, clamped to the grid.  The quotients are
// clamped to [-1, g+1] BEFORE the conversion to int (same rectangle: the int clamp below absorbs the ends), so
// that a far-away centre never converts an out-of-range float — undefined in C, and different on CPU and GPU.
OMFS_HD void ex_tile_rect(float px, float py, int radius, int gx, int gy, int& minx, int& miny, int& maxx,
                          int& maxy) {
    const float rf = (float)radius;
    const float hx = (float)gx + 1.0f, hy = (float)gy + 1.0f;
    const int a = (int)fminf(fmaxf((px - rf) / 16.0f, -1.0f), hx), b = (int)fminf(fmaxf((py - rf) / 16.0f, -1.0f), hy);
    const int c = (int)fminf(fmaxf((px + rf + 16.0f - 1.0f) / 16.0f, -1.0f), hx);
    const int d = (int)fminf(fmaxf((py + rf + 16.0f - 1.0f) / 16.0f, -1.0f), hy);
    minx = a < 0 ? 0 : (a > gx ? gx : a);
    miny = b < 0 ? 0 : (b > gy ? gy : b);
    maxx = c < 0 ? 0 : (c > gx ? gx : c);
    maxy = d < 0 ? 0 : (d > gy ? gy : d);
}

constexpr float kShC0 = 0.28209479177387814f;
constexpr float kShC1 = 0.4886025119029199f;
constexpr float kShC2_0 = 1.0925484305920792f, kShC2_1 = -1.0925484305920792f, kShC2_2 = 0.31539156525252005f,
                kShC2_3 = -1.0925484305920792f, kShC2_4 = 0.5462742152960396f;
constexpr float kShC3_0 = -0.5900435899266435f, kShC3_1 = 2.890611442640554f, kShC3_2 = -0.4570457994644658f,
                kShC3_3 = 0.3731763325901154f, kShC3_4 = -0.4570457994644658f, kShC3_5 = 1.445305721320277f,
                kShC3_6 = -0.5900435899266435f;
constexpr float kConA = -0.72134752044448170368f;  // -0.5 * log2(e)
constexpr float kConB = -1.44269504088896340736f;  // -log2(e)
constexpr float kLog2Inv255 = -7.99435343685885793770f;

// SH basis for a unit direction (x,y,z): bs[16]
OMFS_HD void ex_sh_basis(float x, float y, float z, float bs[16]) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    bs[0] = kShC0;
    bs[1] = -kShC1 * y;
    bs[2] = kShC1 * z;
    bs[3] = -kShC1 * x;
    bs[4] = kShC2_0 * xy;
    bs[5] = kShC2_1 * yz;
    bs[6] = kShC2_2 * (2.0f * zz - xx - yy);
    bs[7] = kShC2_3 * xz;
    bs[8] = kShC2_4 * (xx - yy);
    bs[9] = kShC3_0 * y * (3.0f * xx - yy);
    bs[10] = kShC3_1 * xy * z;
    bs[11] = kShC3_2 * y * (4.0f * zz - xx - yy);
    bs[12] = kShC3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    bs[13] = kShC3_4 * x * (4.0f * zz - xx - yy);
    bs[14] = kShC3_5 * z * (xx - yy);
    bs[15] = kShC3_6 * x * (xx - 3.0f * yy);
}

struct BindPre {
    // geometry that survives culling
    float px, py, depth;
    int radius;
    float ca, cb, cc;
    uint32_t tiles;
    int minx, miny, maxx, maxy;  // the tile rectangle `tiles` is the area of
    // bind result (for SH direction and debugging)
    float mx, my, mz;
};

// U5 + the geometric half of U6.  ff = the parent face's 20-float record.  Returns false when the
// Gaussian is culled (behind the near limit, degenerate covariance or no tile touched).
OMFS_HD bool ex_bind_project(const float ff[20], float lx, float ly, float lz, float sa0, float sa1, float sa2,
                             float w2, float x2, float y2, float z2, const float* cam, int width, int height,
                             int gx, int gy, BindPre& o) {
    const float* Vm = cam;
    const float* Pm = cam + 16;
    const float tanx = cam[35], tany = cam[36], fx = cam[37], fy = cam[38];
    const float fs = ff[3];
    const float tx = ff[8] * lx + ff[9] * ly + ff[10] * lz;
    const float ty = ff[12] * lx + ff[13] * ly + ff[14] * lz;
    const float tz = ff[16] * lx + ff[17] * ly + ff[18] * lz;
    const float mx = tx * fs + ff[0], my = ty * fs + ff[1], mz = tz * fs + ff[2];
    o.mx = mx;
    o.my = my;
    o.mz = mz;
    o.tiles = 0;
    o.radius = 0;
    const float s0 = sa0 * fs, s1 = sa1 * fs, s2 = sa2 * fs;
    const float w1 = ff[4], x1 = ff[5], y1 = ff[6], z1 = ff[7];
    const float qr = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
    const float qx = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
    const float qy = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
    const float qz = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;

    const float vx = Vm[0] * mx + Vm[4] * my + Vm[8] * mz + Vm[12];
    const float vy = Vm[1] * mx + Vm[5] * my + Vm[9] * mz + Vm[13];
    const float vz = Vm[2] * mx + Vm[6] * my + Vm[10] * mz + Vm[14];
    if (!(vz > 0.2f)) return false;  // the published near-plane cull; written so that a NaN depth is culled too
    const float hx = Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12];
    const float hy = Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13];
    const float hw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ppx = hx * pw, ppy = hy * pw;

    const float R00 = 1.0f - 2.0f * (qy * qy + qz * qz), R01 = 2.0f * (qx * qy - qr * qz),
                R02 = 2.0f * (qx * qz + qr * qy);
    const float R10 = 2.0f * (qx * qy + qr * qz), R11 = 1.0f - 2.0f * (qx * qx + qz * qz),
                R12 = 2.0f * (qy * qz - qr * qx);
    const float R20 = 2.0f * (qx * qz - qr * qy), R21 = 2.0f * (qy * qz + qr * qx),
                R22 = 1.0f - 2.0f * (qx * qx + qy * qy);
    const float M00 = R00 * s0, M01 = R01 * s1, M02 = R02 * s2;
    const float M10 = R10 * s0, M11 = R11 * s1, M12 = R12 * s2;
    const float M20 = R20 * s0, M21 = R21 * s1, M22 = R22 * s2;
    const float S00 = M00 * M00 + M01 * M01 + M02 * M02;
    const float S01 = M00 * M10 + M01 * M11 + M02 * M12;
    const float S02 = M00 * M20 + M01 * M21 + M02 * M22;
    const float S11 = M10 * M10 + M11 * M11 + M12 * M12;
    const float S12 = M10 * M20 + M11 * M21 + M12 * M22;
    const float S22 = M20 * M20 + M21 * M21 + M22 * M22;

    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float txtz = vx / vz, tytz = vy / vz;
    const float cxv = fminf(limx, fmaxf(-limx, txtz)) * vz;
    const float cyv = fminf(limy, fmaxf(-limy, tytz)) * vz;
    const float j00 = fx / vz, j02 = -(fx * cxv) / (vz * vz);
    const float j11 = fy / vz, j12 = -(fy * cyv) / (vz * vz);
    const float T00 = j00 * Vm[0] + j02 * Vm[2], T01 = j00 * Vm[4] + j02 * Vm[6], T02 = j00 * Vm[8] + j02 * Vm[10];
    const float T10 = j11 * Vm[1] + j12 * Vm[2], T11 = j11 * Vm[5] + j12 * Vm[6], T12 = j11 * Vm[9] + j12 * Vm[10];
    const float u0 = S00 * T00 + S01 * T01 + S02 * T02;
    const float u1 = S01 * T00 + S11 * T01 + S12 * T02;
    const float u2 = S02 * T00 + S12 * T01 + S22 * T02;
    const float w0 = S00 * T10 + S01 * T11 + S02 * T12;
    const float w1v = S01 * T10 + S11 * T11 + S12 * T12;
    const float w2v = S02 * T10 + S12 * T11 + S22 * T12;
    const float c00 = (T00 * u0 + T01 * u1 + T02 * u2) + 0.3f;
    const float c01 = T10 * u0 + T11 * u1 + T12 * u2;
    const float c11 = (T10 * w0 + T11 * w1v + T12 * w2v) + 0.3f;
    const float det = c00 * c11 - c01 * c01;
    if (!(fabsf(det) > 0.0f)) return false;  // published: det == 0; a NaN determinant is culled as well
    const float det_inv = 1.0f / det;
    const float conx = c11 * det_inv, cony = -c01 * det_inv, conz = c00 * det_inv;
    const float mid = 0.5f * (c00 + c11);
    const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lam1 = mid + disc, lam2 = mid - disc;
    const float rad_f = ceilf(3.0f * sqrtf(fmaxf(lam1, lam2)));
    const float px = ((ppx + 1.0f) * (float)width - 1.0f) * 0.5f;
    const float py = ((ppy + 1.0f) * (float)height - 1.0f) * 0.5f;
    // degenerate inputs (NaN / infinite parameters, a zero-area parent triangle) are culled here, explicitly: the
    // sum is finite exactly when every term is.  Nothing below depends on it for finite inputs.
    if (!(fabsf(px) + fabsf(py) + rad_f + fabsf(conx) + fabsf(cony) + fabsf(conz) + vz < 3.0e38f)) return false;
    const int radius = (int)fminf(rad_f, 1.0e6f);
    int minx, miny, maxx, maxy;
    ex_tile_rect(px, py, radius, gx, gy, minx, miny, maxx, maxy);
    const uint32_t tt = (uint32_t)((maxx - minx) * (maxy - miny));
    if (tt == 0) return false;
    o.px = px;
    o.py = py;
    o.depth = vz;
    o.radius = radius;
    o.ca = kConA * conx;
    o.cb = kConB * cony;
    o.cc = kConA * conz;
    o.tiles = tt;
    o.minx = minx;
    o.miny = miny;
    o.maxx = maxx;
    o.maxy = maxy;
    return true;
}

// view direction for SH: (mu - campos) / |mu - campos|
OMFS_HD void ex_view_dir(float mx, float my, float mz, const float* cam, float& dx, float& dy, float& dz) {
    dx = mx - cam[32];
    dy = my - cam[33];
    dz = mz - cam[34];
    const float dl = sqrtf(ex_dot3(dx, dy, dz, dx, dy, dz));
    dx = dx / dl;
    dy = dy / dl;
    dz = dz / dl;
}

// one pixel x one Gaussian of U10.  Returns 0 = skip, 1 = blended, 2 = pixel saturated (stop).
// exp2 is the only non-reproducible operation (device: ex2.approx, oracle: exp2f).
//
// Canonical order of the exponent e = lo + ca dx^2 + cb dx dy + cc dy^2 (conic pre-scaled by log2 e,
// lo = log2(opacity)): the terms that depend only on the pixel COLUMN, c0 = lo + (ca dx) dx and
// v = cb dx, are formed first, then e = (cc dy + v) dy + c0 — two fused operations per pixel once the
// column terms exist (the compositing kernel shares them between the pixels of one column).  The published
// `power > 0` rejection (a rounding artefact of a negative-definite form) reads `e > lo` here.
template <typename Exp2>
OMFS_HD int ex_blend(float gx, float gy, float ca, float cb, float cc, float lo, float r, float g, float b,
                     float pxf, float pyf, float& T, float& C0, float& C1, float& C2, Exp2 exp2_fn) {
    const float dx = gx - pxf, dy = gy - pyf;
    const float u = ca * dx;
    const float c0 = fmaf(u, dx, lo);
    const float v = cb * dx;
    const float s = fmaf(cc, dy, v);
    const float e = fmaf(s, dy, c0);
    if (!((e <= lo) & (e >= kLog2Inv255))) return 0;  // one combined predicate; a NaN exponent (NaN opacity) skips
    const float alpha = fminf(0.99f, exp2_fn(e));
    const float testT = T * (1.0f - alpha);
    if (testT < 0.0001f) return 2;
    const float w = alpha * T;
    C0 = fmaf(r, w, C0);
    C1 = fmaf(g, w, C1);
    C2 = fmaf(b, w, C2);
    T = testT;
    return 1;
}

}  // namespace omfs
