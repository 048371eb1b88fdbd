// displace.cu — R5/R6: surgical segment masks and the rigid move of the two mobile segments
// (01_Clinical_Engine/surgical_sim.py:25-47 plane normals, :180-204 half-space rule,
// :293-322 rotate-about-bbox-centre then translate), on an arbitrary point set.
//
// Everything is float64 with --fmad=false and a fixed operation order, so masks AND moved points
// are bit-identical to the numpy float64 restatement (oracle/reference_rows.py):
//   side(p)  = (px-ox)*nx + (py-oy)*ny + (pz-oz)*nz            (left to right)
//   moved(p) = ((R q)_i + c_i) + t_i,  q = p - c,  (R q)_i = (R_i0 q0 + R_i1 q1) + R_i2 q2
// Plane normals and rotation matrices are evaluated on the host exactly as the reference does
// (numpy float64); only the per-point work runs here.
//
// The point sets are small (5 143 FLAME vertices, or the reference's 362-point spheres), so one
// CTA does the two bounding-box reductions deterministically; the per-point pass is a plain
// stream.
#include "common.cuh"

namespace omfs {

struct DisplacePlan {
    double planes[3][8];
    double moves[2][12];
};

__device__ __forceinline__ double plane_side(const double* pl, double x, double y, double z) {
    return (x - pl[3]) * pl[0] + (y - pl[4]) * pl[1] + (z - pl[5]) * pl[2];
}

__device__ __forceinline__ uint8_t classify(const DisplacePlan& plan, const float* p, bool is_mand) {
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    uint8_t m = 0;
    if (!(plane_side(plan.planes[0], x, y, z) > 0.0)) m |= 1;  // Le Fort: invert=True side
    if (plane_side(plan.planes[1], x, y, z) > 0.0) m |= 2;     // BSSO-L: invert=False side
    if (!(plane_side(plan.planes[2], x, y, z) > 0.0)) m |= 4;  // BSSO-R: invert=True side
    if (!is_mand && (m & 1)) m |= 8;                           // mobile maxilla
    if (is_mand && (m & 2) && (m & 4)) m |= 16;                // distal mandible
    return m;
}

// one CTA of 1024 threads: masks + the two bounding boxes
__global__ void __launch_bounds__(1024) displace_mask_bbox_kernel(int P, const float* __restrict__ pts,
                                                                  DisplacePlan plan,
                                                                  const float* __restrict__ jaw_weight,
                                                                  int mandible_first, uint8_t* __restrict__ mask,
                                                                  float* __restrict__ bbox) {
    __shared__ float s_red[12][32];
    float lo[2][3], hi[2][3];
    for (int s = 0; s < 2; s++)
        for (int c = 0; c < 3; c++) {
            lo[s][c] = INFINITY;
            hi[s][c] = -INFINITY;
        }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const bool is_mand = jaw_weight ? (jaw_weight[i] > 0.5f) : (i >= mandible_first);
        const uint8_t m = classify(plan, pts + (size_t)i * 3, is_mand);
        mask[i] = m;
        const int s = (m & 8) ? 0 : ((m & 16) ? 1 : -1);
        if (s >= 0)
            for (int c = 0; c < 3; c++) {
                const float v = pts[(size_t)i * 3 + c];
                lo[s][c] = fminf(lo[s][c], v);
                hi[s][c] = fmaxf(hi[s][c], v);
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = 0; s < 2; s++)
        for (int c = 0; c < 3; c++) {
            float a = lo[s][c], b = hi[s][c];
            for (int d = 16; d > 0; d >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xffffffffu, a, d));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, d));
            }
            if (lane == 0) {
                s_red[s * 6 + c][warp] = a;
                s_red[s * 6 + 3 + c][warp] = b;
            }
        }
    __syncthreads();
    if (threadIdx.x < 12) {
        const bool is_min = (threadIdx.x % 6) < 3;
        float a = s_red[threadIdx.x][0];
        for (int w = 1; w < 32; w++) a = is_min ? fminf(a, s_red[threadIdx.x][w]) : fmaxf(a, s_red[threadIdx.x][w]);
        bbox[threadIdx.x] = a;
    }
}

__global__ void __launch_bounds__(256) displace_apply_kernel(int P, const float* __restrict__ pts,
                                                             DisplacePlan plan, const uint8_t* __restrict__ mask,
                                                             const float* __restrict__ bbox,
                                                             float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint8_t m = mask[i];
    const float* p = pts + (size_t)i * 3;
    const int s = (m & 8) ? 0 : ((m & 16) ? 1 : -1);
    if (s < 0) {
        out[(size_t)i * 3] = p[0];
        out[(size_t)i * 3 + 1] = p[1];
        out[(size_t)i * 3 + 2] = p[2];
        return;
    }
    const float* bb = bbox + s * 6;
    double c[3], q[3];
    for (int k = 0; k < 3; k++) {
        c[k] = ((double)bb[k] + (double)bb[3 + k]) * 0.5;
        q[k] = (double)p[k] - c[k];
    }
    const double* M = plan.moves[s];
    for (int r = 0; r < 3; r++) {
        const double rq = (M[r * 3] * q[0] + M[r * 3 + 1] * q[1]) + M[r * 3 + 2] * q[2];
        out[(size_t)i * 3 + r] = (float)((rq + c[r]) + M[9 + r]);
    }
}

}  // namespace omfs

using namespace omfs;

extern "C" int omfs_displace_points(int P, const float* d_points, const double* h_planes, const double* h_moves,
                                    const float* d_jaw_weight, int mandible_first, uint8_t* d_mask, float* d_out,
                                    float* d_bbox, void* stream) {
    OMFS_REQUIRE(P >= 0 && d_points && h_planes && h_moves && d_mask && d_out && d_bbox, "bad arguments");
    if (P == 0) return OMFS_OK;
    DisplacePlan plan;
    memcpy(plan.planes, h_planes, sizeof(plan.planes));
    memcpy(plan.moves, h_moves, sizeof(plan.moves));
    cudaStream_t st = (cudaStream_t)stream;
    displace_mask_bbox_kernel<<<1, 1024, 0, st>>>(P, d_points, plan, d_jaw_weight, mandible_first, d_mask, d_bbox);
    displace_apply_kernel<<<ceil_div(P, 256), 256, 0, st>>>(P, d_points, plan, d_mask, d_bbox, d_out);
    count_launch(2);
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}
