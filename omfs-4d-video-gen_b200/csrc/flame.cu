// flame.cu — the geometry front end (tolerance domain): U1/U2 operand prep, the per-subject shape
// fold, the CUDA-core blendshape contraction (cross-check / tiny-T path), U3 skinning.
// The tensor-core contraction lives in flame_gemm_tc.cu.
//
// Blendshape evaluation as ONE GEMM (the reference's in-tree version expands the
// basis to [B,V,3,K] before an einsum, flame_fitter.py:170-175):
//
//     VP[T, npad] = base[npad] + A[T, K] . B[K, npad]
//
//   * the 300 shape coefficients are constant per subject, so they are folded into `base` once
//     (flame_fold_subject_kernel) and the per-frame K is n_expr + 36 pose-corrective features;
//   * the 5 joint positions are linear in the same coefficients, so they ride along as 15 extra
//     output columns (B's joint columns = J_regressor . dirs, zero for the pose-corrective rows,
//     because joints are regressed from the un-corrected shape);
//   * to reach fp32-class accuracy on tf32 tensor cores both operands are split hi/lo and the
//     three significant products are concatenated along K:  A' = [Ah | Ah | Al],  B' = [Bh | Bl | Bh].
#include "common.cuh"

namespace omfs {

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// SMPL-X style Rodrigues: angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1-cos) K^2
__device__ void rodrigues(const float r[3], float R[9]) {
    const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
    const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    const float x = r[0] / angle, y = r[1] / angle, z = r[2] / angle;
    float s, c;
    sincosf(angle, &s, &c);
    const float oc = 1.0f - c;
    const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) a += K[i * 3 + k] * K[k * 3 + j];
            R[i * 3 + j] = ((i == j) ? 1.0f : 0.0f) + s * K[i * 3 + j] + oc * a;
        }
}

// grid = T, block = 64
__global__ void __launch_bounds__(64) flame_pose_prep_kernel(int n_expr, int kpad, const float* __restrict__ expr,
                                                             const float* __restrict__ rotation,
                                                             const float* __restrict__ neck,
                                                             const float* __restrict__ jaw,
                                                             const float* __restrict__ eyes,
                                                             float* __restrict__ acoef, float* __restrict__ rmats) {
    __shared__ float s_R[5][9];
    const int t = blockIdx.x;
    if (threadIdx.x < 5) {
        const int j = threadIdx.x;
        float r[3];
        const float* src = (j == 0) ? rotation + t * 3 : (j == 1) ? neck + t * 3 : (j == 2) ? jaw + t * 3
                                                                                           : eyes + t * 6 + (j - 3) * 3;
        r[0] = src[0];
        r[1] = src[1];
        r[2] = src[2];
        float R[9];
        rodrigues(r, R);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            s_R[j][i] = R[i];
            rmats[((size_t)t * 5 + j) * 9 + i] = R[i];
        }
    }
    __syncthreads();
    float* row = acoef + (size_t)t * 3 * kpad;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        float c = 0.f;
        if (k < n_expr) {
            c = expr[(size_t)t * n_expr + k];
        } else if (k < n_expr + 36) {
            const int f = k - n_expr, j = 1 + f / 9, i = f % 9;
            c = s_R[j][i] - ((i % 4 == 0) ? 1.0f : 0.0f);
        }
        const float hi = tf32_hi(c);
        const float lo = tf32_hi(c - hi);
        row[k] = hi;
        row[kpad + k] = hi;
        row[2 * kpad + k] = lo;
    }
}

// base[i] = template[i] + sum_k shape[k] * shapedirs[k][i] (+ static) (+ plan), i < 3V
__global__ void __launch_bounds__(256) flame_fold_subject_kernel(int V3, int n_shape,
                                                                 const float* __restrict__ v_template,
                                                                 const float* __restrict__ shapedirs,
                                                                 const float* __restrict__ shape,
                                                                 const float* __restrict__ static_off,
                                                                 const float* __restrict__ plan_off,
                                                                 float* __restrict__ base) {
    __shared__ float s_shape[512];
    for (int k = threadIdx.x; k < n_shape; k += blockDim.x) s_shape[k] = shape[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V3) return;
    float a = v_template[i];
    for (int k = 0; k < n_shape; k++) a = fmaf(s_shape[k], __ldg(shapedirs + (size_t)k * V3 + i), a);
    if (static_off) a += static_off[i];
    if (plan_off) a += plan_off[i];
    base[i] = a;
}

// out[t][j*3+c] = sum_v jreg[j][v] * x[t][v*3+c]; grid = (15, T), block = 256.
// Used for the joint base (T = 1, x = base) and for dynamic offsets.
__global__ void __launch_bounds__(256) flame_joint_regress_kernel(int V, size_t x_stride, size_t out_stride,
                                                                  const float* __restrict__ jreg,
                                                                  const float* __restrict__ x,
                                                                  float* __restrict__ out) {
    __shared__ float s_red[8];
    const int j = blockIdx.x / 3, c = blockIdx.x % 3, t = blockIdx.y;
    const float* xt = x + (size_t)t * x_stride;
    float a = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) a = fmaf(__ldg(jreg + (size_t)j * V + v), xt[v * 3 + c], a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += s_red[w];
        out[(size_t)t * out_stride + blockIdx.x] = s;
    }
}

// CUDA-core contraction on the same operands as the tensor-core kernel.
// C[t][n] = base[n] + sum_k A[t][k] * Bt[n][k];  block = 256 threads, tile 32 (t) x 64 (n), K step 32.
__global__ void __launch_bounds__(256) flame_blend_simt_kernel(int T, int K3, int npad, const float* __restrict__ A,
                                                               const float* __restrict__ Bt,
                                                               const float* __restrict__ base,
                                                               float* __restrict__ C) {
    __shared__ float As[32][33];
    __shared__ float Bs[64][33];
    const int n0 = blockIdx.x * 64, t0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16
    float acc[2][4] = {};
    for (int k0 = 0; k0 < K3; k0 += 32) {
        for (int i = threadIdx.x; i < 32 * 32; i += 256) {
            const int r = i >> 5, c = i & 31;
            As[r][c] = (t0 + r < T && k0 + c < K3) ? A[(size_t)(t0 + r) * K3 + k0 + c] : 0.f;
        }
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {
            const int r = i >> 5, c = i & 31;
            Bs[r][c] = (n0 + r < npad && k0 + c < K3) ? __ldg(Bt + (size_t)(n0 + r) * K3 + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            const float a0 = As[ty * 2][k], a1 = As[ty * 2 + 1][k];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float b = Bs[tx + 16 * j][k];
                acc[0][j] = fmaf(a0, b, acc[0][j]);
                acc[1][j] = fmaf(a1, b, acc[1][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int t = t0 + ty * 2 + i;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx + 16 * j;
            if (n < npad) C[(size_t)t * npad + n] = acc[i][j] + base[n];
        }
    }
}

// U3.  grid = (ceil(V/256), T), block 256.
__global__ void __launch_bounds__(256) flame_lbs_kernel(int V, int npad, const float* __restrict__ vp,
                                                        const float* __restrict__ rmats,
                                                        const float* __restrict__ weights,
                                                        const float* __restrict__ transl,
                                                        const float* __restrict__ dyn,
                                                        const float* __restrict__ jdyn,
                                                        float* __restrict__ verts) {
    __shared__ float s_A[5][12];
    const int t = blockIdx.y;
    const float* row = vp + (size_t)t * npad;
    if (threadIdx.x == 0) {
        float J[15], R[5][9];
        for (int i = 0; i < 15; i++) J[i] = row[3 * V + i] + (jdyn ? jdyn[(size_t)t * 15 + i] : 0.f);
        for (int j = 0; j < 5; j++)
            for (int i = 0; i < 9; i++) R[j][i] = rmats[((size_t)t * 5 + j) * 9 + i];
        const int parent[5] = {-1, 0, 1, 1, 1};
        float G[5][12];
        for (int j = 0; j < 5; j++) {
            float rel[3];
            for (int c = 0; c < 3; c++) rel[c] = (j == 0) ? J[c] : J[j * 3 + c] - J[parent[j] * 3 + c];
            if (j == 0) {
                for (int r = 0; r < 3; r++) {
                    for (int c = 0; c < 3; c++) G[0][r * 4 + c] = R[0][r * 3 + c];
                    G[0][r * 4 + 3] = rel[r];
                }
            } else {
                const float* P = G[parent[j]];
                for (int r = 0; r < 3; r++) {
                    for (int c = 0; c < 3; c++) {
                        float a = 0.f;
                        for (int k = 0; k < 3; k++) a += P[r * 4 + k] * R[j][k * 3 + c];
                        G[j][r * 4 + c] = a;
                    }
                    float a = 0.f;
                    for (int k = 0; k < 3; k++) a += P[r * 4 + k] * rel[k];
                    G[j][r * 4 + 3] = a + P[r * 4 + 3];
                }
            }
        }
        for (int j = 0; j < 5; j++)
            for (int r = 0; r < 3; r++) {
                float a = 0.f;
                for (int c = 0; c < 3; c++) {
                    s_A[j][r * 4 + c] = G[j][r * 4 + c];
                    a += G[j][r * 4 + c] * J[j * 3 + c];
                }
                s_A[j][r * 4 + 3] = G[j][r * 4 + 3] - a;
            }
    }
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float x = row[v * 3], y = row[v * 3 + 1], z = row[v * 3 + 2];
    if (dyn) {
        const float* d = dyn + ((size_t)t * V + v) * 3;
        x += d[0];
        y += d[1];
        z += d[2];
    }
    float w[5];
#pragma unroll
    for (int j = 0; j < 5; j++) w[j] = __ldg(weights + v * 5 + j);
    float M[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 5; j++) a += w[j] * s_A[j][i];
        M[i] = a;
    }
    float* o = verts + ((size_t)t * V + v) * 3;
#pragma unroll
    for (int r = 0; r < 3; r++)
        o[r] = (M[r * 4] * x + M[r * 4 + 1] * y + M[r * 4 + 2] * z + M[r * 4 + 3]) + transl[t * 3 + r];
}

int launch_blend_gemm_tc(int T, int kpad, int npad, const float* d_acoef, const float* d_bt, const float* d_base,
                         float* d_vp, int variant, cudaStream_t stream);  // flame_gemm_tc.cu

}  // namespace omfs

using namespace omfs;

extern "C" int omfs_flame_pose_prep(int T, int n_expr, int kpad, const float* d_expr, const float* d_rotation,
                                    const float* d_neck, const float* d_jaw, const float* d_eyes, float* d_acoef,
                                    float* d_rmats, void* stream) {
    OMFS_REQUIRE(T >= 0 && n_expr >= 0 && kpad >= n_expr + 36 && kpad % 8 == 0, "bad sizes");
    OMFS_REQUIRE(d_expr && d_rotation && d_neck && d_jaw && d_eyes && d_acoef && d_rmats, "null pointer");
    if (T == 0) return OMFS_OK;
    flame_pose_prep_kernel<<<T, 64, 0, (cudaStream_t)stream>>>(n_expr, kpad, d_expr, d_rotation, d_neck, d_jaw,
                                                               d_eyes, d_acoef, d_rmats);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

extern "C" int omfs_flame_fold_subject(int V, int n_shape, int npad, const float* d_template,
                                       const float* d_shapedirs, const float* d_shape, const float* d_static,
                                       const float* d_plan, const float* d_jreg, float* d_base, void* stream) {
    OMFS_REQUIRE(V > 0 && n_shape >= 0 && n_shape <= 512 && npad >= 3 * V + 15, "bad sizes");
    OMFS_REQUIRE(d_template && d_shapedirs && d_shape && d_jreg && d_base, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    OMFS_CUDA(cudaMemsetAsync(d_base, 0, sizeof(float) * npad, st));
    flame_fold_subject_kernel<<<ceil_div(3 * V, 256), 256, 0, st>>>(3 * V, n_shape, d_template, d_shapedirs, d_shape,
                                                                    d_static, d_plan, d_base);
    flame_joint_regress_kernel<<<dim3(15, 1), 256, 0, st>>>(V, 0, 0, d_jreg, d_base, d_base + 3 * V);
    count_launch(2);
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

extern "C" int omfs_flame_joint_dyn(int T, int V, const float* d_jreg, const float* d_dyn, float* d_jdyn,
                                    void* stream) {
    OMFS_REQUIRE(T >= 0 && V > 0 && d_jreg && d_dyn && d_jdyn, "bad arguments");
    if (T == 0) return OMFS_OK;
    flame_joint_regress_kernel<<<dim3(15, T), 256, 0, (cudaStream_t)stream>>>(V, (size_t)V * 3, 15, d_jreg, d_dyn,
                                                                             d_jdyn);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

extern "C" int omfs_flame_blend_gemm(int T, int kpad, int npad, const float* d_acoef, const float* d_bt,
                                     const float* d_base, float* d_vp, int impl, void* stream) {
    OMFS_REQUIRE(T >= 0 && kpad > 0 && kpad % 8 == 0 && npad > 0, "bad sizes");
    OMFS_REQUIRE(d_acoef && d_bt && d_base && d_vp, "null pointer");
    if (T == 0) return OMFS_OK;
    // impl 0: tensor cores (the panel re-use kernel from 1024 rows up, where it is the faster one; the
    // concatenated-K kernel below), 2 / 3: force the concatenated-K / panel re-use kernel, 1: CUDA cores
    if (impl == 0 || impl == 2 || impl == 3) {
        const int variant = impl == 0 ? (T >= 1024 ? 0 : 2) : (impl == 3 ? 0 : 2);
        return launch_blend_gemm_tc(T, kpad, npad, d_acoef, d_bt, d_base, d_vp, variant, (cudaStream_t)stream);
    }
    dim3 grid(ceil_div(npad, 64), ceil_div(T, 32));
    flame_blend_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T, 3 * kpad, npad, d_acoef, d_bt, d_base, d_vp);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

extern "C" int omfs_flame_lbs(int T, int V, int npad, const float* d_vp, const float* d_rmats,
                              const float* d_weights, const float* d_transl, const float* d_dyn,
                              const float* d_jdyn, float* d_verts, void* stream) {
    OMFS_REQUIRE(T >= 0 && V > 0 && npad >= 3 * V + 15, "bad sizes");
    OMFS_REQUIRE(T <= 65535, "at most 65535 frames per call");
    OMFS_REQUIRE(d_vp && d_rmats && d_weights && d_transl && d_verts, "null pointer");
    if (T == 0) return OMFS_OK;
    dim3 grid(ceil_div(V, 256), T);
    flame_lbs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(V, npad, d_vp, d_rmats, d_weights, d_transl, d_dyn,
                                                             d_jdyn, d_verts);
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}
