// Micro-benchmark: issue rate of scalar FP32 vs packed f32x2 arithmetic on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ub_f32x2 tools/ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8];
    unsigned long long p[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    unsigned long long s2 = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = fmaf(a[i], s, 0.25f * s);          // scalar FFMA (reg,reg,reg)
            if (MODE == 1) p[i] = fma2(p[i], s2, s2);                 // packed FFMA2
            if (MODE == 2) a[i] = __fmul_rn(a[i], s);                 // scalar FMUL
            if (MODE == 3) p[i] = mul2(p[i], s2);
            if (MODE == 4) a[i] = __fadd_rn(a[i], s);
            if (MODE == 5) p[i] = add2(p[i], s2);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; i++) r += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * threads / 32 * iters * 8;
    printf("%-12s %.3f ms  %.2f warp-inst/clk/SM (at 1.965 GHz)  %.1f Gelem-op/s\n", name, ms, winst / (ms * 1e-3) / 1.965e9 / 148,
           winst * 32 * ((MODE & 1) ? 2 : 1) / (ms * 1e-3) / 1e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("ffma", d); run<1>("ffma2", d); run<2>("fmul", d); run<3>("fmul2", d); run<4>("fadd", d); run<5>("fadd2", d);
    return 0;
}
