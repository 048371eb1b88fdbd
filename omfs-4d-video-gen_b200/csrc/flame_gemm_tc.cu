// flame_gemm_tc.cu — U1+U2 as a frames x coeffs x verts GEMM on the 5th-gen tensor cores.
//
//     VP[T, npad] = base[npad] + A'[T, K3] . B'[npad, K3]^T          (K3 = 3*kpad, tf32x3 split)
//
// sm_100a design:
//   * operands are K-major fp32 rows; TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages
//     128 x 32 (A) and 256 x 32 (B) tiles into a 2-deep shared-memory ring — 128 bytes per row is
//     exactly one swizzle atom, which is what the UMMA shared-memory descriptor expects;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=256, K=8 per
//     instruction, four per stage); the fp32 accumulator lives in 256 TMEM columns;
//   * tcgen05.commit releases smem stages back to the TMA warp and finally signals the epilogue,
//     which reads TMEM with tcgen05.ld (32 lanes x 32 columns per instruction), transposes each 32x32
//     chunk through shared memory, adds the per-subject base row and writes VP as full 128-byte rows;
//   * rows of A beyond T, rows of B beyond npad and the K tail beyond K3 come back as zeros from TMA's
//     out-of-bounds fill, so no operand padding is needed in HBM.
// Warp roles (192 threads): warps 0-3 epilogue (warp id % 4 selects the TMEM lane quarter),
// warp 4 TMA producer, warp 5 TMEM allocation + MMA issue.  One output tile per CTA, TWO CTAs per SM
// (97 KB of shared memory and 256 TMEM columns each): one CTA's epilogue overlaps the other's main loop.
// Measured at the largest configuration (T = 7 680 plan-frames, 60 x 61 CTAs): the kernel is bound by the
// L2 -> shared-memory operand traffic (K is only 408, every tile re-reads its A and B panels: 9.4 TB/s),
// which is why the wide N tile and the co-resident CTAs pay: 0.44 ms (M128 x N128, 4 stages, 1 CTA/SM,
// row-per-lane stores) -> 0.24 ms, 400 TFLOP/s executed tf32, 1.96 TB/s of output (tools/bench_gemm.py).
// From 1024 rows up the panel re-use kernel below takes over (0.235 ms, tensor pipe 43 % busy).
#include <cuda.h>

#include "common.cuh"

namespace omfs {

#ifndef OMFS_GEMM_BN
#define OMFS_GEMM_BN 256
#endif
constexpr int BM = 128, BN = OMFS_GEMM_BN, BK = 32;  // BK fp32 = 128 bytes = one SWIZZLE_128B atom
#ifndef OMFS_GEMM_STAGES
#define OMFS_GEMM_STAGES 2
#endif
#ifndef OMFS_GEMM_CTAS
#define OMFS_GEMM_CTAS 2
#endif
constexpr int kStages = OMFS_GEMM_STAGES;
constexpr int kUmmaK = 8;                   // tf32: 32 bytes per MMA along K
constexpr uint32_t kStageBytesA = BM * BK * 4, kStageBytesB = BN * BK * 4;
constexpr uint32_t kTmemCols = BN;
constexpr int kGemmThreads = 192;
constexpr size_t kGemmSmem = (size_t)kStages * (kStageBytesA + kStageBytesB) + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 26)) __trap();  // a lost arrival is a bug: fail instead of hanging the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);   // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=BN
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// Epilogue of warps 0-3 (TMEM lanes [32*warp, 32*warp+32)), shared by both kernels.
// tcgen05.ld hands lane l the 32 columns of accumulator ROW l; stored as they come, one STG.128 would touch 32
// different rows (16 of every 32-byte sector).  Each 32x32 chunk is therefore transposed through shared memory
// (the operand ring is idle once the accumulator is complete): rows padded to 144 bytes keep both the row-wise
// STS.128 and the read-back conflict-free, and the global stores become 4 full 128-byte rows per instruction.
__device__ __forceinline__ void epilogue_store(uint32_t bar_tmem_full, uint32_t tmem_base, float* stage_base, int warp,
                                               int lane, int m0, int n0, int T, int npad,
                                               const float* __restrict__ base, float* __restrict__ C) {
    mbar_wait(bar_tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr int kRowPad = 36;  // floats per staged row (32 + 4)
    float* stage = stage_base + warp * 32 * kRowPad;
    const int cq = (lane & 7) * 4, rq = lane >> 3;  // read-back: 8 lanes per row, 4 rows per instruction
#pragma unroll 1
    for (int c = 0; c < BN / 32; c++) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
              "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
              "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
              "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int nb = n0 + c * 32;
        if (nb >= npad) break;  // ragged last N tile (warp-uniform)
        float4* srow = reinterpret_cast<float4*>(stage + lane * kRowPad);
#pragma unroll
        for (int j = 0; j < 8; j++)
            srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const float4 b = __ldg(reinterpret_cast<const float4*>(base + nb + cq));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = rq + 4 * i;
            const float4 a = *reinterpret_cast<const float4*>(stage + r * kRowPad + cq);
            const int row = m0 + warp * 32 + r;
            if (row < T)
                *reinterpret_cast<float4*>(C + (size_t)row * npad + nb + cq) =
                    make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kGemmThreads, OMFS_GEMM_CTAS)
flame_blend_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int T,
                      int K3, int npad, const float* __restrict__ base, float* __restrict__ C) {
    extern __shared__ unsigned char gemm_smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t raw = smem_u32(gemm_smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t smem_a = tiles;
    const uint32_t smem_b = tiles + kStages * kStageBytesA;
    const uint32_t bars = smem_b + kStages * kStageBytesB;  // full[kStages], empty[kStages], tmem_full, tmem_ptr
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tmem_full = bars + 16 * kStages;
    const uint32_t tmem_ptr_addr = bar_tmem_full + 8;
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(gemm_smem_raw + (tmem_ptr_addr - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int num_kb = (K3 + BK - 1) / BK;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
        for (int s = 0; s < kStages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_generic;

    if (warp == 4) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(bar_empty + 8 * s, phase ^ 1u);
                mbar_expect_tx(bar_full + 8 * s, kStageBytesA + kStageBytesB);
                tma_load_2d(smem_a + s * kStageBytesA, &map_a, bar_full + 8 * s, kb * BK, m0);
                tma_load_2d(smem_b + s * kStageBytesB, &map_b, bar_full + 8 * s, kb * BK, n0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ===== MMA issuer =====
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(bar_full + 8 * s, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = umma_desc_sw128(smem_a + s * kStageBytesA);
                const uint64_t db = umma_desc_sw128(smem_b + s * kStageBytesB);
#pragma unroll
                for (int k = 0; k < BK / kUmmaK; k++) {
                    // advance 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                    umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc,
                              (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
            }
            umma_commit(bar_tmem_full);  // accumulator complete
        }
    } else {
        // ===== epilogue: warps 0-3 =====
        epilogue_store(bar_tmem_full, tmem_base, reinterpret_cast<float*>(gemm_smem_raw + (tiles - raw)), warp, lane, m0,
                       n0, T, npad, base, C);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Panel re-use form.  The tf32x3 product A'.B'^T = Ah.Bh^T + Ah.Bl^T + Al.Bh^T reads its four panels Ah, Al, Bh, Bl
// ONCE per K block and issues the three products from the same shared-memory tiles, instead of streaming the
// concatenated operands [Ah|Ah|Al] x [Bh|Bl|Bh] (which stages Ah and Bh twice).  The kernel is bound by the
// L2 -> shared-memory operand traffic, so this is what it saves: per 128x256 tile 9 blocks x 48 KB = 432 KB instead
// of 13 x 48 KB = 624 KB.  K blocks are 16 fp32 = 64 bytes wide (SWIZZLE_64B atoms), so that a stage is still
// 48 KB and two CTAs stay resident per SM.
constexpr int BKP = 16;                                  // fp32 per K block = one 64-byte swizzle row
constexpr uint32_t kPanelBytesA = BM * BKP * 4, kPanelBytesB = BN * BKP * 4;
constexpr uint32_t kStageBytesP = 2 * kPanelBytesA + 2 * kPanelBytesB;
constexpr size_t kGemmSmemP = (size_t)kStages * kStageBytesP + 1024 + 256;

// K-major, SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;  // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(kGemmThreads, OMFS_GEMM_CTAS)
flame_blend_tc_panels_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                             const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                             int T, int kpad, int npad, const float* __restrict__ base, float* __restrict__ C) {
    extern __shared__ unsigned char gemm_smem_raw[];
    const uint32_t raw = smem_u32(gemm_smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + kStages * kStageBytesP;  // full[kStages], empty[kStages], tmem_full, tmem_ptr
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tmem_full = bars + 16 * kStages;
    const uint32_t tmem_ptr_addr = bar_tmem_full + 8;
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(gemm_smem_raw + (tmem_ptr_addr - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int num_kb = (kpad + BKP - 1) / BKP;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_al) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_bh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_bl) : "memory");
        for (int s = 0; s < kStages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_generic;

    if (warp == 4) {
        if (lane == 0) {
            // ===== TMA producer: four panels per K block =====
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                const uint32_t st = tiles + s * kStageBytesP;
                mbar_wait(bar_empty + 8 * s, phase ^ 1u);
                mbar_expect_tx(bar_full + 8 * s, kStageBytesP);
                tma_load_2d(st, &map_ah, bar_full + 8 * s, kb * BKP, m0);
                tma_load_2d(st + kPanelBytesA, &map_al, bar_full + 8 * s, kb * BKP, m0);
                tma_load_2d(st + 2 * kPanelBytesA, &map_bh, bar_full + 8 * s, kb * BKP, n0);
                tma_load_2d(st + 2 * kPanelBytesA + kPanelBytesB, &map_bl, bar_full + 8 * s, kb * BKP, n0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ===== MMA issuer: hi.hi + hi.lo + lo.hi from the same tiles =====
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                const uint32_t st = tiles + s * kStageBytesP;
                mbar_wait(bar_full + 8 * s, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dah = umma_desc_sw64(st), dal = umma_desc_sw64(st + kPanelBytesA);
                const uint64_t dbh = umma_desc_sw64(st + 2 * kPanelBytesA);
                const uint64_t dbl = umma_desc_sw64(st + 2 * kPanelBytesA + kPanelBytesB);
#pragma unroll
                for (int k = 0; k < BKP / kUmmaK; k++) {
                    const uint64_t o = (uint64_t)(2 * k);  // +32 bytes along K inside the swizzle row
                    umma_tf32(tmem_base, dah + o, dbh + o, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dah + o, dbl + o, kIdesc, 1u);
                    umma_tf32(tmem_base, dal + o, dbh + o, kIdesc, 1u);
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_tmem_full);
        }
    } else {
        epilogue_store(bar_tmem_full, tmem_base, reinterpret_cast<float*>(gemm_smem_raw + (tiles - raw)), warp, lane, m0,
                       n0, T, npad, base, C);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 rows of `cols` elements, `row_stride` elements apart -> 2D tensor map with a (box_cols x box_rows) box whose
// rows are one swizzle atom wide (128 or 64 bytes)
static int make_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride,
                    uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return OMFS_ERR_UNSUPPORTED;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {row_stride * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = box_cols * sizeof(float) == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols);
        return OMFS_ERR_CUDA;
    }
    return OMFS_OK;
}

int launch_blend_gemm_tc(int T, int kpad, int npad, const float* d_acoef, const float* d_bt, const float* d_base,
                         float* d_vp, int variant, cudaStream_t stream) {
    const int K3 = 3 * kpad;
    if (npad % 32 != 0) {  // the epilogue writes whole 32-column chunks; a ragged last N tile is zero-filled by TMA
        set_error("blend gemm: npad (%d) must be a multiple of 32", npad);
        return OMFS_ERR_INVALID;
    }
    if (((uintptr_t)d_acoef | (uintptr_t)d_bt | (uintptr_t)d_vp | (uintptr_t)d_base) & 15) {
        set_error("blend gemm: operands must be 16-byte aligned");
        return OMFS_ERR_INVALID;
    }
    dim3 grid(ceil_div(npad, BN), ceil_div(T, BM));
    int rc;
    if (variant == 0 && kpad % 4 == 0) {
        // panel re-use: operands are [Ah | Ah | Al] and [Bh | Bl | Bh] rows of 3*kpad floats (capi.cu bakes them)
        CUtensorMap ah, al, bh, bl;
        if ((rc = make_map(&ah, d_acoef, (uint64_t)T, (uint64_t)kpad, (uint64_t)K3, BKP, BM))) return rc;
        if ((rc = make_map(&al, d_acoef + 2 * kpad, (uint64_t)T, (uint64_t)kpad, (uint64_t)K3, BKP, BM))) return rc;
        if ((rc = make_map(&bh, d_bt, (uint64_t)npad, (uint64_t)kpad, (uint64_t)K3, BKP, BN))) return rc;
        if ((rc = make_map(&bl, d_bt + kpad, (uint64_t)npad, (uint64_t)kpad, (uint64_t)K3, BKP, BN))) return rc;
        static DeviceOnce once_p;
        if ((rc = ensure_dyn_smem(once_p, flame_blend_tc_panels_kernel, (int)kGemmSmemP))) return rc;
        flame_blend_tc_panels_kernel<<<grid, kGemmThreads, kGemmSmemP, stream>>>(ah, al, bh, bl, T, kpad, npad, d_base,
                                                                                 d_vp);
    } else {
        CUtensorMap map_a, map_b;
        if ((rc = make_map(&map_a, d_acoef, (uint64_t)T, (uint64_t)K3, (uint64_t)K3, BK, BM))) return rc;
        if ((rc = make_map(&map_b, d_bt, (uint64_t)npad, (uint64_t)K3, (uint64_t)K3, BK, BN))) return rc;
        static DeviceOnce once_s;
        if ((rc = ensure_dyn_smem(once_s, flame_blend_tc_kernel, (int)kGemmSmem))) return rc;
        flame_blend_tc_kernel<<<grid, kGemmThreads, kGemmSmem, stream>>>(map_a, map_b, T, K3, npad, d_base, d_vp);
    }
    count_launch();
    OMFS_LAUNCH_CHECK();
    return OMFS_OK;
}

}  // namespace omfs
