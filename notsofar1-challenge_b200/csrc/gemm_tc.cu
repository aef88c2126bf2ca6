// tcgen05 GEMM engine for sm_100a: TMA (SWIZZLE_128B) -> shared memory -> tcgen05.mma kind::tf32 with the
// fp32 accumulator in TMEM -> tcgen05.ld epilogue with the fused ops of gemm_common.cuh.
//
//   D[b][m][n] = sum_k A[b][m][k] B[b][n][k]          both operands K-major fp32 (tf32 inputs)
//
// 3xTF32 (n_terms == 3): D = A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in the same TMEM tile; the
// operands arrive pre-split (X_hi has its 13 low mantissa bits clear, X_lo = X - X_hi exactly), so the
// tensor core's truncation of X_hi is a no-op and the result is fp32-grade (~2^-21 relative).
//
// CTA = 6 warps: warp 0 lane 0 TMA producer | warp 1 TMEM allocator + lane 0 MMA issuer | warps 2..5 epilogue
// (each owns the 32 TMEM lanes of its warp-id % 4 quarter).  One 128x128 output tile per CTA,
// BLOCK_K = 32 floats = one 128-byte swizzle row, 3- or 4-stage mbarrier ring.
#include "gemm_common.cuh"
#include <cuda.h>
#include <mutex>

namespace nsf {

constexpr int TBM = 128, TBN = 128, TBK = 32;
constexpr int kTileBytes = TBM * TBK * 4;          // 16 KB, same for A and B tiles (TBM == TBN)
constexpr int kTcThreads = 192;
constexpr uint32_t kTmemCols = 128;

template <int NTERMS> struct TcCfg {
    static constexpr int kTilesPerStage = NTERMS == 3 ? 4 : 2;
    static constexpr int kStages = NTERMS == 3 ? 3 : 4;
    static constexpr int kStageBytes = kTilesPerStage * kTileBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a launch failure), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("nsf gemm_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), LBO unused (=1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = TBN
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
}

template <int NTERMS>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const GemmParams p) {
    using Cfg = TcCfg<NTERMS>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* gen_tiles = smem_raw + (tiles - raw);
    const uint32_t bars = tiles + Cfg::kStages * Cfg::kStageBytes;       // full[S], empty[S], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * Cfg::kStages);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen_tiles + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 1));
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * Cfg::kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TBN, m0 = blockIdx.y * TBM, b = blockIdx.z;
    const int nkb = p.K / TBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
                const uint32_t st = tiles + s * Cfg::kStageBytes;
                const int k0 = kb * TBK;
                tma_load_3d(st, &map_a_hi, k0, m0, b, full_bar(s));
                if (NTERMS == 3) {
                    tma_load_3d(st + kTileBytes, &map_a_lo, k0, m0, b, full_bar(s));
                    tma_load_3d(st + 2 * kTileBytes, &map_b_hi, k0, n0, b, full_bar(s));
                    tma_load_3d(st + 3 * kTileBytes, &map_b_lo, k0, n0, b, full_bar(s));
                } else {
                    tma_load_3d(st + kTileBytes, &map_b_hi, k0, n0, b, full_bar(s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = make_idesc();
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(full_bar(s), ph);
                tcgen05_fence_after();
                const uint32_t st = tiles + s * Cfg::kStageBytes;
                const uint32_t a_hi = st, a_lo = st + kTileBytes;
                const uint32_t b_hi = st + (NTERMS == 3 ? 2 : 1) * kTileBytes, b_lo = st + 3 * kTileBytes;
#pragma unroll
                for (int ks = 0; ks < TBK / 8; ++ks) {                    // UMMA_K = 8 tf32 = 32 bytes
                    const uint32_t koff = ks * 32;
                    const uint64_t da_hi = make_smem_desc(a_hi + koff), db_hi = make_smem_desc(b_hi + koff);
                    if (NTERMS == 3) {
                        const uint64_t da_lo = make_smem_desc(a_lo + koff), db_lo = make_smem_desc(b_lo + koff);
                        tcgen05_mma_tf32(tmem_base, da_lo, db_hi, idesc, (kb | ks) != 0);    // small terms first
                        tcgen05_mma_tf32(tmem_base, da_hi, db_lo, idesc, 1);
                        tcgen05_mma_tf32(tmem_base, da_hi, db_hi, idesc, 1);
                    } else {
                        tcgen05_mma_tf32(tmem_base, da_hi, db_hi, idesc, (kb | ks) != 0);
                    }
                }
                tcgen05_commit(empty_bar(s));          // frees the stage once the MMAs above have read it
            }
            tcgen05_commit(tmem_full_bar);             // accumulator complete
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> fused epilogue -> global
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        const int m = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < TBN; c0 += 32) {
            if (n0 + c0 >= p.N) break;                 // warp-uniform
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c0 + j;
                    if (n < p.N) gemm_epilogue(p, b, m, n, __uint_as_float(r[j]));
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// [batch][rows][K] fp32, K contiguous; box = 32 x 128 x 1, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int64_t batch, int64_t batch_stride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("gemm_tc: cuTensorMapEncodeTiled is not available from the driver"); return NSF_ERR_CUDA; }
    if (batch_stride == 0) batch_stride = rows * ld;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)batch_stride * 4};
    cuuint32_t box[3] = {(cuuint32_t)TBK, (cuuint32_t)TBM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15)) {
        set_error("gemm_tc: operand base/strides must be 16-byte aligned (ld=%lld, batch_stride=%lld)", (long long)ld, (long long)batch_stride);
        return NSF_ERR_INVALID_ARG;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return NSF_ERR_CUDA; }
    return NSF_OK;
}

template <int NTERMS>
static int launch_t(const GemmParams& p, cudaStream_t stream) {
    using Cfg = TcCfg<NTERMS>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(&ma_hi, p.A_hi, p.M, p.K, p.lda, p.batch, p.a_batch_stride))) return rc;
    if ((rc = make_map(&mb_hi, p.B_hi, p.N, p.K, p.ldb, p.batch, p.b_batch_stride))) return rc;
    if (NTERMS == 3) {
        if (!p.A_lo || !p.B_lo) { set_error("gemm_tc: 3xTF32 needs split operands"); return NSF_ERR_INVALID_ARG; }
        if ((rc = make_map(&ma_lo, p.A_lo, p.M, p.K, p.lda, p.batch, p.a_batch_stride))) return rc;
        if ((rc = make_map(&mb_lo, p.B_lo, p.N, p.K, p.ldb, p.batch, p.b_batch_stride))) return rc;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    NSF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<NTERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    dim3 grid(ceil_div(p.N, TBN), ceil_div(p.M, TBM), p.batch);
    gemm_tc_kernel<NTERMS><<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    return check_launch("gemm_tc_kernel");
}

int gemm_tc_launch(const GemmParams& p, int n_terms, cudaStream_t stream) {
    if (p.K % TBK != 0 || p.K <= 0) { set_error("gemm_tc: K=%d must be a positive multiple of %d", p.K, TBK); return NSF_ERR_INVALID_ARG; }
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) { set_error("gemm_tc: empty problem"); return NSF_ERR_INVALID_ARG; }
    return n_terms == 3 ? launch_t<3>(p, stream) : launch_t<1>(p, stream);
}

}  // namespace nsf
