// CUDA-core fp32 GEMM engine (NSF_GEMM_SIMT_FP32): exact fp32 FMA arithmetic on the re-assembled
// operands.  It is the cross-check for the tcgen05 engine (gemm_tc.cu) and the fallback-free way to
// run the mask network bit-faithfully in fp32; not the throughput path.
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-prefetched double buffering.
#include "gemm_common.cuh"

namespace nsf {

constexpr int SBM = 128, SBN = 128, SBK = 16, SPAD = 4;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][SBK][SBM + SPAD];
    __shared__ __align__(16) float Bs[2][SBK][SBN + SPAD];
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const float* Ah = p.A_hi + (size_t)b * p.a_batch_stride;
    const float* Al = p.A_lo ? p.A_lo + (size_t)b * p.a_batch_stride : nullptr;
    const float* Bh = p.B_hi + (size_t)b * p.b_batch_stride;
    const float* Bl = p.B_lo ? p.B_lo + (size_t)b * p.b_batch_stride : nullptr;

    const int lrow = tid >> 2, lkq = tid & 3;           // loader mapping: rows lrow, lrow+64; float4 index along k
    float4 ra[2], rb[2];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + lrow + 64 * j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < p.M) {
                v = __ldg(reinterpret_cast<const float4*>(Ah + (size_t)m * p.lda + k0 + lkq * 4));
                if (Al) {
                    const float4 l = __ldg(reinterpret_cast<const float4*>(Al + (size_t)m * p.lda + k0 + lkq * 4));
                    v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
                }
            }
            ra[j] = v;
            const int n = n0 + lrow + 64 * j;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.N) {
                w = __ldg(reinterpret_cast<const float4*>(Bh + (size_t)n * p.ldb + k0 + lkq * 4));
                if (Bl) {
                    const float4 l = __ldg(reinterpret_cast<const float4*>(Bl + (size_t)n * p.ldb + k0 + lkq * 4));
                    w.x += l.x; w.y += l.y; w.z += l.z; w.w += l.w;
                }
            }
            rb[j] = w;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = lrow + 64 * j;
            As[buf][lkq * 4 + 0][r] = ra[j].x; As[buf][lkq * 4 + 1][r] = ra[j].y;
            As[buf][lkq * 4 + 2][r] = ra[j].z; As[buf][lkq * 4 + 3][r] = ra[j].w;
            Bs[buf][lkq * 4 + 0][r] = rb[j].x; Bs[buf][lkq * 4 + 1][r] = rb[j].y;
            Bs[buf][lkq * 4 + 2][r] = rb[j].z; Bs[buf][lkq * 4 + 3][r] = rb[j].w;
        }
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = p.K / SBK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tile((kb + 1) * SBK);
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < p.n_valid) gemm_epilogue(p, b, m, n, acc[i][j]);
        }
    }
}

int gemm_simt_launch(const GemmParams& p, cudaStream_t stream) {
    if (p.K % SBK != 0 || (p.lda & 3) || (p.ldb & 3)) {
        set_error("gemm_simt: K=%d must be a multiple of %d and lda/ldb multiples of 4", p.K, SBK);
        return NSF_ERR_INVALID_ARG;
    }
    dim3 grid(ceil_div(p.N, SBN), ceil_div(p.M, SBM), p.batch);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
    return check_launch("gemm_simt_kernel");
}

}  // namespace nsf
