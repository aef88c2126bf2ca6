// Non-causal multi-head attention with online softmax on tcgen05 (bf16 operands, fp32 accumulate) -- the Whisper
// encoder's attention (1 500 positions, d_k = 64).
//
// Algorithm [upstream openai-whisper, whisper/model.py MultiHeadAttention.qkv_attention]: w = softmax(q k^T) with q and k
// pre-scaled by d_k^-0.25 each, out = w v.  The scale is folded into the projection weights by the host.
//
// Work item = (batch x head, block of 128 query rows).  Per 128-key tile j:
//   MMA warp      S_j = Q K_j^T into one of two TMEM score buffers (so S_{j+1} is computed while the softmax of tile j
//                 runs), then O_j = P_j V_j (P as the A operand from tensor memory) into a 64-column TMEM tile;
//   softmax warps two threads per query row (64 keys and 32 output columns each; the row maximum is exchanged through
//                 shared memory once per tile, the row sums only at the end): running max / sum in the log2 domain, P_j
//                 written back to TMEM as packed bf16x2, the output accumulator kept in registers:
//                 acc = (acc + O_{j-1}) * 2^(m_old - m_new);
//   TMA warp      Q once per item, K_j / V_j^T through two 3-slot rings.
// q, k, v: [n_bh][T][64] bf16 (rows beyond T are zero-filled by TMA); out: [batch * T][ldo] at column head * 64.
#include "gemm_common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

namespace nsf {

constexpr int kFaThreads = 320;                 // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two per TMEM lane quarter)
constexpr int kFaDk = 64;
constexpr int kFaQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kFaKBytes = 128 * 128;            // 128 keys x 64 bf16
constexpr int kFaVBytes = 128 * 128;            // V tile: 128 keys x 64 d, read by the P V product as an MN-major operand
constexpr int kFaRing = 3;
constexpr int kFaTileBytes = kFaQBytes + kFaRing * (kFaKBytes + kFaVBytes);
constexpr int kFaColS = 0;                      // two 128-column score buffers
constexpr int kFaColP = 256;                    // 64 packed columns = 128 keys
constexpr int kFaColO = 320;                    // 64 columns
constexpr int kFaXchBytes = 2 * 128 * 4;        // row maxima (and, at the end, row sums) of the two key halves
constexpr int kFaSmemBytes = kFaTileBytes + kFaXchBytes + 256 + 1024;

struct FaParams {
    int n_bh, n_heads, T;
    float* out_hi; float* out_lo; int64_t ldo; int out_fmt;
};

__device__ __forceinline__ float fa_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t fa_pack_bf16(float lo_elem, float hi_elem) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kFaThreads, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_v, const FaParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t q_smem = base;
    auto k_smem = [&](int s) { return base + kFaQBytes + s * kFaKBytes; };
    auto v_smem = [&](int s) { return base + kFaQBytes + kFaRing * kFaKBytes + s * kFaVBytes; };
    float* xch = reinterpret_cast<float*>(gen + kFaTileBytes);     // [2][128]
    const uint32_t bars = base + kFaTileBytes + kFaXchBytes;
    const uint32_t q_full = bars, q_empty = bars + 8;
    auto k_full = [&](int s) { return bars + 16u + 8u * s; };
    auto k_empty = [&](int s) { return bars + 16u + 8u * (kFaRing + s); };
    auto v_full = [&](int s) { return bars + 64u + 8u * s; };
    auto v_empty = [&](int s) { return bars + 64u + 8u * (kFaRing + s); };
    auto s_full = [&](int b) { return bars + 112u + 8u * b; };
    auto s_empty = [&](int b) { return bars + 128u + 8u * b; };
    const uint32_t p_full = bars + 144, o_full = bars + 152, o_empty = bars + 160;
    const uint32_t tmem_ptr_addr = bars + 168;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int n_qb = (T + 127) / 128, n_kt = (T + 127) / 128;
    const int total = p.n_bh * n_qb;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int s = 0; s < kFaRing; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(s_empty(b), 8); }
        mbar_init(p_full, 8); mbar_init(o_full, 1); mbar_init(o_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            uint32_t it = 0, kt = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
                const int bh = item / n_qb, qb = item - bh * n_qb;
                mbar_wait(q_empty, (it & 1) ^ 1);
                mbar_expect_tx(q_full, kFaQBytes);
                tma_load_3d(q_smem, &map_q, 0, qb * 128, bh, q_full);
                for (int j = 0; j < n_kt; ++j, ++kt) {
                    const int s = kt % kFaRing;
                    const uint32_t ph = (kt / kFaRing) & 1;
                    mbar_wait(k_empty(s), ph ^ 1);
                    mbar_expect_tx(k_full(s), kFaKBytes);
                    tma_load_3d(k_smem(s), &map_k, 0, j * 128, bh, k_full(s));
                    mbar_wait(v_empty(s), ph ^ 1);
                    mbar_expect_tx(v_full(s), kFaVBytes);
                    tma_load_3d(v_smem(s), &map_v, 0, j * 128, bh, v_full(s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc_s = make_idesc_f16(128, 1), idesc_o = make_idesc_f16_bmn(64, 1);
            uint32_t it = 0, kt = 0, st_cnt = 0, pv_cnt = 0;      // kt: K tiles issued; st_cnt: S buffers used; pv_cnt: P V products issued
            auto issue_s = [&](uint32_t ktile) {
                const int s = ktile % kFaRing;
                mbar_wait(k_full(s), (ktile / kFaRing) & 1);
                const uint32_t b = st_cnt & 1;
                mbar_wait(s_empty(b), ((st_cnt >> 1) & 1) ^ 1);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    tcgen05_mma_f16(tmem_base + kFaColS + 128 * b, make_smem_desc(q_smem + ks * 32), make_smem_desc(k_smem(s) + ks * 32),
                                    idesc_s, ks != 0);
                tcgen05_commit(k_empty(s));
                tcgen05_commit(s_full(b));
                ++st_cnt;
            };
            auto issue_pv = [&](uint32_t ktile) {
                const int s = ktile % kFaRing;
                mbar_wait(v_full(s), (ktile / kFaRing) & 1);
                mbar_wait(p_full, pv_cnt & 1);
                mbar_wait(o_empty, (pv_cnt & 1) ^ 1);
                tcgen05_fence_after();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    tcgen05_mma_f16_ts(tmem_base + kFaColO, tmem_base + kFaColP + 8 * i,
                                       make_smem_desc(v_smem(s) + i * 2048), idesc_o, i != 0);      // 16 key rows of 128 bytes
                tcgen05_commit(v_empty(s));
                tcgen05_commit(o_full);
                ++pv_cnt;
            };
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
                mbar_wait(q_full, it & 1);
                const uint32_t kt0 = kt;
                issue_s(kt0);
                for (int j = 0; j < n_kt; ++j) {
                    if (j + 1 < n_kt) issue_s(kt0 + j + 1);
                    else tcgen05_commit(q_empty);                   // all score MMAs of the item are issued: Q may be replaced
                    issue_pv(kt0 + j);
                }
                kt += n_kt;
            }
        }
    } else {
        // ===== softmax / accumulate warps: two threads per query row (key half hh, output columns 32 hh ..)
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int r = 32 * q + lane;
        const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
        const float c = 1.4426950408889634f;                       // scores are natural-log logits
        uint32_t st_cnt = 0, o_cnt = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int bh = item / n_qb, qb = item - bh * n_qb;
            float m = -INFINITY, l = 0.f;                           // l: this thread's half of the row sum
            float acc[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] = 0.f;
            for (int j = 0; j < n_kt; ++j, ++st_cnt) {
                const uint32_t b = st_cnt & 1;
                mbar_wait(s_full(b), (st_cnt >> 1) & 1);
                tcgen05_fence_after();
                uint32_t sv[64];
                tmem_ld_32x32_nowait(tmem_base + lane_sel + kFaColS + 128 * b + 64 * hh, sv);
                tmem_ld_32x32_nowait(tmem_base + lane_sel + kFaColS + 128 * b + 64 * hh + 32, sv + 32);
                tmem_ld_wait();
                tmem_ld_fence(sv); tmem_ld_fence(sv + 32);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty(b));             // the score buffer can be refilled
                const int n_valid = min(64, T - j * 128 - 64 * hh);
                float mt = -INFINITY;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const float v = i < n_valid ? __uint_as_float(sv[i]) * c : -INFINITY;
                    sv[i] = __float_as_uint(v);
                    mt = fmaxf(mt, v);
                }
                xch[hh * 128 + r] = mt;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                const float m_new = fmaxf(m, fmaxf(mt, xch[(hh ^ 1) * 128 + r]));
                const float alpha = fa_ex2(m - m_new);              // 0 on the first tile (m = -inf)
                float sum = 0.f;
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float p0 = fa_ex2(__uint_as_float(sv[2 * i]) - m_new), p1 = fa_ex2(__uint_as_float(sv[2 * i + 1]) - m_new);
                    pk[i] = fa_pack_bf16(p0, p1);
                    // the row sum uses the rounded probabilities the tensor core will see
                    sum += __uint_as_float(pk[i] << 16) + __uint_as_float(pk[i] & 0xffff0000u);
                }
                l = l * alpha + sum;
                m = m_new;
                if (j > 0) {
                    // O_{j-1} = P_{j-1} V_{j-1} is relative to the previous maximum, like acc
                    mbar_wait(o_full, o_cnt & 1);
                    tcgen05_fence_after();
                    uint32_t ov[32];
                    tmem_ld_32x32(tmem_base + lane_sel + kFaColO + 32 * hh, ov);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_empty);
                    ++o_cnt;
#pragma unroll
                    for (int d = 0; d < 32; ++d) acc[d] = (acc[d] + __uint_as_float(ov[d])) * alpha;
                }
                // P_j -> TMEM (the previous P V product has completed: o_full above, or nothing was issued yet).  The
                // exchange slot is reused next tile: the barrier above orders this tile's reads before those writes
                // because every thread passes it again only after its partner has read.
                tmem_st_32x32(tmem_base + lane_sel + kFaColP + 32 * hh, pk);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");       // partner has read xch: safe to overwrite
            }
            // last tile's product and the other half of the row sum
            mbar_wait(o_full, o_cnt & 1);
            tcgen05_fence_after();
            {
                uint32_t ov[32];
                tmem_ld_32x32(tmem_base + lane_sel + kFaColO + 32 * hh, ov);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
                ++o_cnt;
                xch[hh * 128 + r] = l;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                const float inv = 1.f / (l + xch[(hh ^ 1) * 128 + r]);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
#pragma unroll
                for (int d = 0; d < 32; ++d) acc[d] = (acc[d] + __uint_as_float(ov[d])) * inv;
            }
            const int t1 = qb * 128 + r;
            if (t1 < T) {
                const int bidx = bh / p.n_heads, h = bh - bidx * p.n_heads;
                const size_t o = ((size_t)bidx * T + t1) * p.ldo + (size_t)h * kFaDk + 32 * hh;
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                    float v8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v8[e] = acc[8 * k8 + e];
                    split_store8(p.out_fmt, p.out_hi, p.out_lo, o + 8 * k8, v8);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// q, k, v: [n_bh][T][64] bf16; out [batch * T][ldo] in out_fmt at column head * 64
int flash_attn_launch(const void* q, const void* k, const void* vt, int n_batch, int n_heads, int T,
                      float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream) {
    if (T < 1) { set_error("flash_attn: T=%d", T); return NSF_ERR_INVALID_ARG; }
    if ((ldo & 7) || ((uintptr_t)out_hi & 15)) { set_error("flash_attn: ldo=%lld / output alignment", (long long)ldo); return NSF_ERR_INVALID_ARG; }
    const int n_bh = n_batch * n_heads;
    CUtensorMap mq, mk, mv;
    int rc;
    if ((rc = make_tmap_kmajor16(&mq, q, T, kFaDk, kFaDk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor16(&mk, k, T, kFaDk, kFaDk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor16(&mv, vt, T, kFaDk, kFaDk, n_bh, 0, 128))) return rc;
    FaParams p;
    p.n_bh = n_bh; p.n_heads = n_heads; p.T = T;
    p.out_hi = out_hi; p.out_lo = out_lo; p.ldo = ldo; p.out_fmt = out_fmt;
    NSF_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmemBytes));
    const int64_t total = (int64_t)n_bh * ((T + 127) / 128);
    const int grid = (int)(total < sm_count() ? total : sm_count());
    flash_attn_kernel<<<grid, kFaThreads, kFaSmemBytes, stream>>>(mq, mk, mv, p);
    return check_launch("flash_attn_kernel");
}

// ------------------------------------------------------------------------------------------- test hook
__global__ void __launch_bounds__(256)
fa_test_cvt_kernel(const float* __restrict__ in, int64_t n, uint16_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat16_as_ushort(__float2bfloat16_rn(in[i]));
}
}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_flash_attention_test_workspace_bytes(int n_batch, int n_heads, int T) {
    const int64_t n_bh = (int64_t)n_batch * n_heads;
    return 3 * n_bh * T * kFaDk * 2 + 1024;
}

extern "C" int nsf_flash_attention_test(const float* q, const float* k, const float* v, int n_batch, int n_heads, int T, float* out,
                                        void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(q && k && v && out && workspace, "nsf_flash_attention_test: null pointer");
    NSF_REQUIRE(n_batch > 0 && n_heads > 0 && T > 0, "nsf_flash_attention_test: bad sizes");
    NSF_REQUIRE(workspace_bytes >= nsf_flash_attention_test_workspace_bytes(n_batch, n_heads, T) && ((uintptr_t)workspace & 255) == 0,
                "nsf_flash_attention_test: workspace too small or not 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t n_bh = (int64_t)n_batch * n_heads;
    const int64_t nq = n_bh * T * kFaDk;
    uint16_t* qb = reinterpret_cast<uint16_t*>(workspace);
    uint16_t* kb = qb + nq;
    uint16_t* vb = kb + nq;
    fa_test_cvt_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(q, nq, qb);
    fa_test_cvt_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(k, nq, kb);
    fa_test_cvt_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(v, nq, vb);
    int rc = check_launch("fa_test_cvt_kernel");
    if (rc) return rc;
    ProfScope prof(PROF_ATTN, 4.0 * T * T * kFaDk * (double)n_bh, s);
    return flash_attn_launch(qb, kb, vb, n_batch, n_heads, T, out, nullptr, (int64_t)n_heads * kFaDk, SPLIT_FP32, s);
}
