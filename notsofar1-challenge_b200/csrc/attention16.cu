// Fused relative-position attention on 16-bit head + remainder pairs (the 2xBF16 engine's attention).
//
// Reference: css/css_with_conformer/nnet/conformer.py  MultiHeadedAttention.forward :66-92 (see attention.cu for the
// formula).  Same on-chip pipeline as attention.cu -- S = Q K^T and Bm = Q PEw^T in TMEM, register barrel-shift skew,
// softmax, P back to TMEM as the A operand of P V -- with three changes that cut the bytes per work item from 416 KB
// to 128 KB (the fp32-pair kernel starves its MMAs on L2 -> shared-memory traffic, ncu: 52 % of the softmax warps'
// samples wait for the score MMAs):
//   * q, k, v^T and pe_k arrive as bf16 pairs (hi = bf16(x), lo = bf16(x - hi)); every product is three kind::f16
//     MMAs (lo.hi + hi.lo + hi.hi, fp32 accumulate), half the bytes and twice the rate of the 3xTF32 form;
//   * a CTA works on one 128-row block index for its whole life, so the pe_k window that block can reach
//     (<= 320 rows) is loaded once and stays resident in shared memory;
//   * Q, K and V^T of an item are single bulk tiles (no ring): Q / K of the next item land while the softmax of the
//     current one runs, V^T of the next item while its scores are computed.
// The probabilities are written to TMEM as packed bf16x2 (two keys per 32-bit column, hi and lo planes).
// T <= 192 frames per segment, d_k = 64.
#include "gemm_common.cuh"
#include "tc_ptx.cuh"
#include <math.h>
#include <cstdlib>

namespace nsf {

constexpr int kA2Split = 3;                      // softmax threads per query row (one TMEM lane quarter each: 4 * kA2Split warps)
constexpr int kA2Threads = 64 + 128 * kA2Split; // warp 0 TMA, warp 1 MMA, warps 2..13 softmax / epilogue
constexpr int kA2Dk = 64;
constexpr int kA2QHalf = 128 * 128;             // one plane of Q: 128 rows x 128 B
constexpr int kA2KHalf = 192 * 128;             // one plane of K: 192 rows
constexpr int kA2VHalf = 192 * 128;             // one plane of V: 192 key rows x 64 d (MN-major B operand of P V: no transposed copy)
constexpr int kA2PeHalf = 320 * 128;            // one plane of the pe_k window
constexpr int kA2QOff = 0;
constexpr int kA2KOff = kA2QOff + 2 * kA2QHalf;
constexpr int kA2VOff = kA2KOff + 2 * kA2KHalf;
constexpr int kA2PeOff = kA2VOff + 2 * kA2VHalf;
constexpr int kA2TileBytes = kA2PeOff + 2 * kA2PeHalf;          // 208 KB
constexpr int kA2ColB = 192;                    // first TMEM column of Bm
constexpr int kA2ColPlo = 96;                   // first TMEM column of the remainder plane of P (head plane at 0)
constexpr int kA2ColO = 384;                    // first TMEM column of O
constexpr int kA2Slots = 192 / kA2Split;         // key positions per softmax thread: 64
constexpr int kA2XchBytes = 2 * kA2Split * 128 * 4;             // (max, sum) exchange between the threads of a row
constexpr int kA2SmemBytes = kA2TileBytes + kA2XchBytes + 256 /*barriers*/ + 1024 /*alignment*/;
static_assert(kA2Slots % 32 == 0 && kA2Slots + 32 <= 96, "softmax thread geometry");

struct Attn16Params {
    int n_bh, n_heads, T, Tp;
    int pe_row0;          // first pe_k row of row block 0's window: maxlen - (T - 1)
    int n_rb, g0;         // row blocks per (segment, head); CTAs [0, g0) take block 0, the rest block 1
    float scale_log2e;    // log2(e) / sqrt(d_k)
    float* out_hi; float* out_lo; int64_t ldo; int out_fmt;
};

__device__ __forceinline__ void a2_named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
template <int SH>
__device__ __forceinline__ void a2_barrel_stage(uint32_t (&w)[kA2Slots + 32], int lane) {
    const bool on = (lane & SH) != 0;
#pragma unroll
    for (int i = 0; i < kA2Slots + SH - 1; ++i) w[i] = on ? w[i + SH] : w[i];
}
__device__ __forceinline__ float ex2_approx(float x) {          // MUFU.EX2, 2 ulp; ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);       // .x (low half) = lo_elem
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kA2Threads, 1)
attn16_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
              const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
              const __grid_constant__ CUtensorMap map_pe_hi, const __grid_constant__ CUtensorMap map_pe_lo,
              const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
              const Attn16Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t q_smem = base + kA2QOff, k_smem = base + kA2KOff, v_smem = base + kA2VOff, pe_smem = base + kA2PeOff;
    float* xch = reinterpret_cast<float*>(gen + kA2TileBytes);                 // [2 (max, sum)][kA2Split][128]
    const uint32_t bars = base + kA2TileBytes + kA2XchBytes;
    const uint32_t qk_full = bars, qk_empty = bars + 8, v_full = bars + 16, v_empty = bars + 24, pe_full = bars + 32;
    const uint32_t s_ready = bars + 40, p_ready = bars + 48, o_ready = bars + 56, o_drained = bars + 64;
    const uint32_t tmem_ptr_addr = bars + 72;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    // this CTA's row block and its share of the (segment, head) items
    int rb, first, stride;
    if (p.n_rb == 1 || (int)blockIdx.x < p.g0) { rb = 0; first = blockIdx.x; stride = p.n_rb == 1 ? gridDim.x : p.g0; }
    else { rb = 1; first = blockIdx.x - p.g0; stride = gridDim.x - p.g0; }
    const int R0 = rb * 128;
    const int rows_valid = min(128, T - R0);
    const int nBh = (((rows_valid + T - 1) + 31) / 32) * 16;        // half of the pe_k window columns (multiple of 16)
    const int nPe = (2 * nBh + 63) / 64;                            // 64-row boxes of the window
    const int nks_pv = (T + 15) / 16;                               // 16-key k-steps of P V
    // output accumulator O (64 columns): behind the pe_k window where the window leaves room (row block 1: 256 columns), else on
    // top of the window's second half (row block 0: the window fills the rest of tensor memory) -- only then does the next item's
    // second half of Bm have to wait until O has been read out
    const bool o_aliases = kA2ColB + 2 * nBh + 64 > 512;
    const uint32_t col_o = o_aliases ? (uint32_t)(kA2ColB + 2 * nBh - 64) : (uint32_t)(kA2ColB + 2 * nBh);
    // Bm is issued in two pieces: nBmA window columns right behind the previous item's P V product, and -- only where O sits on the
    // window's last 64 columns -- those 64 once O has been read out
    const int nBmA = o_aliases ? 2 * nBh - 64 : 2 * nBh, nBmB = o_aliases ? 64 : 0;

    if (threadIdx.x == 0) {
        mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1); mbar_init(pe_full, 1);
        mbar_init(s_ready, 1); mbar_init(p_ready, 4 * kA2Split); mbar_init(o_ready, 1); mbar_init(o_drained, 4 * kA2Split);
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
        if (lane == 0 && first < p.n_bh) {
            // ===== TMA producer
            mbar_expect_tx(pe_full, (uint32_t)nPe * 2u * 8192u);
            for (int c = 0; c < nPe; ++c) {
                tma_load_3d(pe_smem + c * 8192, &map_pe_hi, 0, p.pe_row0 + R0 + 64 * c, 0, pe_full);
                tma_load_3d(pe_smem + kA2PeHalf + c * 8192, &map_pe_lo, 0, p.pe_row0 + R0 + 64 * c, 0, pe_full);
            }
            uint32_t it = 0;
            for (int bh = first; bh < p.n_bh; bh += stride, ++it) {
                mbar_wait(qk_empty, (it & 1) ^ 1);
                mbar_expect_tx(qk_full, 2u * kA2QHalf + 2u * kA2KHalf);
                tma_load_3d(q_smem, &map_q_hi, 0, R0, bh, qk_full);
                tma_load_3d(q_smem + kA2QHalf, &map_q_lo, 0, R0, bh, qk_full);
                tma_load_3d(k_smem, &map_k_hi, 0, 0, bh, qk_full);
                tma_load_3d(k_smem + kA2KHalf, &map_k_lo, 0, 0, bh, qk_full);
                mbar_wait(v_empty, (it & 1) ^ 1);
                mbar_expect_tx(v_full, 2u * kA2VHalf);
                tma_load_3d(v_smem, &map_v_hi, 0, 0, bh, v_full);
                tma_load_3d(v_smem + kA2VHalf, &map_v_lo, 0, 0, bh, v_full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && first < p.n_bh) {
            // ===== MMA issuer
            const uint32_t idesc_s = make_idesc_f16(192, 1), idesc_ba = make_idesc_f16(nBmA, 1), idesc_bb = make_idesc_f16(64, 1),
                           idesc_o = make_idesc_f16_bmn(64, 1);
            mbar_wait(pe_full, 0);
            // scores S = Q K^T into columns [0, 192) and the first nBmA columns of the relative-position product Bm: neither touches the
            // output accumulator O of the previous item, so they are issued right behind that item's P V product and run while the
            // softmax warps still read its O; only the window's last 64 columns -- where O sits when the window fills tensor
            // memory (row block 0) -- wait for o_drained.
            auto issue_scores_and_bm = [&](uint32_t it_next, bool wait_drain, uint32_t drain_parity) {
                mbar_wait(qk_full, it_next & 1);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                // d_k = 64 = 4 k-steps of 16
                    const uint32_t ko = ks * 32;
                    const uint64_t a_hi = make_smem_desc(q_smem + ko), a_lo = make_smem_desc(q_smem + kA2QHalf + ko);
                    const uint64_t bk_hi = make_smem_desc(k_smem + ko), bk_lo = make_smem_desc(k_smem + kA2KHalf + ko);
                    tcgen05_mma_f16(tmem_base, a_lo, bk_hi, idesc_s, ks != 0);
                    tcgen05_mma_f16(tmem_base, a_hi, bk_lo, idesc_s, 1);
                    tcgen05_mma_f16(tmem_base, a_hi, bk_hi, idesc_s, 1);
                    const uint64_t bp_hi = make_smem_desc(pe_smem + ko), bp_lo = make_smem_desc(pe_smem + kA2PeHalf + ko);
                    const uint32_t d = tmem_base + kA2ColB;
                    tcgen05_mma_f16(d, a_lo, bp_hi, idesc_ba, ks != 0);
                    tcgen05_mma_f16(d, a_hi, bp_lo, idesc_ba, 1);
                    tcgen05_mma_f16(d, a_hi, bp_hi, idesc_ba, 1);
                }
                if (wait_drain) {
                    mbar_wait(o_drained, drain_parity);         // the previous item's O has been read out of TMEM
                    tcgen05_fence_after();
                }
                if (nBmB > 0) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t ko = ks * 32;
                        const uint64_t a_hi = make_smem_desc(q_smem + ko), a_lo = make_smem_desc(q_smem + kA2QHalf + ko);
                        const uint32_t po = (uint32_t)nBmA * 128u + ko;
                        const uint64_t bp_hi = make_smem_desc(pe_smem + po), bp_lo = make_smem_desc(pe_smem + kA2PeHalf + po);
                        const uint32_t d = tmem_base + kA2ColB + nBmA;
                        tcgen05_mma_f16(d, a_lo, bp_hi, idesc_bb, ks != 0);
                        tcgen05_mma_f16(d, a_hi, bp_lo, idesc_bb, 1);
                        tcgen05_mma_f16(d, a_hi, bp_hi, idesc_bb, 1);
                    }
                }
                tcgen05_commit(qk_empty);
                tcgen05_commit(s_ready);
            };
            issue_scores_and_bm(0, false, 0);
            uint32_t it = 0;
            for (int bh = first; bh < p.n_bh; bh += stride, ++it) {
                mbar_wait(p_ready, it & 1);                     // probabilities are in TMEM
                mbar_wait(v_full, it & 1);
                tcgen05_fence_after();
                // straight-line issue (the trip count is a run-time value <= 12): descriptor arithmetic and the moves into uniform
                // registers are hoisted ahead of the waits, the 36 MMAs go out back to back -- issued one by one from a rolled loop
                // they took about as long to issue as to execute (N = 64: 32 tensor cycles each)
#pragma unroll
                for (int j = 0; j < 192 / 16; ++j) {
                    if (j >= nks_pv) break;
                    const uint32_t vo = (uint32_t)j * 2048u;                         // 16 key rows of 128 bytes
                    const uint64_t bv_hi = make_smem_desc(v_smem + vo), bv_lo = make_smem_desc(v_smem + kA2VHalf + vo);
                    const uint32_t a_col = tmem_base + 8u * j;                       // 16 keys = 8 packed columns
                    tcgen05_mma_f16_ts(tmem_base + col_o, a_col + kA2ColPlo, bv_hi, idesc_o, j != 0);
                    tcgen05_mma_f16_ts(tmem_base + col_o, a_col, bv_lo, idesc_o, 1);
                    tcgen05_mma_f16_ts(tmem_base + col_o, a_col, bv_hi, idesc_o, 1);
                }
                tcgen05_commit(v_empty);
                tcgen05_commit(o_ready);
                // the next item's scores overwrite P only after the P V product above (tensor-pipe order)
                if (bh + stride < p.n_bh) issue_scores_and_bm(it + 1, o_aliases, it & 1);
            }
        }
    } else {
        // ===== softmax / epilogue warps: kA2Split threads per query row, kA2Slots keys each
        const int q = warp & 3;                          // TMEM lane quarter
        const int hh = (warp - 2) >> 2;                  // which part of the key positions
        const int r = 32 * q + lane;                     // row inside the block
        const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
        float* xch_max = xch;                            // [kA2Split][128]
        float* xch_sum = xch + kA2Split * 128;
        const bool warp_valid = (R0 + 32 * q) < T;       // warp-uniform, fixed for the CTA's row block
        uint32_t it = 0;
        for (int bh = first; bh < p.n_bh; bh += stride, ++it) {
            mbar_wait(s_ready, it & 1);
            tcgen05_fence_after();

            uint32_t w[kA2Slots + 32];                   // fp32 bit patterns
            float mx = -INFINITY;
            if (warp_valid) {
                // window of Bm: columns kA2ColB + 32 q + u_lo + [0, kA2Slots + 32); the thread needs element
                // lane + (kA2Slots - 1) - slot
                const int u_lo = (T - 1) - (kA2Slots * hh + kA2Slots - 1);
                const uint32_t wcol = (uint32_t)(kA2ColB + 32 * q + u_lo);
#pragma unroll
                for (int k = 0; k < kA2Slots / 32 + 1; ++k) tmem_ld_32x32_nowait(tmem_base + lane_sel + wcol + 32 * k, w + 32 * k);
                tmem_ld_wait();                         // the window loads overlap
#pragma unroll
                for (int k = 0; k < kA2Slots / 32 + 1; ++k) tmem_ld_fence(w + 32 * k);
                a2_barrel_stage<16>(w, lane);
                a2_barrel_stage<8>(w, lane);
                a2_barrel_stage<4>(w, lane);
                a2_barrel_stage<2>(w, lane);
                a2_barrel_stage<1>(w, lane);
                // scores in the log2 domain: slot s <-> key t2 = kA2Slots hh + s uses w[kA2Slots - 1 - s]
#pragma unroll
                for (int cc = 0; cc < kA2Slots / 32; ++cc) {
                    uint32_t sv[32];
                    tmem_ld_32x32(tmem_base + lane_sel + (uint32_t)(kA2Slots * hh + 32 * cc), sv);
                    if (kA2Slots * hh + 32 * (cc + 1) <= T) {         // warp-uniform: every key of the chunk exists (no masking)
#pragma unroll
                        for (int jx = 0; jx < 32; ++jx) {
                            const int slot = 32 * cc + jx;
                            const float val = (__uint_as_float(sv[jx]) + __uint_as_float(w[kA2Slots - 1 - slot])) * p.scale_log2e;
                            w[kA2Slots - 1 - slot] = __float_as_uint(val);
                            mx = fmaxf(mx, val);
                        }
                    } else {
#pragma unroll
                        for (int jx = 0; jx < 32; ++jx) {
                            const int slot = 32 * cc + jx;
                            const int t2 = kA2Slots * hh + slot;
                            float val = (__uint_as_float(sv[jx]) + __uint_as_float(w[kA2Slots - 1 - slot])) * p.scale_log2e;
                            val = (t2 < T) ? val : -INFINITY;
                            w[kA2Slots - 1 - slot] = __float_as_uint(val);
                            mx = fmaxf(mx, val);
                        }
                    }
                }
            }
            // one exchange per row instead of two: every thread exponentiates against its OWN maximum and publishes (max, sum);
            // the row total is sum_o sum_o 2^(max_o - M) with M the row maximum, and a thread's probabilities are
            // e_i 2^(max_own - M) / total.  (The exponentials no longer wait for the other threads of the row.)
            const float mref = mx == -INFINITY ? 0.f : mx;           // a part with no valid key (short segments): all zeros
            float sum = 0.f;
            if (warp_valid) {
#pragma unroll
                for (int i = 0; i < kA2Slots; ++i) {
                    const float e = ex2_approx(__uint_as_float(w[i]) - mref);
                    w[i] = __float_as_uint(e);
                    sum += e;
                }
            }
            xch_max[hh * 128 + r] = mx;
            xch_sum[hh * 128 + r] = sum;
            a2_named_bar_sync(1 + q, 32 * kA2Split);
            {
                float M = xch_max[r];
#pragma unroll
                for (int o = 1; o < kA2Split; ++o) M = fmaxf(M, xch_max[o * 128 + r]);
                // every thread of the row adds the rescaled partial sums in the same (part) order: identical totals
                float tot = 0.f;
#pragma unroll
                for (int o = 0; o < kA2Split; ++o) tot = fmaf(xch_sum[o * 128 + r], ex2_approx(xch_max[o * 128 + r] - M), tot);
                sum = mx == -INFINITY ? 1.f : tot / ex2_approx(mref - M);   // p_i = e_i / sum (all e_i are zero for a part without keys)
            }
            // the exchange slots are rewritten by the next item only after this item's p_ready / o_ready round trip
            if (warp_valid) {
                const float inv = 1.f / sum;
                // packed column j of this part holds keys kA2Slots hh + 2j (low half) and 2j + 1 (high half)
#pragma unroll
                for (int cc = 0; cc < kA2Slots / 32; ++cc) {
                    uint32_t ph[16], pl[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        const int j = 16 * cc + jj;
                        const float p0 = __uint_as_float(w[kA2Slots - 1 - 2 * j]) * inv;
                        const float p1 = __uint_as_float(w[kA2Slots - 2 - 2 * j]) * inv;
                        ph[jj] = pack_bf16(p0, p1);                                   // one packed convert (F2FP)
                        pl[jj] = pack_bf16(p0 - __uint_as_float(ph[jj] << 16), p1 - __uint_as_float(ph[jj] & 0xffff0000u));
                    }
                    tmem_st_32x16(tmem_base + lane_sel + (uint32_t)(kA2Slots / 2 * hh + 16 * cc), ph);
                    tmem_st_32x16(tmem_base + lane_sel + (uint32_t)(kA2ColPlo + kA2Slots / 2 * hh + 16 * cc), pl);
                }
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            // ---- O = P V: 32 columns per warp (parts 0 and 1), split and stored to the [M][d_model] activation
            mbar_wait(o_ready, it & 1);
            tcgen05_fence_after();
            if (warp_valid && hh < 2) {
                uint32_t ov[32];
                tmem_ld_32x32(tmem_base + lane_sel + col_o + (uint32_t)(32 * hh), ov);
                const int t1 = R0 + r;
                if (t1 < T) {
                    const int seg = bh / p.n_heads, h = bh - seg * p.n_heads;
                    const size_t o = ((size_t)seg * T + t1) * p.ldo + (size_t)h * kA2Dk + 32 * hh;
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        float v8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v8[e] = __uint_as_float(ov[8 * k8 + e]);
                        split_store8(p.out_fmt, p.out_hi, p.out_lo, o + 8 * k8, v8);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_drained);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// q, k, v: [n_bh][T][64] bf16 pairs (v row-major like k: the P V product reads it as an MN-major operand); pe: [2 * maxlen][64] bf16 pairs;
// out: [n_seg * T][ldo] split in out_fmt.  The float pointers are reinterpreted as uint16_t arrays (SplitFmt).
int attn16_launch(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo, const float* vt_hi,
                  const float* vt_lo, const float* pe_hi, const float* pe_lo, int maxlen, int n_seg, int n_heads, int T, int Tp,
                  float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream) {
    if (T < 2 || T > 192 || Tp % 32 != 0 || Tp < T || Tp > 192) { set_error("attn16: T=%d Tp=%d unsupported", T, Tp); return NSF_ERR_UNSUPPORTED; }
    if (maxlen < T || (ldo & 7) || ((uintptr_t)out_hi & 15) || ((uintptr_t)out_lo & 15)) {
        set_error("attn16: maxlen=%d ldo=%lld (ldo must be a multiple of 8, outputs 16-byte aligned)", maxlen, (long long)ldo);
        return NSF_ERR_INVALID_ARG;
    }
    const int n_bh = n_seg * n_heads;
    CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mp_hi, mp_lo, mv_hi, mv_lo;
    int rc;
    if ((rc = make_tmap_kmajor16(&mq_hi, q_hi, T, kA2Dk, kA2Dk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor16(&mq_lo, q_lo, T, kA2Dk, kA2Dk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor16(&mk_hi, k_hi, T, kA2Dk, kA2Dk, n_bh, 0, 192))) return rc;
    if ((rc = make_tmap_kmajor16(&mk_lo, k_lo, T, kA2Dk, kA2Dk, n_bh, 0, 192))) return rc;
    if ((rc = make_tmap_kmajor16(&mp_hi, pe_hi, 2 * (int64_t)maxlen, kA2Dk, kA2Dk, 1, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor16(&mp_lo, pe_lo, 2 * (int64_t)maxlen, kA2Dk, kA2Dk, 1, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor16(&mv_hi, vt_hi, T, kA2Dk, kA2Dk, n_bh, 0, 192))) return rc;
    if ((rc = make_tmap_kmajor16(&mv_lo, vt_lo, T, kA2Dk, kA2Dk, n_bh, 0, 192))) return rc;
    Attn16Params p;
    p.n_bh = n_bh; p.n_heads = n_heads; p.T = T; p.Tp = Tp; p.pe_row0 = maxlen - (T - 1);
    p.n_rb = (T + 127) / 128;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)kA2Dk);
    p.out_hi = out_hi; p.out_lo = out_lo; p.ldo = ldo; p.out_fmt = out_fmt;
    NSF_CUDA(cudaFuncSetAttribute(attn16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2SmemBytes));
    int grid;
    if (p.n_rb == 1) {
        grid = n_bh < sm_count() ? n_bh : sm_count();
        p.g0 = grid;
    } else {
        grid = 2 * n_bh < sm_count() ? 2 * n_bh : sm_count();
        // both row blocks cost a full 128-row MMA pass; block 1 has the narrower pe_k window and fewer live softmax rows
        static const float g0_frac = [] { const char* e = getenv("NSF_ATTN_G0"); const float v = e ? (float)atof(e) : 0.f; return v > 0.f && v < 1.f ? v : 0.55f; }();
        p.g0 = (int)lroundf(grid * g0_frac);
        if (p.g0 < 1) p.g0 = 1;
        if (p.g0 > grid - 1) p.g0 = grid - 1;
    }
    attn16_kernel<<<grid, kA2Threads, kA2SmemBytes, stream>>>(mq_hi, mq_lo, mk_hi, mk_lo, mp_hi, mp_lo, mv_hi, mv_lo, p);
    return check_launch("attn16_kernel");
}

// ------------------------------------------------------------------------------------------- test hook
__global__ void __launch_bounds__(256)
attn16_test_split_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_store(SPLIT_BF16, hi, lo, (size_t)i, in[i]);
}
__global__ void __launch_bounds__(256)
attn16_test_merge_kernel(const float* __restrict__ hi, const float* __restrict__ lo, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hi[i] + lo[i];
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_attention16_test(const float* q, const float* k, const float* v, const float* pe, int maxlen, int n_seg,
                                    int n_heads, int T, float* out, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(q && k && v && pe && out && workspace, "nsf_attention16_test: null pointer");
    NSF_REQUIRE(attn_fused_supported(T, kA2Dk) && n_seg > 0 && n_heads > 0 && maxlen >= T, "nsf_attention16_test: T=%d maxlen=%d", T, maxlen);
    NSF_REQUIRE(workspace_bytes >= nsf_attention_test_workspace_bytes(n_seg, n_heads, T, maxlen) && ((uintptr_t)workspace & 255) == 0,
                "nsf_attention16_test: workspace too small or not 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t n_bh = (int64_t)n_seg * n_heads, Tp = (T + 31) / 32 * 32;
    const int64_t nq = n_bh * T * kA2Dk, nv = n_bh * kA2Dk * Tp, np = 2 * (int64_t)maxlen * kA2Dk;
    // same carve-up as nsf_attention_test (float-sized slots; the bf16 planes use half of each)
    float* w = reinterpret_cast<float*>(workspace);
    float *q_hi = w, *q_lo = q_hi + nq, *k_hi = q_lo + nq, *k_lo = k_hi + nq, *v_hi = k_lo + nq, *v_lo = v_hi + nv;
    float *p_hi = v_lo + nv, *p_lo = p_hi + np, *o_hi = p_lo + np, *o_lo = o_hi + nq;
    attn16_test_split_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(q, nq, q_hi, q_lo);
    attn16_test_split_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(k, nq, k_hi, k_lo);
    attn16_test_split_kernel<<<(unsigned)ceil_div64(np, 256), 256, 0, s>>>(pe, np, p_hi, p_lo);
    attn16_test_split_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(v, nq, v_hi, v_lo);
    int rc = check_launch("attn16_test_split_kernel");
    if (rc) return rc;
    {
        ProfScope prof(PROF_ATTN, 6.0 * T * T * kA2Dk * (double)n_bh, s);
        if ((rc = attn16_launch(q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, p_hi, p_lo, maxlen, n_seg, n_heads, T, (int)Tp, o_hi, o_lo,
                                (int64_t)n_heads * kA2Dk, SPLIT_TF32, s))) return rc;
    }
    attn16_test_merge_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(o_hi, o_lo, nq, out);
    return check_launch("attn16_test_merge_kernel");
}
