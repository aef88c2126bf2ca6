// Fused relative-position multi-head self-attention of the Conformer blocks, one persistent tcgen05 kernel.
//
// Reference: css/css_with_conformer/nnet/conformer.py  MultiHeadedAttention.forward :66-92
//     A = q k^T (:73);  B[t1,t2] = q[t1] . pe_k[t1 - t2] (:74-77);  p = softmax((A + B) / sqrt(d_k)) (:78,87);
//     o = p v (:90)
// with pe_k rows taken from Embedding(2*maxlen, d_k) at clamp(t1 - t2) + maxlen (RelativePositionalEncoding :23-29).
//
// Work item = (segment, head, block of 128 query rows).  Everything between the QKV projection and the output
// projection stays on chip:
//   TMA:      Q block (128 x 64), then 64-row chunks of K, of the pe_k window this row block can reach, and of V^T
//             stream through a 4-slot shared-memory ring (all operands pre-split into TF32 head + remainder).
//   tcgen05:  S = Q K^T into TMEM columns [0, 192), Bm = Q PEw^T into columns [192, 512) (3xTF32, fp32 accumulate).
//   softmax:  8 warps, two threads per query row.  Row r of Bm is needed skewed -- score[t2] uses Bm[r][r + (T-1) - t2]
//             -- and tcgen05.ld gives every lane the same columns, so each thread loads a 128-column window and
//             shifts it by its lane index with a 5-stage barrel shifter in registers.  exp / sum in fp32.
//   P V:      the probabilities go back to TMEM (split again, over the consumed S / Bm columns) and feed the
//             tensor core as the A operand straight from tensor memory; V^T comes from the ring.  O (128 x 64)
//             lands in columns [384, 448), is split and written to the [M][d_model] activation the output
//             projection reads.
// T <= 192 frames per segment (the pipeline uses 186); longer segments take the unfused path in conformer.cu.
#include "gemm_common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

namespace nsf {

constexpr int kAttnThreads = 320;              // warp 0 TMA, warp 1 MMA, warps 2..9 softmax / epilogue
constexpr int kAttnDk = 64;
constexpr int kAttnQBytes = 4 * 16384;         // Q hi kb0 | hi kb1 | lo kb0 | lo kb1, each 128 rows x 128 B
constexpr int kAttnSlotBytes = 4 * 8192;       // 64-row operand chunk: hi kb0 | hi kb1 | lo kb0 | lo kb1
constexpr int kAttnRing = 4;
constexpr int kAttnColB = 192;                 // first TMEM column of Bm (and of P_lo)
constexpr int kAttnColO = 384;                 // first TMEM column of O
constexpr int kAttnSlotsPerThread = 96;        // key positions per softmax thread (two threads per row)
constexpr int kAttnSmemBytes = kAttnQBytes + kAttnRing * kAttnSlotBytes + 2048 /*pair exchange*/ + 256 /*barriers*/ + 1024;

struct AttnParams {
    int n_bh;          // segments * heads
    int n_heads;
    int T, Tp;         // frames per segment, V^T row pitch (multiple of 32)
    int pe_row0;       // first pe_k row of the window: maxlen - (T - 1)
    float inv_sqrt_dk;
    float* out_hi; float* out_lo; int64_t ldo;      // [n_seg * T][d_model], column head * 64 + d
    int out_fmt;                                    // SplitFmt of the output pair
};

__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// One stage of the register barrel shifter: lanes with bit SH set move w[i + SH] down to w[i]; after the stage the
// remaining shift is < SH, so only elements [0, 96 + SH - 1) are still needed.
template <int SH>
__device__ __forceinline__ void barrel_stage(uint32_t (&w)[128], int lane) {
    const bool on = (lane & SH) != 0;
#pragma unroll
    for (int i = 0; i < kAttnSlotsPerThread + SH - 1; ++i) w[i] = on ? w[i + SH] : w[i];
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                  const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                  const __grid_constant__ CUtensorMap map_pe_hi, const __grid_constant__ CUtensorMap map_pe_lo,
                  const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                  const AttnParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t q_smem = base;
    const uint32_t ring_smem = base + kAttnQBytes;
    float* xch = reinterpret_cast<float*>(gen + kAttnQBytes + kAttnRing * kAttnSlotBytes);          // [2 (max, sum)][2 (half)][128]
    const uint32_t bars = base + kAttnQBytes + kAttnRing * kAttnSlotBytes + 2048;
    const uint32_t q_full = bars, q_empty = bars + 8;
    auto ring_full = [&](int s) { return bars + 16u + 8u * s; };
    auto ring_empty = [&](int s) { return bars + 16u + 8u * (kAttnRing + s); };
    const uint32_t s_ready = bars + 16u + 8u * (2 * kAttnRing);
    const uint32_t p_ready = s_ready + 8, o_ready = s_ready + 16, o_drained = s_ready + 24;
    const uint32_t tmem_ptr_addr = s_ready + 32;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int n_rb = (T + 127) / 128;
    const int total = p.n_bh * n_rb;
    const int nS = (T + 63) / 64;                       // K chunks
    const int n_kb_v = p.Tp / 32;                       // 32-frame k-blocks of V^T
    const int nV = (n_kb_v + 1) / 2;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int s = 0; s < kAttnRing; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), 1); }
        mbar_init(s_ready, 1); mbar_init(p_ready, 8); mbar_init(o_ready, 1); mbar_init(o_drained, 8);
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
            uint32_t rit = 0, it = 0;
            auto ring_acquire = [&](uint32_t bytes) -> uint32_t {
                const int s = rit % kAttnRing;
                const uint32_t ph = (rit / kAttnRing) & 1;
                mbar_wait(ring_empty(s), ph ^ 1);
                mbar_expect_tx(ring_full(s), bytes);
                return (uint32_t)s;
            };
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
                const int bh = item / n_rb, rb = item - bh * n_rb;
                const int R0 = rb * 128;
                const int rows_valid = min(128, T - R0);
                const int nB = (rows_valid + T - 1 + 63) / 64;          // pe_k window chunks this row block can reach
                mbar_wait(q_empty, (it & 1) ^ 1);
                mbar_expect_tx(q_full, kAttnQBytes);
                tma_load_3d(q_smem, &map_q_hi, 0, R0, bh, q_full);
                tma_load_3d(q_smem + 16384, &map_q_hi, 32, R0, bh, q_full);
                tma_load_3d(q_smem + 32768, &map_q_lo, 0, R0, bh, q_full);
                tma_load_3d(q_smem + 49152, &map_q_lo, 32, R0, bh, q_full);
                for (int c = 0; c < nS + nB; ++c, ++rit) {
                    const uint32_t s = ring_acquire(kAttnSlotBytes);
                    const uint32_t dst = ring_smem + s * kAttnSlotBytes;
                    const bool is_k = c < nS;
                    const CUtensorMap* mh = is_k ? &map_k_hi : &map_pe_hi;
                    const CUtensorMap* ml = is_k ? &map_k_lo : &map_pe_lo;
                    const int row = is_k ? 64 * c : p.pe_row0 + R0 + 64 * (c - nS);
                    const int bb = is_k ? bh : 0;
                    tma_load_3d(dst, mh, 0, row, bb, ring_full(s));
                    tma_load_3d(dst + 8192, mh, 32, row, bb, ring_full(s));
                    tma_load_3d(dst + 16384, ml, 0, row, bb, ring_full(s));
                    tma_load_3d(dst + 24576, ml, 32, row, bb, ring_full(s));
                }
                for (int v = 0; v < nV; ++v, ++rit) {
                    const int nk = min(2, n_kb_v - 2 * v);
                    const uint32_t s = ring_acquire((uint32_t)nk * 16384u);
                    const uint32_t dst = ring_smem + s * kAttnSlotBytes;
                    for (int kk = 0; kk < nk; ++kk) {
                        tma_load_3d(dst + kk * 8192, &map_v_hi, (2 * v + kk) * 32, 0, bh, ring_full(s));
                        tma_load_3d(dst + 16384 + kk * 8192, &map_v_lo, (2 * v + kk) * 32, 0, bh, ring_full(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = make_idesc_tf32(64);
            uint32_t rit = 0, it = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
                const int bh = item / n_rb, rb = item - bh * n_rb;
                const int rows_valid = min(128, T - rb * 128);
                const int nB = (rows_valid + T - 1 + 63) / 64;
                mbar_wait(o_drained, (it & 1) ^ 1);             // previous item's O has been read out of TMEM
                mbar_wait(q_full, it & 1);
                tcgen05_fence_after();
                for (int c = 0; c < nS + nB; ++c, ++rit) {
                    const int s = rit % kAttnRing;
                    mbar_wait(ring_full(s), (rit / kAttnRing) & 1);
                    tcgen05_fence_after();
                    const uint32_t bt = ring_smem + s * kAttnSlotBytes;
                    const uint32_t d_tmem = tmem_base + (c < nS ? 64 * c : kAttnColB + 64 * (c - nS));
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t ko = ks * 32;
                            const uint64_t a_hi = make_smem_desc(q_smem + kb * 16384 + ko);
                            const uint64_t a_lo = make_smem_desc(q_smem + 32768 + kb * 16384 + ko);
                            const uint64_t b_hi = make_smem_desc(bt + kb * 8192 + ko);
                            const uint64_t b_lo = make_smem_desc(bt + 16384 + kb * 8192 + ko);
                            tcgen05_mma_tf32(d_tmem, a_lo, b_hi, idesc, (kb | ks) != 0);
                            tcgen05_mma_tf32(d_tmem, a_hi, b_lo, idesc, 1);
                            tcgen05_mma_tf32(d_tmem, a_hi, b_hi, idesc, 1);
                        }
                    tcgen05_commit(ring_empty(s));
                }
                tcgen05_commit(q_empty);
                tcgen05_commit(s_ready);
                mbar_wait(p_ready, it & 1);                     // probabilities are in TMEM
                tcgen05_fence_after();
                for (int v = 0; v < nV; ++v, ++rit) {
                    const int s = rit % kAttnRing;
                    mbar_wait(ring_full(s), (rit / kAttnRing) & 1);
                    tcgen05_fence_after();
                    const uint32_t bt = ring_smem + s * kAttnSlotBytes;
                    const int nk = min(2, n_kb_v - 2 * v);
                    for (int kk = 0; kk < nk; ++kk)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t col = (uint32_t)((2 * v + kk) * 32 + ks * 8);
                            const uint64_t b_hi = make_smem_desc(bt + kk * 8192 + ks * 32);
                            const uint64_t b_lo = make_smem_desc(bt + 16384 + kk * 8192 + ks * 32);
                            tcgen05_mma_tf32_ts(tmem_base + kAttnColO, tmem_base + kAttnColB + col, b_hi, idesc, (v | kk | ks) != 0);
                            tcgen05_mma_tf32_ts(tmem_base + kAttnColO, tmem_base + col, b_lo, idesc, 1);
                            tcgen05_mma_tf32_ts(tmem_base + kAttnColO, tmem_base + col, b_hi, idesc, 1);
                        }
                    tcgen05_commit(ring_empty(s));
                }
                tcgen05_commit(o_ready);
            }
        }
    } else {
        // ===== softmax / epilogue warps
        const int q = warp & 3;                          // TMEM lane quarter
        const int hh = (warp - 2) >> 2;                  // which half of the key positions
        const int r = 32 * q + lane;                     // row inside the block
        const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
        float* xch_max = xch;                            // [2][128]
        float* xch_sum = xch + 256;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
            const int bh = item / n_rb, rb = item - bh * n_rb;
            const int R0 = rb * 128;
            const bool warp_valid = (R0 + 32 * q) < T;   // warp-uniform
            mbar_wait(s_ready, it & 1);
            tcgen05_fence_after();

            uint32_t w[128];                             // fp32 bit patterns (kept as b32 for the TMEM round trips)
            float mx = -INFINITY;
            if (warp_valid) {
                // window of Bm: columns kAttnColB + 32 q + u_lo + [0, 128); the thread needs element lane + 95 - slot
                const int u_lo = (T - 1) - (kAttnSlotsPerThread * hh + kAttnSlotsPerThread - 1);
                const uint32_t wcol = (uint32_t)(kAttnColB + 32 * q + u_lo);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t tmp[32];
                    tmem_ld_32x32(tmem_base + lane_sel + wcol + 32 * k, tmp);
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[32 * k + j] = tmp[j];
                }
                // barrel shifter: w[i] <- w[i + lane]
                barrel_stage<16>(w, lane);
                barrel_stage<8>(w, lane);
                barrel_stage<4>(w, lane);
                barrel_stage<2>(w, lane);
                barrel_stage<1>(w, lane);
                // scores: slot s <-> key t2 = 96 hh + s uses w[95 - s]
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    uint32_t sv[32];
                    tmem_ld_32x32(tmem_base + lane_sel + (uint32_t)(kAttnSlotsPerThread * hh + 32 * cc), sv);
#pragma unroll
                    for (int jx = 0; jx < 32; ++jx) {
                        const int slot = 32 * cc + jx;
                        const int t2 = kAttnSlotsPerThread * hh + slot;
                        float val = (__uint_as_float(sv[jx]) + __uint_as_float(w[kAttnSlotsPerThread - 1 - slot])) * p.inv_sqrt_dk;
                        val = (t2 < T) ? val : -INFINITY;
                        w[kAttnSlotsPerThread - 1 - slot] = __float_as_uint(val);
                        mx = fmaxf(mx, val);
                    }
                }
            }
            xch_max[hh * 128 + r] = mx;
            named_bar_sync(1 + q, 64);
            mx = fmaxf(mx, xch_max[(hh ^ 1) * 128 + r]);
            float sum = 0.f;
            if (warp_valid) {
#pragma unroll
                for (int i = 0; i < kAttnSlotsPerThread; ++i) {
                    const float e = expf(__uint_as_float(w[i]) - mx);
                    w[i] = __float_as_uint(e);
                    sum += e;
                }
            }
            xch_sum[hh * 128 + r] = sum;
            named_bar_sync(1 + q, 64);
            sum += xch_sum[(hh ^ 1) * 128 + r];
            if (warp_valid) {
                const float inv = 1.f / sum;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    uint32_t pv[32];
#pragma unroll
                    for (int jx = 0; jx < 32; ++jx) {
                        const int idx = kAttnSlotsPerThread - 1 - (32 * cc + jx);
                        w[idx] = __float_as_uint(__uint_as_float(w[idx]) * inv);
                        pv[jx] = w[idx] & 0xffffe000u;                                     // TF32 head
                    }
                    tmem_st_32x32(tmem_base + lane_sel + (uint32_t)(kAttnSlotsPerThread * hh + 32 * cc), pv);
#pragma unroll
                    for (int jx = 0; jx < 32; ++jx) {
                        const int idx = kAttnSlotsPerThread - 1 - (32 * cc + jx);
                        pv[jx] = __float_as_uint(__uint_as_float(w[idx]) - __uint_as_float(pv[jx]));   // exact remainder
                    }
                    tmem_st_32x32(tmem_base + lane_sel + (uint32_t)(kAttnColB + kAttnSlotsPerThread * hh + 32 * cc), pv);
                }
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            // ---- O = P V: 32 columns per warp, split and stored to the [M][d_model] activation
            mbar_wait(o_ready, it & 1);
            tcgen05_fence_after();
            if (warp_valid) {
                uint32_t ov[32];
                tmem_ld_32x32(tmem_base + lane_sel + (uint32_t)(kAttnColO + 32 * hh), ov);
                const int t1 = R0 + r;
                if (t1 < T) {
                    const int seg = bh / p.n_heads, h = bh - seg * p.n_heads;
                    const size_t o = ((size_t)seg * T + t1) * p.ldo + (size_t)h * kAttnDk + 32 * hh;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4)
                        split_store4(p.out_fmt, p.out_hi, p.out_lo, o + 4 * k4,
                                     make_float4(__uint_as_float(ov[4 * k4]), __uint_as_float(ov[4 * k4 + 1]),
                                                 __uint_as_float(ov[4 * k4 + 2]), __uint_as_float(ov[4 * k4 + 3])));
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

// q, k: [n_bh][T][64] split; vt: [n_bh][64][Tp] split; pe: [2 * maxlen][64] split; out: [n_seg * T][ldo] split.
int attn_fused_launch(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo, const float* vt_hi,
                      const float* vt_lo, const float* pe_hi, const float* pe_lo, int maxlen, int n_seg, int n_heads, int T, int Tp,
                      float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream) {
    if (T < 2 || T > 192 || Tp % 32 != 0 || Tp < T || Tp > 192) { set_error("attn_fused: T=%d Tp=%d unsupported", T, Tp); return NSF_ERR_UNSUPPORTED; }
    if (maxlen < T || (ldo & 3)) { set_error("attn_fused: maxlen=%d ldo=%lld", maxlen, (long long)ldo); return NSF_ERR_INVALID_ARG; }
    const int n_bh = n_seg * n_heads;
    CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mp_hi, mp_lo, mv_hi, mv_lo;
    int rc;
    if ((rc = make_tmap_kmajor(&mq_hi, q_hi, T, kAttnDk, kAttnDk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor(&mq_lo, q_lo, T, kAttnDk, kAttnDk, n_bh, 0, 128))) return rc;
    if ((rc = make_tmap_kmajor(&mk_hi, k_hi, T, kAttnDk, kAttnDk, n_bh, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor(&mk_lo, k_lo, T, kAttnDk, kAttnDk, n_bh, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor(&mp_hi, pe_hi, 2 * (int64_t)maxlen, kAttnDk, kAttnDk, 1, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor(&mp_lo, pe_lo, 2 * (int64_t)maxlen, kAttnDk, kAttnDk, 1, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor(&mv_hi, vt_hi, kAttnDk, Tp, Tp, n_bh, 0, 64))) return rc;
    if ((rc = make_tmap_kmajor(&mv_lo, vt_lo, kAttnDk, Tp, Tp, n_bh, 0, 64))) return rc;
    AttnParams p;
    p.n_bh = n_bh; p.n_heads = n_heads; p.T = T; p.Tp = Tp; p.pe_row0 = maxlen - (T - 1);
    p.inv_sqrt_dk = 1.f / sqrtf((float)kAttnDk);
    p.out_hi = out_hi; p.out_lo = out_lo; p.ldo = ldo; p.out_fmt = out_fmt;
    NSF_CUDA(cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    const int total = n_bh * ((T + 127) / 128);
    const int grid = total < sm_count() ? total : sm_count();
    attn_fused_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(mq_hi, mq_lo, mk_hi, mk_lo, mp_hi, mp_lo, mv_hi, mv_lo, p);
    return check_launch("attn_fused_kernel");
}


// ------------------------------------------------------------------------------------------- test hook
__global__ void __launch_bounds__(256)
attn_test_split_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float h, l; split_tf32(in[i], h, l); hi[i] = h; lo[i] = l; }
}
// v [n_bh][T][64] -> vt [n_bh][64][Tp] split (frames >= T zero)
__global__ void __launch_bounds__(256)
attn_test_vt_kernel(const float* __restrict__ v, int n_bh, int T, int Tp, float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_bh * kAttnDk * Tp) return;
    const int t = (int)(i % Tp);
    const int d = (int)((i / Tp) % kAttnDk);
    const int64_t bh = i / ((int64_t)Tp * kAttnDk);
    float h = 0.f, l = 0.f;
    if (t < T) split_tf32(v[(bh * T + t) * kAttnDk + d], h, l);
    hi[i] = h; lo[i] = l;
}
__global__ void __launch_bounds__(256)
attn_test_merge_kernel(const float* __restrict__ hi, const float* __restrict__ lo, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hi[i] + lo[i];
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_attention_test_workspace_bytes(int n_seg, int n_heads, int T, int maxlen) {
    const int64_t n_bh = (int64_t)n_seg * n_heads, Tp = (T + 31) / 32 * 32;
    const int64_t floats = 4 * n_bh * T * kAttnDk + 2 * n_bh * kAttnDk * Tp + 4 * (int64_t)maxlen * kAttnDk + 2 * n_bh * T * kAttnDk;
    return floats * 4 + 1024;
}

extern "C" int nsf_attention_test(const float* q, const float* k, const float* v, const float* pe, int maxlen, int n_seg,
                                  int n_heads, int T, float* out, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(q && k && v && pe && out && workspace, "nsf_attention_test: null pointer");
    NSF_REQUIRE(attn_fused_supported(T, kAttnDk) && n_seg > 0 && n_heads > 0 && maxlen >= T, "nsf_attention_test: T=%d maxlen=%d", T, maxlen);
    NSF_REQUIRE(workspace_bytes >= nsf_attention_test_workspace_bytes(n_seg, n_heads, T, maxlen) && ((uintptr_t)workspace & 255) == 0,
                "nsf_attention_test: workspace too small or not 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t n_bh = (int64_t)n_seg * n_heads, Tp = (T + 31) / 32 * 32;
    const int64_t nq = n_bh * T * kAttnDk, nv = n_bh * kAttnDk * Tp, np = 2 * (int64_t)maxlen * kAttnDk;
    float* w = reinterpret_cast<float*>(workspace);
    float *q_hi = w, *q_lo = q_hi + nq, *k_hi = q_lo + nq, *k_lo = k_hi + nq, *v_hi = k_lo + nq, *v_lo = v_hi + nv;
    float *p_hi = v_lo + nv, *p_lo = p_hi + np, *o_hi = p_lo + np, *o_lo = o_hi + nq;
    attn_test_split_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(q, nq, q_hi, q_lo);
    attn_test_split_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(k, nq, k_hi, k_lo);
    attn_test_split_kernel<<<(unsigned)ceil_div64(np, 256), 256, 0, s>>>(pe, np, p_hi, p_lo);
    attn_test_vt_kernel<<<(unsigned)ceil_div64(nv, 256), 256, 0, s>>>(v, (int)n_bh, T, (int)Tp, v_hi, v_lo);
    int rc = check_launch("attn_test_split_kernel");
    if (rc) return rc;
    {
        ProfScope prof(PROF_ATTN, 6.0 * T * T * kAttnDk * (double)n_bh, s);
        if ((rc = attn_fused_launch(q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, p_hi, p_lo, maxlen, n_seg, n_heads, T, (int)Tp, o_hi, o_lo,
                                    (int64_t)n_heads * kAttnDk, SPLIT_TF32, s))) return rc;
    }
    attn_test_merge_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, s>>>(o_hi, o_lo, nq, out);
    return check_launch("attn_test_merge_kernel");
}
