// tcgen05 GEMM engine for sm_100a: persistent, warp-specialised.  Two kernels: gemm_tc_kernel (one CTA per tile, described first)
// and gemm_tc2_kernel (CTA pairs, tcgen05.mma.cta_group::2 -- the default for the 16-bit engines' unbatched GEMMs, see the section
// "CTA pairs" below); they share the epilogue code.
//
//   D[b][m][n] = sum_k A[b][m][k] B[b][n][k]          both operands K-major
//
// TMA (SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma with two fp32 accumulators in TMEM (double buffered,
// 2 x 256 columns) -> tcgen05.ld epilogue staged through a swizzled shared-memory transpose so that every lane
// owns 8 consecutive output columns (16-byte global accesses), with the fused ops of gemm_common.cuh.
//
// MODE selects the arithmetic (template parameter; operands arrive pre-split by the producing kernels):
//   16   2xBF16 / 2xF16: D = A_lo B_hi + A_hi B_lo + A_hi B_hi, three kind::f16 MMAs on 16-bit head / remainder planes
//        (bf16: 16 mantissa bits, fp32 range; fp16: 22 bits on power-of-two-scaled values), BLOCK_K = 64
//   116  plain bf16, one kind::f16 MMA (the Whisper encoder), BLOCK_K = 64, 4 stages
//   3    3xTF32: the same three-term form on kind::tf32 (TF32 head with 13 cleared bits + exact fp32 remainder), BLOCK_K = 32
//   1    single-pass TF32
//
// CTA = 10 warps, one CTA per SM, grid = min(#tiles, #SMs), tiles handed out round-robin (n fastest, so the
// CTAs of a wave share A row-blocks through L2):
//   warp 0 lane 0  TMA producer            warp 1  TMEM allocator + lane 0 MMA issuer
//   warps 2..9     epilogue (warp w reads the 32 TMEM lanes of quarter w % 4; w and w + 4 split the columns)
// Tile 128 x 256, one 128-byte swizzle row of K per stage; 2 stages of 96 KB for the split modes, 4 of 48 KB for the
// single-pass ones.  While the epilogue warps drain accumulator `acc`, the MMA warp already fills `acc^1`.
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace nsf {

constexpr int TBM = 128;                           // tile rows; tile columns TN = 256 (default) or 32 (small-M GEMMs, see gemm_tc_launch)
constexpr int kTcThreads = 320;                   // 1 TMA + 1 MMA + 8 epilogue warps
constexpr uint32_t kTmemCols = 512;                // two 128 x 256 fp32 accumulators
constexpr int kEpiPitch = 33;                      // staging row pitch (floats): conflict-free both ways
constexpr int kEpiBytes = 8 * 32 * kEpiPitch * 4;  // one 32 x 32 staging tile per epilogue warp

// MODE 1: one kind::tf32 pass; 3: 3xTF32; 16: three kind::f16 MMAs on 16-bit head/remainder pairs (2xBF16 / 2xF16)
// RB: bytes of K per shared-memory row = swizzle span (128: SWIZZLE_128B, the default; 64: SWIZZLE_64B -- half-size stages,
// twice as many of them: the same bytes in flight at a finer grain; an experiment, see gemm_tc_launch).
template <int MODE, int TN, int RB = 128> struct TcCfg {
    static constexpr bool kSplit = MODE == 3 || MODE == 16;
    static constexpr int kRowBytes = RB;
    static constexpr int kBlockK = RB / (MODE >= 16 ? 2 : 4);     // elements of K per stage (one swizzle row)
    static constexpr int kATileBytes = TBM * RB;                  // 16 KB (RB = 128)
    static constexpr int kBTileBytes = TN * RB;                   // 32 KB (TN = 256) or 4 KB (TN = 32)
    static constexpr int kStageBytes = (kSplit ? 2 : 1) * (kATileBytes + kBTileBytes);
    static constexpr int kStages = (TN == 32 ? 8 : (kSplit ? 2 : 4)) * (128 / RB);
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 256 /*barriers*/ + 1024 /*alignment slack*/;
};




// ------------------------------------------------------------------------------------------- fused epilogue
// One 32 (rows r0..) x 32 (columns nc..) chunk of the accumulator, r[j] = D[r0 + lane][nc + j].
// Outputs whose fastest index is the column (row-major activations) go through a shared-memory transpose so
// that each store instruction writes one full 128-byte row segment; outputs whose fastest index is the row
// (V^T, the [F][T] masks) are stored straight from the TMEM register layout.  All loops are fully unrolled:
// a single warp per scheduler has to cover its own latencies.
struct RowWalker {      // (segment, frame) of consecutive rows m = seg * T + t without divisions in the loop
    int seg, t, T;
    __device__ __forceinline__ RowWalker(int m, int T_) : seg(m / T_), t(m - (m / T_) * T_), T(T_) {}
    __device__ __forceinline__ void next() { if (++t == T) { t = 0; ++seg; } }
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk_t(const GemmParams& p, int b, int r0, int nc, const uint32_t (&r)[32],
                                               float* stage, int lane, float (&ln_s)[4], float (&ln_q)[4],
                                               const size_t (&qkv_row)[4], const float4 (&xo)[4][2]) {
    const bool lane_is_row = EPI == EPI_MASK || (EPI == EPI_QKV && nc >= 2 * p.d_model && !p.v_rowmajor);
    if (lane_is_row) {
        const int m = r0 + lane;
        if (m >= p.M) return;
        const int seg = m / p.T, t = m - seg * p.T;
        if (EPI == EPI_MASK) {
            const int n_masks = p.n_valid / kBins;
            int k = nc / kBins, f = nc - k * kBins;
            float* base = p.out0 + (size_t)seg * n_masks * kBins * p.T + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (nc + j < p.N) {
                    const float v = __uint_as_float(r[j]) * p.acc_scale + (p.bias ? __ldg(p.bias + nc + j) : 0.f);
                    base[((size_t)k * kBins + f) * p.T] = __fdividef(1.f, 1.f + __expf(-v));   // torch.sigmoid (conformer.py:304), MUFU exp / rcp (2 ulp)
                }
                if (++f == kBins) { f = 0; ++k; }
            }
        } else {    // V^T of EPI_QKV: [seg][head][d][Tp]; a 32-column chunk stays inside one head (d_k % 32 == 0)
            const int c = nc - 2 * p.d_model;
            const int h = c / p.d_k, d0 = c - h * p.d_k;
            const size_t o = (((size_t)seg * p.n_heads + h) * p.d_k + d0) * p.Tp + t;
            const float2 st = p.ln_stats ? __ldg(p.ln_stats + m) : make_float2(0.f, 1.f);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float v = __uint_as_float(r[j]) * p.acc_scale;
                if (p.ln_stats) v = (v - st.x * __ldg(p.ln_csum + nc + j)) * st.y;
                split_store(p.qkv_fmt, p.vt_hi, p.vt_lo, o + (size_t)j * p.Tp, v + (p.bias ? __ldg(p.bias + nc + j) : 0.f));
            }
        }
        return;
    }
    if (p.vec8) {
        // ---- vector path: swizzled 32 x 32 staging (float4 group g of row l at group g ^ (l & 7): conflict-free both
        // ways), then every lane owns 8 consecutive columns (cg) of rows rq, rq + 8, rq + 16, rq + 24: 16-byte accesses
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(stage + lane * 32 + 4 * (jj ^ (lane & 7))) =
                make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]), __uint_as_float(r[4 * jj + 2]),
                            __uint_as_float(r[4 * jj + 3]));
        __syncwarp();
        const int cg = lane & 3, rq = lane >> 2;
        const int n0 = nc + 8 * cg;
        if (n0 < p.N) {
            float bs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) bs[e] = 0.f;
            if (p.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4));
                bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
            }
            // folded LayerNorm of the operand rows: per-row (mean, rstd), per-column sums of the scaled weights
            float cs[8];
            float2 lst[4];
            if ((EPI == EPI_QKV || EPI == EPI_RELU_SPLIT) && p.ln_stats) {
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.ln_csum + n0)), c1 = __ldg(reinterpret_cast<const float4*>(p.ln_csum + n0 + 4));
                cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int m = r0 + rq + 8 * k;
                    lst[k] = __ldg(p.ln_stats + (m < p.M ? m : r0));
                }
            }
            const int g0 = (2 * cg) ^ rq, g1 = (2 * cg + 1) ^ rq;
            // EPI_QKV geometry (a 32-column chunk stays inside q or k and inside one head)
            // (no integer divisions per chunk: q / k / v by comparison, head index by shift for the usual d_k = 64)
            const int which = EPI == EPI_QKV ? (nc >= 2 * p.d_model ? 2 : (nc >= p.d_model ? 1 : 0)) : 0;
            const int cq = n0 - which * p.d_model;
            const int hq = EPI == EPI_QKV ? (p.d_k == 64 ? cq >> 6 : cq / p.d_k) : 0, dq = cq - hq * p.d_k;
            // in-place residual stream: the old values of the lane's four rows (xo) were requested by the caller before the
            // accumulator chunk was read from tensor memory (resid_load), so their latency overlaps that read and the transpose
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rr = rq + 8 * k;
                const int m = r0 + rr;
                if (m >= p.M) break;
                const float4 a0 = *reinterpret_cast<const float4*>(stage + rr * 32 + 4 * g0);
                const float4 a1 = *reinterpret_cast<const float4*>(stage + rr * 32 + 4 * g1);
                float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                if ((EPI == EPI_QKV || EPI == EPI_RELU_SPLIT) && p.ln_stats) {
                    const float ar = p.acc_scale * lst[k].y, nmr = -lst[k].x * lst[k].y;       // rstd (acc - mean csum) + bias
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], ar, fmaf(nmr, cs[e], bs[e]));
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], p.acc_scale, bs[e]);
                }
                switch (EPI) {
                    case EPI_STORE: {
                        float* o = p.out0 + (size_t)b * p.o_batch_stride + (size_t)m * p.ldo + n0;
                        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        break;
                    }
                    case EPI_RELU_SPLIT: {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                        split_store8(p.out_fmt, p.out0, p.out1, (size_t)m * p.ldo + n0, v);
                        break;
                    }
                    case EPI_GELU_SPLIT: {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
                        split_store8(p.out_fmt, p.out0, p.out1, (size_t)b * p.o_batch_stride + (size_t)m * p.ldo + n0, v);
                        break;
                    }
                    case EPI_GELU_POS: {
                        float* o = p.out0 + ((size_t)b * p.M + m) * p.ldo + n0;
                        const float* ps = p.out1 + (size_t)m * p.ldo + n0;
                        const float4 p0 = __ldg(reinterpret_cast<const float4*>(ps)), p1 = __ldg(reinterpret_cast<const float4*>(ps + 4));
                        *reinterpret_cast<float4*>(o) = make_float4(gelu_erf(v[0]) + p0.x, gelu_erf(v[1]) + p0.y, gelu_erf(v[2]) + p0.z, gelu_erf(v[3]) + p0.w);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(gelu_erf(v[4]) + p1.x, gelu_erf(v[5]) + p1.y, gelu_erf(v[6]) + p1.z, gelu_erf(v[7]) + p1.w);
                        break;
                    }
                    case EPI_RESID: {
                        float* o = p.out0 + (size_t)m * p.ldo + n0;
                        const float4 x0 = xo[k][0], x1 = xo[k][1];
                        const float xn[8] = {x0.x + p.alpha * v[0], x0.y + p.alpha * v[1], x0.z + p.alpha * v[2], x0.w + p.alpha * v[3],
                                             x1.x + p.alpha * v[4], x1.y + p.alpha * v[5], x1.z + p.alpha * v[6], x1.w + p.alpha * v[7]};
                        *reinterpret_cast<float4*>(o) = make_float4(xn[0], xn[1], xn[2], xn[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(xn[4], xn[5], xn[6], xn[7]);
                        if (p.ln_part) {
                            // LayerNorm source of the next GEMM: raw planes of the new row + its partial moments
                            if (p.ln_hi) split_store8(p.out_fmt, p.ln_hi, p.ln_lo, (size_t)m * p.ldo + n0, xn);
                            float s = 0.f, q = 0.f;
#pragma unroll
                            for (int e = 0; e < 8; ++e) { s += xn[e]; q = fmaf(xn[e], xn[e], q); }
                            ln_s[k] += s;
                            ln_q[k] += q;
                        }
                        break;
                    }
                    case EPI_QKV: {     // q, k: [seg][head][t][d_k]; qkv_row[k] = (seg n_heads T + t) d_k of the lane's k-th row (once per tile)
                        split_store8(p.qkv_fmt, which == 0 ? p.q_hi : which == 1 ? p.k_hi : p.vt_hi, which == 0 ? p.q_lo : which == 1 ? p.k_lo : p.vt_lo,
                                     qkv_row[k] + (size_t)(hq * p.T) * p.d_k + dq, v);
                        break;
                    }
                    case EPI_PV: {      // batch = (seg, head); rows are frames of that segment
                        const int seg = b / p.n_heads, h = b - seg * p.n_heads;
                        split_store8(p.out_fmt, p.out0, p.out1, ((size_t)seg * p.T + m) * p.ldo + (size_t)h * p.d_k + n0, v);
                        break;
                    }
                    default: break;
                }
            }
        }
        __syncwarp();
        return;
    }
    // lane == column (scalar fallback for shapes that do not allow 16-byte accesses)
#pragma unroll
    for (int j = 0; j < 32; ++j) stage[lane * kEpiPitch + j] = __uint_as_float(r[j]);
    __syncwarp();
    const int n = nc + lane;
    const int rows = min(32, p.M - r0);
    if (n < p.N) {
        const float bs = p.bias ? __ldg(p.bias + n) : 0.f;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = stage[i * kEpiPitch + lane] * p.acc_scale + bs;
        switch (EPI) {
            case EPI_STORE: {
                float* o = p.out0 + (size_t)b * p.o_batch_stride + (size_t)r0 * p.ldo + n;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows) o[(size_t)i * p.ldo] = v[i];
                break;
            }
            case EPI_GELU_SPLIT: {
                const size_t o = (size_t)b * p.o_batch_stride + (size_t)r0 * p.ldo + n;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows) split_store(p.out_fmt, p.out0, p.out1, o + (size_t)i * p.ldo, gelu_erf(v[i]));
                break;
            }
            case EPI_GELU_POS: {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows)
                        p.out0[((size_t)b * p.M + r0 + i) * p.ldo + n] = gelu_erf(v[i]) + __ldg(p.out1 + (size_t)(r0 + i) * p.ldo + n);
                break;
            }
            case EPI_RELU_SPLIT: {
                const size_t o = (size_t)r0 * p.ldo + n;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows) split_store(p.out_fmt, p.out0, p.out1, o + (size_t)i * p.ldo, fmaxf(v[i], 0.f));
                break;
            }
            case EPI_RESID: {
                // in-place residual stream: all loads of the chunk are issued before the first store
                float* o = p.out0 + (size_t)r0 * p.ldo + n;
                float old[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) old[i] = (i < rows) ? o[(size_t)i * p.ldo] : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows) o[(size_t)i * p.ldo] = old[i] + p.alpha * v[i];
                break;
            }
            case EPI_QKV: {     // q, k: [seg][head][t][d_k]
                const int which = nc / p.d_model, c = n - which * p.d_model;
                const int h = c / p.d_k, d = c - h * p.d_k;
                float* o_hi = which == 0 ? p.q_hi : which == 1 ? p.k_hi : p.vt_hi;
                float* o_lo = which == 0 ? p.q_lo : which == 1 ? p.k_lo : p.vt_lo;
                RowWalker w(r0, p.T);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < rows)
                        split_store(p.qkv_fmt, o_hi, o_lo, (((size_t)w.seg * p.n_heads + h) * p.T + w.t) * p.d_k + d, v[i]);
                    w.next();
                }
                break;
            }
            case EPI_PV: {      // batch = (seg, head); rows are frames of that segment
                const int seg = b / p.n_heads, h = b - seg * p.n_heads;
                const size_t o = ((size_t)seg * p.T + r0) * p.ldo + (size_t)h * p.d_k + n;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows) split_store(p.out_fmt, p.out0, p.out1, o + (size_t)i * p.ldo, v[i]);
                break;
            }
            default: break;
        }
    }
    __syncwarp();
}

// p.epi is uniform: one dispatch per chunk, everything inside is resolved at compile time
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int b, int r0, int nc, const uint32_t (&r)[32],
                                               float* stage, int lane, float (&ln_s)[4], float (&ln_q)[4],
                                               const size_t (&qkv_row)[4], const float4 (&xo)[4][2]) {
    switch (p.epi) {
        case EPI_STORE:      epilogue_chunk_t<EPI_STORE>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_RELU_SPLIT: epilogue_chunk_t<EPI_RELU_SPLIT>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_RESID:      epilogue_chunk_t<EPI_RESID>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_QKV:        epilogue_chunk_t<EPI_QKV>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_PV:         epilogue_chunk_t<EPI_PV>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_MASK:       epilogue_chunk_t<EPI_MASK>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_GELU_SPLIT: epilogue_chunk_t<EPI_GELU_SPLIT>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        case EPI_GELU_POS:   epilogue_chunk_t<EPI_GELU_POS>(p, b, r0, nc, r, stage, lane, ln_s, ln_q, qkv_row, xo); break;
        default: break;
    }
}

// EPI >= 0: the epilogue kind is a compile-time constant of the instantiation (the hot GEMMs of the mask network: smaller
// code -- the generic kernel is 28 k SASS instructions, `no_inst` was 8-9 % of its stall samples -- and registers allocated
// for one epilogue instead of the worst of eight); EPI < 0: p.epi is read at run time.
// old values of the in-place residual stream for the chunk at column nc (vector epilogue): the lane's 8 columns of rows
// r0 + (lane >> 2) + 8 k, all eight loads in flight at once
__device__ __forceinline__ void resid_load(const GemmParams& p, int r0, int nc, int lane, float4 (&xo)[4][2]) {
    const int n0 = nc + 8 * (lane & 3), rq = lane >> 2;
    if (n0 < p.N) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = r0 + rq + 8 * k;
            const float* o = p.out0 + (size_t)(m < p.M ? m : r0) * p.ldo + n0;
            xo[k][0] = *reinterpret_cast<const float4*>(o);
            xo[k][1] = *reinterpret_cast<const float4*>(o + 4);
        }
    }
}

template <int MODE, int TN, int RB, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const GemmParams p, const int tiles_m, const int tiles_n, const int total_tiles) {
    using Cfg = TcCfg<MODE, TN, RB>;
    constexpr int TBN = TN;
    constexpr int kBTileBytes = Cfg::kBTileBytes, kATileBytes = Cfg::kATileBytes, kRowBytes = Cfg::kRowBytes;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* gen_tiles = smem_raw + (tiles - raw);
    float* epi_stage = reinterpret_cast<float*>(gen_tiles + Cfg::kStages * Cfg::kStageBytes);
    const uint32_t bars = tiles + Cfg::kStages * Cfg::kStageBytes + kEpiBytes;
    // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * Cfg::kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * Cfg::kStages + 2 + a); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * Cfg::kStages + 4);
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(gen_tiles + Cfg::kStages * Cfg::kStageBytes + kEpiBytes + 8 * (2 * Cfg::kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int epi = EPI >= 0 ? EPI : p.epi;
    const int nkb = (p.K + Cfg::kBlockK - 1) / Cfg::kBlockK;      // a K tail is zero-filled by TMA

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
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
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_blk = tile % tiles_n, rest = tile / tiles_n;
                const int m_blk = rest % tiles_m, b = rest / tiles_m;
                const int m0 = m_blk * TBM, n0 = n_blk * TBN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (it / Cfg::kStages) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
                    const uint32_t st = tiles + s * Cfg::kStageBytes;
                    const int k0 = kb * Cfg::kBlockK;
                    const int bb = p.b_shared ? 0 : b;
                    tma_load_3d(st, &map_a_hi, k0, m0, b, full_bar(s));
                    if (Cfg::kSplit) {
                        tma_load_3d(st + kATileBytes, &map_a_lo, k0, m0, b, full_bar(s));
                        tma_load_3d(st + 2 * kATileBytes, &map_b_hi, k0, n0, bb, full_bar(s));
                        tma_load_3d(st + 2 * kATileBytes + kBTileBytes, &map_b_lo, k0, n0, bb, full_bar(s));
                    } else {
                        tma_load_3d(st + kATileBytes, &map_b_hi, k0, n0, bb, full_bar(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = MODE >= 16 ? make_idesc_f16(TBN, p.op_fmt != SPLIT_F16 && p.op_fmt != SPLIT_F16_1) : make_idesc_tf32(TBN);
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1);       // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TBN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (it / Cfg::kStages) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t st = tiles + s * Cfg::kStageBytes;
                    const uint32_t a_hi = st, a_lo = st + kATileBytes;
                    const uint32_t b_hi = st + (Cfg::kSplit ? 2 : 1) * kATileBytes, b_lo = b_hi + kBTileBytes;
#pragma unroll
                    for (int ks = 0; ks < kRowBytes / 32; ++ks) {             // UMMA_K = 8 tf32 = 16 halves = 32 bytes
                        const uint32_t koff = ks * 32;
                        const uint64_t da_hi = make_smem_desc_rb<RB>(a_hi + koff), db_hi = make_smem_desc_rb<RB>(b_hi + koff);
                        if (MODE == 16) {
                            const uint64_t da_lo = make_smem_desc_rb<RB>(a_lo + koff), db_lo = make_smem_desc_rb<RB>(b_lo + koff);
                            tcgen05_mma_f16(d_tmem, da_lo, db_hi, idesc, (kb | ks) != 0);     // small terms first
                            tcgen05_mma_f16(d_tmem, da_hi, db_lo, idesc, 1);
                            tcgen05_mma_f16(d_tmem, da_hi, db_hi, idesc, 1);
                        } else if (MODE == 116) {
                            tcgen05_mma_f16(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);
                        } else if (MODE == 3) {
                            const uint64_t da_lo = make_smem_desc_rb<RB>(a_lo + koff), db_lo = make_smem_desc_rb<RB>(b_lo + koff);
                            tcgen05_mma_tf32(d_tmem, da_lo, db_hi, idesc, (kb | ks) != 0);    // small terms first
                            tcgen05_mma_tf32(d_tmem, da_hi, db_lo, idesc, 1);
                            tcgen05_mma_tf32(d_tmem, da_hi, db_hi, idesc, 1);
                        } else {
                            tcgen05_mma_tf32(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);
                        }
                    }
                    tcgen05_commit(empty_bar(s));          // frees the stage once the MMAs above have read it
                }
                tcgen05_commit(tmem_full_bar(acc));        // accumulator complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> (shared-memory transpose) -> fused epilogue -> global.
        // Warps w and w + 4 share TMEM lane quarter w % 4 and take alternate 32-column chunks.
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                  // 0: chunks 0, 2, 4, ...   1: chunks 1, 3, 5, ...
        float* stage = epi_stage + (warp - 2) * 32 * kEpiPitch;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const int n_blk = tile % tiles_n, rest = tile / tiles_n;
            const int m_blk = rest % tiles_m, b = rest / tiles_m;
            const int n0 = n_blk * TBN;
            const int r0 = m_blk * TBM + q * 32;           // first row of this warp's quarter
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            const int n_end = min(p.N - n0, TBN);           // valid columns of this tile (> 0)
            const int c_first = half * 32;
            mbar_wait(tmem_full_bar(acc), aph);
            tcgen05_fence_after();
            const bool ln_src = epi == EPI_RESID && p.ln_part != nullptr;
            if (c_first >= n_end) {                         // nothing to read for this warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
                if (ln_src && r0 + lane < p.M) p.ln_part[(size_t)(r0 + lane) * p.ln_slots + 2 * n_blk + half] = make_float2(0.f, 0.f);
                continue;
            }
            float ln_s[4] = {0.f, 0.f, 0.f, 0.f}, ln_q[4] = {0.f, 0.f, 0.f, 0.f};
            // per-tile row work of the vector epilogues (a lane owns rows r0 + (lane >> 2) + 8 k of every chunk)
            size_t qkv_row[4] = {0, 0, 0, 0};
            if (epi == EPI_QKV) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int m = min(r0 + (lane >> 2) + 8 * k, p.M - 1);
                    const int seg = m / p.T, t = m - seg * p.T;
                    qkv_row[k] = ((size_t)seg * p.n_heads * p.T + t) * p.d_k;
                }
            }
#pragma unroll 1
            for (int c0 = c_first; c0 < n_end; c0 += 64) {
                uint32_t r[32];
                float4 xo[4][2];
                if (epi == EPI_RESID && p.vec8 && r0 < p.M) resid_load(p, r0, n0 + c0, lane, xo);
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * TBN + (uint32_t)c0, r);
                if (c0 + 64 >= n_end) {
                    // last read of this accumulator by this warp: hand it back to the MMA warp before the stores
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
                }
                if (r0 >= p.M) continue;                    // warp-uniform: quarter entirely out of range
                if constexpr (EPI >= 0) epilogue_chunk_t<EPI>(p, b, r0, n0 + c0, r, stage, lane, ln_s, ln_q, qkv_row, xo);
                else epilogue_chunk(p, b, r0, n0 + c0, r, stage, lane, ln_s, ln_q, qkv_row, xo);
            }
            if (ln_src && r0 < p.M) {
                // row moments of this warp's columns of the tile: the four lanes that share a row (cg = lane & 3) add up
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float s = ln_s[k], q2 = ln_q[k];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);  q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);  q2 += __shfl_xor_sync(0xffffffffu, q2, 2);
                    const int m = r0 + (lane >> 2) + 8 * k;
                    if ((lane & 3) == 0 && m < p.M) p.ln_part[(size_t)m * p.ln_slots + 2 * n_blk + half] = make_float2(s, q2);
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

// ------------------------------------------------------------------------------------------- CTA pairs (cta_group::2)
// The 2xBF16 GEMMs of the mask network on pairs of CTAs (clusters of two, one per SM): a pair owns a 256 x 256 tile, each CTA
// its 128 rows; `tcgen05.mma.cta_group::2` (M = 256) is issued by the leader and reads A from each CTA's own shared memory and
// the two 128-column halves of B from both -- every CTA stages only HALF of B.  Per K block of 64 a CTA moves 64 KB (A hi/lo 32,
// half B hi/lo 32) instead of 96 KB from L2 for the same MMAs, and three stages fit where two did: the single-CTA kernel's main
// loop is bound by that traffic / its latency (58-80 % tensor-active whatever the epilogue does,
// profiles/r02_ncu_full_gemm16_block_per_epilogue_kernels_summary.csv).
//   full[s]       leader's barrier: the leader's producer arms it with the bytes of BOTH CTAs, both producers' TMA loads credit it
//   empty[s]      per CTA: the leader's commit is multicast to both
//   tmem_full[a]  per CTA (multicast commit); tmem_empty[a]: leader's barrier, 8 epilogue warps of each CTA arrive on it
// SPLIT: head + remainder planes, three MMAs per product (2xBF16 / 2xF16: 64 KB stages, 3 of them); !SPLIT: plain bf16 operands,
// one MMA (the Whisper encoder's GEMMs: 32 KB stages, 6 of them).
constexpr int k2ATile = TBM * 128;                  // 16 KB: 128 rows x 64 16-bit elements
constexpr int k2BHalf = 128 * 128;                  // 16 KB: 128 of the tile's 256 columns x 64 elements
template <bool SPLIT> struct Tc2Cfg {
    static constexpr int kStageBytes = (SPLIT ? 2 : 1) * (k2ATile + k2BHalf);
    static constexpr int kStages = SPLIT ? 3 : 6;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 256 + 1024;
};

template <int EPI, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const GemmParams p, const int tiles_m, const int tiles_n, const int total_pairs) {
    constexpr int TBN = 256;
    constexpr int k2Stages = Tc2Cfg<SPLIT>::kStages, k2StageBytes = Tc2Cfg<SPLIT>::kStageBytes;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    unsigned char* gen_tiles = smem_raw + (tiles - raw);
    float* epi_stage = reinterpret_cast<float*>(gen_tiles + k2Stages * k2StageBytes);
    const uint32_t bars = tiles + k2Stages * k2StageBytes + kEpiBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (k2Stages + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * k2Stages + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * k2Stages + 2 + a); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * k2Stages + 4);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen_tiles + k2Stages * k2StageBytes + kEpiBytes + 8 * (2 * k2Stages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nkb = (p.K + 63) / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < k2Stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer's barriers exist before anything is signalled across
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): own A rows, own half of B
            uint32_t it = 0;
            for (int item = pair; item < total_pairs; item += n_pairs) {
                const int n_blk = item % tiles_n, mp = item / tiles_n;
                const int m0 = (2 * mp + (int)rank) * TBM, n0 = n_blk * TBN + (int)rank * 128;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % k2Stages;
                    const uint32_t ph = (it / k2Stages) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2u * k2StageBytes);
                    const uint32_t st = tiles + s * k2StageBytes;
                    const int k0 = kb * 64;
                    tma_load_3d_2sm(st, &map_a_hi, k0, m0, 0, full_bar(s));
                    if (SPLIT) {
                        tma_load_3d_2sm(st + k2ATile, &map_a_lo, k0, m0, 0, full_bar(s));
                        tma_load_3d_2sm(st + 2 * k2ATile, &map_b_hi, k0, n0, 0, full_bar(s));
                        tma_load_3d_2sm(st + 2 * k2ATile + k2BHalf, &map_b_lo, k0, n0, 0, full_bar(s));
                    } else {
                        tma_load_3d_2sm(st + k2ATile, &map_b_hi, k0, n0, 0, full_bar(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ===== MMA issuer (leader only)
            const uint32_t idesc = make_idesc_f16_2sm(TBN, p.op_fmt != SPLIT_F16 && p.op_fmt != SPLIT_F16_1);       // bf16 operands unless the scaled-fp16 pairs
            uint32_t it = 0, tl = 0;
            for (int item = pair; item < total_pairs; item += n_pairs, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1);       // both CTAs' epilogues have drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TBN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % k2Stages;
                    const uint32_t ph = (it / k2Stages) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t st = tiles + s * k2StageBytes;
                    const uint32_t a_hi = st, a_lo = st + k2ATile, b_hi = st + (SPLIT ? 2 : 1) * k2ATile, b_lo = b_hi + k2BHalf;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t koff = ks * 32;
                        const uint64_t da_hi = make_smem_desc(a_hi + koff), db_hi = make_smem_desc(b_hi + koff);
                        if (SPLIT) {
                            const uint64_t da_lo = make_smem_desc(a_lo + koff), db_lo = make_smem_desc(b_lo + koff);
                            tcgen05_mma_f16_2sm(d_tmem, da_lo, db_hi, idesc, (kb | ks) != 0);     // small terms first
                            tcgen05_mma_f16_2sm(d_tmem, da_hi, db_lo, idesc, 1);
                            tcgen05_mma_f16_2sm(d_tmem, da_hi, db_hi, idesc, 1);
                        } else {
                            tcgen05_mma_f16_2sm(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);
                        }
                    }
                    tcgen05_commit_2sm(empty_bar(s));
                }
                tcgen05_commit_2sm(tmem_full_bar(acc));
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): as in gemm_tc_kernel, on this CTA's 128 rows of the pair's tile
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stage = epi_stage + (warp - 2) * 32 * kEpiPitch;
        uint32_t tl = 0;
        for (int item = pair; item < total_pairs; item += n_pairs, ++tl) {
            const int n_blk = item % tiles_n, mp = item / tiles_n;
            const int n0 = n_blk * TBN;
            const int r0 = (2 * mp + (int)rank) * TBM + q * 32;
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            const int n_end = min(p.N - n0, TBN);
            const int c_first = half * 32;
            if (EPI == EPI_RESID && r0 < p.M) {
                // in-place residual: ask L2 for the old values of this warp's chunks while the pair's MMAs for the tile still run
                // (the loads in the chunk loop then see L2 latency; with the main loop at 97 % tensor-active the K = 512
                // out-projection was left at 75 % by this round trip)
                for (int c0 = c_first; c0 < n_end; c0 += 64) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int m = r0 + (lane >> 2) + 8 * k;
                        if (m < p.M && n0 + c0 + 8 * (lane & 3) < p.N)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.out0 + (size_t)m * p.ldo + n0 + c0 + 8 * (lane & 3)));
                    }
                }
            }
            mbar_wait(tmem_full_bar(acc), aph);
            tcgen05_fence_after();
            const bool ln_src = EPI == EPI_RESID && p.ln_part != nullptr;
            if (c_first >= n_end) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(tmem_empty_bar(acc));
                if (ln_src && r0 + lane < p.M) p.ln_part[(size_t)(r0 + lane) * p.ln_slots + 2 * n_blk + half] = make_float2(0.f, 0.f);
                continue;
            }
            float ln_s[4] = {0.f, 0.f, 0.f, 0.f}, ln_q[4] = {0.f, 0.f, 0.f, 0.f};
            size_t qkv_row[4] = {0, 0, 0, 0};
            if (EPI == EPI_QKV) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int m = min(r0 + (lane >> 2) + 8 * k, p.M - 1);
                    const int seg = m / p.T, t = m - seg * p.T;
                    qkv_row[k] = ((size_t)seg * p.n_heads * p.T + t) * p.d_k;
                }
            }
#pragma unroll 1
            for (int c0 = c_first; c0 < n_end; c0 += 64) {
                uint32_t r[32];
                float4 xo[4][2];
                if (EPI == EPI_RESID && r0 < p.M) resid_load(p, r0, n0 + c0, lane, xo);
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * TBN + (uint32_t)c0, r);
                if (c0 + 64 >= n_end) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tmem_empty_bar(acc));
                }
                if (r0 >= p.M) continue;
                epilogue_chunk_t<EPI>(p, 0, r0, n0 + c0, r, stage, lane, ln_s, ln_q, qkv_row, xo);
            }
            if (ln_src && r0 < p.M) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float s = ln_s[k], q2 = ln_q[k];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);  q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);  q2 += __shfl_xor_sync(0xffffffffu, q2, 2);
                    const int m = r0 + (lane >> 2) + 8 * k;
                    if ((lane & 3) == 0 && m < p.M) p.ln_part[(size_t)m * p.ln_slots + 2 * n_blk + half] = make_float2(s, q2);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // neither CTA leaves (or frees tensor memory) while the pair's MMAs / arrivals are in flight
    tcgen05_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------- host side
// NSF_GEMM_2SM=0 keeps the single-CTA kernels (A/B measurements)
static bool use_cta_pairs() {
    static const bool on = [] { const char* e = getenv("NSF_GEMM_2SM"); return !(e && e[0] == '0'); }();
    return on;
}

template <int MODE, int TN, int RB = 128>
static int launch_t(const GemmParams& p, cudaStream_t stream) {
    using Cfg = TcCfg<MODE, TN, RB>;
    constexpr int TBN = TN;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    auto make = [&](CUtensorMap* m, const float* base, int64_t rows, int64_t ld, int64_t bstride, int box_rows, int64_t batch) {
        return MODE >= 16 ? make_tmap_kmajor16(m, base, rows, p.K, ld, batch, bstride, box_rows, RB)
                          : make_tmap_kmajor(m, base, rows, p.K, ld, batch, bstride, box_rows);
    };
    const int64_t b_batch = p.b_shared ? 1 : p.batch;
    if ((rc = make(&ma_hi, p.A_hi, p.M, p.lda, p.a_batch_stride, TBM, p.batch))) return rc;
    if ((rc = make(&mb_hi, p.B_hi, p.N, p.ldb, p.b_shared ? 0 : p.b_batch_stride, TBN, b_batch))) return rc;
    if (Cfg::kSplit) {
        if (!p.A_lo || !p.B_lo) { set_error("gemm_tc: split engines need head and remainder operands"); return NSF_ERR_INVALID_ARG; }
        if ((rc = make(&ma_lo, p.A_lo, p.M, p.lda, p.a_batch_stride, TBM, p.batch))) return rc;
        if ((rc = make(&mb_lo, p.B_lo, p.N, p.ldb, p.b_shared ? 0 : p.b_batch_stride, TBN, b_batch))) return rc;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    GemmParams pv = p;
    {
        auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
        bool ok = p.N % 8 == 0 && (!p.bias || al16(p.bias));
        switch (p.epi) {
            case EPI_STORE: case EPI_RESID:
                ok = ok && p.ldo % 4 == 0 && p.o_batch_stride % 4 == 0 && al16(p.out0); break;
            case EPI_RELU_SPLIT: case EPI_PV:
                ok = ok && p.ldo % 8 == 0 && al16(p.out0) && (p.out_fmt == SPLIT_BF16_1 || p.out_fmt == SPLIT_F16_1 || al16(p.out1)) && p.d_k % 8 == 0; break;
            case EPI_GELU_SPLIT:
                ok = ok && p.ldo % 8 == 0 && p.o_batch_stride % 8 == 0 && al16(p.out0) && (p.out_fmt == SPLIT_BF16_1 || p.out_fmt == SPLIT_F16_1 || al16(p.out1)); break;
            case EPI_GELU_POS:
                ok = ok && p.ldo % 4 == 0 && al16(p.out0) && al16(p.out1); break;
            case EPI_QKV:
                ok = ok && p.d_k % 8 == 0 && p.d_model % 32 == 0 && al16(p.q_hi) && al16(p.q_lo) && al16(p.k_hi) && al16(p.k_lo) &&
                     (!p.v_rowmajor || (al16(p.vt_hi) && al16(p.vt_lo))); break;
            default: ok = false; break;     // EPI_MASK writes along rows
        }
        pv.vec8 = ok ? 1 : 0;
        if ((p.ln_part || p.ln_stats) && !ok) { set_error("gemm_tc: the folded-LayerNorm epilogues need the vector path (N, pitches, alignment)"); return NSF_ERR_INVALID_ARG; }
        if (p.ln_part && !(p.epi == EPI_RESID && p.ldo % 8 == 0 && (!p.ln_hi || (al16(p.ln_hi) && al16(p.ln_lo))) && p.ln_slots >= 2 * ceil_div(p.N, TBN))) {
            set_error("gemm_tc: LayerNorm source outputs need EPI_RESID, ldo %% 8 == 0, aligned planes and 2 slots per column tile");
            return NSF_ERR_INVALID_ARG;
        }
        if (p.ln_stats && !((p.epi == EPI_QKV || p.epi == EPI_RELU_SPLIT) && p.ln_csum && al16(p.ln_csum))) {
            set_error("gemm_tc: folded LayerNorm is implemented for EPI_QKV / EPI_RELU_SPLIT with an aligned column-sum vector");
            return NSF_ERR_INVALID_ARG;
        }
    }

    const int tiles_m = ceil_div(p.M, TBM), tiles_n = ceil_div(p.N, TBN);
    const int64_t total = (int64_t)tiles_m * tiles_n * p.batch;
    if (total > 0x7fffffff) { set_error("gemm_tc: too many tiles"); return NSF_ERR_INVALID_ARG; }
    const int grid = (int)(total < sm_count() ? total : sm_count());
#define NSF_GEMM_LAUNCH(E)                                                                                                      \
    do {                                                                                                                        \
        NSF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE, TN, RB, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); \
        gemm_tc_kernel<MODE, TN, RB, E><<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, pv, tiles_m, tiles_n, (int)total); \
    } while (0)
    if constexpr ((MODE == 16 || MODE == 116) && TN == 256 && RB == 128) {
    if ((pv.vec8 || (MODE == 16 && p.epi == EPI_MASK)) && p.batch == 1 && sm_count() >= 2 && use_cta_pairs() &&
        (p.epi == EPI_RESID || p.epi == EPI_QKV || p.epi == EPI_STORE || p.epi == EPI_RELU_SPLIT || (MODE == 116 && p.epi == EPI_GELU_SPLIT) ||
         (MODE == 16 && p.epi == EPI_MASK))) {
        // CTA pairs: B is staged in halves of 128 columns
        CUtensorMap mb2_hi, mb2_lo;
        if ((rc = make_tmap_kmajor16(&mb2_hi, p.B_hi, p.N, p.K, p.ldb, 1, 0, 128))) return rc;
        if (Cfg::kSplit) { if ((rc = make_tmap_kmajor16(&mb2_lo, p.B_lo, p.N, p.K, p.ldb, 1, 0, 128))) return rc; }
        else mb2_lo = mb2_hi;
        const int pairs_m = (tiles_m + 1) / 2;
        const int64_t total_pairs = (int64_t)pairs_m * tiles_n;
        int grid2 = (int)(2 * total_pairs < sm_count() ? 2 * total_pairs : (sm_count() & ~1));
#define NSF_GEMM2_LAUNCH(E)                                                                                                     \
        do {                                                                                                                    \
            constexpr bool kSp = MODE == 16;                                                                                    \
            NSF_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<E, kSp>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg<kSp>::kSmemBytes)); \
            gemm_tc2_kernel<E, kSp><<<grid2, kTcThreads, Tc2Cfg<kSp>::kSmemBytes, stream>>>(ma_hi, ma_lo, mb2_hi, mb2_lo, pv, tiles_m, tiles_n, (int)total_pairs); \
        } while (0)
        switch (p.epi) {
            case EPI_RESID:      NSF_GEMM2_LAUNCH(EPI_RESID); break;
            case EPI_QKV:        NSF_GEMM2_LAUNCH(EPI_QKV); break;
            case EPI_STORE:      NSF_GEMM2_LAUNCH(EPI_STORE); break;
            case EPI_RELU_SPLIT: NSF_GEMM2_LAUNCH(EPI_RELU_SPLIT); break;
            default:             if constexpr (MODE == 116) { NSF_GEMM2_LAUNCH(EPI_GELU_SPLIT); } else { NSF_GEMM2_LAUNCH(EPI_MASK); } break;   // the mask head writes along rows
        }
#undef NSF_GEMM2_LAUNCH
        return check_launch("gemm_tc2_kernel");
    }
    }
    if (MODE == 16 && TN == 256 && RB == 128 && pv.vec8) {
        // the mask network's GEMMs: one instantiation per epilogue kind
        switch (p.epi) {
            case EPI_RELU_SPLIT: NSF_GEMM_LAUNCH(EPI_RELU_SPLIT); break;
            case EPI_RESID:      NSF_GEMM_LAUNCH(EPI_RESID); break;
            case EPI_QKV:        NSF_GEMM_LAUNCH(EPI_QKV); break;
            case EPI_STORE:      NSF_GEMM_LAUNCH(EPI_STORE); break;
            default:             NSF_GEMM_LAUNCH(-1); break;
        }
    } else {
        NSF_GEMM_LAUNCH(-1);
    }
#undef NSF_GEMM_LAUNCH
    return check_launch("gemm_tc_kernel");
}

int gemm_tc_launch(const GemmParams& p, int mode, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0 || p.K <= 0) { set_error("gemm_tc: empty problem"); return NSF_ERR_INVALID_ARG; }
    if (mode == 16) {
        if (p.op_fmt != SPLIT_BF16 && p.op_fmt != SPLIT_F16) { set_error("gemm_tc: 16-bit engine needs SPLIT_BF16 / SPLIT_F16 operands"); return NSF_ERR_INVALID_ARG; }
        if (p.K % 8 != 0) { set_error("gemm_tc: K=%d must be a multiple of 8", p.K); return NSF_ERR_INVALID_ARG; }
        // NSF_GEMM_RB=64 selects four 48 KB stages of 64-byte rows (SWIZZLE_64B) instead of two 96 KB stages of 128-byte rows.
        // Measured (profiles/r02_bench_m*.json): parity identical, GEMM time 72.3 vs 70.1 ms per step -- the step runs against the
        // 1 kW power cap (SM clock ~1.6 GHz), where the extra barrier / descriptor / commit work per byte costs more than the finer
        // prefetch grain hides; kept as a measurement switch.
        static const bool rb64 = [] { const char* e = getenv("NSF_GEMM_RB"); return e && atoi(e) == 64; }();
        return rb64 ? launch_t<16, 256, 64>(p, stream) : launch_t<16, 256, 128>(p, stream);
    }
    if (mode == 116) {
        if (p.op_fmt != SPLIT_BF16_1 && p.op_fmt != SPLIT_F16_1) { set_error("gemm_tc: the single-pass engine needs SPLIT_BF16_1 / SPLIT_F16_1 operands"); return NSF_ERR_INVALID_ARG; }
        if (p.K % 8 != 0) { set_error("gemm_tc: K=%d must be a multiple of 8", p.K); return NSF_ERR_INVALID_ARG; }
        // small M (the Whisper decode step: M = sequences in flight): 128 x 32 tiles spread the weight stream over N / 32 CTAs
        // instead of N / 256, with an 8-deep TMA ring per CTA
        if (p.M <= 256 && p.batch == 1) return launch_t<116, 32>(p, stream);
        return launch_t<116, 256>(p, stream);
    }
    if (p.op_fmt != SPLIT_TF32) { set_error("gemm_tc: tf32 engines need SPLIT_TF32 operands"); return NSF_ERR_INVALID_ARG; }
    if (p.K % 32 != 0) { set_error("gemm_tc: K=%d must be a multiple of 32", p.K); return NSF_ERR_INVALID_ARG; }
    return mode == 3 ? launch_t<3, 256>(p, stream) : launch_t<1, 256>(p, stream);
}

}  // namespace nsf
