// Conformer mask network forward (eval mode), batched over segments.
//
// Reference: css/css_with_conformer/nnet/conformer.py
//   ConformerCSS.forward :287-310, ConformerEncoder :189-239 (embed Linear->LN->ReLU :205-210, relative
//   positions :229-233 / RelativePositionalEncoding :12-29), EncoderLayer :172-186
//   (x += .5 FF; x += MHSA; x += Conv; x += .5 FF; LN), MultiHeadedAttention :57-92,
//   ConvModule :113-127 (scalar 1x1 "pointwise" convs, GLU, depthwise k=33, BatchNorm eval, ReLU),
//   FeedForward :146-150.
// All matrix products go through gemm_launch (default engine 2xBF16: tcgen05 kind::f16 on bf16 head + remainder planes, CTA pairs);
// everything else is fused into a handful of streaming kernels here.  Activations that feed a GEMM are written "split" in the
// engine's format (common.cuh SplitFmt); with the 2xBF16 engine two LayerNorms per block are folded into the GEMMs (ln_fold_rule).
#include "gemm_common.cuh"
#include <new>

namespace nsf {

constexpr int kPerLayer = 34;
constexpr int kGlobalOffsets = 12;
// pe_k is packed twice: TF32 pairs for attention.cu / the unfused path, bf16 pairs for attention16.cu (which one is used
// depends on the engine AND on the segment length, which the handle only learns at forward time)
enum GlobalOff { G_EMB_W_HI = 0, G_EMB_W_LO, G_EMB_B, G_EMB_LN_G, G_EMB_LN_B, G_PE_HI, G_PE_LO, G_HEAD_W_HI, G_HEAD_W_LO, G_HEAD_B,
                 G_PE16_HI, G_PE16_LO };
enum LayerOff {
    L_FFI_LN_G = 0, L_FFI_LN_B, L_FFI_W1_HI, L_FFI_W1_LO, L_FFI_B1, L_FFI_W2_HI, L_FFI_W2_LO, L_FFI_B2,
    L_ATT_LN_G, L_ATT_LN_B, L_WQKV_HI, L_WQKV_LO, L_BQKV, L_WO_HI, L_WO_LO, L_BO,
    L_CONV_LN_G, L_CONV_LN_B, L_CONV_SCALARS, L_DW_W, L_BN_SCALE, L_BN_SHIFT,
    L_FFO_LN_G, L_FFO_LN_B, L_FFO_W1_HI, L_FFO_W1_LO, L_FFO_B1, L_FFO_W2_HI, L_FFO_W2_LO, L_FFO_B2,
    L_OUT_LN_G, L_OUT_LN_B,
    L_QKV_CSUM, L_FFO_CSUM          // column sums of the gamma-scaled Wqkv / feed_forward_out W1 (folded LayerNorms; zeros otherwise)
};

}  // namespace nsf

struct nsf_conformer {
    nsf_conformer_dims dims;
    const float* blob;
    int64_t blob_floats;
    int n_offsets;
    int64_t* offsets;
    bool ln_fold;               // the blob was packed for folded LayerNorms (nsf_conformer_ln_fold)
    const float* g(int i) const { return blob + offsets[i]; }
    const float* l(int layer, int i) const { return blob + offsets[nsf::kGlobalOffsets + layer * nsf::kPerLayer + i]; }
};

namespace nsf {

// ------------------------------------------------------------------------------------------- LayerNorm family
// One warp per row, the row held in registers as float4s (d = 128 * NV, or the scalar tail kernel for other
// multiples of 32); all loads of a row are in flight before the first reduction.  HBM-bound streaming kernels.
constexpr int kLnMaxPerLane = 32;   // scalar kernels: d_model <= 1024

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int NV>
__device__ __forceinline__ void ln_stats(const float4 (&v)[NV], float inv_d, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
    }
    rstd = 1.f / sqrtf(warp_sum(q) * inv_d + 1e-5f);
}
__device__ __forceinline__ float4 ln_apply(float4 v, float mean, float rstd, float4 g, float4 b) {
    return make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y, (v.z - mean) * rstd * g.z + b.z,
                       (v.w - mean) * rstd * g.w + b.w);
}

// y1 = LN(x; g1, b1) [relu]; optional store; y2 = LN(y1; g2, b2) if g2 else y1; optional split store.
template <int NV>
__global__ void __launch_bounds__(256)
ln_vec_kernel(const float* __restrict__ x, int M, const float* __restrict__ g1, const float* __restrict__ b1, int relu1,
              float* __restrict__ out_x, const float* __restrict__ g2, const float* __restrict__ b2,
              float* __restrict__ out_hi, float* __restrict__ out_lo, int fmt) {
    constexpr int d = 128 * NV;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float inv_d = 1.f / (float)d;
    const size_t ro = (size_t)row * d;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ld4(x + ro + (i * 32 + lane) * 4);
    float mean, rstd;
    ln_stats<NV>(v, inv_d, mean, rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 y = ln_apply(v[i], mean, rstd, ldg4(g1 + c), ldg4(b1 + c));
        if (relu1) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
        v[i] = y;
        if (out_x) st4(out_x + ro + c, y);
    }
    if (g2) {
        ln_stats<NV>(v, inv_d, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            v[i] = ln_apply(v[i], mean, rstd, ldg4(g2 + c), ldg4(b2 + c));
        }
    }
    if (out_hi) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const size_t o = ro + (i * 32 + lane) * 4;
            if (out_lo) {
                split_store4(fmt, out_hi, out_lo, o, v[i]);
            } else {
                st4(out_hi + o, v[i]);
            }
        }
    }
}

// scalar variant for d not a multiple of 128
__global__ void __launch_bounds__(256)
ln_kernel(const float* __restrict__ x, int M, int d, const float* __restrict__ g1, const float* __restrict__ b1, int relu1,
          float* __restrict__ out_x, const float* __restrict__ g2, const float* __restrict__ b2,
          float* __restrict__ out_hi, float* __restrict__ out_lo, int fmt) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int n_per = d >> 5;
    const float inv_d = 1.f / (float)d;
    float v[kLnMaxPerLane];
    const float* xr = x + (size_t)row * d;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) { v[i] = xr[i * 32 + lane]; s += v[i]; }
    float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) { const float c = v[i] - mean; q += c * c; }
    float rstd = 1.f / sqrtf(warp_sum(q) * inv_d + 1e-5f);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) {
            const int c = i * 32 + lane;
            float y = (v[i] - mean) * rstd * __ldg(g1 + c) + __ldg(b1 + c);
            if (relu1) y = fmaxf(y, 0.f);
            v[i] = y;
            s += y;
            if (out_x) out_x[(size_t)row * d + c] = y;
        }
    if (g2) {
        mean = warp_sum(s) * inv_d;
        q = 0.f;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i)
            if (i < n_per) { const float c = v[i] - mean; q += c * c; }
        rstd = 1.f / sqrtf(warp_sum(q) * inv_d + 1e-5f);
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i)
            if (i < n_per) {
                const int c = i * 32 + lane;
                v[i] = (v[i] - mean) * rstd * __ldg(g2 + c) + __ldg(b2 + c);
            }
    }
    if (out_hi) {
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i)
            if (i < n_per) {
                const size_t o = (size_t)row * d + i * 32 + lane;
                if (out_lo) {
                    split_store(fmt, out_hi, out_lo, o, v[i]);
                } else {
                    out_hi[o] = v[i];
                }
            }
    }
}

int ln_launch(const float* x, int M, int d, const float* g1, const float* b1, int relu1, float* out_x, const float* g2,
              const float* b2, float* out_hi, float* out_lo, int fmt, cudaStream_t s) {
    const int grid = ceil_div(M, 8);
#define NSF_LN_CASE(NV) case NV: ln_vec_kernel<NV><<<grid, 256, 0, s>>>(x, M, g1, b1, relu1, out_x, g2, b2, out_hi, out_lo, fmt); break;
    switch ((d % 128 == 0) ? d / 128 : 0) {
        NSF_LN_CASE(1) NSF_LN_CASE(2) NSF_LN_CASE(3) NSF_LN_CASE(4) NSF_LN_CASE(5) NSF_LN_CASE(6) NSF_LN_CASE(8) NSF_LN_CASE(10)
        default:
            if (d > 32 * kLnMaxPerLane || d % 32 != 0) { set_error("ln_launch: d=%d unsupported", d); return NSF_ERR_UNSUPPORTED; }
            ln_kernel<<<grid, 256, 0, s>>>(x, M, d, g1, b1, relu1, out_x, g2, b2, out_hi, out_lo, fmt);
            break;
    }
#undef NSF_LN_CASE
    return check_launch("ln_kernel");
}

// conv module front: h = LN(x); u = (w1a h + b1a) * sigmoid(w1g h + b1g)      conformer.py:115-117
__device__ __forceinline__ float glu1(float h, float w1a, float b1a, float w1g, float b1g) {
    const float a = w1a * h + b1a, gt = w1g * h + b1g;
    return a * (1.f / (1.f + expf(-gt)));
}
// the same with MUFU-based exp and reciprocal (2 ulp): the fused conv kernel evaluates 57 M sigmoids per block
__device__ __forceinline__ float glu1_fast(float h, float w1a, float b1a, float w1g, float b1g) {
    const float a = w1a * h + b1a, gt = w1g * h + b1g;
    return a * __fdividef(1.f, 1.f + __expf(-gt));
}
template <int NV>
__global__ void __launch_bounds__(256)
ln_glu_vec_kernel(const float* __restrict__ x, int M, const float* __restrict__ g, const float* __restrict__ b,
                  const float* __restrict__ scalars, float* __restrict__ u) {
    constexpr int d = 128 * NV;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const size_t ro = (size_t)row * d;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ld4(x + ro + (i * 32 + lane) * 4);
    float mean, rstd;
    ln_stats<NV>(v, 1.f / (float)d, mean, rstd);
    const float w1a = __ldg(scalars + 0), b1a = __ldg(scalars + 1), w1g = __ldg(scalars + 2), b1g = __ldg(scalars + 3);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 h = ln_apply(v[i], mean, rstd, ldg4(g + c), ldg4(b + c));
        st4(u + ro + c, make_float4(glu1(h.x, w1a, b1a, w1g, b1g), glu1(h.y, w1a, b1a, w1g, b1g), glu1(h.z, w1a, b1a, w1g, b1g),
                                    glu1(h.w, w1a, b1a, w1g, b1g)));
    }
}
__global__ void __launch_bounds__(256)
ln_glu_kernel(const float* __restrict__ x, int M, int d, const float* __restrict__ g, const float* __restrict__ b,
              const float* __restrict__ scalars, float* __restrict__ u) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int n_per = d >> 5;
    const float inv_d = 1.f / (float)d;
    float v[kLnMaxPerLane];
    const float* xr = x + (size_t)row * d;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) { v[i] = xr[i * 32 + lane]; s += v[i]; }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) { const float c = v[i] - mean; q += c * c; }
    const float rstd = 1.f / sqrtf(warp_sum(q) * inv_d + 1e-5f);
    const float w1a = __ldg(scalars + 0), b1a = __ldg(scalars + 1), w1g = __ldg(scalars + 2), b1g = __ldg(scalars + 3);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
        if (i < n_per) {
            const int c = i * 32 + lane;
            u[(size_t)row * d + c] = glu1((v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c), w1a, b1a, w1g, b1g);
        }
}
static int ln_glu_launch(const float* x, int M, int d, const float* g, const float* b, const float* scalars, float* u, cudaStream_t s) {
    const int grid = ceil_div(M, 8);
    switch ((d % 128 == 0) ? d / 128 : 0) {
        case 1: ln_glu_vec_kernel<1><<<grid, 256, 0, s>>>(x, M, g, b, scalars, u); break;
        case 2: ln_glu_vec_kernel<2><<<grid, 256, 0, s>>>(x, M, g, b, scalars, u); break;
        case 4: ln_glu_vec_kernel<4><<<grid, 256, 0, s>>>(x, M, g, b, scalars, u); break;
        case 8: ln_glu_vec_kernel<8><<<grid, 256, 0, s>>>(x, M, g, b, scalars, u); break;
        default: ln_glu_kernel<<<grid, 256, 0, s>>>(x, M, d, g, b, scalars, u); break;
    }
    return check_launch("ln_glu_kernel");
}

// conv module back: depthwise conv over time (zero padded inside the segment), BN (folded), ReLU, scalar affine,
// residual add.                                                                     conformer.py:118-126, 181
// One CTA stages a whole segment of 64 channels in shared memory (coalesced float4 rows); each thread then slides a
// 40-value register window down its channel and produces 8 outputs per window (33 taps from registers).
constexpr int kDwCh = 64, kDwOut = 8, kDwMaxK = 33;
__global__ void __launch_bounds__(256, 2)
dwconv_kernel(const float* __restrict__ u, float* __restrict__ x, int T, int d, int ks, const float* __restrict__ dw_w,
              const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, const float* __restrict__ scalars) {
    extern __shared__ __align__(16) float tile[];     // [T + ks - 1][kDwCh], row r <-> frame r - pad
    const int seg = blockIdx.y, c0 = blockIdx.x * kDwCh;
    const int pad = (ks - 1) / 2;
    const int rows = T + ks - 1;
    for (int idx = threadIdx.x; idx < rows * (kDwCh / 4); idx += blockDim.x) {
        const int r = idx / (kDwCh / 4), c4 = idx - r * (kDwCh / 4);
        const int t = r - pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T && c0 + c4 * 4 < d) v = ldg4(u + ((size_t)seg * T + t) * d + c0 + c4 * 4);
        st4(tile + r * kDwCh + c4 * 4, v);
    }
    __syncthreads();
    const int c = threadIdx.x & (kDwCh - 1), grp = threadIdx.x / kDwCh;       // 4 frame groups
    if (c0 + c >= d) return;
    float w[kDwMaxK];
#pragma unroll
    for (int j = 0; j < kDwMaxK; ++j) w[j] = j < ks ? __ldg(dw_w + (size_t)(c0 + c) * ks + j) : 0.f;
    const float sc = __ldg(bn_scale + c0 + c), sh = __ldg(bn_shift + c0 + c);
    const float w2 = __ldg(scalars + 4), b2 = __ldg(scalars + 5);
    const int per = ceil_div(ceil_div(T, 4), kDwOut) * kDwOut;
    const int t_end = min(T, (grp + 1) * per);
    for (int t0 = grp * per; t0 < t_end; t0 += kDwOut) {
        float win[kDwOut + kDwMaxK - 1];
#pragma unroll
        for (int i = 0; i < kDwOut + kDwMaxK - 1; ++i) win[i] = (t0 + i < rows) ? tile[(t0 + i) * kDwCh + c] : 0.f;
        float xo[kDwOut];
#pragma unroll
        for (int o = 0; o < kDwOut; ++o) xo[o] = (t0 + o < t_end) ? x[((size_t)seg * T + t0 + o) * d + c0 + c] : 0.f;
#pragma unroll
        for (int o = 0; o < kDwOut; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kDwMaxK; ++j) acc = fmaf(w[j], win[o + j], acc);
            const float y = fmaxf(acc * sc + sh, 0.f);
            if (t0 + o < t_end) x[((size_t)seg * T + t0 + o) * d + c0 + c] = xo[o] + (w2 * y + b2);
        }
    }
}

// Fused conv module: per-row LayerNorm statistics in one streaming pass, then LN + GLU are applied while the
// depthwise-conv kernel stages its tile -- u never exists in HBM (2.3 GB -> 1.4 GB of traffic per block at 605 segments).
template <int NV>
__global__ void __launch_bounds__(256)
ln_stats_vec_kernel(const float* __restrict__ x, int M, float2* __restrict__ stats) {
    constexpr int d = 128 * NV;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ld4(x + (size_t)row * d + (i * 32 + lane) * 4);
    float mean, rstd;
    ln_stats<NV>(v, 1.f / (float)d, mean, rstd);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// CTA = (segment, 32 channels), 128 threads.  Staging: one float4 of x per thread and row, LN + GLU with the per-channel
// constants folded (xn = x rstd - mean rstd; a = xn (g w1a) + (b w1a + b1a); gate exponent = xn (-log2e g w1g) - log2e (b w1g + b1g)):
// three FMAs, two MUFUs, one add and one multiply per element.  Convolution: a thread owns a channel PAIR and a run of
// frames; the two channels ride in the halves of packed fp32 FMAs (fma.rn.f32x2: one issue slot for two products), the
// 33 taps are taken in three passes of 11 so that the register window is 18 rows, tap weights come from shared memory
// (one 8-byte load per 8 packed FMAs).  Small CTAs (5 resident per SM) keep the staging loads of one CTA under the FMAs
// of the others.  Summation order per output = taps ascending, as before.
constexpr int kDw2Ch = 32, kDw2Threads = 128, kDw2Groups = kDw2Threads / (kDw2Ch / 2), kDw2Pass = 11, kDw2Slack = 8;
static_assert(kDwMaxK == 3 * kDw2Pass, "three tap passes");
__global__ void __launch_bounds__(kDw2Threads, 5)
dwconv_fused_kernel(float* __restrict__ x, const float2* __restrict__ stats, const float* __restrict__ ln_g,
                    const float* __restrict__ ln_b, int T, int d, int ks, const float* __restrict__ dw_w,
                    const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, const float* __restrict__ scalars,
                    uint32_t* __restrict__ ln_hi, uint32_t* __restrict__ ln_lo, float* __restrict__ ln_part, int ln_stride) {
    extern __shared__ __align__(16) float tile[];     // [T + ks - 1][kDw2Ch] u = GLU(LN(x)), row r <-> frame r - pad; then [kDwMaxK][kDw2Ch] taps
    const int seg = blockIdx.y, c0 = blockIdx.x * kDw2Ch;
    const int pad = (ks - 1) / 2;
    const int rows = T + kDwMaxK - 1 + kDw2Slack;     // rows past the halo are zero: the last window of a frame run reads up to row t0 + 39
    float* wt = tile + (size_t)rows * kDw2Ch;
    const float w1a = __ldg(scalars + 0), b1a = __ldg(scalars + 1), w1g = __ldg(scalars + 2), b1g = __ldg(scalars + 3);
    // tap weights of the CTA's channels: four threads per channel, taps (tid & 3) + 4 m (no index division; the run of the CTA's
    // weights is contiguous in dw_w); the loads are issued now and land in shared memory after the tile has been staged
    constexpr int kWPer = (kDwMaxK + 3) / 4;
    static_assert(kDw2Threads == 4 * kDw2Ch, "four threads per channel load its taps");
    float wreg[kWPer];
    const int wc = threadIdx.x >> 2, wj0 = threadIdx.x & 3;
#pragma unroll
    for (int k = 0; k < kWPer; ++k) {
        const int j = wj0 + 4 * k;
        wreg[k] = (j < ks && c0 + wc < d) ? __ldg(dw_w + (size_t)(c0 + wc) * ks + j) : 0.f;
    }
    {
        constexpr int kC4 = kDw2Ch / 4;                               // float4 columns of a tile row
        constexpr int kRowStep = kDw2Threads / kC4;                   // rows covered by the CTA per load
        constexpr int kDeep = 6;                                      // rows per thread in flight
        const int c4 = threadIdx.x & (kC4 - 1), tr = threadIdx.x / kC4;
        const bool c_ok = c0 + c4 * 4 < d;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // halo and slack rows
        for (int r = tr; r < rows - T; r += kRowStep) st4(tile + (r < pad ? r : r + T) * kDw2Ch + c4 * 4, zero4);
        const float4 g4 = c_ok ? ldg4(ln_g + c0 + c4 * 4) : zero4;
        const float4 b4 = c_ok ? ldg4(ln_b + c0 + c4 * 4) : zero4;
        constexpr float kNegLog2e = -1.4426950408889634f;
        const float4 gw = make_float4(g4.x * w1a, g4.y * w1a, g4.z * w1a, g4.w * w1a);
        const float4 bw = make_float4(fmaf(b4.x, w1a, b1a), fmaf(b4.y, w1a, b1a), fmaf(b4.z, w1a, b1a), fmaf(b4.w, w1a, b1a));
        const float4 gg = make_float4(kNegLog2e * g4.x * w1g, kNegLog2e * g4.y * w1g, kNegLog2e * g4.z * w1g, kNegLog2e * g4.w * w1g);
        const float4 bg = make_float4(kNegLog2e * fmaf(b4.x, w1g, b1g), kNegLog2e * fmaf(b4.y, w1g, b1g),
                                      kNegLog2e * fmaf(b4.z, w1g, b1g), kNegLog2e * fmaf(b4.w, w1g, b1g));
        auto glu = [](float xn, float gw_, float bw_, float gg_, float bg_) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(xn, gg_, bg_)));
            return __fdividef(fmaf(xn, gw_, bw_), 1.f + e);
        };
        const float* xs = x + (size_t)seg * T * d + c0 + c4 * 4;       // row t at xs + t * d
        const float2* ss = stats + (size_t)seg * T;
        float* ts = tile + pad * kDw2Ch + c4 * 4;                     // row t at ts + t * kDw2Ch
        // kDeep rows per trip: all global loads of a trip are in flight before the first use
        for (int t0 = tr; t0 < T; t0 += kDeep * kRowStep) {
            float4 xv[kDeep];
            float2 sv[kDeep];
#pragma unroll
            for (int u = 0; u < kDeep; ++u) {
                const int t = t0 + u * kRowStep;
                const bool ok = t < T && c_ok;
                xv[u] = ok ? ld4(xs + t * d) : zero4;
                sv[u] = ok ? __ldg(ss + t) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kDeep; ++u) {
                const int t = t0 + u * kRowStep;
                if (t < T) {
                    const float A = sv[u].y, B = -sv[u].x * sv[u].y;
                    const float4 v = c_ok ? make_float4(glu(fmaf(xv[u].x, A, B), gw.x, bw.x, gg.x, bg.x), glu(fmaf(xv[u].y, A, B), gw.y, bw.y, gg.y, bg.y),
                                                        glu(fmaf(xv[u].z, A, B), gw.z, bw.z, gg.z, bg.z), glu(fmaf(xv[u].w, A, B), gw.w, bw.w, gg.w, bg.w))
                                          : zero4;
                    st4(ts + t * kDw2Ch, v);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kWPer; ++k) {
        const int j = wj0 + 4 * k;
        if (j < kDwMaxK) wt[j * kDw2Ch + wc] = wreg[k];               // taps >= ks and channels >= d are zeros
    }
    __syncthreads();
    const int cp = threadIdx.x & (kDw2Ch / 2 - 1), grp = threadIdx.x / (kDw2Ch / 2);
    const int c = c0 + 2 * cp;
    if (c >= d) return;
    const float2 sc = __ldg(reinterpret_cast<const float2*>(bn_scale + c)), sh = __ldg(reinterpret_cast<const float2*>(bn_shift + c));
    const float w2 = __ldg(scalars + 4), b2 = __ldg(scalars + 5);
    const int per = ceil_div(ceil_div(T, kDw2Groups), kDwOut) * kDwOut;
    const int t_end = min(T, (grp + 1) * per);
    const float2* tile2 = reinterpret_cast<const float2*>(tile) + cp;         // row stride kDw2Ch / 2
    const float2* wt2 = reinterpret_cast<const float2*>(wt) + cp;
    float* xc = x + (size_t)seg * T * d + c;                                  // frame t of the channel pair at xc + t * d
    // every thread runs the same number of windows (the row moments below are warp collectives); a window past the
    // thread's frame run is skipped but for the shuffles
    for (int wi = 0; wi < per / kDwOut; ++wi) {
        const int t0 = grp * per + wi * kDwOut;
        const bool live = t0 < t_end;
        float2 xn[kDwOut];
#pragma unroll
        for (int o = 0; o < kDwOut; ++o) xn[o] = make_float2(0.f, 0.f);
        if (live) {
            float2 xo[kDwOut];
#pragma unroll
            for (int o = 0; o < kDwOut; ++o)
                xo[o] = (t0 + o < t_end) ? *reinterpret_cast<const float2*>(xc + (t0 + o) * d) : make_float2(0.f, 0.f);
            float2 acc[kDwOut];
#pragma unroll
            for (int o = 0; o < kDwOut; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int p = 0; p < 3; ++p) {
                float2 win[kDwOut + kDw2Pass - 1];
                const float2* tp = tile2 + (t0 + kDw2Pass * p) * (kDw2Ch / 2);   // rows up to t0 + 39 exist (zero filled past the halo)
#pragma unroll
                for (int i = 0; i < kDwOut + kDw2Pass - 1; ++i) win[i] = tp[i * (kDw2Ch / 2)];
                const float2* wp = wt2 + kDw2Pass * p * (kDw2Ch / 2);
#pragma unroll
                for (int jj = 0; jj < kDw2Pass; ++jj) {
                    const float2 wj = wp[jj * (kDw2Ch / 2)];
#pragma unroll
                    for (int o = 0; o < kDwOut; ++o) acc[o] = __ffma2_rn(wj, win[o + jj], acc[o]);
                }
            }
            // running pointers: one 64-bit add per row instead of an index computation per store
            float* xw = xc + (size_t)t0 * d;
            const size_t e0 = (((size_t)seg * T + t0) * d + c) >> 1;
            uint32_t* hw = ln_hi + e0;
            uint32_t* lw = ln_lo + e0;
            const int d2 = d >> 1;
#pragma unroll
            for (int o = 0; o < kDwOut; ++o) {
                if (t0 + o < t_end) {
                    const float y0 = fmaxf(fmaf(acc[o].x, sc.x, sh.x), 0.f), y1 = fmaxf(fmaf(acc[o].y, sc.y, sh.y), 0.f);
                    xn[o] = make_float2(xo[o].x + fmaf(w2, y0, b2), xo[o].y + fmaf(w2, y1, b2));
                    *reinterpret_cast<float2*>(xw) = xn[o];
                    if (ln_part) {
                        // LayerNorm source of feed_forward_out: the new rows as raw bf16 head / remainder planes
                        const __nv_bfloat162 hq = __floats2bfloat162_rn(xn[o].x, xn[o].y);
                        const __nv_bfloat162 lq = __floats2bfloat162_rn(xn[o].x - __low2float(hq), xn[o].y - __high2float(hq));
                        *hw = *reinterpret_cast<const uint32_t*>(&hq);
                        *lw = *reinterpret_cast<const uint32_t*>(&lq);
                    }
                }
                xw += d; hw += d2; lw += d2;
            }
        }
        if (ln_part) {
            // (sum, sum of squares) of the CTA's 32 channels for the window's 8 rows: 16 values per thread, a transposing
            // butterfly over the 16 lanes of the frame group (15 shuffles instead of 64); lane cp ends with value cp
            float v8[8], v4[4], v2[2];
            {
                const bool up = (cp & 8) != 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int o = j >> 1;
                    // value index i = 2 o + kind; lower half i < 8 <-> outputs 0..3, upper half <-> outputs 4..7
                    const float lo_v = (j & 1) ? fmaf(xn[o].x, xn[o].x, xn[o].y * xn[o].y) : xn[o].x + xn[o].y;
                    const float hi_v = (j & 1) ? fmaf(xn[o + 4].x, xn[o + 4].x, xn[o + 4].y * xn[o + 4].y) : xn[o + 4].x + xn[o + 4].y;
                    const float keep = up ? hi_v : lo_v, send = up ? lo_v : hi_v;
                    v8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            {
                const bool up = (cp & 4) != 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float keep = up ? v8[j + 4] : v8[j], send = up ? v8[j] : v8[j + 4];
                    v4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
            }
            {
                const bool up = (cp & 2) != 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float keep = up ? v4[j + 2] : v4[j], send = up ? v4[j] : v4[j + 2];
                    v2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
            }
            const bool up = (cp & 1) != 0;
            const float tot = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 1);
            // lane cp holds value index cp: output row t0 + (cp >> 1), kind cp & 1 (0: sum, 1: sum of squares)
            const int t = t0 + (cp >> 1);
            if (t < t_end) ln_part[(((size_t)seg * T + t) * ln_stride + blockIdx.x) * 2 + (cp & 1)] = tot;
        }
    }
}

// (mean, rstd) of every row from the partial moments the producer of the rows left behind (slots per row, d columns)
__global__ void __launch_bounds__(256)
ln_finalize_kernel(const float2* __restrict__ part, int M, int stride, int used, float inv_d, float2* __restrict__ stats) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < used; ++i) {
        const float2 v = __ldg(part + (size_t)m * stride + i);
        s += v.x;
        q += v.y;
    }
    const float mean = s * inv_d;
    const float var = fmaxf(fmaf(-mean, mean, q * inv_d), 0.f);
    stats[m] = make_float2(mean, 1.f / sqrtf(var + 1e-5f));
}

// softmax over t2 of (S1[bh][t1][t2] + S2[bh*T + t1][t1 - t2 + T - 1]) / sqrt(d_k)       conformer.py:73-87
constexpr int kSmMaxPerLane = 8;    // T <= 256
__global__ void __launch_bounds__(256)
relpos_softmax_kernel(const float* __restrict__ S1, const float* __restrict__ S2, int rows_total, int T, int Tp, int ld2,
                      float inv_sqrt_dk, float* __restrict__ P_hi, float* __restrict__ P_lo) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);       // bh * T + t1
    const int lane = threadIdx.x & 31;
    if (row >= rows_total) return;
    const int bh = row / T, t1 = row - bh * T;
    const float* s1 = S1 + ((size_t)bh * T + t1) * Tp;
    const float* s2 = S2 + (size_t)row * ld2 + (t1 + T - 1);
    float v[kSmMaxPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kSmMaxPerLane; ++i) {
        const int t2 = i * 32 + lane;
        v[i] = -INFINITY;
        if (t2 < T) { v[i] = (__ldg(s1 + t2) + __ldg(s2 - t2)) * inv_sqrt_dk; mx = fmaxf(mx, v[i]); }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kSmMaxPerLane; ++i) {
        const int t2 = i * 32 + lane;
        if (t2 < T) { v[i] = expf(v[i] - mx); sum += v[i]; }
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < kSmMaxPerLane; ++i) {
        const int t2 = i * 32 + lane;
        if (t2 < Tp) {
            const float pv = t2 < T ? v[i] * inv : 0.f;
            float hi, lo;
            split_tf32(pv, hi, lo);
            const size_t o = ((size_t)bh * T + t1) * Tp + t2;
            P_hi[o] = hi;
            P_lo[o] = lo;
        }
    }
}

// scale: extra factor for SPLIT_F16 operands that are not activations (weights: kF16WeightScale / kF16ActScale)
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ hi, float* __restrict__ lo, int fmt, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_store(fmt, hi, lo, (size_t)i, in[i] * scale);
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Folded LayerNorms (the attention LayerNorm into the QKV projection, the feed_forward_out LayerNorm into its first linear
// layer): LN(x) W^T = rstd (x (gamma W)^T - mean colsum(gamma W)) + (b + W beta), so the producer of x (the residual
// epilogue of the previous GEMM / the conv-module kernel) writes x itself as the next GEMM's split operand together with
// per-row partial moments, and the consumer's epilogue applies the row statistics: the two LayerNorm passes over the
// activations (920 MB each per block at 1 209 segments) disappear.  Needs the vector epilogues and the fused conv-module
// kernel: 2xBF16 engine, d_model = 128 * {1, 2, 4, 8}.  The weight blob is packed accordingly (separator.py asks
// nsf_conformer_ln_fold); NSF_LN_FOLD=0 switches it off (A/B measurements).
constexpr int kLnSlots = 16;        // partial moments per row: 2 per 256-column GEMM tile, 1 per 32-channel conv CTA
static bool ln_fold_rule(const nsf_conformer_dims& D) {
    const char* e = getenv("NSF_LN_FOLD");
    if (e && e[0] == '0') return false;
    const int g = D.d_model / 128;
    return D.gemm_engine == NSF_GEMM_TC_2XBF16 && D.d_model % 128 == 0 && g <= 8 && (g & (g - 1)) == 0 &&
           D.d_model / kDw2Ch <= kLnSlots;
}

struct Workspace {
    float *x, *h_hi, *h_lo, *u_hi, *u_lo, *q_hi, *q_lo, *k_hi, *k_lo, *vt_hi, *vt_lo, *s1, *s2, *p_hi, *p_lo;
    float *ln_part, *ln_stats;      // folded LayerNorms: [M][kLnSlots] partial (sum, sum of squares), [M] (mean, rstd)
    int64_t total_floats;
};

static Workspace carve(const nsf_conformer_dims& D, int n_seg, float* base) {
    const int64_t M = (int64_t)n_seg * D.T;
    const int Tp = (int)align_up(D.T, 32);
    const int ld2 = (int)align_up(2 * D.T - 1, 32);
    const int64_t BH = (int64_t)n_seg * D.n_heads;
    const int d_k = D.d_model / D.n_heads;
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += align_up(n, 64); return p; };
    Workspace w;
    w.x = take(M * D.d_model);
    w.h_hi = take(M * D.d_model);
    w.h_lo = take(M * D.d_model);
    w.u_hi = take(M * D.d_ff);
    w.u_lo = take(M * D.d_ff);
    w.q_hi = take(M * D.d_model);
    w.q_lo = take(M * D.d_model);
    w.k_hi = take(M * D.d_model);
    w.k_lo = take(M * D.d_model);
    w.vt_hi = take(BH * d_k * Tp);
    w.vt_lo = take(BH * d_k * Tp);
    const bool unfused = !attn_fused_supported(D.T, d_k) || D.gemm_engine == NSF_GEMM_SIMT_FP32;   // score buffers of the unfused path
    w.s1 = take(unfused ? BH * D.T * Tp : 0);
    w.s2 = take(unfused ? BH * D.T * ld2 : 0);
    w.p_hi = take(unfused ? BH * D.T * Tp : 0);
    w.p_lo = take(unfused ? BH * D.T * Tp : 0);
    w.ln_part = take(M * kLnSlots * 2);
    w.ln_stats = take(M * 2);
    w.total_floats = off;
    return w;
}

static int validate_dims(const nsf_conformer_dims& D) {
    NSF_REQUIRE(D.d_model >= 32 && D.d_model <= 1024 && D.d_model % 32 == 0, "conformer: d_model=%d", D.d_model);
    NSF_REQUIRE(D.n_heads >= 1 && D.d_model % D.n_heads == 0 && (D.d_model / D.n_heads) % 32 == 0,
                "conformer: d_model/n_heads must be a multiple of 32");
    NSF_REQUIRE(D.d_ff >= 32 && D.d_ff % 32 == 0, "conformer: d_ff=%d", D.d_ff);
    NSF_REQUIRE(D.n_blocks >= 1 && D.kernel_size >= 1 && D.kernel_size <= kDwMaxK && (D.kernel_size & 1),
                "conformer: kernel_size=%d", D.kernel_size);
    NSF_REQUIRE(D.T >= 2 && D.T <= 32 * kSmMaxPerLane && D.T <= D.maxlen, "conformer: T=%d", D.T);
    NSF_REQUIRE(D.n_out >= kBins && D.n_out % kBins == 0, "conformer: n_out=%d", D.n_out);
    NSF_REQUIRE(D.gemm_engine >= NSF_GEMM_SIMT_FP32 && D.gemm_engine <= NSF_GEMM_TC_2XF16, "conformer: gemm_engine");
    return NSF_OK;
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_conformer_ln_fold(const nsf_conformer_dims* dims) { return dims && ln_fold_rule(*dims) ? 1 : 0; }

extern "C" int64_t nsf_conformer_num_offsets(const nsf_conformer_dims* dims) {
    return dims ? kGlobalOffsets + (int64_t)kPerLayer * dims->n_blocks : 0;
}

extern "C" int nsf_conformer_create(const nsf_conformer_dims* dims, const float* blob, int64_t blob_floats,
                                    const int64_t* offsets, int n_offsets, nsf_conformer** out) {
    NSF_REQUIRE(dims && blob && offsets && out, "nsf_conformer_create: null pointer");
    int rc = validate_dims(*dims);
    if (rc) return rc;
    NSF_REQUIRE(n_offsets == nsf_conformer_num_offsets(dims), "nsf_conformer_create: expected %lld offsets, got %d",
                (long long)nsf_conformer_num_offsets(dims), n_offsets);
    for (int i = 0; i < n_offsets; ++i)
        NSF_REQUIRE(offsets[i] >= 0 && offsets[i] < blob_floats && (offsets[i] & 3) == 0,
                    "nsf_conformer_create: offset %d = %lld out of range or not 16-byte aligned", i, (long long)offsets[i]);
    nsf_conformer* h = new (std::nothrow) nsf_conformer;
    NSF_REQUIRE(h, "nsf_conformer_create: out of memory");
    h->dims = *dims;
    h->ln_fold = ln_fold_rule(*dims);
    h->blob = blob;
    h->blob_floats = blob_floats;
    h->n_offsets = n_offsets;
    h->offsets = new (std::nothrow) int64_t[n_offsets];
    if (!h->offsets) { delete h; set_error("nsf_conformer_create: out of memory"); return NSF_ERR_INVALID_ARG; }
    for (int i = 0; i < n_offsets; ++i) h->offsets[i] = offsets[i];
    *out = h;
    return NSF_OK;
}

extern "C" void nsf_conformer_destroy(nsf_conformer* h) {
    if (!h) return;
    delete[] h->offsets;
    delete h;
}

extern "C" int64_t nsf_conformer_workspace_bytes(const nsf_conformer_dims* dims, int n_seg) {
    if (!dims || n_seg <= 0) return 0;
    return carve(*dims, n_seg, nullptr).total_floats * (int64_t)sizeof(float);
}

extern "C" int nsf_conformer_forward(nsf_conformer* h, const float* feat, const float* feat_lo, int64_t ldf, int n_seg,
                                     float* masks, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(h && feat && masks && workspace, "nsf_conformer_forward: null pointer");
    if (n_seg <= 0) return NSF_OK;
    const nsf_conformer_dims& D = h->dims;
    const int eng = D.gemm_engine;
    NSF_REQUIRE(eng == NSF_GEMM_SIMT_FP32 || feat_lo, "nsf_conformer_forward: tensor-core engines need split features");
    const int Kf = (int)align_up(D.in_features, 32);
    NSF_REQUIRE(ldf >= Kf && (ldf & 3) == 0, "nsf_conformer_forward: ldf=%lld must be >= %d and a multiple of 4", (long long)ldf, Kf);
    NSF_REQUIRE(((uintptr_t)workspace & 255) == 0, "nsf_conformer_forward: workspace must be 256-byte aligned");
    Workspace w = carve(D, n_seg, reinterpret_cast<float*>(workspace));
    NSF_REQUIRE(workspace_bytes >= w.total_floats * (int64_t)sizeof(float), "nsf_conformer_forward: workspace too small");
    cudaStream_t s = (cudaStream_t)stream_;

    const int T = D.T, d = D.d_model, H = D.n_heads, d_k = d / H, dff = D.d_ff;
    const int M = n_seg * T;
    // activations that feed the linear layers are stored in the engine's split format; q / k / v^T / pe_k and the
    // attention-internal products stay SPLIT_TF32 (3xTF32) under the 16-bit engines
    const int fmt = split_fmt_of_engine(eng);
    const int eng_attn = fmt == SPLIT_TF32 ? eng : NSF_GEMM_TC_3XTF32;
    // 2xBF16 runs the attention on bf16 pairs too (attention16.cu); 2xF16 keeps the 3xTF32 attention (fp32-grade throughout)
    const bool attn16 = eng == NSF_GEMM_TC_2XBF16 && attn_fused_supported(T, d_k);
    const int qkv_fmt = attn16 ? SPLIT_BF16 : SPLIT_TF32;
    const float lin_scale = fmt == SPLIT_F16 ? 1.f / (kF16ActScale * kF16WeightScale) : 1.f;
    const int Tp = (int)align_up(T, 32), ld2 = (int)align_up(2 * T - 1, 32);
    const int BH = n_seg * H;
    int rc;

    auto base_params = [&]() {
        GemmParams p = {};
        p.batch = 1;
        p.alpha = 1.f;
        p.acc_scale = 1.f;
        p.op_fmt = SPLIT_TF32; p.out_fmt = fmt; p.qkv_fmt = qkv_fmt; p.v_rowmajor = attn16 ? 1 : 0;
        p.T = T; p.Tp = Tp; p.n_heads = H; p.d_k = d_k; p.d_model = d;
        return p;
    };
    const bool fold = h->ln_fold;
    float2* ln_part = reinterpret_cast<float2*>(w.ln_part);
    float2* ln_stats = reinterpret_cast<float2*>(w.ln_stats);
    // ln_src: the residual epilogue also writes the rows as the next GEMM's operand (h planes) + partial moments;
    // ln_csum: the operand rows are raw, LayerNorm is applied in the epilogue (ln_stats, column sums of the scaled weights)
    auto linear = [&](const float* a_hi, const float* a_lo, int64_t lda, int K, const float* w_hi, const float* w_lo,
                      const float* bias, int N, int epi, float alpha, float* o0, float* o1, int64_t ldo,
                      int ln_src = 0, const float* ln_csum = nullptr) {
        GemmParams p = base_params();
        if (ln_src) { p.ln_part = ln_part; p.ln_slots = kLnSlots; }
        if (ln_src == 1) { p.ln_hi = w.h_hi; p.ln_lo = w.h_lo; }         // 2: row moments only (the conv module's LayerNorm statistics)
        if (ln_csum) { p.ln_stats = ln_stats; p.ln_csum = ln_csum; }
        p.A_hi = a_hi; p.A_lo = a_lo; p.lda = lda;
        p.B_hi = w_hi; p.B_lo = w_lo; p.ldb = K;
        p.M = M; p.N = N; p.K = K; p.n_valid = N;
        p.op_fmt = fmt; p.acc_scale = lin_scale;
        p.bias = bias; p.epi = epi; p.alpha = alpha;
        p.out0 = o0; p.out1 = o1; p.ldo = ldo;
        p.q_hi = w.q_hi; p.q_lo = w.q_lo; p.k_hi = w.k_hi; p.k_lo = w.k_lo; p.vt_hi = w.vt_hi; p.vt_lo = w.vt_lo;
        return gemm_launch(eng, p, s);
    };

    // the time padding of V^T (columns T..Tp-1) must be zero for the P V product
    if (Tp != T && !attn16) {
        NSF_CUDA(cudaMemsetAsync(w.vt_hi, 0, sizeof(float) * (size_t)BH * d_k * Tp, s));
        NSF_CUDA(cudaMemsetAsync(w.vt_lo, 0, sizeof(float) * (size_t)BH * d_k * Tp, s));
    }

    // embed: Linear -> LayerNorm -> ReLU (conformer.py:205-210), fused with layer 0's first LayerNorm
    rc = linear(feat, feat_lo, ldf, Kf, h->g(G_EMB_W_HI), h->g(G_EMB_W_LO), h->g(G_EMB_B), d, EPI_STORE, 1.f, w.x, nullptr, d);
    if (rc) return rc;
    { ProfScope prof(PROF_NET_OTHER, 0.0, s);
    rc = ln_launch(w.x, M, d, h->g(G_EMB_LN_G), h->g(G_EMB_LN_B), 1, w.x, h->l(0, L_FFI_LN_G), h->l(0, L_FFI_LN_B), w.h_hi, w.h_lo, fmt, s); }
    if (rc) return rc;

    const float inv_sqrt_dk = 1.f / sqrtf((float)d_k);
    for (int L = 0; L < D.n_blocks; ++L) {
        // x += 0.5 * FF_in(x)                                                                  conformer.py:179
        rc = linear(w.h_hi, w.h_lo, d, d, h->l(L, L_FFI_W1_HI), h->l(L, L_FFI_W1_LO), h->l(L, L_FFI_B1), dff, EPI_RELU_SPLIT,
                    1.f, w.u_hi, w.u_lo, dff);
        if (rc) return rc;
        rc = linear(w.u_hi, w.u_lo, dff, dff, h->l(L, L_FFI_W2_HI), h->l(L, L_FFI_W2_LO), h->l(L, L_FFI_B2), d, EPI_RESID, 0.5f,
                    w.x, nullptr, d, fold ? 1 : 0);
        if (rc) return rc;
        // x += MHSA(x)                                                                         conformer.py:180
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          if (fold) {
              ln_finalize_kernel<<<ceil_div(M, 256), 256, 0, s>>>(ln_part, M, kLnSlots, 2 * ceil_div(d, 256), 1.f / (float)d, ln_stats);
              rc = check_launch("ln_finalize_kernel");
          } else {
              rc = ln_launch(w.x, M, d, h->l(L, L_ATT_LN_G), h->l(L, L_ATT_LN_B), 0, nullptr, nullptr, nullptr, w.h_hi, w.h_lo, fmt, s);
          } }
        if (rc) return rc;
        rc = linear(w.h_hi, w.h_lo, d, d, h->l(L, L_WQKV_HI), h->l(L, L_WQKV_LO), h->l(L, L_BQKV), 3 * d, EPI_QKV, 1.f, nullptr,
                    nullptr, 0, 0, fold ? h->l(L, L_QKV_CSUM) : nullptr);
        if (rc) return rc;
        if (attn_fused_supported(T, d_k) && eng != NSF_GEMM_SIMT_FP32) {
            // scores, relative-position skew, softmax and P V in one tcgen05 kernel (attention.cu)
            ProfScope prof(PROF_ATTN, 6.0 * T * T * d_k * (double)BH, s);
            if (attn16) {
                if ((rc = attn16_launch(w.q_hi, w.q_lo, w.k_hi, w.k_lo, w.vt_hi, w.vt_lo, h->g(G_PE16_HI), h->g(G_PE16_LO), D.maxlen,
                                        n_seg, H, T, Tp, w.h_hi, w.h_lo, d, fmt, s))) return rc;
            } else if ((rc = attn_fused_launch(w.q_hi, w.q_lo, w.k_hi, w.k_lo, w.vt_hi, w.vt_lo, h->g(G_PE_HI), h->g(G_PE_LO), D.maxlen,
                                        n_seg, H, T, Tp, w.h_hi, w.h_lo, d, fmt, s))) return rc;
        } else {
            {   // A = q k^T per (segment, head)
                GemmParams p = base_params();
                p.A_hi = w.q_hi; p.A_lo = w.q_lo; p.lda = d_k; p.a_batch_stride = (int64_t)T * d_k;
                p.B_hi = w.k_hi; p.B_lo = w.k_lo; p.ldb = d_k; p.b_batch_stride = (int64_t)T * d_k;
                p.M = T; p.N = T; p.K = d_k; p.n_valid = T; p.batch = BH;
                p.epi = EPI_STORE; p.out0 = w.s1; p.ldo = Tp; p.o_batch_stride = (int64_t)T * Tp;
                if ((rc = gemm_launch(eng_attn, p, s))) return rc;
            }
            {   // B' = q pe_k[maxlen-(T-1) .. maxlen+(T-1)]^T for every (segment, head, t1) row at once
                GemmParams p = base_params();
                p.A_hi = w.q_hi; p.A_lo = w.q_lo; p.lda = d_k;
                const int64_t pe_off = (int64_t)(D.maxlen - (T - 1)) * d_k;
                p.B_hi = h->g(G_PE_HI) + pe_off; p.B_lo = h->g(G_PE_LO) + pe_off; p.ldb = d_k;
                p.M = BH * T; p.N = 2 * T - 1; p.K = d_k; p.n_valid = 2 * T - 1;
                p.epi = EPI_STORE; p.out0 = w.s2; p.ldo = ld2;
                if ((rc = gemm_launch(eng_attn, p, s))) return rc;
            }
            { ProfScope prof(PROF_NET_OTHER, 0.0, s); relpos_softmax_kernel<<<ceil_div(BH * T, 8), 256, 0, s>>>(w.s1, w.s2, BH * T, T, Tp, ld2, inv_sqrt_dk, w.p_hi, w.p_lo); }
            if ((rc = check_launch("relpos_softmax_kernel"))) return rc;
            {   // o = p v per (segment, head), gathered back to [M, d_model]
                GemmParams p = base_params();
                p.A_hi = w.p_hi; p.A_lo = w.p_lo; p.lda = Tp; p.a_batch_stride = (int64_t)T * Tp;
                p.B_hi = w.vt_hi; p.B_lo = w.vt_lo; p.ldb = Tp; p.b_batch_stride = (int64_t)d_k * Tp;
                p.M = T; p.N = d_k; p.K = Tp; p.n_valid = d_k; p.batch = BH;
                p.epi = EPI_PV; p.out0 = w.h_hi; p.out1 = w.h_lo; p.ldo = d;
                if ((rc = gemm_launch(eng_attn, p, s))) return rc;
            }
        }
        rc = linear(w.h_hi, w.h_lo, d, d, h->l(L, L_WO_HI), h->l(L, L_WO_LO), h->l(L, L_BO), d, EPI_RESID, 1.f, w.x, nullptr, d, fold ? 2 : 0);
        if (rc) return rc;
        // x += Conv(x)                                                                         conformer.py:181
        if (d % 128 == 0 && d / 128 <= 8 && ((d / 128) & (d / 128 - 1)) == 0) {
            // LN statistics, then LN + GLU + depthwise conv + BN + ReLU + affine + residual in one pass over x
            ProfScope prof(PROF_NET_OTHER, 0.0, s);
            float2* stats = reinterpret_cast<float2*>(w.u_hi);          // [M] (mean, rstd); u_hi is free between the FFNs
            const int grid_s = ceil_div(M, 8);
            if (fold) {
                // the out-projection's residual epilogue left the row moments behind
                ln_finalize_kernel<<<ceil_div(M, 256), 256, 0, s>>>(ln_part, M, kLnSlots, 2 * ceil_div(d, 256), 1.f / (float)d, stats);
            } else
            switch (d / 128) {
                case 1: ln_stats_vec_kernel<1><<<grid_s, 256, 0, s>>>(w.x, M, stats); break;
                case 2: ln_stats_vec_kernel<2><<<grid_s, 256, 0, s>>>(w.x, M, stats); break;
                case 4: ln_stats_vec_kernel<4><<<grid_s, 256, 0, s>>>(w.x, M, stats); break;
                default: ln_stats_vec_kernel<8><<<grid_s, 256, 0, s>>>(w.x, M, stats); break;
            }
            if ((rc = check_launch("ln_stats_vec_kernel"))) return rc;
            dim3 grid(ceil_div(d, kDw2Ch), n_seg);
            const size_t smem = (size_t)(T + kDwMaxK - 1 + kDw2Slack + kDwMaxK) * kDw2Ch * sizeof(float);
            NSF_CUDA(cudaFuncSetAttribute(dwconv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dwconv_fused_kernel<<<grid, kDw2Threads, smem, s>>>(w.x, stats, h->l(L, L_CONV_LN_G), h->l(L, L_CONV_LN_B), T, d, D.kernel_size,
                                                                h->l(L, L_DW_W), h->l(L, L_BN_SCALE), h->l(L, L_BN_SHIFT), h->l(L, L_CONV_SCALARS),
                                                                fold ? reinterpret_cast<uint32_t*>(w.h_hi) : nullptr,
                                                                fold ? reinterpret_cast<uint32_t*>(w.h_lo) : nullptr, fold ? w.ln_part : nullptr, kLnSlots);
            if ((rc = check_launch("dwconv_fused_kernel"))) return rc;
        } else {
            { ProfScope prof(PROF_NET_OTHER, 0.0, s);
              rc = ln_glu_launch(w.x, M, d, h->l(L, L_CONV_LN_G), h->l(L, L_CONV_LN_B), h->l(L, L_CONV_SCALARS), w.u_hi, s); }
            if (rc) return rc;
            dim3 grid(ceil_div(d, kDwCh), n_seg);
            const size_t smem = (size_t)(T + D.kernel_size - 1) * kDwCh * sizeof(float);
            NSF_CUDA(cudaFuncSetAttribute(dwconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            { ProfScope prof(PROF_NET_OTHER, 0.0, s); dwconv_kernel<<<grid, 256, smem, s>>>(w.u_hi, w.x, T, d, D.kernel_size, h->l(L, L_DW_W), h->l(L, L_BN_SCALE),
                                                 h->l(L, L_BN_SHIFT), h->l(L, L_CONV_SCALARS)); }
            if ((rc = check_launch("dwconv_kernel"))) return rc;
        }
        // x += 0.5 * FF_out(x)                                                                 conformer.py:182
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          if (fold) {
              ln_finalize_kernel<<<ceil_div(M, 256), 256, 0, s>>>(ln_part, M, kLnSlots, d / kDw2Ch, 1.f / (float)d, ln_stats);
              rc = check_launch("ln_finalize_kernel");
          } else {
              rc = ln_launch(w.x, M, d, h->l(L, L_FFO_LN_G), h->l(L, L_FFO_LN_B), 0, nullptr, nullptr, nullptr, w.h_hi, w.h_lo, fmt, s);
          } }
        if (rc) return rc;
        rc = linear(w.h_hi, w.h_lo, d, d, h->l(L, L_FFO_W1_HI), h->l(L, L_FFO_W1_LO), h->l(L, L_FFO_B1), dff, EPI_RELU_SPLIT,
                    1.f, w.u_hi, w.u_lo, dff, 0, fold ? h->l(L, L_FFO_CSUM) : nullptr);
        if (rc) return rc;
        rc = linear(w.u_hi, w.u_lo, dff, dff, h->l(L, L_FFO_W2_HI), h->l(L, L_FFO_W2_LO), h->l(L, L_FFO_B2), d, EPI_RESID, 0.5f,
                    w.x, nullptr, d);
        if (rc) return rc;
        // x = LN(x) (conformer.py:184), fused with the next block's first LayerNorm (or the split for the mask head)
        const bool last = (L == D.n_blocks - 1);
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          rc = ln_launch(w.x, M, d, h->l(L, L_OUT_LN_G), h->l(L, L_OUT_LN_B), 0, last ? nullptr : w.x, last ? nullptr : h->l(L + 1, L_FFI_LN_G),
                         last ? nullptr : h->l(L + 1, L_FFI_LN_B), w.h_hi, w.h_lo, fmt, s); }
        if (rc) return rc;
    }
    // mask head: sigmoid(Linear), transposed into [seg][mask][F][T]                             conformer.py:302-309
    rc = linear(w.h_hi, w.h_lo, d, d, h->g(G_HEAD_W_HI), h->g(G_HEAD_W_LO), h->g(G_HEAD_B), D.n_out, EPI_MASK, 1.f, masks, nullptr, 0);
    return rc;
}

extern "C" int nsf_gemm_test(int engine, const float* A, const float* W, const float* bias, float* Cout, int M, int N,
                             int K, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(A && W && Cout && workspace, "nsf_gemm_test: null pointer");
    NSF_REQUIRE(K % 32 == 0 && M > 0 && N > 0, "nsf_gemm_test: K must be a multiple of 32");
    NSF_REQUIRE(engine >= NSF_GEMM_SIMT_FP32 && engine <= NSF_GEMM_TC_BF16, "nsf_gemm_test: engine");
    const int64_t na = (int64_t)M * K, nw = (int64_t)N * K;
    const int64_t need = (align_up(na, 64) * 2 + align_up(nw, 64) * 2) * (int64_t)sizeof(float);
    NSF_REQUIRE(workspace_bytes >= need, "nsf_gemm_test: workspace needs %lld bytes", (long long)need);
    cudaStream_t s = (cudaStream_t)stream_;
    float* a_hi = reinterpret_cast<float*>(workspace);
    float* a_lo = a_hi + align_up(na, 64);
    float* w_hi = a_lo + align_up(na, 64);
    float* w_lo = w_hi + align_up(nw, 64);
    const int fmt = split_fmt_of_engine(engine);
    split_kernel<<<(unsigned)ceil_div64(na, 256), 256, 0, s>>>(A, na, a_hi, a_lo, fmt, 1.f);
    split_kernel<<<(unsigned)ceil_div64(nw, 256), 256, 0, s>>>(W, nw, w_hi, w_lo, fmt, fmt == SPLIT_F16 ? kF16WeightScale / kF16ActScale : 1.f);
    int rc = check_launch("split_kernel");
    if (rc) return rc;
    GemmParams p = {};
    p.A_hi = a_hi; p.A_lo = a_lo; p.lda = K;
    p.B_hi = w_hi; p.B_lo = w_lo; p.ldb = K;
    p.M = M; p.N = N; p.K = K; p.n_valid = N; p.batch = 1;
    p.op_fmt = fmt; p.out_fmt = fmt; p.acc_scale = fmt == SPLIT_F16 ? 1.f / (kF16ActScale * kF16WeightScale) : 1.f;
    p.bias = bias; p.epi = EPI_STORE; p.alpha = 1.f; p.out0 = Cout; p.ldo = N;
    return gemm_launch(engine, p, s);
}
