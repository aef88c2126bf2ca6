// Segment feature extraction (HBM-bound elementwise + per-(segment, bin) time reductions).
//
// Reference: ConformerCssWrapper.separate front half (css/training/conformer_wrapper.py:91-94:
// stft.abs(), stft.angle()) + FeatureExtractor.forward (css_with_conformer/executor/feature.py:543-569):
//   spectral  f0 = clamp(|X_0|, eps); (f0 - mean_T) / (std_T(unbiased) + eps)          feature.py:496-507
//   spatial   d_m = angle(X_m) - angle(X_0); yr = cos d, yi = sin d;
//             ipd_m = atan2(yi - mean_T yi, yr - mean_T yr)   (ipd_mean_normalize_version 1)  feature.py:214-221
// Generic bins take cos/sin of the phase difference straight from X_m conj(X_0) / (|X_m||X_0|)
// (no atan2 -> sincos round trip).  The DC and Nyquist bins are real up to the sin(pi_f32) residue
// th.polar leaves, so there the *sign* of every IPD hangs on the last bit of the reference's
// float32 atan2/sin: those two bins replay the reference's arithmetic literally
// (angle -> -3.1415925f for negative real parts, sinf/cosf of the float32 difference).
#include "common.cuh"

namespace nsf {

constexpr int kFeatBinsPerCta = 8;     // one warp per bin
constexpr int kFeatMaxIter = 8;        // T <= 256
constexpr float kEps32 = 1.1920928955078125e-07f;   // th.finfo(th.float32).eps, feature.py:15



// atan2 without branches: |error| < 1e-7 rad (1 ulp at pi/4).  atan(a) = a P(a^2) on [0, 1] (degree-8 minimax fit in a^2,
// 9.2e-8 evaluated in fp32), the octant folded back with pi/2 - r and pi - r, the sign of y copied on: the +-pi cut is
// decided by the sign bit of y exactly as in atan2f (atan2f itself is ~70 instructions with branches; 36 calls per lane
// were 45 % of this kernel's instructions, profiles/r02_ncu_full_hbm_kernels_summary.csv).
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;             // atan2(0, 0) = 0
    const float s = a * a;
    float r = 0.0024567132350057364f;
    r = fmaf(r, s, -0.014401310123503208f);
    r = fmaf(r, s, 0.03978114202618599f);
    r = fmaf(r, s, -0.07234849780797958f);
    r = fmaf(r, s, 0.1049894243478775f);
    r = fmaf(r, s, -0.14161227643489838f);
    r = fmaf(r, s, 0.19985906779766083f);
    r = fmaf(r, s, -0.33332598209381104f);
    r = fmaf(r, s, 0.9999998807907104f);
    r *= a;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = (__float_as_uint(x) >> 31) ? 3.14159265358979324f - r : r;   // x < 0 or x = -0
    return copysignf(r, y);
}

// NITER = ceil(T / 32) register slots per lane (6 for the pipeline's T = 186: two CTAs per SM instead of one)
template <int C, int NITER>
__global__ void __launch_bounds__(kFeatBinsPerCta * 32, NITER <= 6 ? 4 : 2)
css_features_kernel(const float2* __restrict__ X, int64_t T_long, int64_t T_valid, int64_t seg_first, int T, int hop,
                    const float* __restrict__ in_bias, const float* __restrict__ in_scale,
                    float* __restrict__ feat, float* __restrict__ feat_lo, int64_t ldf, int fmt) {
    extern __shared__ float tile[];     // [T][C][kFeatBinsPerCta]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.y;
    const int f0bin = blockIdx.x * kFeatBinsPerCta;
    const int f = f0bin + warp;
    const int64_t st = (seg_first + seg) * (int64_t)hop;

    if (f < kBins) {
        const bool edge_bin = (f == 0 || f == kBins - 1);
        const float2* Xf = X + ((size_t)f * T_long + st) * C;
        constexpr int CM = C - 1 > 0 ? C - 1 : 1;
        // cos / sin of the phase differences and the floored magnitude of one frame.  Evaluated twice per frame -- once for the
        // time means, once for the features -- instead of keeping 2 x 6 x NITER values in registers: 120 -> ~64 registers, four
        // CTAs per SM instead of two for a kernel that is bound by the latency of its own dependent chains.
        auto frame = [&](int t, float& fm, float (&cr)[CM], float (&sr)[CM]) {
            float2 x[C];
            const bool valid = (st + t) < T_valid;           // zero-padded tail of the last segment, css.py:185-190
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = valid ? __ldg(Xf + (size_t)t * C + c) : make_float2(0.f, 0.f);
            // |X_0| and 1 / |X_0| from one MUFU.RSQ (2 ulp; the IEEE sqrt + division pair is ~16 instructions per channel)
            const float p0sq = x[0].x * x[0].x + x[0].y * x[0].y;
            const float inv0r = rsqrtf(p0sq);
            const float a0 = p0sq > 0.f ? p0sq * inv0r : 0.f;
            fm = fmaxf(a0, kEps32);
            if (edge_bin) {
                // literal replay: phase in {0, -3.1415925f} (torch angle of (re<0, tiny negative imag))
                const float p0 = (x[0].x < 0.f) ? -3.1415925f : 0.f;
#pragma unroll
                for (int m = 0; m < C - 1; ++m) {
                    const float pm = (x[m + 1].x < 0.f) ? -3.1415925f : 0.f;
                    const float d = pm - p0;
                    // cosf/sinf of {0, +-3.1415925f}: exact table of the correctly rounded values
                    cr[m] = (d == 0.f) ? 1.f : -1.f;
                    sr[m] = (d == 0.f) ? 0.f : (d > 0.f ? 1.509958e-07f : -1.509958e-07f);
                }
            } else {
                // unit phasor of X_0 (angle(0) == 0 -> (1, 0))
                const float inv0 = a0 > 0.f ? inv0r : 0.f;
                const float u0x = a0 > 0.f ? x[0].x * inv0 : 1.f, u0y = x[0].y * inv0;
#pragma unroll
                for (int m = 0; m < C - 1; ++m) {
                    const float am = x[m + 1].x * x[m + 1].x + x[m + 1].y * x[m + 1].y;      // |X_m|^2: only its sign test and rsqrt are used
                    const float invm = am > 0.f ? rsqrtf(am) : 0.f;
                    const float umx = am > 0.f ? x[m + 1].x * invm : 1.f, umy = x[m + 1].y * invm;
                    // u_m * conj(u_0) = exp(i (angle_m - angle_0))
                    cr[m] = umx * u0x + umy * u0y;
                    sr[m] = umy * u0x - umx * u0y;
                }
            }
        };
        float mag0[NITER];
        float s_mag = 0.f;
        float s_yr[CM], s_yi[CM];
#pragma unroll
        for (int m = 0; m < CM; ++m) { s_yr[m] = 0.f; s_yi[m] = 0.f; }
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
            const int t = it * 32 + lane;
            mag0[it] = 0.f;
            if (t < T) {
                float cr[CM], sr[CM];
                frame(t, mag0[it], cr, sr);
                s_mag += mag0[it];
#pragma unroll
                for (int m = 0; m < C - 1; ++m) { s_yr[m] += cr[m]; s_yi[m] += sr[m]; }
            }
        }
        const float mean = warp_sum(s_mag) / (float)T;
        float ssq = 0.f;
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
            const int t = it * 32 + lane;
            if (t < T) { const float d = mag0[it] - mean; ssq += d * d; }
        }
        const float var = warp_sum(ssq) / (float)(T - 1);       // unbiased, torch.std default
        const float rstd = 1.f / (sqrtf(var) + kEps32);
        float m_yr[CM], m_yi[CM];
#pragma unroll
        for (int m = 0; m < C - 1; ++m) {
            m_yr[m] = warp_sum(s_yr[m]) / (float)T;
            m_yi[m] = warp_sum(s_yi[m]) / (float)T;
        }
#pragma unroll 1
        for (int it = 0; it < NITER; ++it) {
            const int t = it * 32 + lane;
            if (t < T) {
                float fm, cr[CM], sr[CM];
                frame(t, fm, cr, sr);                             // the same arithmetic as in the first pass: identical values
                float* o = tile + ((size_t)t * C) * kFeatBinsPerCta + warp;
                o[0] = (fm - mean) * rstd;
#pragma unroll
                for (int m = 0; m < C - 1; ++m)
                    o[(m + 1) * kFeatBinsPerCta] = atan2_poly(sr[m] - m_yi[m], cr[m] - m_yr[m]);
            }
        }
    }
    __syncthreads();
    // the K padding columns [C * 257, ldf) of the segment's rows are zero-filled by the CTA of the last (one-bin) group
    if (blockIdx.x == gridDim.x - 1) {
        const int pad0 = C * kBins, npad = (int)ldf - pad0;
        const bool wide = !(feat_lo && (fmt == SPLIT_BF16 || fmt == SPLIT_F16));       // fp32 planes, else 16-bit planes
        for (int idx = threadIdx.x; idx < T * npad; idx += blockDim.x) {
            const int t = idx / npad, c = pad0 + idx - t * npad;
            const size_t o = ((size_t)seg * T + t) * ldf + c;
            if (wide) {
                feat[o] = 0.f;
                if (feat_lo) feat_lo[o] = 0.f;
            } else {
                reinterpret_cast<uint16_t*>(feat)[o] = 0;
                reinterpret_cast<uint16_t*>(feat_lo)[o] = 0;
            }
        }
    }
    // write-out: row (seg*T + t), column m*257 + f, runs of up to 8 consecutive bins
    const int nb = min(kFeatBinsPerCta, kBins - f0bin);
    if ((fmt == SPLIT_BF16 || fmt == SPLIT_F16) && feat_lo) {
        // 16-bit planes: warp m writes channel m; 8 lanes cover the run of 8 bins of one frame (the eight 2-byte stores of a
        // run leave the SM as one 16-byte request), 4 frames per instruction.  Bin, channel, bias and scale are fixed per
        // thread, the address advances by a constant: ~10 instructions per element instead of ~30 for the generic loop below.
        if (warp < C) {
            const int m = warp, b = lane & 7, col = m * kBins + f0bin + b;
            if (b < nb) {
                const float bi = in_bias ? __ldg(in_bias + col) : 0.f, sc = in_scale ? __ldg(in_scale + col) : 1.f;
                uint16_t* hp = reinterpret_cast<uint16_t*>(feat) + (size_t)seg * T * ldf + col;
                uint16_t* lp = reinterpret_cast<uint16_t*>(feat_lo) + (size_t)seg * T * ldf + col;
                const float* src = tile + m * kFeatBinsPerCta + b;
                for (int t = lane >> 3; t < T; t += 4) {
                    const float v = (src[t * (C * kFeatBinsPerCta)] + bi) * sc;                // conformer.py:297-299 (bi = 0, sc = 1: exact)
                    uint16_t h, l;
                    if (fmt == SPLIT_BF16) split_bf16(v, h, l); else split_f16(v, kF16ActScale, h, l);
                    hp[(size_t)t * ldf] = h;
                    lp[(size_t)t * ldf] = l;
                }
            }
        }
        return;
    }
    const int total = T * C * kFeatBinsPerCta;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int b = idx % kFeatBinsPerCta;
        const int m = (idx / kFeatBinsPerCta) % C;
        const int t = idx / (kFeatBinsPerCta * C);
        if (b >= nb) continue;
        const int col = m * kBins + f0bin + b;
        float v = tile[idx];
        if (in_bias) v = (v + __ldg(in_bias + col)) * __ldg(in_scale + col);      // conformer.py:297-299
        const size_t o = ((size_t)seg * T + t) * ldf + col;
        if (feat_lo) {
            split_store(fmt, feat, feat_lo, o, v);
        } else {
            feat[o] = v;
        }
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_css_features(const float* X, int64_t T_long, int64_t T_valid, int n_ch, int64_t seg_first, int n_seg,
                                int T, int hop, const float* in_bias, const float* in_scale, float* feat,
                                float* feat_lo, int64_t ldf, int split_fmt, void* stream) {
    NSF_REQUIRE(X && feat, "nsf_css_features: null pointer");
    NSF_REQUIRE(n_ch == 7 || n_ch == 1, "nsf_css_features: n_ch=%d (supported: 7, 1)", n_ch);
    NSF_REQUIRE(T >= 2 && T <= 32 * kFeatMaxIter, "nsf_css_features: T=%d not in [2,%d]", T, 32 * kFeatMaxIter);
    NSF_REQUIRE(ldf >= (int64_t)kBins * n_ch, "nsf_css_features: ldf too small");
    NSF_REQUIRE((in_bias == nullptr) == (in_scale == nullptr), "nsf_css_features: bias/scale must come together");
    NSF_REQUIRE(split_fmt >= SPLIT_TF32 && split_fmt <= SPLIT_F16 && (split_fmt == SPLIT_TF32 || feat_lo),
                "nsf_css_features: split_fmt=%d (16-bit formats need feat_lo)", split_fmt);
    // frames >= T_valid are never read (they are the zero padding of the last segment), so segments may extend past X
    NSF_REQUIRE(T_valid >= 0 && T_valid <= T_long && seg_first >= 0 && hop >= 1, "nsf_css_features: T_valid=%lld T_long=%lld seg_first=%lld hop=%d",
                (long long)T_valid, (long long)T_long, (long long)seg_first, hop);
    if (n_seg <= 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(ceil_div(kBins, kFeatBinsPerCta), n_seg);
    ProfScope prof(PROF_FEATURES, (double)n_seg * T * kBins * n_ch * (8.0 + (feat_lo ? (split_fmt == SPLIT_TF32 ? 8.0 : 4.0) : 4.0)), s);
    const size_t smem = (size_t)T * n_ch * kFeatBinsPerCta * sizeof(float);
#define NSF_FEAT_LAUNCH(CH, NI)                                                                                              \
    do {                                                                                                                     \
        NSF_CUDA(cudaFuncSetAttribute(css_features_kernel<CH, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        css_features_kernel<CH, NI><<<grid, kFeatBinsPerCta * 32, smem, s>>>(reinterpret_cast<const float2*>(X), T_long,       \
            T_valid, seg_first, T, hop, in_bias, in_scale, feat, feat_lo, ldf, split_fmt);                                   \
    } while (0)
    const bool small = T <= 32 * 6;
    if (n_ch == 7) { if (small) NSF_FEAT_LAUNCH(7, 6); else NSF_FEAT_LAUNCH(7, 8); }
    else           { if (small) NSF_FEAT_LAUNCH(1, 6); else NSF_FEAT_LAUNCH(1, 8); }
#undef NSF_FEAT_LAUNCH
    return check_launch("css_features_kernel");
}
