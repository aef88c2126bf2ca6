// Mask-only separation and segment power normalisation (HBM-bound streaming kernels).
//
// Reference: css/css.py:205-247.  Without MVDR (single-channel input, or mc_mvdr = False) the separated segment is
// the reference channel times the floored mask (:218-227); with normalize_segment_power every segment is rescaled so
// that the power of the sum of its streams matches the power of the mixture's reference channel (:233-247).
#include "common.cuh"

namespace nsf {

// Y[seg][s][f][t] = X[f][st + t][0] * max(mask[seg][s][f][t], floor); frames >= T_valid read as zeros (css.py:185-190)
__global__ void __launch_bounds__(256)
mask_apply_kernel(const float* __restrict__ masks, int n_spk, int n_masks, const float2* __restrict__ X, int64_t T_long,
                  int64_t T_valid, int n_ch, int64_t seg_first, int T, int hop, int n_bins, float floor_, float2* __restrict__ Y) {
    const int seg = blockIdx.y;
    const int64_t st = (seg_first + seg) * (int64_t)hop;
    const int n = n_bins * T;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int f = e / T, t = e - f * T;
        float2 x = make_float2(0.f, 0.f);
        if (st + t < T_valid) x = __ldg(X + ((size_t)f * T_long + st + t) * n_ch);
        for (int s = 0; s < n_spk; ++s) {
            const float m = fmaxf(__ldg(masks + (((size_t)seg * n_masks + s) * n_bins + f) * T + t), floor_);
            Y[(((size_t)seg * n_spk + s) * n_bins + f) * T + t] = make_float2(x.x * m, x.y * m);
        }
    }
}

// one CTA per segment: ratio[seg] = sqrt(mean |X_ref|^2) / sqrt(mean |sum_s Y_s|^2) over the segment's t_seg real frames
__global__ void __launch_bounds__(512)
segment_power_kernel(const float2* __restrict__ Y, int n_spk, const float2* __restrict__ X, int64_t T_long, int64_t T_valid,
                     int n_ch, int64_t seg_first, int T, int hop, int n_bins, int64_t mix_frames, float* __restrict__ ratio) {
    const int seg = blockIdx.x;
    const int64_t st = (seg_first + seg) * (int64_t)hop;
    const int t_seg = (int)min((int64_t)T, mix_frames - st);
    double mix = 0.0, sep = 0.0;
    const int n = n_bins * t_seg;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int f = e / t_seg, t = e - f * t_seg;
        if (st + t < T_valid) {
            const float2 x = __ldg(X + ((size_t)f * T_long + st + t) * n_ch);
            mix += (double)(x.x * x.x + x.y * x.y);
        }
        float sr = 0.f, si = 0.f;
        for (int s = 0; s < n_spk; ++s) {
            const float2 y = Y[(((size_t)seg * n_spk + s) * n_bins + f) * T + t];
            sr += y.x; si += y.y;
        }
        sep += (double)(sr * sr + si * si);
    }
    __shared__ double red[2][16];
    mix = warp_sum(mix); sep = warp_sum(sep);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = mix; red[1][warp] = sep; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        const float mix_e = sqrtf((float)(a / (double)n)), sep_e = sqrtf((float)(b / (double)n));
        ratio[seg] = mix_e / sep_e;
    }
}

__global__ void __launch_bounds__(256)
segment_scale_kernel(float2* __restrict__ Y, int64_t per_seg, const float* __restrict__ ratio) {
    const float r = ratio[blockIdx.y];
    float2* y = Y + (size_t)blockIdx.y * per_seg;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per_seg; e += (int64_t)gridDim.x * blockDim.x) {
        float2 v = y[e];
        y[e] = make_float2(r * v.x, r * v.y);
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_mask_apply(const float* masks, int n_spk, int n_noise, const float* X, int64_t T_long, int64_t T_valid, int n_ch,
                              int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor, float* Y, void* stream) {
    NSF_REQUIRE(masks && X && Y, "nsf_mask_apply: null pointer");
    NSF_REQUIRE(n_spk >= 1 && n_noise >= 0 && n_ch >= 1 && T >= 1 && n_bins >= 1 && T_valid <= T_long, "nsf_mask_apply: bad sizes");
    if (n_seg <= 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PROF_MVDR, (double)n_seg * n_bins * T * (8.0 + n_spk * 12.0), s);
    dim3 grid((unsigned)min(ceil_div(n_bins * T, 256), 48), (unsigned)n_seg);
    mask_apply_kernel<<<grid, 256, 0, s>>>(masks, n_spk, n_spk + n_noise, reinterpret_cast<const float2*>(X), T_long, T_valid, n_ch,
                                           seg_first, T, hop, n_bins, mask_floor, reinterpret_cast<float2*>(Y));
    return check_launch("mask_apply_kernel");
}

extern "C" int nsf_segment_power_norm(float* Y, int n_spk, const float* X, int64_t T_long, int64_t T_valid, int n_ch, int64_t seg_first,
                                      int n_seg, int T, int hop, int n_bins, int64_t mix_frames, float* ratio, void* stream) {
    NSF_REQUIRE(Y && X && ratio, "nsf_segment_power_norm: null pointer");
    NSF_REQUIRE(n_spk >= 1 && n_ch >= 1 && T >= 1 && n_bins >= 1 && T_valid <= T_long, "nsf_segment_power_norm: bad sizes");
    if (n_seg <= 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PROF_MVDR, (double)n_seg * n_bins * T * (8.0 + n_spk * 24.0), s);
    segment_power_kernel<<<n_seg, 512, 0, s>>>(reinterpret_cast<const float2*>(Y), n_spk, reinterpret_cast<const float2*>(X), T_long,
                                               T_valid, n_ch, seg_first, T, hop, n_bins, mix_frames, ratio);
    int rc = check_launch("segment_power_kernel");
    if (rc) return rc;
    const int64_t per_seg = (int64_t)n_spk * n_bins * T;
    dim3 grid((unsigned)min((int64_t)ceil_div64(per_seg, 256), (int64_t)64), (unsigned)n_seg);
    segment_scale_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float2*>(Y), per_seg, ratio);
    return check_launch("segment_scale_kernel");
}
