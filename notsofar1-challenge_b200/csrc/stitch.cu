// Stitching of the per-segment results: PIT costs, weighted overlap-add, activity gate, PCM16.
// All HBM-bound streaming kernels.
//
// Reference: css/css.py:254-312 (stage II/III of separate_and_stitch), css/training/losses.py:50-106
// (PitWrapper cost matrix), utils/numpy_utils.py:4-13 (dilate / erode), utils/audio_utils.py:44-49.
#include "common.cuh"

namespace nsf {

constexpr int kMaxSpk = 4;

// ------------------------------------------------------------------------------------------- PIT cost
// cost[i][a][b] = mean_{f, t < ov} loss(left_a[f][T-ov+t], right_b[f][t]); left = segment i-1, right = segment i.
template <int INPUT_KIND, int LOSS_KIND>
__global__ void __launch_bounds__(256)
pit_cost_kernel(const void* __restrict__ in, int n_ch_total, int n_spk, int n_bins, int T, int ov, float* __restrict__ cost, int seg_begin) {
    const int i = seg_begin + blockIdx.x;          // segment index, >= 1 does work
    __shared__ double red[8][kMaxSpk * kMaxSpk];
    double acc[kMaxSpk][kMaxSpk];
#pragma unroll
    for (int a = 0; a < kMaxSpk; ++a)
#pragma unroll
        for (int b = 0; b < kMaxSpk; ++b) acc[a][b] = 0.0;
    if (i >= 1) {
        const size_t seg_stride = (size_t)n_ch_total * n_bins * T;
        const size_t ch_stride = (size_t)n_bins * T;
        const int n = n_bins * ov;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int f = e / ov, t = e - f * ov;
            float l[kMaxSpk], r[kMaxSpk];
#pragma unroll
            for (int a = 0; a < kMaxSpk; ++a) {
                l[a] = 0.f; r[a] = 0.f;
                if (a < n_spk) {
                    const size_t il = (size_t)(i - 1) * seg_stride + a * ch_stride + (size_t)f * T + (T - ov + t);
                    const size_t ir = (size_t)i * seg_stride + a * ch_stride + (size_t)f * T + t;
                    if (INPUT_KIND == 0) {
                        l[a] = __ldg(reinterpret_cast<const float*>(in) + il);
                        r[a] = __ldg(reinterpret_cast<const float*>(in) + ir);
                    } else {
                        const float2 zl = __ldg(reinterpret_cast<const float2*>(in) + il);
                        const float2 zr = __ldg(reinterpret_cast<const float2*>(in) + ir);
                        l[a] = sqrtf(zl.x * zl.x + zl.y * zl.y);
                        r[a] = sqrtf(zr.x * zr.x + zr.y * zr.y);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < kMaxSpk; ++a)
#pragma unroll
                for (int b = 0; b < kMaxSpk; ++b) {
                    const float d = l[a] - r[b];
                    acc[a][b] += (double)(LOSS_KIND == 0 ? fabsf(d) : d * d);
                }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < kMaxSpk; ++a)
#pragma unroll
        for (int b = 0; b < kMaxSpk; ++b) {
            const double v = warp_sum(acc[a][b]);
            if (lane == 0) red[warp][a * kMaxSpk + b] = v;
        }
    __syncthreads();
    if (threadIdx.x < n_spk * n_spk) {
        const int a = threadIdx.x / n_spk, b = threadIdx.x % n_spk;
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][a * kMaxSpk + b];
        cost[((size_t)i * n_spk + a) * n_spk + b] = (i >= 1) ? (float)(v / ((double)n_bins * ov)) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------- mask WOLA
// mask_st[f][t][k] = (sum_i w_i[t - i hop] * m_i[perm_i[k]][f][t - i hop]) / wsum[t], i ascending as css.py:266-295.
__global__ void __launch_bounds__(256)
stitch_masks_kernel(const float* __restrict__ masks, int n_ch_total, const int32_t* __restrict__ perms,
                    const float* __restrict__ seg_w, const float* __restrict__ wsum, int n_seg, int n_spk, int n_bins,
                    int T, int hop, int64_t T_long, float* __restrict__ mask_st, int64_t t_begin, int64_t t_end) {
    const int64_t t = t_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (t >= t_end) return;
    int64_t i_min = (t - T + 1 + hop - 1) / hop;      // ceil((t-T+1)/hop) for positive numerator
    if (t - T + 1 <= 0) i_min = 0;
    int64_t i_max = t / hop;
    if (i_max > n_seg - 1) i_max = n_seg - 1;
    float acc[kMaxSpk] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t i = i_min; i <= i_max; ++i) {
        const int tl = (int)(t - i * hop);
        const float w = __ldg(seg_w + i * T + tl);
        const float* mi = masks + ((size_t)i * n_ch_total * n_bins + f) * T + tl;
#pragma unroll
        for (int k = 0; k < kMaxSpk; ++k)
            if (k < n_spk) {
                const int src = __ldg(perms + i * n_spk + k);
                const float m = __ldg(mi + (size_t)src * n_bins * T);
                acc[k] = __fadd_rn(acc[k], __fmul_rn(w, m));        // separate multiply and add, like the reference
            }
    }
    const float ws = __ldg(wsum + t);
    float* o = mask_st + ((size_t)f * T_long + t) * n_spk;
#pragma unroll
    for (int k = 0; k < kMaxSpk; ++k)
        if (k < n_spk) o[k] = acc[k] / ws;
}

// activity[t][k] = mean_f mask_st[f][t][k]   (css.py:304)
__global__ void __launch_bounds__(256)
activity_mean_kernel(const float* __restrict__ mask_st, int n_bins, int64_t n_tk, float* __restrict__ activity, int64_t j_begin,
                     int64_t j_end) {
    const int64_t j = j_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // t * n_spk + k
    if (j >= j_end) return;
    double s = 0.0;
    for (int f = 0; f < n_bins; ++f) s += (double)__ldg(mask_st + (size_t)f * n_tk + j);
    activity[j] = (float)(s / n_bins);
}

// ------------------------------------------------------------------------------------------- activity gate
// act_b = activity >= th ; tmp = dilate(act_b, dil) (zero padding)
__global__ void __launch_bounds__(256)
activity_threshold_dilate_kernel(const float* __restrict__ activity, int64_t T_long, int n_spk, float th, int dil,
                                 uint8_t* __restrict__ act_b, uint8_t* __restrict__ tmp, int64_t j_begin, int64_t j_end) {
    const int64_t j = j_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j_end) return;
    const int64_t t = j / n_spk;
    const int k = (int)(j - t * n_spk);
    act_b[j] = __ldg(activity + j) >= th;
    int64_t lo = t - dil, hi = t + dil;
    if (lo < 0) lo = 0;
    if (hi > T_long - 1) hi = T_long - 1;
    uint8_t any = 0;
    for (int64_t u = lo; u <= hi; ++u) any |= (uint8_t)(__ldg(activity + u * n_spk + k) >= th);
    tmp[j] = any;
}
// act_final = erode(tmp, ero) (one padding)
__global__ void __launch_bounds__(256)
activity_erode_kernel(const uint8_t* __restrict__ tmp, int64_t T_long, int n_spk, int ero, uint8_t* __restrict__ act_final,
                      int64_t j_begin, int64_t j_end) {
    const int64_t j = j_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j_end) return;
    const int64_t t = j / n_spk;
    const int k = (int)(j - t * n_spk);
    int64_t lo = t - ero, hi = t + ero;
    if (lo < 0) lo = 0;
    if (hi > T_long - 1) hi = T_long - 1;
    uint8_t all = 1;
    for (int64_t u = lo; u <= hi; ++u) all &= tmp[u * n_spk + k];
    act_final[j] = all;
}

// ------------------------------------------------------------------------------------------- STFT WOLA
// S_st[k][t][f] = gate[t][k] * (sum_i w_i Y_i[perm_i[k]][f][t - i hop]) / wsum[t]; tile-transposed through smem so
// that both the reads (t contiguous) and the writes (f contiguous) are coalesced.
__global__ void __launch_bounds__(256)
stitch_stft_kernel(const float2* __restrict__ Y, const int32_t* __restrict__ perms, const float* __restrict__ seg_w,
                   const float* __restrict__ wsum, const uint8_t* __restrict__ act_final, int n_seg, int n_spk,
                   int n_bins, int T, int hop, int64_t T_long, float2* __restrict__ S_st, int64_t t_begin, int64_t t_end) {
    __shared__ float2 tile[kMaxSpk][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const int64_t t0 = t_begin + (int64_t)blockIdx.x * 32;
    const int f0 = blockIdx.y * 32;
    {
        const int64_t t = t0 + tx;
        if (t < t_end) {
            int64_t i_min = (t - T + 1 <= 0) ? 0 : (t - T + 1 + hop - 1) / hop;
            int64_t i_max = t / hop;
            if (i_max > n_seg - 1) i_max = n_seg - 1;
            const float ws = __ldg(wsum + t);
            for (int fy = ty; fy < 32; fy += 8) {
                const int f = f0 + fy;
                if (f >= n_bins) break;
                float2 acc[kMaxSpk];
#pragma unroll
                for (int k = 0; k < kMaxSpk; ++k) acc[k] = make_float2(0.f, 0.f);
                for (int64_t i = i_min; i <= i_max; ++i) {
                    const int tl = (int)(t - i * hop);
                    const float w = __ldg(seg_w + i * T + tl);
#pragma unroll
                    for (int k = 0; k < kMaxSpk; ++k)
                        if (k < n_spk) {
                            const int src = __ldg(perms + i * n_spk + k);
                            const float2 y = __ldg(Y + (((size_t)i * n_spk + src) * n_bins + f) * T + tl);
                            acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(w, y.x));
                            acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(w, y.y));
                        }
                }
#pragma unroll
                for (int k = 0; k < kMaxSpk; ++k)
                    if (k < n_spk) {
                        const bool on = act_final ? (__ldg(act_final + t * n_spk + k) != 0) : true;
                        tile[k][fy][tx] = on ? make_float2(acc[k].x / ws, acc[k].y / ws) : make_float2(0.f, 0.f);
                    }
            }
        }
    }
    __syncthreads();
    {
        const int f = f0 + tx;
        if (f < n_bins) {
            for (int tyy = ty; tyy < 32; tyy += 8) {
                const int64_t t = t0 + tyy;
                if (t >= t_end) break;
#pragma unroll
                for (int k = 0; k < kMaxSpk; ++k)
                    if (k < n_spk) S_st[((size_t)k * T_long + t) * n_bins + f] = tile[k][tx][tyy];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- PCM16
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ wav, int64_t n, float* __restrict__ peak) {
    const float* w = wav + (size_t)blockIdx.y * n;
    float m = 0.f;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(w + j)));
    m = warp_max(m);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        atomicMax(reinterpret_cast<unsigned int*>(peak + blockIdx.y), __float_as_uint(m));   // non-negative floats order as uints
    }
}
__global__ void __launch_bounds__(256)
pcm16_kernel(const float* __restrict__ wav, int64_t n, const float* __restrict__ peak, int16_t* __restrict__ pcm) {
    const float* w = wav + (size_t)blockIdx.y * n;
    int16_t* q = pcm + (size_t)blockIdx.y * n;
    const float den = __fadd_rn(__ldg(peak + blockIdx.y), 1e-7f);            // np.max(np.abs(samps)) + 1e-7  (float32)
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float v = __fdiv_rn(__fmul_rn(__ldg(w + j), 0.99f), den);      // samps * 0.99 / (...)
        int r = __float2int_rn(__fmul_rn(v, 32767.f));                       // libsndfile f2s: lrintf(x * 0x7FFF)
        r = max(-32768, min(32767, r));
        q[j] = (int16_t)r;
    }
}

// mono PCM16 channel files -> the [N][C] float32 recording the STFT reads (css/helpers.py:40-65 + soundfile's int16 -> float32
// scaling x / 32768); tiled through shared memory so that both the per-channel reads and the interleaved writes coalesce
__global__ void __launch_bounds__(256)
pcm16_interleave_kernel(const int16_t* __restrict__ pcm, int64_t n, int C, float* __restrict__ out) {
    __shared__ float tile[8][257];
    const int64_t n0 = (int64_t)blockIdx.x * 256;
    for (int c = 0; c < C; ++c) {
        const int64_t j = n0 + threadIdx.x;
        tile[c][threadIdx.x] = j < n ? (float)pcm[(size_t)c * n + j] * (1.f / 32768.f) : 0.f;
    }
    __syncthreads();
    const int total = 256 * C;
    for (int e = threadIdx.x; e < total; e += 256) {
        const int t = e / C, c = e - t * C;
        if (n0 + t < n) out[(size_t)(n0 + t) * C + c] = tile[c][t];
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_pcm16_to_float_interleaved(const int16_t* pcm, int n_ch, int64_t n, float* out, void* stream) {
    NSF_REQUIRE(pcm && out, "nsf_pcm16_to_float_interleaved: null pointer");
    NSF_REQUIRE(n_ch >= 1 && n_ch <= 8 && n >= 0, "nsf_pcm16_to_float_interleaved: n_ch=%d n=%lld", n_ch, (long long)n);
    if (n == 0) return NSF_OK;
    pcm16_interleave_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(pcm, n, n_ch, out);
    return check_launch("pcm16_interleave_kernel");
}

extern "C" int nsf_pit_cost_range(const void* in, int input_kind, int loss_kind, int seg_begin, int seg_end, int n_ch_total, int n_spk,
                                  int n_bins, int T, int overlap, float* cost, void* stream) {
    NSF_REQUIRE(in && cost, "nsf_pit_cost: null pointer");
    NSF_REQUIRE(seg_begin >= 0, "nsf_pit_cost: seg_begin=%d", seg_begin);
    const int n_seg = seg_end;
    NSF_REQUIRE(n_spk >= 1 && n_spk <= kMaxSpk && n_ch_total >= n_spk, "nsf_pit_cost: n_spk=%d", n_spk);
    NSF_REQUIRE(overlap >= 1 && overlap <= T, "nsf_pit_cost: overlap=%d T=%d", overlap, T);
    NSF_REQUIRE((input_kind == 0 || input_kind == 1) && (loss_kind == 0 || loss_kind == 1), "nsf_pit_cost: bad kind");
    if (seg_end <= seg_begin) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PROF_PIT, (double)(n_seg - max(seg_begin, 1)) * n_bins * overlap * 2.0 * n_spk * (input_kind == 0 ? 4.0 : 8.0), s);
    if (input_kind == 0 && loss_kind == 0) pit_cost_kernel<0, 0><<<seg_end - seg_begin, 256, 0, s>>>(in, n_ch_total, n_spk, n_bins, T, overlap, cost, seg_begin);
    else if (input_kind == 0) pit_cost_kernel<0, 1><<<seg_end - seg_begin, 256, 0, s>>>(in, n_ch_total, n_spk, n_bins, T, overlap, cost, seg_begin);
    else if (loss_kind == 0) pit_cost_kernel<1, 0><<<seg_end - seg_begin, 256, 0, s>>>(in, n_ch_total, n_spk, n_bins, T, overlap, cost, seg_begin);
    else pit_cost_kernel<1, 1><<<seg_end - seg_begin, 256, 0, s>>>(in, n_ch_total, n_spk, n_bins, T, overlap, cost, seg_begin);
    return check_launch("pit_cost_kernel");
}

extern "C" int nsf_pit_cost(const void* in, int input_kind, int loss_kind, int n_seg, int n_ch_total, int n_spk,
                            int n_bins, int T, int overlap, float* cost, void* stream) {
    return nsf_pit_cost_range(in, input_kind, loss_kind, 0, n_seg, n_ch_total, n_spk, n_bins, T, overlap, cost, stream);
}

namespace {
// The stages behind the permutation chain on a frame range (each of them is frame-local on global indices, so a
// range call writes exactly what the whole-recording call writes there).
int stitch_masks_range(const float* masks, int n_ch_total, const int32_t* perms, const float* seg_w, const float* wsum, int n_seg,
                       int n_spk, int n_bins, int T, int hop, int64_t T_long, float* mask_st, float* activity, int64_t t0, int64_t t1,
                       cudaStream_t s) {
    if (t1 <= t0) return NSF_OK;
    dim3 grid((unsigned)ceil_div64(t1 - t0, 256), n_bins);
    ProfScope prof(PROF_STITCH, (double)(t1 - t0) * n_bins * n_spk * 4.0 * ((double)T / hop + 2.0), s);
    stitch_masks_kernel<<<grid, 256, 0, s>>>(masks, n_ch_total, perms, seg_w, wsum, n_seg, n_spk, n_bins, T, hop, T_long, mask_st, t0, t1);
    int rc = check_launch("stitch_masks_kernel");
    if (rc) return rc;
    activity_mean_kernel<<<(unsigned)ceil_div64((t1 - t0) * n_spk, 256), 256, 0, s>>>(mask_st, n_bins, T_long * n_spk, activity,
                                                                                      t0 * n_spk, t1 * n_spk);
    return check_launch("activity_mean_kernel");
}
int stitch_stft_range(const float* Y, const int32_t* perms, const float* seg_w, const float* wsum, const uint8_t* act_final, int n_seg,
                      int n_spk, int n_bins, int T, int hop, int64_t T_long, float* S_st, int64_t t0, int64_t t1, cudaStream_t s) {
    if (t1 <= t0) return NSF_OK;
    dim3 grid((unsigned)ceil_div64(t1 - t0, 32), ceil_div(n_bins, 32));
    ProfScope prof(PROF_STITCH, (double)(t1 - t0) * n_bins * n_spk * 8.0 * ((double)T / hop + 1.0), s);
    stitch_stft_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(Y), perms, seg_w, wsum, act_final, n_seg, n_spk, n_bins, T,
                                            hop, T_long, reinterpret_cast<float2*>(S_st), t0, t1);
    return check_launch("stitch_stft_kernel");
}
}  // namespace

extern "C" int nsf_stitch_masks(const float* masks, int n_ch_total, const int32_t* perms, const float* seg_w,
                                const float* wsum, int n_seg, int n_spk, int n_bins, int T, int hop, int64_t T_long,
                                float* mask_st, float* activity, void* stream) {
    NSF_REQUIRE(masks && perms && seg_w && wsum && mask_st && activity, "nsf_stitch_masks: null pointer");
    NSF_REQUIRE(n_spk >= 1 && n_spk <= kMaxSpk && n_ch_total >= n_spk && hop >= 1 && T >= 1, "nsf_stitch_masks: bad sizes");
    if (T_long <= 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    return stitch_masks_range(masks, n_ch_total, perms, seg_w, wsum, n_seg, n_spk, n_bins, T, hop, T_long, mask_st, activity, 0, T_long,
                              (cudaStream_t)stream);
}

extern "C" int nsf_activity(const float* activity, int64_t T_long, int n_spk, float th, int dil, int ero,
                            uint8_t* act_b, uint8_t* tmp, uint8_t* act_final, void* stream) {
    NSF_REQUIRE(activity && act_b && tmp && act_final, "nsf_activity: null pointer");
    NSF_REQUIRE(dil >= 0 && ero >= 0 && n_spk >= 1, "nsf_activity: bad sizes");
    if (T_long <= 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = T_long * n_spk;
    ProfScope prof(PROF_ACTIVITY, (double)n * 7.0, s);
    activity_threshold_dilate_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(activity, T_long, n_spk, th, dil, act_b, tmp, 0, n);
    int rc = check_launch("activity_threshold_dilate_kernel");
    if (rc) return rc;
    activity_erode_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(tmp, T_long, n_spk, ero, act_final, 0, n);
    return check_launch("activity_erode_kernel");
}

extern "C" int nsf_stitch_stft(const float* Y, const int32_t* perms, const float* seg_w, const float* wsum,
                               const uint8_t* act_final, int n_seg, int n_spk, int n_bins, int T, int hop,
                               int64_t T_long, float* S_st, void* stream) {
    NSF_REQUIRE(Y && perms && seg_w && wsum && S_st, "nsf_stitch_stft: null pointer");
    NSF_REQUIRE(n_spk >= 1 && n_spk <= kMaxSpk && hop >= 1 && T >= 1, "nsf_stitch_stft: bad sizes");
    if (T_long <= 0) return NSF_OK;
    return stitch_stft_range(Y, perms, seg_w, wsum, act_final, n_seg, n_spk, n_bins, T, hop, T_long, S_st, 0, T_long, (cudaStream_t)stream);
}

// Progressive tail: see include/nsf_b200.h.  Bounds of what is final once segments [0, sd) exist (hop = segment hop in frames):
//   stitched masks / activity  t < m = sd * hop          (every segment i <= t / hop is there)
//   act_b / dilated            t < d = m - dil
//   act_final / S_st           t < e = d - ero
//   waveform hops              j < h = floor(e / 8) * 8   (hop j reads frames j-1 and j; multiples of the iSTFT's CTA tile keep its
//                                                          frame pairing, hence its rounding, that of the one-shot launch)
// and everything once sd == n_seg.
extern "C" int nsf_stitch_progress(const float* masks, int n_ch_total, const float* Y, const int32_t* perms, const float* seg_w,
                                   const float* wsum, int n_seg, int seg_prev, int seg_done, int n_spk, int n_bins, int T, int hop,
                                   int64_t T_long, float th, int dil, int ero, float* mask_st, float* activity, uint8_t* act_b,
                                   uint8_t* tmp, uint8_t* act_final, float* S_st, float* wav, int64_t* hops_written, void* stream) {
    NSF_REQUIRE(masks && Y && perms && seg_w && wsum && mask_st && activity && act_b && tmp && act_final && S_st && wav && hops_written,
                "nsf_stitch_progress: null pointer");
    NSF_REQUIRE(n_spk >= 1 && n_spk <= kMaxSpk && n_ch_total >= n_spk && hop >= 1 && T >= 1 && dil >= 0 && ero >= 0 && T_long >= 1,
                "nsf_stitch_progress: bad sizes");
    NSF_REQUIRE(0 <= seg_prev && seg_prev <= seg_done && seg_done <= n_seg, "nsf_stitch_progress: segments %d -> %d of %d", seg_prev,
                seg_done, n_seg);
    struct Bounds { int64_t m, d, e, h; };
    auto bounds = [&](int sd) {
        Bounds b;
        if (sd >= n_seg) { b.m = b.d = b.e = T_long; b.h = T_long + 1; return b; }
        b.m = std::min<int64_t>(T_long, (int64_t)sd * hop);
        b.d = std::max<int64_t>(0, b.m - dil);
        b.e = std::max<int64_t>(0, b.d - ero);
        b.h = b.e / 8 * 8;
        return b;
    };
    const Bounds p = bounds(seg_prev), q = bounds(seg_done);
    hops_written[0] = p.h;
    hops_written[1] = q.h;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = stitch_masks_range(masks, n_ch_total, perms, seg_w, wsum, n_seg, n_spk, n_bins, T, hop, T_long, mask_st, activity, p.m, q.m, s);
    if (rc) return rc;
    if (q.d > p.d) {
        activity_threshold_dilate_kernel<<<(unsigned)ceil_div64((q.d - p.d) * n_spk, 256), 256, 0, s>>>(activity, T_long, n_spk, th, dil,
                                                                                                        act_b, tmp, p.d * n_spk, q.d * n_spk);
        if ((rc = check_launch("activity_threshold_dilate_kernel"))) return rc;
    }
    if (q.e > p.e) {
        activity_erode_kernel<<<(unsigned)ceil_div64((q.e - p.e) * n_spk, 256), 256, 0, s>>>(tmp, T_long, n_spk, ero, act_final, p.e * n_spk,
                                                                                             q.e * n_spk);
        if ((rc = check_launch("activity_erode_kernel"))) return rc;
    }
    rc = stitch_stft_range(Y, perms, seg_w, wsum, act_final, n_seg, n_spk, n_bins, T, hop, T_long, S_st, p.e, q.e, s);
    if (rc) return rc;
    return nsf_istft_range(S_st, n_spk, T_long, wav, p.h, q.h, stream);
}

extern "C" int nsf_peaknorm_pcm16(const float* wav, int n_streams, int64_t n, float* peak, int16_t* pcm, void* stream) {
    NSF_REQUIRE(wav && peak && pcm, "nsf_peaknorm_pcm16: null pointer");
    NSF_REQUIRE(n_streams >= 1 && n >= 0, "nsf_peaknorm_pcm16: bad sizes");
    if (n == 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    NSF_CUDA(cudaMemsetAsync(peak, 0, sizeof(float) * n_streams, s));
    ProfScope prof(PROF_PCM16, (double)n * n_streams * 10.0, s);
    int bx = (int)min((int64_t)148 * 4, ceil_div64(n, 256));
    dim3 grid(bx, n_streams);
    absmax_kernel<<<grid, 256, 0, s>>>(wav, n, peak);
    int rc = check_launch("absmax_kernel");
    if (rc) return rc;
    pcm16_kernel<<<grid, 256, 0, s>>>(wav, n, peak, pcm);
    return check_launch("pcm16_kernel");
}

// ------------------------------------------------------------------------------------------- word crops (CSS -> diarization hand-off)
// The reference re-reads the 16-bit WAVs the CSS stage wrote (read_wav(normalize=True): int16 / 32767,
// utils/audio_utils.py:10-34), cuts wavs[channel][start:end] per (word, scale) and zero-pads the batch
// (pad_sequence, diarization/word_based_diarization.py:98-104).  Here the PCM16 streams never leave HBM.
namespace nsf {
__global__ void __launch_bounds__(256)
gather_crops_kernel(const int16_t* __restrict__ pcm, int64_t n, const int32_t* __restrict__ stream_id,
                    const int64_t* __restrict__ start, const int32_t* __restrict__ len, int64_t max_len,
                    float* __restrict__ out) {
    const int crop = blockIdx.y;
    const int16_t* src = pcm + (size_t)stream_id[crop] * n + start[crop];
    const int l = len[crop];
    float* dst = out + (size_t)crop * max_len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < max_len; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = i < l ? (float)src[i] / 32767.f : 0.f;
}
}  // namespace nsf

extern "C" int nsf_gather_crops(const int16_t* pcm, int n_streams, int64_t n, const int32_t* stream_id, const int64_t* start,
                                const int32_t* len, int n_crops, int64_t max_len, float* out, void* stream) {
    NSF_REQUIRE(pcm && stream_id && start && len && out, "nsf_gather_crops: null pointer");
    NSF_REQUIRE(n_streams >= 1 && n >= 0 && n_crops >= 0 && max_len >= 0, "nsf_gather_crops: bad sizes");
    if (n_crops == 0 || max_len == 0) return NSF_OK;
    NSF_REQUIRE(n_crops <= 65535, "nsf_gather_crops: at most 65535 crops per call");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t gx = nsf::ceil_div64(max_len, 256 * 4);           // four samples per thread and pass, at most 64 CTAs per crop
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    dim3 grid((unsigned)gx, (unsigned)n_crops);
    nsf::gather_crops_kernel<<<grid, 256, 0, s>>>(pcm, n, stream_id, start, len, max_len, out);
    return nsf::check_launch("gather_crops_kernel");
}
