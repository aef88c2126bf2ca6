// TitaNet speaker-embedding forward (row a16: diarization/word_based_diarization.py:26,105 call NeMo's
// EncDecSpeakerLabelModel "titanet_large", third-party and unpinned; the published architecture is restated in
// oracle/titanet_oracle.py, which is what this file is checked against -- parity unpinned w.r.t. NeMo itself).
//
//   nsf_titanet_features   crops [n][max_len] f32 + lengths -> per-feature normalised log-mel, time-major bf16 head / remainder
//                          planes [n][t_pad][80] (pre-emphasis, centred reflect-padded 512-point DFT of the 400-sample hann
//                          frame, 80 slaney mel bands, log(. + 2^-24), mean / unbiased std over the valid frames)
//   nsf_titanet_forward    5 Jasper blocks: depthwise conv over time (tn_dwconv_kernel, masked input) -> pointwise conv as a
//                          tcgen05 GEMM with the BatchNorm folded into weights and bias (2xBF16 split engine of gemm_tc.cu:
//                          three kind::f16 MMAs per product, fp32 accumulation; ReLU + operand split in the epilogue) ->
//                          squeeze-excite (masked mean, two small dense layers, gate) fused with the residual add, ReLU, length
//                          mask and operand split (tn_se_apply_kernel); attentive statistics pooling (context statistics,
//                          two GEMMs, masked softmax over time, weighted mean / std) and the 192-d embedding layer.
// Activations are time-major [n * t_pad][C]: every 1x1 convolution is one GEMM over all crops of the batch; frames beyond a
// crop's length are carried along as finite values and masked wherever the reference masks (conv inputs, pooling).
#include "gemm_common.cuh"
#include "fft512.cuh"

namespace nsf {

constexpr int kTnNfft = 512, kTnWin = 400, kTnHop = 160, kTnBins = 257;
constexpr int kTnMaxBlocks = 8;

// ------------------------------------------------------------------------------------------- front end
// One CTA per (crop, 8 frames): four groups of 64 threads, each transforming two real frames packed as the real and
// imaginary part of one 512-point complex FFT (fft512.cuh, shared-memory radix 8x8x8).  A frame is the pre-emphasised,
// reflect-padded signal under the 400-sample symmetric hann window centred in the 512-point transform; its power spectrum
// goes through the mel filterbank and the log.  lm [n][t_pad][n_mels] f32 (frames >= n_frames are not written).
constexpr int kTnFramesPerCta = 8;
struct TnLogmelSmem {
    float2 tw[kTnNfft];
    float scratch[4][kFftScratchFloats];
    float pw[kTnFramesPerCta][kTnBins + 3];
};

__global__ void __launch_bounds__(256)
tn_logmel_kernel(const float* __restrict__ crops, const int* __restrict__ lengths, int64_t max_len, int t_pad,
                 const float* __restrict__ filters, int n_mels, float* __restrict__ lm,
                 int* __restrict__ n_frames) {
    __shared__ __align__(16) TnLogmelSmem sm;
    const int b = blockIdx.y, t0 = blockIdx.x * kTnFramesPerCta;
    const int len = lengths[b];
    const int nf = len > 0 ? len / kTnHop + 1 : 0;                     // FilterbankFeatures.get_seq_len, centred frames
    if (blockIdx.x == 0 && threadIdx.x == 0) n_frames[b] = nf;
    if (t0 >= nf) return;
    const float* x = crops + (size_t)b * max_len;
    constexpr int off = (kTnNfft - kTnWin) / 2;
    for (int i = threadIdx.x; i < kTnNfft; i += blockDim.x) {
        float sn, cs;
        sincospif((float)i / 256.f, &sn, &cs);                          // (cos, sin)(2 pi i / 512)
        sm.tw[i] = make_float2(cs, sn);
    }
    __syncthreads();
    const int group = threadIdx.x >> 6, lane64 = threadIdx.x & 63;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory"); };
    auto sample = [&](int t, int n) -> float {
        if (n < off || n >= off + kTnWin) return 0.f;
        int j = t * kTnHop + n - kTnNfft / 2;                           // centred frame, reflect padding (torch.stft center=True)
        if (j < 0) j = -j;
        if (j >= len) j = 2 * (len - 1) - j;
        j = max(0, min(j, len - 1));
        const float v = j > 0 ? x[j] - 0.97f * x[j - 1] : x[0];        // pre-emphasis happens before the padding
        return v * (0.5f - 0.5f * cospif(2.f * (float)(n - off) / (float)(kTnWin - 1)));   // hann_window(400, periodic=False)
    };
    const int ta = t0 + 2 * group, tb = ta + 1;
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = make_float2(sample(ta, 64 * a + lane64), sample(tb, 64 * a + lane64));
    float* scratch = sm.scratch[group];
    fft512_group<-1>(v, lane64, scratch, sm.tw, group_sync);
    float* zre = scratch;
    float* zim = scratch + 8 * 72;
    {
        const int k0 = lane64 >> 3, k1 = lane64 & 7;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) { zre[k0 + 8 * k1 + 64 * k2] = v[k2].x; zim[k0 + 8 * k1 + 64 * k2] = v[k2].y; }
    }
    group_sync();
    for (int k = lane64; k < kTnBins; k += 64) {      // A[k] = (Z[k] + conj(Z[-k])) / 2,  B[k] = (Z[k] - conj(Z[-k])) / (2i)
        const int kn = (kTnNfft - k) & (kTnNfft - 1);
        const float ar = zre[k] + zre[kn], ai = zim[k] - zim[kn], br = zim[k] + zim[kn], bi = zre[kn] - zre[k];
        sm.pw[2 * group][k] = 0.25f * (ar * ar + ai * ai);
        sm.pw[2 * group + 1][k] = 0.25f * (br * br + bi * bi);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_mels * kTnFramesPerCta; e += blockDim.x) {
        const int m = e % n_mels, fr = e / n_mels;
        if (t0 + fr >= nf) continue;
        float acc = 0.f;
        for (int k = 0; k < kTnBins; ++k) acc = fmaf(__ldg(filters + (size_t)k * n_mels + m), sm.pw[fr][k], acc);   // filters [257][n_mels]: coalesced over m
        lm[((size_t)b * t_pad + t0 + fr) * n_mels + m] = logf(acc + 5.9604644775390625e-08f);     // log(x + 2^-24)
    }
}

// per (crop, feature): mean and unbiased std over the valid frames (+1e-5), normalise, zero beyond, operand split
__global__ void tn_featnorm_kernel(const float* __restrict__ lm, const int* __restrict__ n_frames, int t_pad, int n_mels,
                                   float* __restrict__ out_hi, float* __restrict__ out_lo) {
    const int b = blockIdx.x, m = threadIdx.x;
    if (m >= n_mels) return;
    const int nf = n_frames[b];
    const float* p = lm + (size_t)b * t_pad * n_mels + m;
    float s = 0.f;
    for (int t = 0; t < nf; ++t) s += p[(size_t)t * n_mels];
    const float mean = nf > 0 ? s / nf : 0.f;
    float q = 0.f;
    for (int t = 0; t < nf; ++t) { const float d = p[(size_t)t * n_mels] - mean; q = fmaf(d, d, q); }
    const float inv = 1.f / (sqrtf(q / fmaxf((float)(nf - 1), 1.f)) + 1e-5f);
    for (int t = 0; t < t_pad; ++t) {
        const float v = t < nf ? (p[(size_t)t * n_mels] - mean) * inv : 0.f;
        split_store(SPLIT_BF16, out_hi, out_lo, ((size_t)b * t_pad + t) * n_mels + m, v);
    }
}

// ------------------------------------------------------------------------------------------- encoder pieces
__device__ __forceinline__ void bf16x8_to_f32(const uint4& h, const uint4& l, float (&v)[8]) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        v[2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
        v[2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
    }
}

__device__ __forceinline__ void f16x8_to_f32(const uint4& h, float (&v)[8]) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
        v[2 * e] = f.x;
        v[2 * e + 1] = f.y;
    }
}
// bf16 head / remainder planes -> one fp16 plane (the features enter the fp16 engine through this)
__global__ void __launch_bounds__(256)
tn_pairs_to_f16_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo, __half* __restrict__ out, int64_t n8) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n8) return;
    float v[8];
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(in_hi + e * 8), *reinterpret_cast<const uint4*>(in_lo + e * 8), v);
    split_store8(SPLIT_F16_1, reinterpret_cast<float*>(out), nullptr, (size_t)e * 8, v);
}

// depthwise convolution over time ('same' padding, cross-correlation), input masked beyond the crop's length (MaskedConv1d).
// CTA = (crop, 64 frames, 64 channels), 128 threads: the 64 + k - 1 input rows are staged once in shared memory as fp32 (each
// input element crosses L2 1.0 - 1.5 times instead of k times; two planes of 4-channel halves so that the 16-byte reads of a
// quarter warp are contiguous), thread = (4 frames, 8 channels) so that a tap's weights are fetched once per four outputs.
// With k = 15 the kernel issues 1.9 FMA per HBM byte: bound by instruction issue unless the overhead per FMA stays small.
// in / out: bf16 head + remainder planes [n * t_pad][C]; w [k][C] f32.
constexpr int kDwFrames = 64, kDwCh = 64, kDwMaxK = 31, kDwRowsMax = kDwFrames + kDwMaxK - 1;
__global__ void __launch_bounds__(128)
tn_dwconv_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo, const int* __restrict__ n_frames, int t_pad,
                 int C, int k, const float* __restrict__ w, float* __restrict__ out_hi, float* __restrict__ out_lo, int fmt) {
    __shared__ __align__(16) float4 tile[2][kDwRowsMax][8];
    const int b = blockIdx.z, t0 = blockIdx.y * kDwFrames, c0 = blockIdx.x * kDwCh;
    const int nf = n_frames[b];
    const int half = k >> 1, rows = kDwFrames + k - 1;
    const int nvec = min(kDwCh, C - c0) >> 3;                         // 8-channel vectors in this CTA's channel slice
    for (int e = threadIdx.x; e < rows * 8; e += blockDim.x) {
        const int r = e >> 3, v8 = e & 7;
        const int tt = t0 + r - half;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (v8 < nvec && tt >= 0 && tt < nf) {
            const size_t o = ((size_t)b * t_pad + tt) * C + c0 + v8 * 8;
            if (fmt == SPLIT_F16_1) f16x8_to_f32(*reinterpret_cast<const uint4*>(in_hi + o), v);
            else bf16x8_to_f32(*reinterpret_cast<const uint4*>(in_hi + o), *reinterpret_cast<const uint4*>(in_lo + o), v);
        }
        tile[0][r][v8] = make_float4(v[0], v[1], v[2], v[3]);
        tile[1][r][v8] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    const int fg = threadIdx.x >> 3, v8 = threadIdx.x & 7;
    if (v8 >= nvec) return;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
    for (int j = 0; j < k; ++j) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * C + c0 + v8 * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * C + c0 + v8 * 8 + 4));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 x0 = tile[0][fg + 16 * i + j][v8], x1 = tile[1][fg + 16 * i + j][v8];
            acc[i][0] = fmaf(w0.x, x0.x, acc[i][0]); acc[i][1] = fmaf(w0.y, x0.y, acc[i][1]);
            acc[i][2] = fmaf(w0.z, x0.z, acc[i][2]); acc[i][3] = fmaf(w0.w, x0.w, acc[i][3]);
            acc[i][4] = fmaf(w1.x, x1.x, acc[i][4]); acc[i][5] = fmaf(w1.y, x1.y, acc[i][5]);
            acc[i][6] = fmaf(w1.z, x1.z, acc[i][6]); acc[i][7] = fmaf(w1.w, x1.w, acc[i][7]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + fg + 16 * i;
        if (t < t_pad) split_store8(fmt, out_hi, out_lo, ((size_t)b * t_pad + t) * C + c0 + v8 * 8, acc[i]);
    }
}

// masked statistics over time per (crop, channel): mean (and optionally std = sqrt(max(mean((x - mean)^2), 1e-10))) of
// x [n * t_pad][C] f32 over the valid frames.  CTA = 32 channels x 8 time slices.
__global__ void __launch_bounds__(256)
tn_masked_stats_kernel(const float* __restrict__ x, const int* __restrict__ n_frames, int t_pad, int C, float* __restrict__ mean_out,
                       float* __restrict__ std_out, int out_pitch) {
    __shared__ float red[8][33];
    const int b = blockIdx.y, cl = threadIdx.x & 31, ts = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int nf = n_frames[b];
    const bool ok = c < C;
    const float* p = x + (size_t)b * t_pad * C + c;
    float s = 0.f;
    if (ok) for (int t = ts; t < nf; t += 8) s += p[(size_t)t * C];
    red[ts][cl] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mean += red[i][cl];
    mean = nf > 0 ? mean / nf : 0.f;
    if (ts == 0 && ok) mean_out[(size_t)b * out_pitch + c] = mean;
    if (!std_out) return;
    __syncthreads();
    float q = 0.f;
    if (ok) for (int t = ts; t < nf; t += 8) { const float d = p[(size_t)t * C] - mean; q = fmaf(d, d, q); }
    red[ts][cl] = q;
    __syncthreads();
    if (ts == 0 && ok) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += red[i][cl];
        std_out[(size_t)b * out_pitch + c] = sqrtf(fmaxf(nf > 0 ? v / nf : 0.f, 1e-10f));
    }
}

// small dense layer, one warp per output: out[b][j] = act(sum_k W[j][k] in[b][k] + bias[j]);  act 0 none, 1 relu, 2 sigmoid
__global__ void __launch_bounds__(256)
tn_fc_kernel(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ bias, int N, int n, int act,
             float* __restrict__ out) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)n * N) return;
    const int b = (int)(wid / N), j = (int)(wid - (int64_t)b * N);
    const float* a = in + (size_t)b * K;
    const float* w = W + (size_t)j * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(__ldg(w + k), a[k], s);
    s = warp_sum(s);
    if (lane == 0) {
        s += bias ? bias[j] : 0.f;
        out[(size_t)b * N + j] = act == 1 ? fmaxf(s, 0.f) : act == 2 ? 1.f / (1.f + expf(-s)) : s;
    }
}

// the same for short rows (K <= 512: the squeeze-excite expansion), one thread per output on transposed weights Wt [K][N]
__global__ void __launch_bounds__(256)
tn_fc_t_kernel(const float* __restrict__ in, int K, const float* __restrict__ Wt, int N, int n, int act, float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * N) return;
    const int b = (int)(e / N), j = (int)(e - (int64_t)b * N);
    const float* a = in + (size_t)b * K;
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(__ldg(Wt + (size_t)k * N + j), a[k], s);
    out[e] = act == 1 ? fmaxf(s, 0.f) : act == 2 ? 1.f / (1.f + expf(-s)) : s;
}

// y = relu(pre * gate[crop][c] + res), zero beyond the crop's length; bf16 planes for the next block, optionally fp32 in place
__global__ void __launch_bounds__(256)
tn_se_apply_kernel(float* __restrict__ pre, const float* __restrict__ gate, const float* __restrict__ res, const int* __restrict__ n_frames,
                   int t_pad, int C, float* __restrict__ out_hi, float* __restrict__ out_lo, int write_f32, int64_t total, int fmt) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c8 = C >> 3;
    const int cg = (int)(e % c8);
    const int64_t row = e / c8;
    const int b = (int)(row / t_pad), t = (int)(row - (int64_t)b * t_pad);
    const size_t o = (size_t)row * C + (size_t)cg * 8;
    float v[8];
    if (t < n_frames[b]) {
        const float4 p0 = *reinterpret_cast<const float4*>(pre + o), p1 = *reinterpret_cast<const float4*>(pre + o + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(gate + (size_t)b * C + cg * 8), g1 = *reinterpret_cast<const float4*>(gate + (size_t)b * C + cg * 8 + 4);
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (res) { r0 = *reinterpret_cast<const float4*>(res + o); r1 = *reinterpret_cast<const float4*>(res + o + 4); }
        v[0] = fmaxf(fmaf(p0.x, g0.x, r0.x), 0.f); v[1] = fmaxf(fmaf(p0.y, g0.y, r0.y), 0.f);
        v[2] = fmaxf(fmaf(p0.z, g0.z, r0.z), 0.f); v[3] = fmaxf(fmaf(p0.w, g0.w, r0.w), 0.f);
        v[4] = fmaxf(fmaf(p1.x, g1.x, r1.x), 0.f); v[5] = fmaxf(fmaf(p1.y, g1.y, r1.y), 0.f);
        v[6] = fmaxf(fmaf(p1.z, g1.z, r1.z), 0.f); v[7] = fmaxf(fmaf(p1.w, g1.w, r1.w), 0.f);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    split_store8(fmt, out_hi, out_lo, o, v);
    if (write_f32) {
        *reinterpret_cast<float4*>(pre + o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(pre + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// ------------------------------------------------------------------------------------------- decoder pieces
// h2 = tanh(a * relu(h1 + bvec[crop]) + c)  (TDNN: conv -> ReLU -> BatchNorm, then Tanh) -> bf16 planes [M][A]
__global__ void tn_att_act_kernel(const float* __restrict__ h1, const float* __restrict__ bvec, const float* __restrict__ a,
                                  const float* __restrict__ c, int t_pad, int A, float* __restrict__ out_hi, float* __restrict__ out_lo,
                                  int64_t total, int fmt) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int j = (int)(e % A);
    const int b = (int)((e / A) / t_pad);
    const float v = tanhf(fmaf(a[j], fmaxf(h1[e] + bvec[(size_t)b * A + j], 0.f), c[j]));
    split_store(fmt, out_hi, out_lo, (size_t)e, v);
}

// attentive statistics per (crop, channel): alpha = softmax over the valid frames of logit, mu = sum alpha x,
// sg = sqrt(max(sum alpha (x - mu)^2, 1e-10)); pool [n][2C] = [mu | sg].  CTA = 32 channels x 8 time slices.
__global__ void __launch_bounds__(256)
tn_att_pool_kernel(const float* __restrict__ x, const float* __restrict__ logit, const int* __restrict__ n_frames, int t_pad, int C,
                   float* __restrict__ pool) {
    __shared__ float red[8][33];
    const int b = blockIdx.y, cl = threadIdx.x & 31, ts = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int nf = n_frames[b];
    const bool ok = c < C;
    const size_t base = (size_t)b * t_pad * C + c;
    auto reduce = [&](float v, bool is_max) {
        __syncthreads();
        red[ts][cl] = v;
        __syncthreads();
        float r = red[0][cl];
#pragma unroll
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i][cl]) : r + red[i][cl];
        return r;
    };
    float mx = -INFINITY;
    if (ok) for (int t = ts; t < nf; t += 8) mx = fmaxf(mx, logit[base + (size_t)t * C]);
    mx = reduce(mx, true);
    float den = 0.f, num = 0.f;
    if (ok) for (int t = ts; t < nf; t += 8) {
        const float w = expf(logit[base + (size_t)t * C] - mx);
        den += w;
        num = fmaf(w, x[base + (size_t)t * C], num);
    }
    den = reduce(den, false);
    num = reduce(num, false);
    const float mu = den > 0.f ? num / den : 0.f;
    float var = 0.f;
    if (ok) for (int t = ts; t < nf; t += 8) {
        const float w = expf(logit[base + (size_t)t * C] - mx);
        const float d = x[base + (size_t)t * C] - mu;
        var = fmaf(w, d * d, var);
    }
    var = reduce(var, false);
    if (ts == 0 && ok) {
        pool[(size_t)b * 2 * C + c] = mu;
        pool[(size_t)b * 2 * C + C + c] = sqrtf(fmaxf(den > 0.f ? var / den : 0.f, 1e-10f));
    }
}

// cosine affinity of one scale (getCosAffinityMatrix [upstream]): rows normalised by (norm + 3.5e-4)
__global__ void tn_rownorm_kernel(const float* __restrict__ emb, int64_t row_pitch, int D, int n, float* __restrict__ out) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= n) return;
    const float* e = emb + wid * row_pitch;
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s = fmaf(e[k], e[k], s);
    s = warp_sum(s);
    const float inv = 1.f / (sqrtf(s) + 3.5e-4f);
    for (int k = lane; k < D; k += 32) out[wid * D + k] = e[k] * inv;
}
// sim[i][j] = en[i] . en[j], unit diagonal; per-CTA min / max folded into mm[0..1] (order-encoded uints)
__device__ __forceinline__ unsigned tn_order(float v) { const unsigned u = __float_as_uint(v); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float tn_unorder(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__global__ void __launch_bounds__(256)
tn_cossim_kernel(const float* __restrict__ en, int D, int n, float* __restrict__ sim, unsigned* __restrict__ mm) {
    __shared__ float a[16][65], bt[16][65];
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
    float s = 0.f;
    for (int k0 = 0; k0 < D; k0 += 64) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int r = e >> 6, k = e & 63;
            a[r][k] = (i0 + r < n && k0 + k < D) ? en[(size_t)(i0 + r) * D + k0 + k] : 0.f;
            bt[r][k] = (j0 + r < n && k0 + k < D) ? en[(size_t)(j0 + r) * D + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 16
        for (int k = 0; k < 64; ++k) s = fmaf(a[ti][k], bt[tj][k], s);
        __syncthreads();
    }
    const int i = i0 + ti, j = j0 + tj;
    float lo = INFINITY, hi = -INFINITY;
    if (i < n && j < n) {
        if (i == j) s = 1.f;                                  // res.fill_diagonal_(1)
        sim[(size_t)i * n + j] = s;
        lo = hi = s;
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(mm, tn_order(lo)); atomicMax(mm + 1, tn_order(hi)); }
}
// acc += (sim - min) / (max - min) / n_scales   (ScalerMinMax, then the mean over scales of word_based_diarization.py:174-177)
__global__ void tn_affinity_accum_kernel(const float* __restrict__ sim, const unsigned* __restrict__ mm, float scale, int64_t total,
                                         float* __restrict__ acc) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const float lo = tn_unorder(mm[0]), hi = tn_unorder(mm[1]);
    acc[e] += (sim[e] - lo) / (hi - lo) * scale;
}

}  // namespace nsf

using namespace nsf;

struct nsf_titanet {
    nsf_titanet_dims dims;
    const float* blob;
    int64_t* offsets;
    int n_offsets;
};

namespace {

struct TnPlan {                      // offset indices, in the order notsofar_b200/titanet.py::pack_titanet emits them
    int dw[kTnMaxBlocks][8], pw_hi[kTnMaxBlocks][8], pw_lo[kTnMaxBlocks][8], pw_b[kTnMaxBlocks][8];
    int fc0[kTnMaxBlocks], fc2[kTnMaxBlocks], res_hi[kTnMaxBlocks], res_lo[kTnMaxBlocks], res_b[kTnMaxBlocks];
    int w1x_hi, w1x_lo, w1ms, b1, a1, c1, w2_hi, w2_lo, b2, we, be, num;
    int c_max, c_last;
};

bool tn_dims_ok(const nsf_titanet_dims& d) {
    if (d.precision != 0 && d.precision != 1) return false;
    if (d.n_blocks < 1 || d.n_blocks > kTnMaxBlocks || d.feat_in < 8 || d.feat_in % 8 || d.att_ch < 8 || d.att_ch % 8 || d.emb < 1) return false;
    for (int b = 0; b < d.n_blocks; ++b)
        if (d.filters[b] < 64 || d.filters[b] % 64 || d.repeat[b] < 1 || d.repeat[b] > 8 || d.kernel[b] < 1 || d.kernel[b] > kDwMaxK || !(d.kernel[b] & 1)) return false;
    return true;
}

TnPlan tn_plan(const nsf_titanet_dims& d) {
    TnPlan p = {};
    int o = 0;
    p.c_max = d.feat_in;
    for (int b = 0; b < d.n_blocks; ++b) {
        for (int r = 0; r < d.repeat[b]; ++r) { p.dw[b][r] = o++; p.pw_hi[b][r] = o++; p.pw_lo[b][r] = o++; p.pw_b[b][r] = o++; }
        p.fc0[b] = o++; p.fc2[b] = o++;
        if (d.residual[b]) { p.res_hi[b] = o++; p.res_lo[b] = o++; p.res_b[b] = o++; }
        p.c_max = d.filters[b] > p.c_max ? d.filters[b] : p.c_max;
    }
    p.c_last = d.filters[d.n_blocks - 1];
    p.w1x_hi = o++; p.w1x_lo = o++; p.w1ms = o++; p.b1 = o++; p.a1 = o++; p.c1 = o++; p.w2_hi = o++; p.w2_lo = o++; p.b2 = o++;
    p.we = o++; p.be = o++;
    p.num = o;
    return p;
}

inline int64_t tn_align(int64_t v) { return (v + 255) / 256 * 256; }

struct TnWorkspace {
    float *a_hi, *a_lo, *b_hi, *b_lo, *d_hi, *d_lo;      // bf16 planes [M][c_max]: block input, block output / sub-block output, depthwise output
    float *pre, *res;                                    // fp32 [M][c_max]
    float *pooled, *hid, *gate, *stat, *bvec, *pool;     // per-crop vectors
    float *h1, *h2_hi, *h2_lo;                           // attention hidden [M][att]
    int64_t total_bytes;
};

TnWorkspace tn_carve(const nsf_titanet_dims& d, int n, int t_pad, unsigned char* base) {
    const TnPlan p = tn_plan(d);
    const int64_t M = (int64_t)n * t_pad;
    TnWorkspace w = {};
    int64_t cur = 0;
    auto take = [&](int64_t bytes) { unsigned char* q = base ? base + cur : nullptr; cur += tn_align(bytes); return reinterpret_cast<float*>(q); };
    const int64_t plane = M * p.c_max * 2;
    w.a_hi = take(plane); w.a_lo = take(plane); w.b_hi = take(plane); w.b_lo = take(plane); w.d_hi = take(plane); w.d_lo = take(plane);
    w.pre = take(M * p.c_max * 4); w.res = take(M * p.c_max * 4);
    w.pooled = take((int64_t)n * p.c_max * 4); w.hid = take((int64_t)n * p.c_max * 4); w.gate = take((int64_t)n * p.c_max * 4);
    w.stat = take((int64_t)n * 2 * p.c_max * 4); w.bvec = take((int64_t)n * d.att_ch * 4); w.pool = take((int64_t)n * 2 * p.c_max * 4);
    w.h1 = take(M * d.att_ch * 4); w.h2_hi = take(M * d.att_ch * 2); w.h2_lo = take(M * d.att_ch * 2);
    w.total_bytes = cur;
    return w;
}

}  // namespace

extern "C" int64_t nsf_titanet_num_offsets(const nsf_titanet_dims* d) { return d && tn_dims_ok(*d) ? tn_plan(*d).num : 0; }

extern "C" int nsf_titanet_create(const nsf_titanet_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets,
                                  int n_offsets, nsf_titanet** out) {
    NSF_REQUIRE(dims && blob && offsets && out, "nsf_titanet_create: null pointer");
    NSF_REQUIRE(tn_dims_ok(*dims), "nsf_titanet_create: unsupported dims (feat_in / att_ch multiples of 8, filters multiples of 64, odd kernels <= 31, <= 8 blocks)");
    NSF_REQUIRE(n_offsets == tn_plan(*dims).num, "nsf_titanet_create: expected %d offsets, got %d", tn_plan(*dims).num, n_offsets);
    for (int i = 0; i < n_offsets; ++i)
        NSF_REQUIRE(offsets[i] >= 0 && offsets[i] < blob_floats && offsets[i] % 4 == 0, "nsf_titanet_create: offset %d out of range / unaligned", i);
    nsf_titanet* h = new nsf_titanet;
    h->dims = *dims;
    h->blob = blob;
    h->n_offsets = n_offsets;
    h->offsets = new int64_t[n_offsets];
    for (int i = 0; i < n_offsets; ++i) h->offsets[i] = offsets[i];
    *out = h;
    return NSF_OK;
}

extern "C" void nsf_titanet_destroy(nsf_titanet* h) {
    if (!h) return;
    delete[] h->offsets;
    delete h;
}

extern "C" int64_t nsf_titanet_workspace_bytes(const nsf_titanet_dims* dims, int n_crops, int t_pad) {
    if (!dims || !tn_dims_ok(*dims) || n_crops <= 0 || t_pad <= 0) return 0;
    return tn_carve(*dims, n_crops, t_pad, nullptr).total_bytes;
}

extern "C" int nsf_titanet_features(const float* crops, const int32_t* lengths, int n_crops, int64_t max_len, int t_pad,
                                    const float* mel_filters, int n_mels, float* lm_scratch, void* feat_hi, void* feat_lo,
                                    int32_t* n_frames, void* stream_) {
    NSF_REQUIRE(crops && lengths && mel_filters && lm_scratch && feat_hi && feat_lo && n_frames, "nsf_titanet_features: null pointer");
    NSF_REQUIRE(n_crops >= 1 && n_crops <= 65535 && max_len >= 1 && n_mels >= 1 && n_mels <= 256, "nsf_titanet_features: bad sizes");
    NSF_REQUIRE(t_pad >= max_len / kTnHop + 1, "nsf_titanet_features: t_pad=%d cannot hold %lld frames", t_pad, (long long)(max_len / kTnHop + 1));
    cudaStream_t s = (cudaStream_t)stream_;
    ProfScope prof(PROF_FEATURES, (double)n_crops * max_len * 4.0, s);
    tn_logmel_kernel<<<dim3((t_pad + kTnFramesPerCta - 1) / kTnFramesPerCta, n_crops), 256, 0, s>>>(crops, lengths, max_len, t_pad, mel_filters, n_mels, lm_scratch, n_frames);
    int rc = check_launch("tn_logmel_kernel");
    if (rc) return rc;
    tn_featnorm_kernel<<<n_crops, (n_mels + 31) / 32 * 32, 0, s>>>(lm_scratch, n_frames, t_pad, n_mels, reinterpret_cast<float*>(feat_hi),
                                                                   reinterpret_cast<float*>(feat_lo));
    return check_launch("tn_featnorm_kernel");
}

extern "C" int nsf_titanet_forward(nsf_titanet* h, const void* feat_hi, const void* feat_lo, const int32_t* n_frames, int n_crops,
                                   int t_pad, float* emb, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(h && feat_hi && feat_lo && n_frames && emb && workspace, "nsf_titanet_forward: null pointer");
    if (n_crops <= 0) return NSF_OK;
    NSF_REQUIRE(t_pad >= 1 && n_crops <= 65535, "nsf_titanet_forward: bad sizes");
    NSF_REQUIRE(((uintptr_t)workspace & 255) == 0, "nsf_titanet_forward: workspace must be 256-byte aligned");
    const nsf_titanet_dims& D = h->dims;
    const TnPlan P = tn_plan(D);
    TnWorkspace w = tn_carve(D, n_crops, t_pad, reinterpret_cast<unsigned char*>(workspace));
    NSF_REQUIRE(workspace_bytes >= w.total_bytes, "nsf_titanet_forward: workspace too small");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t M64 = (int64_t)n_crops * t_pad;
    NSF_REQUIRE(M64 < 0x7fffffff / 8, "nsf_titanet_forward: batch too large");
    const int M = (int)M64;
    auto g = [&](int i) { return h->blob + h->offsets[i]; };
    int rc;

    // precision 0: fp32-grade (bf16 head / remainder planes, three MMAs per product); 1: one fp16 plane, one MMA -- the arithmetic
    // of the reference's autocast() region (word_based_diarization.py:102-105): fp16 operands, fp32 accumulation
    const int fmt = D.precision == 1 ? SPLIT_F16_1 : SPLIT_BF16;
    auto gemm = [&](const float* a_hi, const float* a_lo, int K, int w_hi, int w_lo, const float* bias, int N, int epi, float* o0, float* o1,
                    int out_fmt) {
        GemmParams p = {};
        p.batch = 1; p.alpha = 1.f; p.acc_scale = 1.f;
        p.op_fmt = fmt; p.out_fmt = out_fmt == SPLIT_BF16 ? fmt : out_fmt; p.qkv_fmt = fmt; p.d_k = 64;
        p.A_hi = a_hi; p.A_lo = a_lo; p.lda = K; p.B_hi = g(w_hi); p.B_lo = g(w_lo); p.ldb = K;
        p.M = M; p.N = N; p.K = K; p.n_valid = N; p.bias = bias; p.epi = epi; p.out0 = o0; p.out1 = o1; p.ldo = N;
        return gemm_launch(fmt == SPLIT_F16_1 ? NSF_GEMM_TC_BF16 : NSF_GEMM_TC_2XBF16, p, s);      // NSF_GEMM_TC_BF16 = the single-pass engine
    };
    auto fc = [&](const float* in, int K, const float* W, const float* bias, int N, int act, float* out) {
        const int64_t threads = (int64_t)n_crops * N * 32;
        tn_fc_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(in, K, W, bias, N, n_crops, act, out);
        return check_launch("tn_fc_kernel");
    };

    const float* in_hi = reinterpret_cast<const float*>(feat_hi);
    const float* in_lo = reinterpret_cast<const float*>(feat_lo);
    if (fmt == SPLIT_F16_1) {
        // the features arrive as bf16 pairs: one fp16 plane of them in a remainder plane this engine never writes
        const int64_t n8 = M64 * D.feat_in / 8;
        tn_pairs_to_f16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(feat_hi),
                                                                           reinterpret_cast<const uint16_t*>(feat_lo), reinterpret_cast<__half*>(w.b_lo), n8);
        if ((rc = check_launch("tn_pairs_to_f16_kernel"))) return rc;
        in_hi = w.b_lo;
        in_lo = w.b_lo;
    }
    int c_in = D.feat_in;
    for (int b = 0; b < D.n_blocks; ++b) {
        const int co = D.filters[b], k = D.kernel[b], rep = D.repeat[b];
        const bool last_block = b == D.n_blocks - 1;
        // the pair the block does not read its input from holds the sub-block outputs and, at the end, the block output
        float* y_hi = (in_hi == w.a_hi) ? w.b_hi : w.a_hi;
        float* y_lo = (in_hi == w.a_hi) ? w.b_lo : w.a_lo;
        const float* x_hi = in_hi;
        const float* x_lo = in_lo;
        int c = c_in;
        for (int r = 0; r < rep; ++r) {
            { ProfScope prof(PROF_NET_OTHER, 0.0, s);
              tn_dwconv_kernel<<<dim3((c + kDwCh - 1) / kDwCh, (t_pad + kDwFrames - 1) / kDwFrames, n_crops), 128, 0, s>>>(
                  reinterpret_cast<const uint16_t*>(x_hi), reinterpret_cast<const uint16_t*>(x_lo), n_frames, t_pad, c, k, g(P.dw[b][r]), w.d_hi, w.d_lo, fmt);
              if ((rc = check_launch("tn_dwconv_kernel"))) return rc; }
            if (r < rep - 1) {       // pointwise conv + BatchNorm + ReLU -> planes (the block's scratch output buffer)
                if ((rc = gemm(w.d_hi, w.d_lo, c, P.pw_hi[b][r], P.pw_lo[b][r], g(P.pw_b[b][r]), co, EPI_RELU_SPLIT, y_hi, y_lo, SPLIT_BF16))) return rc;
                x_hi = y_hi; x_lo = y_lo;
            } else {                 // last repeat: pointwise conv + BatchNorm -> fp32 (squeeze-excite needs the whole crop first)
                if ((rc = gemm(w.d_hi, w.d_lo, c, P.pw_hi[b][r], P.pw_lo[b][r], g(P.pw_b[b][r]), co, EPI_STORE, w.pre, nullptr, SPLIT_FP32))) return rc;
            }
            c = co;
        }
        if (D.residual[b])           // residual branch: 1x1 conv + BatchNorm of the block input
            if ((rc = gemm(in_hi, in_lo, c_in, P.res_hi[b], P.res_lo[b], g(P.res_b[b]), co, EPI_STORE, w.res, nullptr, SPLIT_FP32))) return rc;
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          tn_masked_stats_kernel<<<dim3((co + 31) / 32, n_crops), 256, 0, s>>>(w.pre, n_frames, t_pad, co, w.pooled, nullptr, co);
          if ((rc = check_launch("tn_masked_stats_kernel"))) return rc;
          if ((rc = fc(w.pooled, co, g(P.fc0[b]), nullptr, co / 8, 1, w.hid))) return rc;
          { const int64_t outs = (int64_t)n_crops * co;               // fc.2, stored transposed [co / 8][co]
            tn_fc_t_kernel<<<(unsigned)((outs + 255) / 256), 256, 0, s>>>(w.hid, co / 8, g(P.fc2[b]), co, n_crops, 2, w.gate);
            if ((rc = check_launch("tn_fc_t_kernel"))) return rc; }
          const int64_t total = M64 * (co / 8);
          tn_se_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w.pre, w.gate, D.residual[b] ? w.res : nullptr, n_frames, t_pad, co,
                                                                             y_hi, y_lo, last_block ? 1 : 0, total, fmt);
          if ((rc = check_launch("tn_se_apply_kernel"))) return rc;
          in_hi = y_hi; in_lo = y_lo; }
        c_in = co;
    }

    // ---- decoder: attentive statistics pooling with global context, embedding layer
    const int C = P.c_last, A = D.att_ch;
    { ProfScope prof(PROF_NET_OTHER, 0.0, s);
      tn_masked_stats_kernel<<<dim3((C + 31) / 32, n_crops), 256, 0, s>>>(w.pre, n_frames, t_pad, C, w.stat, w.stat + C, 2 * C);
      if ((rc = check_launch("tn_masked_stats_kernel"))) return rc;
      if ((rc = fc(w.stat, 2 * C, g(P.w1ms), g(P.b1), A, 0, w.bvec))) return rc; }
    if ((rc = gemm(in_hi, in_lo, C, P.w1x_hi, P.w1x_lo, nullptr, A, EPI_STORE, w.h1, nullptr, SPLIT_FP32))) return rc;
    { ProfScope prof(PROF_NET_OTHER, 0.0, s);
      const int64_t total = M64 * A;
      tn_att_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w.h1, w.bvec, g(P.a1), g(P.c1), t_pad, A, w.h2_hi, w.h2_lo, total, fmt);
      if ((rc = check_launch("tn_att_act_kernel"))) return rc; }
    if ((rc = gemm(w.h2_hi, w.h2_lo, A, P.w2_hi, P.w2_lo, g(P.b2), C, EPI_STORE, w.res, nullptr, SPLIT_FP32))) return rc;
    { ProfScope prof(PROF_NET_OTHER, 0.0, s);
      tn_att_pool_kernel<<<dim3((C + 31) / 32, n_crops), 256, 0, s>>>(w.pre, w.res, n_frames, t_pad, C, w.pool);
      if ((rc = check_launch("tn_att_pool_kernel"))) return rc;
      if ((rc = fc(w.pool, 2 * C, g(P.we), g(P.be), D.emb, 0, emb))) return rc; }
    return NSF_OK;
}

extern "C" int nsf_cos_affinity_accum(const float* emb, int64_t row_pitch, int dim, int n, float scale, float* en_scratch,
                                      float* sim_scratch, uint32_t* minmax, float* acc, void* stream_) {
    NSF_REQUIRE(emb && en_scratch && sim_scratch && minmax && acc, "nsf_cos_affinity_accum: null pointer");
    NSF_REQUIRE(n >= 2 && dim >= 1 && row_pitch >= dim, "nsf_cos_affinity_accum: bad sizes (n >= 2)");
    cudaStream_t s = (cudaStream_t)stream_;
    const unsigned init[2] = {0xffffffffu, 0u};
    NSF_CUDA(cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, s));
    tn_rownorm_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, s>>>(emb, row_pitch, dim, n, en_scratch);
    int rc = check_launch("tn_rownorm_kernel");
    if (rc) return rc;
    tn_cossim_kernel<<<dim3((n + 15) / 16, (n + 15) / 16), 256, 0, s>>>(en_scratch, dim, n, sim_scratch, minmax);
    if ((rc = check_launch("tn_cossim_kernel"))) return rc;
    const int64_t total = (int64_t)n * n;
    tn_affinity_accum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(sim_scratch, minmax, scale, total, acc);
    return check_launch("tn_affinity_accum_kernel");
}
