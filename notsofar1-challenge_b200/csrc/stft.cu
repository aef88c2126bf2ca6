// Multichannel STFT and iSTFT kernels (HBM-bound; shared-memory radix-8 FFTs, no tensor cores).
//
// Reference semantics:
//   STFT  : css_with_conformer/executor/feature.py:88-128 (conv1d with the hann*DFT kernel of
//           init_kernel :19-45, stride 256, no padding) + th.polar round trip of
//           css/training/conformer_wrapper.py:120-124.
//   iSTFT : feature.py:138-167 (conv_transpose1d with the sqrt-hann/16 kernel; FeatureExtractor
//           does not forward `window`, feature.py:422-425) as called from conformer_wrapper.py:131-146.
#include "common.cuh"
#include "fft512.cuh"
#include <mutex>

namespace nsf {

// ------------------------------------------------------------------------------------------- tables
__device__ float2 g_twiddle[512];
__device__ float g_hann[512];
__device__ float g_sqrt_hann16[512];

__global__ void init_tables_kernel() {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 512) {
        double s, c;
        sincospi(j / 256.0, &s, &c);
        // keep the exact zeros of the quarter circle exact
        if (j == 128 || j == 384) c = 0.0;
        if (j == 0 || j == 256) s = 0.0;
        g_twiddle[j] = make_float2((float)c, (float)s);
        double h = 0.5 - 0.5 * cospi(j / 256.0);          // periodic hann (th.hann_window default)
        g_hann[j] = (float)h;
        g_sqrt_hann16[j] = (float)(sqrt(h) / 16.0);       // S = 0.5*sqrt(N*N/hop) = 16, feature.py:32-34
    }
}

static std::mutex g_tab_mutex;
static bool g_tab_ready[64] = {};

static int ensure_tables(cudaStream_t stream) {
    int dev = 0;
    NSF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return NSF_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (g_tab_ready[dev]) return NSF_OK;
    init_tables_kernel<<<2, 256, 0, stream>>>();
    int rc = check_launch("init_tables_kernel");
    if (rc) return rc;
    NSF_CUDA(cudaStreamSynchronize(stream));
    g_tab_ready[dev] = true;
    return NSF_OK;
}

// ------------------------------------------------------------------------------------------- STFT
constexpr int kStftTT = 4;          // frames per CTA
constexpr int kStftThreads = 512;   // 8 groups of 64 threads = 4 channel pairs x 2 frames in flight (the FFT stages are
                                    // barrier-latency bound: 32 warps per SM instead of 16)

struct StftSmem {
    float2 tw[512];
    float hann[512];
    float scratch[kStftThreads / 64][kFftScratchFloats];
    // out[k][tt][c] follows (dynamic, 257 * TT * n_ch float2)
};

__global__ void __launch_bounds__(kStftThreads, 2)
stft_mc_kernel(const float* __restrict__ x, int n_ch, float2* __restrict__ X, int64_t T_long, int64_t n_frames) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
    float2* out = reinterpret_cast<float2*>(smem_raw + sizeof(StftSmem));

    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += kStftThreads) {
        sm.tw[i] = g_twiddle[i];
        sm.hann[i] = g_hann[i];
    }
    __syncthreads();

    const int group = tid >> 6;          // (frame parity, channel pair)
    const int lane64 = tid & 63;
    const int ch_a = 2 * (group & 3), ch_b = 2 * (group & 3) + 1;
    const int64_t t0 = (int64_t)blockIdx.x * kStftTT;
    const int n_tt = (int)min((int64_t)kStftTT, n_frames - t0);
    const bool active = ch_a < n_ch;
    float* scratch = sm.scratch[group];
    auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory"); };

    if (active) {
        for (int tt = group >> 2; tt < n_tt; tt += kStftThreads / 256) {
            const float* xf = x + (t0 + tt) * (int64_t)kHop * n_ch;
            float2 v[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int n = 64 * a + lane64;
                const float w = sm.hann[n];
                const float va = __ldg(xf + (int64_t)n * n_ch + ch_a);
                const float vb = (ch_b < n_ch) ? __ldg(xf + (int64_t)n * n_ch + ch_b) : 0.f;
                v[a] = make_float2(va * w, vb * w);
            }
            fft512_group<-1>(v, lane64, scratch, sm.tw, group_sync);
            // natural-order spectrum of the packed pair into the group's scratch
            float* zre = scratch;
            float* zim = scratch + 8 * 72;
            {
                const int k0 = lane64 >> 3, k1 = lane64 & 7;
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    zre[k0 + 8 * k1 + 64 * k2] = v[k2].x;
                    zim[k0 + 8 * k1 + 64 * k2] = v[k2].y;
                }
            }
            group_sync();
            // A[k] = (Z[k] + conj(Z[-k]))/2 ; B[k] = (Z[k] - conj(Z[-k]))/(2i)
            for (int k = lane64; k <= 256; k += 64) {
                const int kn = (512 - k) & 511;
                const float zr = zre[k], zi = zim[k], yr = zre[kn], yi = zim[kn];
                float2 A = make_float2(0.5f * (zr + yr), 0.5f * (zi - yi));
                float2 B = make_float2(0.5f * (zi + yi), 0.5f * (yr - zr));
                if (k == 0 || k == 256) {
                    // reference: imag is exactly +0 here, then X = polar(|re|, atan2(+0, re)):
                    // re < 0 -> phase pi_f32 -> imag = |re| * sinf(pi_f32) = re * 8.742278e-08f
                    const float s = 8.742278e-08f;
                    A.y = A.x < 0.f ? A.x * s : 0.f;
                    B.y = B.x < 0.f ? B.x * s : 0.f;
                }
                float2* o = out + ((size_t)k * kStftTT + tt) * n_ch;
                o[ch_a] = A;
                if (ch_b < n_ch) o[ch_b] = B;
            }
            group_sync();
        }
    }
    __syncthreads();
    // write-out: for every bin a run of n_tt * n_ch contiguous complex values
    const int run = n_tt * n_ch;
    const int warp = tid >> 5, lane = tid & 31;
    for (int k = warp; k < kBins; k += kStftThreads / 32) {
        const float2* src = out + (size_t)k * kStftTT * n_ch;
        float2* dst = X + ((size_t)k * T_long + t0) * n_ch;
        for (int j = lane; j < run; j += 32) dst[j] = src[j];
    }
}

// ------------------------------------------------------------------------------------------- iSTFT
constexpr int kIstftTT = 8;          // output hops (256 samples each) per CTA
constexpr int kIstftThreads = 256;   // 4 groups of 64 threads

struct IstftSmem {
    float2 tw[512];
    float win[512];
    float scratch[4][kFftScratchFloats];
    float frames[kIstftTT + 2][512];   // windowed time-domain frames t0-1 .. t0+TT (slot j <-> frame t0-1+j)
};

// One CTA produces samples [256*t0, 256*(t0+TT)) of one stream from frames t0-1 .. t0+TT-1; the CTA
// holding the last frame also writes the 256-sample tail.  Frame pairs are packed into one complex
// transform: u1 + i u2 = IDFT(Hs1 + i Hs2), Hs = Hermitian-symmetrised half spectrum (so that the
// real part of the one-sided sum of feature.py:157-162 falls out without doubling).
__global__ void __launch_bounds__(kIstftThreads, 2)
istft_kernel(const float2* __restrict__ S, int64_t T_long, float* __restrict__ wav, int64_t n_out, int64_t hop_begin, int64_t hop_end) {
    __shared__ IstftSmem sm;
    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += kIstftThreads) {
        sm.tw[i] = g_twiddle[i];
        sm.win[i] = g_sqrt_hann16[i];
    }
    const int stream_id = blockIdx.y;
    const int64_t t0 = hop_begin + (int64_t)blockIdx.x * kIstftTT;
    const float2* Ss = S + (size_t)stream_id * T_long * kBins;
    float* ws = wav + (size_t)stream_id * n_out;
    __syncthreads();

    const int group = tid >> 6, lane64 = tid & 63;
    float* scratch = sm.scratch[group];
    auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory"); };

    // slots 0..TT (frames t0-1 .. t0+TT-1) -> (TT+1) frames -> ceil((TT+1)/2) pairs over 4 groups
    constexpr int kPairs = (kIstftTT + 2) / 2;
    for (int pair = group; pair < kPairs; pair += 4) {
        const int slot_a = 2 * pair, slot_b = 2 * pair + 1;
        const int64_t ta = t0 - 1 + slot_a, tb = t0 - 1 + slot_b;
        const bool va = (ta >= 0 && ta < T_long), vb = (tb >= 0 && tb < T_long && slot_b <= kIstftTT);
        const float2* Sa = Ss + (size_t)(va ? ta : 0) * kBins;
        const float2* Sb = Ss + (size_t)(vb ? tb : 0) * kBins;
        float2 v[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int k = 64 * a + lane64;            // 0..511
            const int kk = k <= 256 ? k : 512 - k;    // source bin
            float2 ha = va ? __ldg(Sa + kk) : make_float2(0.f, 0.f);
            float2 hb = vb ? __ldg(Sb + kk) : make_float2(0.f, 0.f);
            // reference: r = |S| cos(angle S), i = |S| sin(angle S) (feature.py:157-158) == (re, im) to 1 ulp
            if (k == 0 || k == 256) { ha.y = 0.f; hb.y = 0.f; }                 // Hs[0] = Re S[0], Hs[256] = Re S[256]
            else if (k < 256) { ha.x *= 0.5f; ha.y *= 0.5f; hb.x *= 0.5f; hb.y *= 0.5f; }
            else { ha.x *= 0.5f; ha.y *= -0.5f; hb.x *= 0.5f; hb.y *= -0.5f; }  // conj(S[512-k])/2
            v[a] = make_float2(ha.x - hb.y, ha.y + hb.x);                        // Hs_a + i Hs_b
        }
        fft512_group<+1>(v, lane64, scratch, sm.tw, group_sync);
        const int k0 = lane64 >> 3, k1 = lane64 & 7;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            const int m = k0 + 8 * k1 + 64 * k2;
            const float g = sm.win[m];
            sm.frames[slot_a][m] = v[k2].x * g;
            if (slot_b <= kIstftTT) sm.frames[slot_b][m] = v[k2].y * g;
        }
    }
    __syncthreads();
    // overlap-add: hop j (samples 256*(t0+j) ..) = first half of frame t0+j (slot j+1) + second half of frame t0+j-1 (slot j)
    for (int idx = tid; idx < kIstftTT * 256; idx += kIstftThreads) {
        const int j = idx >> 8, m = idx & 255;
        const int64_t hop_idx = t0 + j;
        if (hop_idx >= hop_end) break;
        // order of the two addends follows conv_transpose1d's accumulation over t (older frame first)
        float acc = 0.f;
        if (hop_idx - 1 >= 0) acc += sm.frames[j][256 + m];
        if (hop_idx < T_long) acc += sm.frames[j + 1][m];
        ws[hop_idx * 256 + m] = acc;
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_num_frames(int64_t n_samples) {
    return n_samples < kFrame ? 0 : (n_samples - kFrame) / kHop + 1;
}

extern "C" int nsf_stft_mc(const float* x, int64_t n_samples, int n_ch, float* X, int64_t T_long, int64_t n_frames,
                           void* stream) {
    NSF_REQUIRE(x && X, "nsf_stft_mc: null pointer");
    NSF_REQUIRE(n_ch >= 1 && n_ch <= 8, "nsf_stft_mc: n_ch=%d not in [1,8]", n_ch);
    NSF_REQUIRE(n_frames >= 0 && n_frames <= nsf_num_frames(n_samples) && n_frames <= T_long,
                "nsf_stft_mc: n_frames=%lld exceeds signal (%lld samples) or pitch %lld", (long long)n_frames,
                (long long)n_samples, (long long)T_long);
    if (n_frames == 0) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_tables(s);
    if (rc) return rc;
    const size_t smem = sizeof(StftSmem) + (size_t)kBins * kStftTT * n_ch * sizeof(float2);
    NSF_CUDA(cudaFuncSetAttribute(stft_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = ceil_div64(n_frames, kStftTT);
    ProfScope prof(PROF_STFT, (double)n_frames * n_ch * (kHop * 4.0 + kBins * 8.0), s);
    stft_mc_kernel<<<(unsigned)grid, kStftThreads, smem, s>>>(x, n_ch, reinterpret_cast<float2*>(X), T_long, n_frames);
    return check_launch("stft_mc_kernel");
}

extern "C" int nsf_istft_range(const float* S_st, int n_streams, int64_t T_long, float* wav, int64_t hop_begin, int64_t hop_end,
                               void* stream) {
    NSF_REQUIRE(S_st && wav, "nsf_istft: null pointer");
    NSF_REQUIRE(n_streams >= 1 && T_long >= 1, "nsf_istft: bad sizes");
    NSF_REQUIRE(hop_begin >= 0 && hop_begin % kIstftTT == 0 && hop_end <= T_long + 1,
                "nsf_istft_range: hops [%lld, %lld) of %lld (the first must be a multiple of %d)", (long long)hop_begin, (long long)hop_end,
                (long long)(T_long + 1), kIstftTT);
    if (hop_end <= hop_begin) return NSF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_tables(s);
    if (rc) return rc;
    const int64_t n_out = (T_long - 1) * kHop + kFrame;
    dim3 grid((unsigned)ceil_div64(hop_end - hop_begin, kIstftTT), (unsigned)n_streams);
    ProfScope prof(PROF_ISTFT, (double)(hop_end - hop_begin) * n_streams * (kBins * 8.0 + kHop * 4.0), s);
    istft_kernel<<<grid, kIstftThreads, 0, s>>>(reinterpret_cast<const float2*>(S_st), T_long, wav, n_out, hop_begin, hop_end);
    return check_launch("istft_kernel");
}

extern "C" int nsf_istft(const float* S_st, int n_streams, int64_t T_long, float* wav, void* stream) {
    NSF_REQUIRE(T_long >= 1, "nsf_istft: bad sizes");
    return nsf_istft_range(S_st, n_streams, T_long, wav, 0, T_long + 1, stream);
}
