// Whisper audio encoder forward (bf16 tensor cores, fp32 residual stream) and its log-mel front end.
//
// Algorithm [upstream openai-whisper, unpinned HEAD -- requirements.txt:2 of the reference; call site asr/asr.py:69-74]:
//   whisper/audio.py log_mel_spectrogram: STFT (n_fft 400, hop 160, periodic hann, centred with reflect padding), power,
//     mel filterbank, log10(max(., 1e-10)), max(., global_max - 8), (. + 4) / 4
//   whisper/model.py AudioEncoder: x = gelu(conv1(mel)); x = gelu(conv2(x)) (k = 3, stride 2); x += positional_embedding;
//     n_layer x { x += attn(ln(x)); x += mlp(ln(x)) }; ln_post
//   ResidualAttentionBlock / MultiHeadAttention: q, k scaled by d_k^-0.25 each (folded into the packed weights here),
//     key projection without bias, softmax(q k^T) v, output projection; mlp = Linear(d, 4d) -> GELU(erf) -> Linear(4d, d).
// Parity is against the transformers re-implementation of the same published algorithm (tests/test_whisper.py): unpinned
// with respect to openai-whisper itself.
//
// Both convolutions are GEMMs on *strided views*, no im2col copy: with the activations stored time-major with one zero
// row in front, the operand row of output frame t is the contiguous run of 3 input rows starting at row t (conv1) or
// 2t (conv2); the TMA tensor map simply uses a row pitch smaller than the row length.
#include "gemm_common.cuh"
#include <mutex>
#include <new>

namespace nsf {

constexpr int kWhFrames = 3000;          // mel frames per 30-s chunk
constexpr int kWhPadRows = 3002;         // time-major rows per chunk incl. the zero row in front (even: batch pitch % row pitch == 0)
constexpr int kWhNfft = 400, kWhHop = 160, kWhBins = 201;
constexpr int kWhSamples = 480000;

enum WhGlobal { WG_CONV1_W_HI = 0, WG_CONV1_W_LO, WG_CONV1_B, WG_CONV2_W, WG_CONV2_B, WG_POS, WG_LNP_G, WG_LNP_B, WG_NUM };
enum WhLayer { WL_LN1_G = 0, WL_LN1_B, WL_WQKV, WL_BQKV, WL_WO, WL_BO, WL_LN2_G, WL_LN2_B, WL_W1, WL_B1, WL_W2, WL_B2, WL_NUM };

// ------------------------------------------------------------------------------------------- log-mel front end
__device__ float2 g_wh_twiddle[kWhNfft];     // (cos, sin)(2 pi j / 400)
__device__ float g_wh_hann[kWhNfft];
__global__ void wh_init_tables_kernel() {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < kWhNfft) {
        double s, c;
        sincospi(2.0 * j / (double)kWhNfft, &s, &c);
        g_wh_twiddle[j] = make_float2((float)c, (float)s);
        g_wh_hann[j] = (float)(0.5 - 0.5 * cospi(2.0 * j / (double)kWhNfft));      // torch.hann_window(400) (periodic)
    }
}

__device__ __forceinline__ unsigned wh_order(float v) {           // order-preserving float -> uint (for atomicMax)
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float wh_unorder(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// One CTA per (chunk, frame): windowed frame in shared memory, direct 400-point DFT (one bin per thread), power, mel
// filterbank, log10.  log_spec [B][n_mels][3000] f32; gmax [B] order-encoded running maximum.
__global__ void __launch_bounds__(256)
wh_logmel_kernel(const float* __restrict__ audio, int64_t n_samples, const float* __restrict__ filters, int n_mels,
                 float* __restrict__ log_spec, unsigned* __restrict__ gmax, int64_t n_frames) {
    __shared__ float xw[kWhNfft];
    __shared__ float2 tw[kWhNfft];
    __shared__ float pw[kWhBins];
    __shared__ float red[8];
    const int b = blockIdx.y, t = blockIdx.x;
    const float* x = audio + (size_t)b * n_samples;
    for (int n = threadIdx.x; n < kWhNfft; n += blockDim.x) {
        int64_t i = (int64_t)t * kWhHop + n - kWhNfft / 2;           // centred frame, reflect padding (torch.stft center=True)
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2 * (n_samples - 1) - i;
        xw[n] = x[i] * g_wh_hann[n];
        tw[n] = g_wh_twiddle[n];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kWhBins; k += blockDim.x) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < kWhNfft; ++n) {
            const float2 w = tw[idx];
            re = fmaf(xw[n], w.x, re);
            im = fmaf(-xw[n], w.y, im);
            idx += k;
            if (idx >= kWhNfft) idx -= kWhNfft;
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
        const float* f = filters + (size_t)m * kWhBins;
        float acc = 0.f;
        for (int k = 0; k < kWhBins; ++k) acc = fmaf(__ldg(f + k), pw[k], acc);
        const float v = log10f(fmaxf(acc, 1e-10f));
        log_spec[((size_t)b * n_mels + m) * n_frames + t] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
        atomicMax(gmax + b, wh_order(mx));
    }
}

// max(., gmax - 8), (. + 4) / 4, transpose to time-major bf16 head / remainder planes [B][3002][n_mels] (row 0 and row 3001 zero)
__global__ void __launch_bounds__(256)
wh_logmel_finish_kernel(const float* __restrict__ log_spec, const unsigned* __restrict__ gmax, int n_mels, float* __restrict__ mel_hi,
                        float* __restrict__ mel_lo) {
    const int b = blockIdx.y;
    const float floor_ = wh_unorder(gmax[b]) - 8.f;
    const int64_t n = (int64_t)kWhPadRows * n_mels;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / n_mels), m = (int)(e - (int64_t)row * n_mels);
        float v = 0.f;
        if (row >= 1 && row <= kWhFrames) v = (fmaxf(log_spec[((size_t)b * n_mels + m) * kWhFrames + row - 1], floor_) + 4.f) * 0.25f;
        split_store(SPLIT_BF16, mel_hi, mel_lo, (size_t)b * n + e, v);
    }
}

// one 30-s window of a whole-recording log-mel (normalised with the recording's maximum, whisper/audio.py:150-155 [upstream]):
// max(., gmax - 8), (. + 4) / 4 of frames [seek, seek + 3000) (zero beyond the recording = pad_or_trim), time-major bf16 planes
__global__ void __launch_bounds__(256)
wh_mel_window_kernel(const float* __restrict__ log_spec, int64_t n_frames, const unsigned* __restrict__ gmax, int n_mels,
                     const int32_t* __restrict__ seeks, const int32_t* __restrict__ sizes, float* __restrict__ mel_hi, float* __restrict__ mel_lo) {
    const int b = blockIdx.y;
    const float floor_ = wh_unorder(gmax[0]) - 8.f;
    const int64_t seek = seeks[b];
    const int size = sizes ? sizes[b] : kWhFrames;                  // frames of the window that hold content; the rest is zero (pad_or_trim)
    const int64_t n = (int64_t)kWhPadRows * n_mels;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / n_mels), m = (int)(e - (int64_t)row * n_mels);
        float v = 0.f;
        const int64_t t = seek + row - 1;
        if (row >= 1 && row <= size && t < n_frames) v = (fmaxf(log_spec[(size_t)m * n_frames + t], floor_) + 4.f) * 0.25f;
        split_store(SPLIT_BF16, mel_hi, mel_lo, (size_t)b * n + e, v);
    }
}

}  // namespace nsf

struct nsf_whisper_encoder {
    nsf_whisper_dims dims;
    const float* blob;
    int64_t* offsets;
    int n_offsets;
    const float* g(int i) const { return blob + offsets[i]; }
    const float* l(int layer, int i) const { return blob + offsets[nsf::WG_NUM + layer * nsf::WL_NUM + i]; }
};

namespace nsf {

static inline int64_t wh_align(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct WhWorkspace {
    float *x, *h, *u, *q, *k, *vt, *x1p;       // x fp32; the rest bf16 planes addressed through float pointers
    int64_t total_bytes;
};
static WhWorkspace wh_carve(const nsf_whisper_dims& D, int n_batch, unsigned char* base) {
    const int64_t M = (int64_t)n_batch * D.n_ctx;
    const int Tp = (int)wh_align(D.n_ctx, 8);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { unsigned char* p = base ? base + off : nullptr; off += wh_align(bytes, 256); return reinterpret_cast<float*>(p); };
    WhWorkspace w;
    w.x = take(M * D.d_model * 4);
    w.h = take(M * D.d_model * 2);
    w.u = take(M * D.d_ff * 2);
    w.q = take(M * D.d_model * 2);
    w.k = take(M * D.d_model * 2);
    w.vt = take((int64_t)n_batch * D.d_model * Tp * 2);
    w.x1p = take((int64_t)n_batch * kWhPadRows * D.d_model * 2);
    w.total_bytes = off;
    return w;
}

static std::mutex g_wh_tab_mutex;
static bool g_wh_tab_ready[64] = {};

// the DFT twiddles and the window live in __device__ globals: one initialisation per device of the process
static int wh_ensure_tables(cudaStream_t s) {
    int dev = 0;
    NSF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return NSF_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lock(g_wh_tab_mutex);
    if (g_wh_tab_ready[dev]) return NSF_OK;
    wh_init_tables_kernel<<<2, 256, 0, s>>>();
    int rc = check_launch("wh_init_tables_kernel");
    if (rc) return rc;
    NSF_CUDA(cudaStreamSynchronize(s));
    g_wh_tab_ready[dev] = true;
    return NSF_OK;
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_whisper_encoder_num_offsets(const nsf_whisper_dims* d) { return d ? WG_NUM + (int64_t)WL_NUM * d->n_layers : 0; }

extern "C" int nsf_whisper_encoder_create(const nsf_whisper_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets,
                                          int n_offsets, nsf_whisper_encoder** out) {
    NSF_REQUIRE(dims && blob && offsets && out, "nsf_whisper_encoder_create: null pointer");
    NSF_REQUIRE(dims->d_model % 128 == 0 && dims->d_model == dims->n_heads * 64, "whisper: d_model=%d must be n_heads * 64 and a multiple of 128", dims->d_model);
    NSF_REQUIRE(dims->n_mels % 8 == 0 && dims->n_mels >= 8 && dims->d_ff % 8 == 0 && dims->n_layers >= 1, "whisper: n_mels / d_ff / n_layers");
    NSF_REQUIRE(dims->n_ctx == kWhFrames / 2, "whisper: n_ctx=%d (a 30-s chunk has 1500 positions)", dims->n_ctx);
    NSF_REQUIRE(n_offsets == nsf_whisper_encoder_num_offsets(dims), "nsf_whisper_encoder_create: expected %lld offsets", (long long)nsf_whisper_encoder_num_offsets(dims));
    for (int i = 0; i < n_offsets; ++i)
        NSF_REQUIRE(offsets[i] >= 0 && offsets[i] < blob_floats && (offsets[i] & 3) == 0, "nsf_whisper_encoder_create: offset %d", i);
    nsf_whisper_encoder* h = new (std::nothrow) nsf_whisper_encoder;
    NSF_REQUIRE(h, "out of memory");
    h->dims = *dims; h->blob = blob; h->n_offsets = n_offsets;
    h->offsets = new (std::nothrow) int64_t[n_offsets];
    if (!h->offsets) { delete h; set_error("out of memory"); return NSF_ERR_INVALID_ARG; }
    for (int i = 0; i < n_offsets; ++i) h->offsets[i] = offsets[i];
    *out = h;
    return NSF_OK;
}

extern "C" void nsf_whisper_encoder_destroy(nsf_whisper_encoder* h) {
    if (!h) return;
    delete[] h->offsets;
    delete h;
}

extern "C" int64_t nsf_whisper_encoder_workspace_bytes(const nsf_whisper_dims* dims, int n_batch) {
    if (!dims || n_batch <= 0) return 0;
    return wh_carve(*dims, n_batch, nullptr).total_bytes;
}

extern "C" int64_t nsf_whisper_mel_plane_elems(int n_mels, int n_batch) { return (int64_t)n_batch * kWhPadRows * n_mels; }

extern "C" int nsf_whisper_logmel(const float* audio, int n_batch, int64_t n_samples, const float* filters, int n_mels, float* log_spec,
                                  uint32_t* gmax, void* mel_hi, void* mel_lo, void* stream_) {
    NSF_REQUIRE(audio && filters && log_spec && gmax && mel_hi && mel_lo, "nsf_whisper_logmel: null pointer");
    NSF_REQUIRE(n_samples == kWhSamples, "nsf_whisper_logmel: chunks are %d samples (30 s at 16 kHz), got %lld", kWhSamples, (long long)n_samples);
    NSF_REQUIRE(n_batch >= 1 && n_batch <= 65535 && n_mels >= 1, "nsf_whisper_logmel: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_;
    int rc = wh_ensure_tables(s);
    if (rc) return rc;
    NSF_CUDA(cudaMemsetAsync(gmax, 0, sizeof(uint32_t) * n_batch, s));        // order-encoded: 0 is below every float
    ProfScope prof(PROF_FEATURES, (double)n_batch * (kWhSamples * 4.0 + 2.0 * kWhPadRows * n_mels * 2.0), s);
    wh_logmel_kernel<<<dim3(kWhFrames, n_batch), 256, 0, s>>>(audio, n_samples, filters, n_mels, log_spec, gmax, kWhFrames);
    if ((rc = check_launch("wh_logmel_kernel"))) return rc;
    wh_logmel_finish_kernel<<<dim3(64, n_batch), 256, 0, s>>>(log_spec, gmax, n_mels, reinterpret_cast<float*>(mel_hi), reinterpret_cast<float*>(mel_lo));
    return check_launch("wh_logmel_finish_kernel");
}

extern "C" int nsf_whisper_logmel_recording(const float* audio, int64_t n_samples, const float* filters, int n_mels, int64_t n_frames,
                                            float* log_spec, uint32_t* gmax, void* stream_) {
    NSF_REQUIRE(audio && filters && log_spec && gmax, "nsf_whisper_logmel_recording: null pointer");
    NSF_REQUIRE(n_samples >= kWhNfft && n_mels >= 1 && n_frames >= 1 && n_frames <= n_samples / kWhHop && n_frames <= 0x7fffffff,
                "nsf_whisper_logmel_recording: n_samples=%lld n_frames=%lld", (long long)n_samples, (long long)n_frames);
    cudaStream_t s = (cudaStream_t)stream_;
    int rc = wh_ensure_tables(s);
    if (rc) return rc;
    NSF_CUDA(cudaMemsetAsync(gmax, 0, sizeof(uint32_t), s));
    ProfScope prof(PROF_FEATURES, (double)n_samples * 4.0 + (double)n_frames * n_mels * 4.0, s);
    wh_logmel_kernel<<<dim3((unsigned)n_frames, 1), 256, 0, s>>>(audio, n_samples, filters, n_mels, log_spec, gmax, n_frames);
    return check_launch("wh_logmel_kernel");
}

extern "C" int nsf_whisper_mel_windows(const float* log_spec, int64_t n_frames, const uint32_t* gmax, int n_mels, const int32_t* seeks,
                                       const int32_t* sizes, int n_windows, void* mel_hi, void* mel_lo, void* stream_) {
    NSF_REQUIRE(log_spec && gmax && seeks && mel_hi && mel_lo, "nsf_whisper_mel_windows: null pointer");
    NSF_REQUIRE(n_windows >= 1 && n_windows <= 65535 && n_mels >= 1 && n_frames >= 1, "nsf_whisper_mel_windows: bad sizes");
    wh_mel_window_kernel<<<dim3(64, n_windows), 256, 0, (cudaStream_t)stream_>>>(log_spec, n_frames, gmax, n_mels, seeks, sizes,
                                                                                 reinterpret_cast<float*>(mel_hi), reinterpret_cast<float*>(mel_lo));
    return check_launch("wh_mel_window_kernel");
}

extern "C" int nsf_whisper_encoder_forward(nsf_whisper_encoder* h, const void* mel_hi, const void* mel_lo, int n_batch, float* out,
                                           void* out_bf16, void* workspace, int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(h && mel_hi && mel_lo && out && workspace, "nsf_whisper_encoder_forward: null pointer");
    if (n_batch <= 0) return NSF_OK;
    const nsf_whisper_dims& D = h->dims;
    NSF_REQUIRE(((uintptr_t)workspace & 255) == 0, "nsf_whisper_encoder_forward: workspace must be 256-byte aligned");
    WhWorkspace w = wh_carve(D, n_batch, reinterpret_cast<unsigned char*>(workspace));
    NSF_REQUIRE(workspace_bytes >= w.total_bytes, "nsf_whisper_encoder_forward: workspace too small");
    cudaStream_t s = (cudaStream_t)stream_;
    const int d = D.d_model, H = D.n_heads, T = D.n_ctx, dff = D.d_ff, nm = D.n_mels;
    const int M = n_batch * T;
    const int Tp = (int)wh_align(T, 8);
    int rc;

    auto base = [&]() {
        GemmParams p = {};
        p.batch = 1; p.alpha = 1.f; p.acc_scale = 1.f;
        p.op_fmt = SPLIT_BF16_1; p.out_fmt = SPLIT_BF16_1; p.qkv_fmt = SPLIT_BF16_1; p.v_rowmajor = 1;
        p.T = T; p.Tp = Tp; p.n_heads = H; p.d_k = 64; p.d_model = d;
        p.q_hi = w.q; p.q_lo = w.q; p.k_hi = w.k; p.k_lo = w.k; p.vt_hi = w.vt; p.vt_lo = w.vt;
        return p;
    };
    // zero row in front of every chunk of the conv1 output
    NSF_CUDA(cudaMemsetAsync(w.x1p, 0, (size_t)n_batch * kWhPadRows * d * 2, s));

    {   // conv1 (k = 3, pad 1) + GELU: rows t .. t+2 of the zero-framed time-major mel, 2xBF16 engine (K = 3 n_mels is tiny)
        GemmParams p = base();
        p.op_fmt = SPLIT_BF16;
        p.A_hi = reinterpret_cast<const float*>(mel_hi); p.A_lo = reinterpret_cast<const float*>(mel_lo);
        p.lda = nm; p.a_batch_stride = (int64_t)kWhPadRows * nm;
        p.B_hi = h->g(WG_CONV1_W_HI); p.B_lo = h->g(WG_CONV1_W_LO); p.ldb = 3 * nm; p.b_shared = 1;
        p.M = kWhFrames; p.N = d; p.K = 3 * nm; p.n_valid = d; p.batch = n_batch;
        p.bias = h->g(WG_CONV1_B); p.epi = EPI_GELU_SPLIT;
        p.out0 = reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(w.x1p) + d);      // row 1 of every chunk
        p.out1 = p.out0; p.ldo = d; p.o_batch_stride = (int64_t)kWhPadRows * d;
        if ((rc = gemm_launch(NSF_GEMM_TC_2XBF16, p, s))) return rc;
    }
    {   // conv2 (k = 3, stride 2, pad 1) + GELU + positional embedding: rows 2t .. 2t+2 of the zero-framed conv1 output
        GemmParams p = base();
        p.A_hi = w.x1p; p.A_lo = nullptr; p.lda = 2 * d; p.a_batch_stride = (int64_t)kWhPadRows * d;
        p.B_hi = h->g(WG_CONV2_W); p.B_lo = nullptr; p.ldb = 3 * d; p.b_shared = 1;
        p.M = T; p.N = d; p.K = 3 * d; p.n_valid = d; p.batch = n_batch;
        p.bias = h->g(WG_CONV2_B); p.epi = EPI_GELU_POS;
        p.out0 = w.x; p.out1 = const_cast<float*>(h->g(WG_POS)); p.ldo = d;
        if ((rc = gemm_launch(NSF_GEMM_TC_BF16, p, s))) return rc;
    }
    auto linear = [&](const float* a, int K, const float* wt, const float* bias, int N, int epi, float* o0, int64_t ldo) {
        GemmParams p = base();
        p.A_hi = a; p.lda = K; p.B_hi = wt; p.ldb = K;
        p.M = M; p.N = N; p.K = K; p.n_valid = N;
        p.bias = bias; p.epi = epi; p.out0 = o0; p.out1 = o0; p.ldo = ldo;
        return gemm_launch(NSF_GEMM_TC_BF16, p, s);
    };
    for (int L = 0; L < D.n_layers; ++L) {
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          rc = ln_launch(w.x, M, d, h->l(L, WL_LN1_G), h->l(L, WL_LN1_B), 0, nullptr, nullptr, nullptr, w.h, w.h, SPLIT_BF16_1, s); }
        if (rc) return rc;
        if ((rc = linear(w.h, d, h->l(L, WL_WQKV), h->l(L, WL_BQKV), 3 * d, EPI_QKV, nullptr, 0))) return rc;
        { ProfScope prof(PROF_ATTN, 4.0 * T * T * 64 * (double)n_batch * H, s);
          if ((rc = flash_attn_launch(w.q, w.k, w.vt, n_batch, H, T, w.h, w.h, d, SPLIT_BF16_1, s))) return rc; }
        if ((rc = linear(w.h, d, h->l(L, WL_WO), h->l(L, WL_BO), d, EPI_RESID, w.x, d))) return rc;
        { ProfScope prof(PROF_NET_OTHER, 0.0, s);
          rc = ln_launch(w.x, M, d, h->l(L, WL_LN2_G), h->l(L, WL_LN2_B), 0, nullptr, nullptr, nullptr, w.h, w.h, SPLIT_BF16_1, s); }
        if (rc) return rc;
        if ((rc = linear(w.h, d, h->l(L, WL_W1), h->l(L, WL_B1), dff, EPI_GELU_SPLIT, w.u, dff))) return rc;
        if ((rc = linear(w.u, dff, h->l(L, WL_W2), h->l(L, WL_B2), d, EPI_RESID, w.x, d))) return rc;
    }
    { ProfScope prof(PROF_NET_OTHER, 0.0, s);
      rc = ln_launch(w.x, M, d, h->g(WG_LNP_G), h->g(WG_LNP_B), 0, out, nullptr, nullptr, reinterpret_cast<float*>(out_bf16),
                     reinterpret_cast<float*>(out_bf16), SPLIT_BF16_1, s); }
    return rc;
}
