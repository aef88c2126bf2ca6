// Token-level timestamps from cross-attention weights: the numeric core of openai-whisper's word timestamps
// [upstream whisper/timing.py::find_alignment, median_filter, dtw; the reference asks for them with word_timestamps=True,
// asr/asr.py:52-56, and diarization consumes the word boundaries, word_based_diarization.py:78-101].
//
//   weights [B][A][N][M]  softmax cross-attention rows of the A alignment heads for N token positions over M audio positions
//   1. per (b, a, m): mean / population std over the tokens, w <- (w - mean) / std
//   2. per (b, a, n): median filter of width 7 along the audio axis (reflect padding)
//   3. mean over the heads, negated -> cost [B][N][M]
//   4. dynamic time warping per sequence (monotone path from (0, 0) to (N-1, M-1); ties resolved like upstream's dtw_cpu),
//      anti-diagonal wavefront in one CTA, fp32 like upstream; the path is walked back by one thread
//   5. start_frame[b][n] = audio position at which the path first enters token n   (time = start_frame * 0.02 s)
// Only the first m_valid audio positions take part (upstream crops the weights to num_frames // 2 before everything else).
#include "common.cuh"

namespace nsf {

constexpr int kWaMedian = 7;

// mean and 1/std over the tokens for every (b, a, m)
__global__ void wa_stats_kernel(const float* __restrict__ w, int A, int N_max, const int32_t* __restrict__ n_tokens, int M, int m_valid,
                                float2* __restrict__ stats) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int ba = blockIdx.y;
    if (m >= m_valid) return;
    const int N = n_tokens ? max(1, min(n_tokens[ba / A], N_max)) : N_max;      // statistics over the sequence's own tokens
    const float* p = w + (size_t)ba * N_max * M + m;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += p[(size_t)n * M];
    const float mean = s / N;
    float q = 0.f;
    for (int n = 0; n < N; ++n) { const float d = p[(size_t)n * M] - mean; q = fmaf(d, d, q); }
    stats[(size_t)ba * M + m] = make_float2(mean, 1.f / sqrtf(q / N));
}

__device__ __forceinline__ float wa_median7(float (&v)[kWaMedian]) {
    // partial selection sort: the 4th smallest of 7
#pragma unroll
    for (int i = 0; i <= kWaMedian / 2; ++i) {
#pragma unroll
        for (int j = i + 1; j < kWaMedian; ++j) {
            const float lo = fminf(v[i], v[j]), hi = fmaxf(v[i], v[j]);
            v[i] = lo; v[j] = hi;
        }
    }
    return v[kWaMedian / 2];
}

// cost[b][n][m] = -mean_a median7_m((w - mean) / std)
__global__ void wa_cost_kernel(const float* __restrict__ w, const float2* __restrict__ stats, int A, int N, int M, int m_valid,
                               float* __restrict__ cost) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y, b = blockIdx.z;
    if (m >= m_valid) return;
    float acc = 0.f;
    for (int a = 0; a < A; ++a) {
        const float* row = w + (((size_t)b * A + a) * N + n) * M;
        const float2* st = stats + ((size_t)b * A + a) * M;
        float v[kWaMedian];
#pragma unroll
        for (int k = 0; k < kWaMedian; ++k) {
            int mm = m + k - kWaMedian / 2;
            if (m_valid > kWaMedian / 2) {                        // reflect padding (no filtering at all for very short inputs)
                if (mm < 0) mm = -mm;
                if (mm >= m_valid) mm = 2 * (m_valid - 1) - mm;
            } else {
                mm = m;
            }
            const float2 s = st[mm];
            v[k] = (row[mm] - s.x) * s.y;
        }
        acc += wa_median7(v);
    }
    cost[((size_t)b * N + n) * m_valid + m] = -(acc / A);
}

// One CTA per sequence.  D [(N+1)][(M+1)] f32 and trace [(N+1)][(M+1)] u8 live in global scratch (L2-resident).
__global__ void __launch_bounds__(512)
wa_dtw_kernel(const float* __restrict__ cost, const int32_t* __restrict__ n_tokens, int N_max, int M, float* __restrict__ Dall,
              uint8_t* __restrict__ Tall, int32_t* __restrict__ start_frame) {
    const int b = blockIdx.x;
    const int N = n_tokens ? min(n_tokens[b], N_max) : N_max;
    const float* x = cost + (size_t)b * N_max * M;
    float* D = Dall + (size_t)b * (N_max + 1) * (M + 1);
    uint8_t* T = Tall + (size_t)b * (N_max + 1) * (M + 1);
    const int W = M + 1;
    for (int e = threadIdx.x; e < (N + 1) * W; e += blockDim.x) D[e] = INFINITY;
    __syncthreads();
    if (threadIdx.x == 0) D[0] = 0.f;
    __syncthreads();
    for (int d = 2; d <= N + M; ++d) {                               // cells (i, j) with i + j = d, 1 <= i <= N, 1 <= j <= M
        const int i_lo = max(1, d - M), i_hi = min(N, d - 1);
        for (int i = i_lo + threadIdx.x; i <= i_hi; i += blockDim.x) {
            const int j = d - i;
            const float c0 = D[(i - 1) * W + j - 1], c1 = D[(i - 1) * W + j], c2 = D[i * W + j - 1];
            float c;
            uint8_t t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            D[i * W + j] = x[(size_t)(i - 1) * M + j - 1] + c;
            T[i * W + j] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // walk back from (N, M); first row moves left, first column moves up; the path enters token n at the smallest j on it
        for (int n = 0; n < N_max; ++n) start_frame[(size_t)b * N_max + n] = 0;
        int i = N, j = M;
        while (i > 0 || j > 0) {
            if (i >= 1 && j >= 1) start_frame[(size_t)b * N_max + i - 1] = j - 1;      // overwritten until the smallest j of token i - 1
            const int t = (i == 0) ? 2 : (j == 0) ? 1 : T[i * W + j];
            if (t == 0) { --i; --j; }
            else if (t == 1) --i;
            else --j;
        }
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_whisper_alignment_workspace_bytes(int n_batch, int n_heads, int n_tokens, int n_frames) {
    if (n_batch <= 0 || n_heads <= 0 || n_tokens <= 0 || n_frames <= 0) return 0;
    const int64_t B = n_batch, A = n_heads, N = n_tokens, M = n_frames;
    auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
    return al(B * A * M * 8) + al(B * N * M * 4) + al(B * (N + 1) * (M + 1) * 4) + al(B * (N + 1) * (M + 1));
}

extern "C" int nsf_whisper_alignment(const float* weights, int n_batch, int n_heads, int n_tokens, int n_frames, int m_valid,
                                     const int32_t* n_tokens_per_seq, int32_t* start_frame, float* cost_out, void* workspace,
                                     int64_t workspace_bytes, void* stream_) {
    NSF_REQUIRE(weights && start_frame && workspace, "nsf_whisper_alignment: null pointer");
    NSF_REQUIRE(n_batch >= 1 && n_batch <= 65535 && n_heads >= 1 && n_tokens >= 1 && n_tokens <= 65535 && n_frames >= 1,
                "nsf_whisper_alignment: bad sizes");
    NSF_REQUIRE(m_valid >= 1 && m_valid <= n_frames, "nsf_whisper_alignment: m_valid=%d outside [1, %d]", m_valid, n_frames);
    NSF_REQUIRE(((uintptr_t)workspace & 255) == 0 && workspace_bytes >= nsf_whisper_alignment_workspace_bytes(n_batch, n_heads, n_tokens, n_frames),
                "nsf_whisper_alignment: workspace too small or unaligned");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t B = n_batch, A = n_heads, N = n_tokens, M = n_frames;
    auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
    unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
    float2* stats = reinterpret_cast<float2*>(base); base += al(B * A * M * 8);
    float* cost = reinterpret_cast<float*>(base); base += al(B * N * M * 4);
    float* D = reinterpret_cast<float*>(base); base += al(B * (N + 1) * (M + 1) * 4);
    uint8_t* T = base;
    wa_stats_kernel<<<dim3((m_valid + 127) / 128, n_batch * n_heads), 128, 0, s>>>(weights, n_heads, n_tokens, n_tokens_per_seq, n_frames, m_valid, stats);
    int rc = check_launch("wa_stats_kernel");
    if (rc) return rc;
    wa_cost_kernel<<<dim3((m_valid + 127) / 128, n_tokens, n_batch), 128, 0, s>>>(weights, stats, n_heads, n_tokens, n_frames, m_valid, cost);
    if ((rc = check_launch("wa_cost_kernel"))) return rc;
    if (cost_out) NSF_CUDA(cudaMemcpyAsync(cost_out, cost, (size_t)(B * N * m_valid) * 4, cudaMemcpyDeviceToDevice, s));
    wa_dtw_kernel<<<n_batch, 512, 0, s>>>(cost, n_tokens_per_seq, n_tokens, m_valid, D, T, start_frame);
    return check_launch("wa_dtw_kernel");
}
