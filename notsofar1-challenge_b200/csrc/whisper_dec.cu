// Whisper text decoder, greedy step (bf16 tensor-core GEMMs, fp32 residual stream, bf16 key / value caches).
//
// Algorithm [upstream openai-whisper, unpinned -- whisper/model.py TextDecoder / ResidualAttentionBlock; see whisper.cu]:
//   x = token_embedding[tok] + positional_embedding[pos]
//   n_layer x { x += self_attn(ln(x), causal, kv cache); x += cross_attn(ln(x), audio features); x += mlp(ln(x)) }
//   logits = ln(x) @ token_embedding^T;  greedy: next token = argmax(logits)
// The d_k^-0.25 scales of q and k are folded into the packed weights.  Cross-attention keys / values of all layers are
// projected once per batch of chunks (nsf_whisper_decoder_prefill_cross); every decode step then is a fixed sequence of
// small-M tcgen05 GEMMs (M = sequences in flight) and two bandwidth-bound cache-attention kernels.
// Not built: beam search, temperature fallback, the logit filters of whisper/decoding.py (SuppressBlank, SuppressTokens,
// timestamp rules) and word timestamps -- the host loop is plain greedy arg-max.
#include "gemm_common.cuh"
#include <new>

namespace nsf {

enum WdGlobal { WD_TOK_EMB = 0, WD_POS, WD_LN_G, WD_LN_B, WD_NUM };
enum WdLayer { WDL_LN1_G = 0, WDL_LN1_B, WDL_WQKV, WDL_BQKV, WDL_WO, WDL_BO, WDL_LNC_G, WDL_LNC_B, WDL_WCQ, WDL_BCQ, WDL_WCKV, WDL_BCKV,
               WDL_WCO, WDL_BCO, WDL_LN2_G, WDL_LN2_B, WDL_W1, WDL_B1, WDL_W2, WDL_B2, WDL_NUM };

__global__ void wd_set_int_kernel(int32_t* p, int v) { *p = v; }

__global__ void __launch_bounds__(256)
wd_embed_kernel(const int32_t* __restrict__ tokens, const uint16_t* __restrict__ emb, const float* __restrict__ pos_emb,
                const int32_t* __restrict__ pos_ptr, int d, int vocab, float* __restrict__ x) {
    const int b = blockIdx.x;
    const float* pos_row = pos_emb + (size_t)(*pos_ptr) * d;
    int tok = tokens[b];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        x[(size_t)b * d + c] = __bfloat162float(__ushort_as_bfloat16(emb[(size_t)tok * d + c])) + __ldg(pos_row + c);
}

// One CTA per (sequence, head): optional append of the new key / value to the cache, scores against n_keys cached keys,
// softmax, weighted sum of the cached values.  q / k_new / v_new: fp32 rows of pitch ldq (column head * 64 + d).
// Kc, Vc: [n_bh][t_max][64] bf16.  out: bf16 plane [n_batch][d_model].
template <bool APPEND>
__global__ void __launch_bounds__(256)
wd_attn_kernel(const float* __restrict__ q, const float* __restrict__ k_new, const float* __restrict__ v_new, int64_t ldq,
               uint16_t* __restrict__ Kc, uint16_t* __restrict__ Vc, int n_keys_host, const int32_t* __restrict__ pos_ptr, int t_max,
               int n_heads, uint16_t* __restrict__ out, int d_model, float* __restrict__ probs_out = nullptr,
               const int32_t* __restrict__ align_row = nullptr, int n_align = 0, int probs_len = 0) {
    extern __shared__ float sm[];                     // q[64] | p[t_max rounded] | red[32] | acc[8][64]
    const int n_keys = APPEND ? (*pos_ptr + 1) : n_keys_host;       // self-attention: keys 0 .. pos (the position lives on the device)
    float* qs = sm;
    float* p = sm + 64;
    float* red = p + ((t_max + 31) & ~31);
    float* accs = red + 32;
    const int bh = blockIdx.x, b = bh / n_heads, h = bh - b * n_heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint16_t* K = Kc + (size_t)bh * t_max * 64;
    uint16_t* V = Vc + (size_t)bh * t_max * 64;
    if (tid < 64) {
        qs[tid] = q[(size_t)b * ldq + h * 64 + tid];
        if (APPEND) {
            K[(size_t)(n_keys - 1) * 64 + tid] = __bfloat16_as_ushort(__float2bfloat16_rn(k_new[(size_t)b * ldq + h * 64 + tid]));
            V[(size_t)(n_keys - 1) * 64 + tid] = __bfloat16_as_ushort(__float2bfloat16_rn(v_new[(size_t)b * ldq + h * 64 + tid]));
        }
    }
    __syncthreads();
    // scores: 8 lanes per key (16 bytes = 8 dims each), 4 keys per warp instruction, 32 keys per CTA pass
    const int sub = lane & 7, krow = lane >> 3;
    float qv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qv[e] = qs[8 * sub + e];
    float mx = -INFINITY;
    for (int t0 = 0; t0 < n_keys; t0 += 32) {
        const int t = t0 + 4 * warp + krow;
        float s = 0.f;
        if (t < n_keys) {
            const uint4 kk = *reinterpret_cast<const uint4*>(K + (size_t)t * 64 + 8 * sub);
            const uint32_t w4[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                s = fmaf(qv[2 * e], __uint_as_float(w4[e] << 16), fmaf(qv[2 * e + 1], __uint_as_float(w4[e] & 0xffff0000u), s));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (t < n_keys) {
            if (sub == 0) p[t] = s;
            mx = fmaxf(mx, s);
        }
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int t = tid; t < n_keys; t += 256) {
        const float e = __expf(p[t] - mx);
        p[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    if (!APPEND && probs_out) {
        // alignment heads: the softmax row of this position goes to probs [n_batch][n_align][probs_len][n_keys] (word timestamps)
        const int slot = align_row[h];
        const int pos = *pos_ptr;
        if (slot >= 0 && pos < probs_len) {
            float* dst = probs_out + (((size_t)b * n_align + slot) * probs_len + pos) * n_keys;
            const float inv = 1.f / sum;
            for (int t = tid; t < n_keys; t += 256) dst[t] = p[t] * inv;
        }
    }
    // weighted values: thread = (key slot tid / 8 of 32, 8 dims tid % 8), 16-byte loads, partial sums reduced through smem
    const int vsub = tid & 7, vrow = tid >> 3;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int t = vrow; t < n_keys; t += 32) {
        const uint4 vv = *reinterpret_cast<const uint4*>(V + (size_t)t * 64 + 8 * vsub);
        const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
        const float pt = p[t];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[2 * e] = fmaf(pt, __uint_as_float(w4[e] << 16), acc[2 * e]);
            acc[2 * e + 1] = fmaf(pt, __uint_as_float(w4[e] & 0xffff0000u), acc[2 * e + 1]);
        }
    }
    // reduce the 4 key slots that share a lane group inside each warp, then the 8 warps through shared memory
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    __syncthreads();                                   // p[] is no longer needed by anyone: accs may alias nothing, but order the phases
    if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) accs[warp * 64 + 8 * lane + e] = acc[e];
    }
    __syncthreads();
    if (tid < 64) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) o += accs[w * 64 + tid];
        out[(size_t)b * d_model + h * 64 + tid] = __bfloat16_as_ushort(__float2bfloat16_rn(o / sum));
    }
}

// first maximum of every row (torch.argmax)
__global__ void __launch_bounds__(1024)
wd_argmax_kernel(const float* __restrict__ logits, int vocab, int32_t* __restrict__ next) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    const float* row = logits + (size_t)blockIdx.x * vocab;
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = row[i];
        if (v > best) { best = v; idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
        next[blockIdx.x] = idx;
    }
}

// Logit filters of the decoding loop, applied in place before the arg-max, one block per sequence
// [upstream whisper/decoding.py: SuppressBlank, SuppressTokens, ApplyTimestampRules -- DecodingTask applies them in this order]:
//   * at the first sampled position: blank tokens and eot are suppressed; always: the suppress list
//   * <|notimestamps|> is suppressed; timestamps come in pairs (after one timestamp: no text; after two: no timestamp);
//     timestamps do not decrease (and a segment has non-zero length); the first sampled token is a timestamp not later than
//     max_initial_timestamp_index; if the probability mass of all timestamps exceeds the most likely text token, text is
//     suppressed (log-softmax normalisers cancel: logsumexp(timestamp logits) > max(text logits)).
// tokens [n_batch][total_len]: the fed tokens, valid up to position *pos_ptr (the token consumed by this step).
__global__ void __launch_bounds__(1024)
wd_logit_rules_kernel(float* __restrict__ logits, int vocab, const int32_t* __restrict__ tokens, int total_len,
                      const int32_t* __restrict__ pos_ptr, nsf_whisper_rules R, const int32_t* __restrict__ suppress,
                      const int32_t* __restrict__ suppress_first) {
    __shared__ float redf[32];
    __shared__ int redi[32];
    float* row = logits + (size_t)blockIdx.x * vocab;
    const int n_sampled = *pos_ptr + 1 - R.sample_begin;
    if (n_sampled < 0) return;                                   // still inside the prompt: the next token is forced
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (n_sampled == 0)
        for (int i = tid; i < R.n_suppress_first; i += blockDim.x) { const int t = suppress_first[i]; if (t >= 0 && t < vocab) row[t] = -INFINITY; }
    for (int i = tid; i < R.n_suppress; i += blockDim.x) { const int t = suppress[i]; if (t >= 0 && t < vocab) row[t] = -INFINITY; }
    const int tb = R.timestamp_begin;
    if (tb < 0) return;
    __syncthreads();
    auto block_max_i = [&](int v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) redi[wid] = v;
        __syncthreads();
        int r = redi[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = max(r, redi[w]);
        return r;
    };
    auto block_red_f = [&](float v, bool is_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float ov = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, ov) : v + ov; }
        __syncthreads();
        if (lane == 0) redf[wid] = v;
        __syncthreads();
        float r = redf[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, redf[w]) : r + redf[w];
        return r;
    };
    const int32_t* seq = tokens + (size_t)blockIdx.x * total_len + R.sample_begin;
    const bool last_ts = n_sampled >= 1 && seq[n_sampled - 1] >= tb;
    const bool penult_ts = n_sampled < 2 || seq[n_sampled - 2] >= tb;
    int li = -1;
    for (int i = tid; i < n_sampled; i += blockDim.x) if (seq[i] >= tb) li = max(li, i);
    li = block_max_i(li);
    const int forbid_end = li < 0 ? tb : (last_ts && !penult_ts) ? seq[li] : seq[li] + 1;     // timestamps [tb, forbid_end) are forbidden
    float mx_text = -INFINITY, mx_ts = -INFINITY;
    for (int i = tid; i < vocab; i += blockDim.x) {
        bool f = i == R.no_timestamps;
        f |= last_ts && (penult_ts ? i >= tb : i < R.eot);
        f |= i >= tb && i < forbid_end;
        f |= n_sampled == 0 && (i < tb || (R.max_initial_timestamp_index >= 0 && i > tb + R.max_initial_timestamp_index));
        float v = row[i];
        if (f) { v = -INFINITY; row[i] = v; }
        if (i < tb) mx_text = fmaxf(mx_text, v); else mx_ts = fmaxf(mx_ts, v);
    }
    mx_text = block_red_f(mx_text, true);
    mx_ts = block_red_f(mx_ts, true);
    if (mx_ts == -INFINITY) return;                              // no timestamp allowed: nothing to compare (uniform)
    float sum = 0.f;
    for (int i = max(tb, 0) + tid; i < vocab; i += blockDim.x) sum += expf(row[i] - mx_ts);
    sum = block_red_f(sum, false);
    if (mx_ts + logf(sum) > mx_text)
        for (int i = tid; i < tb; i += blockDim.x) row[i] = -INFINITY;
}

// Bookkeeping of one greedy step, one block: which token is fed next (the prompt / teacher-forced token if forced >= 0, the
// end-of-text token once a sequence is done, else the arg-max), the record of fed tokens and arg-maxes, the position.
__global__ void __launch_bounds__(1024)
wd_advance_kernel(const int32_t* __restrict__ next, const int32_t* __restrict__ forced, int total_len, int eot, int32_t* __restrict__ cur,
                  int32_t* __restrict__ out_tokens, int32_t* __restrict__ argmaxes, uint8_t* __restrict__ done, int32_t* __restrict__ pos_ptr,
                  int n_batch) {
    const int p = *pos_ptr;
    __syncthreads();
    for (int b = threadIdx.x; b < n_batch; b += blockDim.x) {
        const int a = next[b];
        int fed = a;
        const int f = (forced && p + 1 < total_len) ? forced[(size_t)b * total_len + p + 1] : -1;
        if (f >= 0) fed = f;
        else {
            if (done[b]) fed = eot;
            else if (eot >= 0 && a == eot) done[b] = 1;
        }
        if (p + 1 < total_len) {
            out_tokens[(size_t)b * total_len + p + 1] = fed;
            argmaxes[(size_t)b * total_len + p + 1] = a;
        }
        cur[b] = fed;
    }
    __syncthreads();
    if (threadIdx.x == 0) *pos_ptr = p + 1;
}

}  // namespace nsf

struct nsf_whisper_decoder {
    nsf_whisper_dec_dims dims;
    const float* blob;
    int64_t* offsets;
    const float* g(int i) const { return blob + offsets[i]; }
    const float* l(int layer, int i) const { return blob + offsets[nsf::WD_NUM + layer * nsf::WDL_NUM + i]; }
};

namespace nsf {

static inline int64_t wd_align(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// beam search: sequence b continues the hypothesis that sat in slot src[b] (whisper/decoding.py PyTorchInference.rearrange_kv_cache
// [upstream]).  Rows [0, *n_pos) of every (layer, sequence, head) self-attention cache go through a scratch copy (gather, then
// copy back: a permutation cannot be applied in place without cycles).
__global__ void __launch_bounds__(256)
wd_cache_gather_kernel(const uint16_t* __restrict__ src_cache, uint16_t* __restrict__ dst_cache, const int32_t* __restrict__ src,
                       const int32_t* __restrict__ n_pos, int n_heads, int n_text_ctx, int64_t layer_stride, int from_scratch) {
    const int bh = blockIdx.x, L = blockIdx.y;
    const int b = bh / n_heads, hd = bh - b * n_heads;
    const int sb = from_scratch ? b : src[b];
    const int n = min(*n_pos, n_text_ctx) * 64 / 8;                 // 16-byte words
    const uint4* s4 = reinterpret_cast<const uint4*>(src_cache + (size_t)L * layer_stride + ((size_t)sb * n_heads + hd) * n_text_ctx * 64);
    uint4* d4 = reinterpret_cast<uint4*>(dst_cache + (size_t)L * layer_stride + ((size_t)b * n_heads + hd) * n_text_ctx * 64);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d4[i] = s4[i];
}

struct WdState {
    uint16_t *ck, *cv, *sk, *sv;       // cross / self caches: [L][n_bh][t][64]
    float *x, *qkv, *qc, *logits;
    float *h, *u;                      // bf16 planes
    int32_t* next;
    int32_t* pos;                      // device-side position of nsf_whisper_decoder_step (host-position entry point)
    int64_t total_bytes;
};
static WdState wd_carve(const nsf_whisper_dec_dims& D, int n_batch, unsigned char* base) {
    int64_t off = 0;
    auto take = [&](int64_t bytes) { unsigned char* p = base ? base + off : nullptr; off += wd_align(bytes, 256); return p; };
    const int64_t bh = (int64_t)n_batch * D.n_heads;
    WdState s;
    s.ck = (uint16_t*)take(D.n_layers * bh * D.n_audio_ctx * 64 * 2);
    s.cv = (uint16_t*)take(D.n_layers * bh * D.n_audio_ctx * 64 * 2);
    s.sk = (uint16_t*)take(D.n_layers * bh * D.n_text_ctx * 64 * 2);
    s.sv = (uint16_t*)take(D.n_layers * bh * D.n_text_ctx * 64 * 2);
    s.x = (float*)take((int64_t)n_batch * D.d_model * 4);
    s.qkv = (float*)take((int64_t)n_batch * 3 * D.d_model * 4);
    s.qc = (float*)take((int64_t)n_batch * D.d_model * 4);
    s.logits = (float*)take((int64_t)n_batch * D.vocab * 4);
    s.h = (float*)take((int64_t)n_batch * D.d_model * 2);
    s.u = (float*)take((int64_t)n_batch * D.d_ff * 2);
    s.next = (int32_t*)take((int64_t)n_batch * 4);
    s.pos = (int32_t*)take(256);
    s.total_bytes = off;
    return s;
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_whisper_decoder_num_offsets(const nsf_whisper_dec_dims* d) { return d ? WD_NUM + (int64_t)WDL_NUM * d->n_layers : 0; }

extern "C" int nsf_whisper_decoder_create(const nsf_whisper_dec_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets,
                                          int n_offsets, nsf_whisper_decoder** out) {
    NSF_REQUIRE(dims && blob && offsets && out, "nsf_whisper_decoder_create: null pointer");
    NSF_REQUIRE(dims->d_model % 128 == 0 && dims->d_model == dims->n_heads * 64 && dims->d_ff % 8 == 0 && dims->n_layers >= 1 &&
                dims->vocab >= 2 && dims->n_text_ctx >= 1 && dims->n_audio_ctx >= 1, "nsf_whisper_decoder_create: bad dims");
    NSF_REQUIRE(n_offsets == nsf_whisper_decoder_num_offsets(dims), "nsf_whisper_decoder_create: expected %lld offsets",
                (long long)nsf_whisper_decoder_num_offsets(dims));
    for (int i = 0; i < n_offsets; ++i)
        NSF_REQUIRE(offsets[i] >= 0 && offsets[i] < blob_floats && (offsets[i] & 3) == 0, "nsf_whisper_decoder_create: offset %d", i);
    nsf_whisper_decoder* h = new (std::nothrow) nsf_whisper_decoder;
    NSF_REQUIRE(h, "out of memory");
    h->dims = *dims; h->blob = blob;
    h->offsets = new (std::nothrow) int64_t[n_offsets];
    if (!h->offsets) { delete h; set_error("out of memory"); return NSF_ERR_INVALID_ARG; }
    for (int i = 0; i < n_offsets; ++i) h->offsets[i] = offsets[i];
    *out = h;
    return NSF_OK;
}

extern "C" void nsf_whisper_decoder_destroy(nsf_whisper_decoder* h) {
    if (!h) return;
    delete[] h->offsets;
    delete h;
}

extern "C" int64_t nsf_whisper_decoder_state_bytes(const nsf_whisper_dec_dims* dims, int n_batch) {
    if (!dims || n_batch <= 0) return 0;
    return wd_carve(*dims, n_batch, nullptr).total_bytes;
}

static GemmParams wd_base(const nsf_whisper_dec_dims& D, int M) {
    GemmParams p = {};
    p.batch = 1; p.alpha = 1.f; p.acc_scale = 1.f;
    p.op_fmt = SPLIT_BF16_1; p.out_fmt = SPLIT_BF16_1; p.qkv_fmt = SPLIT_BF16_1;
    p.n_heads = D.n_heads; p.d_k = 64; p.d_model = D.d_model;
    p.M = M;
    return p;
}

extern "C" int nsf_whisper_decoder_prefill_cross(nsf_whisper_decoder* h, const void* enc_bf16, int n_batch, void* state, int64_t state_bytes,
                                                 void* stream_) {
    NSF_REQUIRE(h && enc_bf16 && state, "nsf_whisper_decoder_prefill_cross: null pointer");
    const nsf_whisper_dec_dims& D = h->dims;
    NSF_REQUIRE(((uintptr_t)state & 255) == 0, "nsf_whisper_decoder_prefill_cross: state must be 256-byte aligned");
    WdState st = wd_carve(D, n_batch, reinterpret_cast<unsigned char*>(state));
    NSF_REQUIRE(state_bytes >= st.total_bytes, "nsf_whisper_decoder_prefill_cross: state too small");
    cudaStream_t s = (cudaStream_t)stream_;
    const int d = D.d_model, Ta = D.n_audio_ctx;
    const int64_t per_layer = (int64_t)n_batch * D.n_heads * Ta * 64;
    for (int L = 0; L < D.n_layers; ++L) {
        // [Wk; Wv] against the audio features; the q / k slots of the QKV scatter are the layer's cross key / value caches
        GemmParams p = wd_base(D, n_batch * Ta);
        p.A_hi = reinterpret_cast<const float*>(enc_bf16); p.lda = d;
        p.B_hi = h->l(L, WDL_WCKV); p.ldb = d;
        p.N = 2 * d; p.K = d; p.n_valid = 2 * d;
        p.bias = h->l(L, WDL_BCKV); p.epi = EPI_QKV;
        p.T = Ta; p.Tp = Ta;
        p.q_hi = p.q_lo = reinterpret_cast<float*>(st.ck + L * per_layer);
        p.k_hi = p.k_lo = reinterpret_cast<float*>(st.cv + L * per_layer);
        p.vt_hi = p.vt_lo = nullptr;
        int rc = gemm_launch(NSF_GEMM_TC_BF16, p, s);
        if (rc) return rc;
    }
    return NSF_OK;
}

struct WdRulesArgs {
    const nsf_whisper_rules* rules; const int32_t* suppress; const int32_t* suppress_first; const int32_t* tokens; int total_len;
    float* probs; const int32_t* align_map; int n_align;        // optional capture of the alignment heads' cross-attention rows
};

static int wd_step_impl(nsf_whisper_decoder* h, const int32_t* tokens, const int32_t* pos_dev, int n_batch, void* state, int64_t state_bytes,
                        float* logits_out, int32_t* next_tokens, cudaStream_t s, const WdRulesArgs* ra = nullptr) {
    const nsf_whisper_dec_dims& D = h->dims;
    NSF_REQUIRE(((uintptr_t)state & 255) == 0, "nsf_whisper_decoder_step: state must be 256-byte aligned");
    WdState st = wd_carve(D, n_batch, reinterpret_cast<unsigned char*>(state));
    NSF_REQUIRE(state_bytes >= st.total_bytes, "nsf_whisper_decoder_step: state too small");
    const int d = D.d_model, H = D.n_heads, B = n_batch, dff = D.d_ff;
    const int64_t bh = (int64_t)B * H;
    int rc;
    uint16_t* hb = reinterpret_cast<uint16_t*>(st.h);

    wd_embed_kernel<<<B, 256, 0, s>>>(tokens, reinterpret_cast<const uint16_t*>(h->g(WD_TOK_EMB)), h->g(WD_POS), pos_dev, d, D.vocab, st.x);
    if ((rc = check_launch("wd_embed_kernel"))) return rc;
    auto linear = [&](const float* a, int K, const float* wt, const float* bias, int N, int epi, float* o0, int64_t ldo) {
        GemmParams p = wd_base(D, B);
        p.A_hi = a; p.lda = K; p.B_hi = wt; p.ldb = K;
        p.N = N; p.K = K; p.n_valid = N;
        p.bias = bias; p.epi = epi; p.out0 = o0; p.out1 = o0; p.ldo = ldo;
        return gemm_launch(NSF_GEMM_TC_BF16, p, s);
    };
    auto attn_smem = [](int t_max) { return (size_t)(64 + ((t_max + 31) & ~31) + 32 + 512) * sizeof(float); };
    for (int L = 0; L < D.n_layers; ++L) {
        if ((rc = ln_launch(st.x, B, d, h->l(L, WDL_LN1_G), h->l(L, WDL_LN1_B), 0, nullptr, nullptr, nullptr, st.h, st.h, SPLIT_BF16_1, s))) return rc;
        if ((rc = linear(st.h, d, h->l(L, WDL_WQKV), h->l(L, WDL_BQKV), 3 * d, EPI_STORE, st.qkv, 3 * d))) return rc;
        { ProfScope prof(PROF_ATTN, 0.0, s);
        wd_attn_kernel<true><<<(unsigned)bh, 256, attn_smem(D.n_text_ctx), s>>>(st.qkv, st.qkv + d, st.qkv + 2 * d, 3 * d,
            st.sk + (size_t)L * bh * D.n_text_ctx * 64, st.sv + (size_t)L * bh * D.n_text_ctx * 64, 0, pos_dev, D.n_text_ctx, H, hb, d); }
        if ((rc = check_launch("wd_attn_kernel<self>"))) return rc;
        if ((rc = linear(st.h, d, h->l(L, WDL_WO), h->l(L, WDL_BO), d, EPI_RESID, st.x, d))) return rc;
        if ((rc = ln_launch(st.x, B, d, h->l(L, WDL_LNC_G), h->l(L, WDL_LNC_B), 0, nullptr, nullptr, nullptr, st.h, st.h, SPLIT_BF16_1, s))) return rc;
        if ((rc = linear(st.h, d, h->l(L, WDL_WCQ), h->l(L, WDL_BCQ), d, EPI_STORE, st.qc, d))) return rc;
        { ProfScope prof(PROF_MVDR, 4.0 * bh * D.n_audio_ctx * 64 * 2.0 / 2.0, s);       // cross-attention cache reads (bytes), reported under "mvdr"
        wd_attn_kernel<false><<<(unsigned)bh, 256, attn_smem(D.n_audio_ctx), s>>>(st.qc, nullptr, nullptr, d,
            st.ck + (size_t)L * bh * D.n_audio_ctx * 64, st.cv + (size_t)L * bh * D.n_audio_ctx * 64, D.n_audio_ctx, pos_dev, D.n_audio_ctx, H, hb, d,
            (ra && ra->probs) ? ra->probs : nullptr, (ra && ra->probs) ? ra->align_map + (size_t)L * H : nullptr, ra ? ra->n_align : 0,
            ra ? ra->total_len : 0); }
        if ((rc = check_launch("wd_attn_kernel<cross>"))) return rc;
        if ((rc = linear(st.h, d, h->l(L, WDL_WCO), h->l(L, WDL_BCO), d, EPI_RESID, st.x, d))) return rc;
        if ((rc = ln_launch(st.x, B, d, h->l(L, WDL_LN2_G), h->l(L, WDL_LN2_B), 0, nullptr, nullptr, nullptr, st.h, st.h, SPLIT_BF16_1, s))) return rc;
        if ((rc = linear(st.h, d, h->l(L, WDL_W1), h->l(L, WDL_B1), dff, EPI_GELU_SPLIT, st.u, dff))) return rc;
        if ((rc = linear(st.u, dff, h->l(L, WDL_W2), h->l(L, WDL_B2), d, EPI_RESID, st.x, d))) return rc;
    }
    if ((rc = ln_launch(st.x, B, d, h->g(WD_LN_G), h->g(WD_LN_B), 0, nullptr, nullptr, nullptr, st.h, st.h, SPLIT_BF16_1, s))) return rc;
    float* lg = logits_out ? logits_out : st.logits;
    if ((rc = linear(st.h, d, h->g(WD_TOK_EMB), nullptr, D.vocab, EPI_STORE, lg, D.vocab))) return rc;
    if (ra && ra->rules) {
        wd_logit_rules_kernel<<<B, 1024, 0, s>>>(lg, D.vocab, ra->tokens, ra->total_len, pos_dev, *ra->rules, ra->suppress, ra->suppress_first);
        if ((rc = check_launch("wd_logit_rules_kernel"))) return rc;
    }
    wd_argmax_kernel<<<B, 1024, 0, s>>>(lg, D.vocab, next_tokens);
    return check_launch("wd_argmax_kernel");
}

extern "C" int nsf_whisper_decoder_step(nsf_whisper_decoder* h, const int32_t* tokens, int pos, int n_batch, void* state, int64_t state_bytes,
                                        float* logits_out, int32_t* next_tokens, void* stream_) {
    NSF_REQUIRE(h && tokens && state && next_tokens, "nsf_whisper_decoder_step: null pointer");
    NSF_REQUIRE(pos >= 0 && pos < h->dims.n_text_ctx, "nsf_whisper_decoder_step: pos=%d outside the text context %d", pos, h->dims.n_text_ctx);
    NSF_REQUIRE(((uintptr_t)state & 255) == 0, "nsf_whisper_decoder_step: state must be 256-byte aligned");
    WdState st = wd_carve(h->dims, n_batch, reinterpret_cast<unsigned char*>(state));
    NSF_REQUIRE(state_bytes >= st.total_bytes, "nsf_whisper_decoder_step: state too small");
    cudaStream_t s = (cudaStream_t)stream_;
    wd_set_int_kernel<<<1, 1, 0, s>>>(st.pos, pos);
    int rc = check_launch("wd_set_int_kernel");
    if (rc) return rc;
    return wd_step_impl(h, tokens, st.pos, n_batch, state, state_bytes, logits_out, next_tokens, s);
}

extern "C" int nsf_whisper_decoder_forward(nsf_whisper_decoder* h, const int32_t* tokens, const int32_t* pos_dev, int n_batch, void* state,
                                           int64_t state_bytes, float* logits_out, void* stream_) {
    NSF_REQUIRE(h && tokens && pos_dev && state && logits_out, "nsf_whisper_decoder_forward: null pointer");
    NSF_REQUIRE(n_batch >= 1, "nsf_whisper_decoder_forward: n_batch=%d", n_batch);
    WdState st = wd_carve(h->dims, n_batch, reinterpret_cast<unsigned char*>(state));
    return wd_step_impl(h, tokens, pos_dev, n_batch, state, state_bytes, logits_out, st.next, (cudaStream_t)stream_);
}

extern "C" int64_t nsf_whisper_decoder_reorder_scratch_bytes(const nsf_whisper_dec_dims* dims, int n_batch) {
    if (!dims || n_batch <= 0) return 0;
    return 2 * (int64_t)dims->n_layers * n_batch * dims->n_heads * dims->n_text_ctx * 64 * 2;
}

extern "C" int nsf_whisper_decoder_reorder(nsf_whisper_decoder* h, const int32_t* src, const int32_t* n_pos_dev, int n_batch, void* state,
                                           int64_t state_bytes, void* scratch, int64_t scratch_bytes, void* stream_) {
    NSF_REQUIRE(h && src && n_pos_dev && state && scratch, "nsf_whisper_decoder_reorder: null pointer");
    const nsf_whisper_dec_dims& D = h->dims;
    NSF_REQUIRE(n_batch >= 1 && ((uintptr_t)state & 255) == 0 && ((uintptr_t)scratch & 15) == 0, "nsf_whisper_decoder_reorder: bad arguments");
    WdState st = wd_carve(D, n_batch, reinterpret_cast<unsigned char*>(state));
    NSF_REQUIRE(state_bytes >= st.total_bytes && scratch_bytes >= nsf_whisper_decoder_reorder_scratch_bytes(&D, n_batch),
                "nsf_whisper_decoder_reorder: state or scratch too small");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t layer_stride = (int64_t)n_batch * D.n_heads * D.n_text_ctx * 64;
    uint16_t* tk = reinterpret_cast<uint16_t*>(scratch);
    uint16_t* tv = tk + (size_t)D.n_layers * layer_stride;
    const dim3 grid((unsigned)(n_batch * D.n_heads), (unsigned)D.n_layers);
    wd_cache_gather_kernel<<<grid, 256, 0, s>>>(st.sk, tk, src, n_pos_dev, D.n_heads, D.n_text_ctx, layer_stride, 0);
    wd_cache_gather_kernel<<<grid, 256, 0, s>>>(st.sv, tv, src, n_pos_dev, D.n_heads, D.n_text_ctx, layer_stride, 0);
    wd_cache_gather_kernel<<<grid, 256, 0, s>>>(tk, st.sk, src, n_pos_dev, D.n_heads, D.n_text_ctx, layer_stride, 1);
    wd_cache_gather_kernel<<<grid, 256, 0, s>>>(tv, st.sv, src, n_pos_dev, D.n_heads, D.n_text_ctx, layer_stride, 1);
    return check_launch("wd_cache_gather_kernel");
}

extern "C" int nsf_whisper_decoder_step_dev(nsf_whisper_decoder* h, int32_t* cur_tokens, int32_t* pos_dev, int n_batch, void* state,
                                            int64_t state_bytes, const int32_t* forced, int total_len, int eot, int32_t* out_tokens,
                                            int32_t* argmaxes, uint8_t* done, void* stream_) {
    NSF_REQUIRE(h && cur_tokens && pos_dev && state && out_tokens && argmaxes && done, "nsf_whisper_decoder_step_dev: null pointer");
    NSF_REQUIRE(n_batch >= 1 && total_len >= 1 && total_len <= h->dims.n_text_ctx, "nsf_whisper_decoder_step_dev: bad sizes");
    WdState st = wd_carve(h->dims, n_batch, reinterpret_cast<unsigned char*>(state));
    cudaStream_t s = (cudaStream_t)stream_;
    int rc = wd_step_impl(h, cur_tokens, pos_dev, n_batch, state, state_bytes, nullptr, st.next, s);
    if (rc) return rc;
    wd_advance_kernel<<<1, 1024, 0, s>>>(st.next, forced, total_len, eot, cur_tokens, out_tokens, argmaxes, done, pos_dev, n_batch);
    return check_launch("wd_advance_kernel");
}

static int wd_check_rules(const nsf_whisper_rules* r, const int32_t* suppress, const int32_t* suppress_first, int vocab) {
    NSF_REQUIRE(r, "whisper rules: null pointer");
    NSF_REQUIRE(r->sample_begin >= 1 && r->n_suppress >= 0 && r->n_suppress_first >= 0, "whisper rules: bad sizes");
    NSF_REQUIRE((r->n_suppress == 0 || suppress) && (r->n_suppress_first == 0 || suppress_first), "whisper rules: missing suppress list");
    NSF_REQUIRE(r->timestamp_begin < vocab && (r->timestamp_begin < 0 || (r->eot >= 0 && r->eot <= r->timestamp_begin)),
                "whisper rules: need eot <= timestamp_begin < vocab");
    return NSF_OK;
}

extern "C" int nsf_whisper_logit_rules(float* logits, int n_batch, int vocab, const int32_t* tokens, int total_len, const int32_t* pos_dev,
                                       const nsf_whisper_rules* rules, const int32_t* suppress, const int32_t* suppress_first, void* stream_) {
    NSF_REQUIRE(logits && tokens && pos_dev && n_batch >= 1 && vocab >= 1 && total_len >= 1, "nsf_whisper_logit_rules: bad arguments");
    int rc = wd_check_rules(rules, suppress, suppress_first, vocab);
    if (rc) return rc;
    wd_logit_rules_kernel<<<n_batch, 1024, 0, (cudaStream_t)stream_>>>(logits, vocab, tokens, total_len, pos_dev, *rules, suppress, suppress_first);
    return check_launch("wd_logit_rules_kernel");
}

extern "C" int nsf_whisper_decoder_step_rules(nsf_whisper_decoder* h, int32_t* cur_tokens, int32_t* pos_dev, int n_batch, void* state,
                                              int64_t state_bytes, const int32_t* forced, int total_len, int eot, int32_t* out_tokens,
                                              int32_t* argmaxes, uint8_t* done, const nsf_whisper_rules* rules, const int32_t* suppress,
                                              const int32_t* suppress_first, float* xattn_probs, const int32_t* align_map, int n_align,
                                              void* stream_) {
    NSF_REQUIRE(h && cur_tokens && pos_dev && state && out_tokens && argmaxes && done, "nsf_whisper_decoder_step_rules: null pointer");
    NSF_REQUIRE(n_batch >= 1 && total_len >= 1 && total_len <= h->dims.n_text_ctx, "nsf_whisper_decoder_step_rules: bad sizes");
    NSF_REQUIRE(!xattn_probs || (align_map && n_align >= 1), "nsf_whisper_decoder_step_rules: capture needs the alignment-head map");
    int rc = rules ? wd_check_rules(rules, suppress, suppress_first, h->dims.vocab) : NSF_OK;
    if (rc) return rc;
    WdState st = wd_carve(h->dims, n_batch, reinterpret_cast<unsigned char*>(state));
    cudaStream_t s = (cudaStream_t)stream_;
    const WdRulesArgs ra = {rules, suppress, suppress_first, out_tokens, total_len, xattn_probs, align_map, n_align};
    if ((rc = wd_step_impl(h, cur_tokens, pos_dev, n_batch, state, state_bytes, nullptr, st.next, s, &ra))) return rc;
    wd_advance_kernel<<<1, 1024, 0, s>>>(st.next, forced, total_len, eot, cur_tokens, out_tokens, argmaxes, done, pos_dev, n_batch);
    return check_launch("wd_advance_kernel");
}
