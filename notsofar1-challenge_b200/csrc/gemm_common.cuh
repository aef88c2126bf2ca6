// GEMM problem description and the fused epilogues shared by the two engines
// (gemm_simt.cu: CUDA-core fp32, gemm_tc.cu: tcgen05 kind::tf32 with TMEM accumulators).
//
//   D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]        A: [batch][M][K], B: [batch][N][K], K contiguous
//
// Operands are stored "split": X_hi holds the TF32-representable head of every fp32 value (low 13
// mantissa bits zero), X_lo the exact remainder, X_hi + X_lo == X.  The tensor-core engine computes
// A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32, ~2^-21 relative) or A_hi B_hi alone; the CUDA-core
// engine multiplies the re-assembled fp32 values.
#pragma once
#include "common.cuh"

namespace nsf {

enum GemmEpilogue : int {
    EPI_STORE = 0,        // out0[b][m][n] = acc + bias[n]
    EPI_RELU_SPLIT = 1,   // v = relu(acc + bias[n]); (out0, out1)[m][n] = split(v)
    EPI_RESID = 2,        // out0[m][n] += alpha * (acc + bias[n])                 (residual stream, in place)
    EPI_QKV = 3,          // scatter q, k (head-major, split) and v (head-major, transposed, split)
    EPI_PV = 4,           // attention output of batch (seg, head) -> (out0, out1)[seg*T + m][head*d_k + n] split
    EPI_MASK = 5,         // out0[seg][n / 257][n % 257][t] = sigmoid(acc + bias[n]),  m = seg*T + t, n < n_valid
    EPI_GELU_SPLIT = 6,   // v = gelu(acc + bias[n]) (exact, erf); (out0, out1)[b][m][n] = split(v), batch stride o_batch_stride
    EPI_GELU_POS = 7,     // out0[b*M + m][n] = gelu(acc + bias[n]) + pos[m][n]   (fp32; pos = out1, row pitch ldo)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

struct GemmParams {
    const float* A_hi; const float* A_lo; int64_t lda; int64_t a_batch_stride;
    const float* B_hi; const float* B_lo; int64_t ldb; int64_t b_batch_stride;
    int M, N, K, batch;
    int op_fmt;               // SplitFmt of A and B (the 16-bit formats reinterpret the float pointers as uint16_t arrays)
    int out_fmt;              // SplitFmt of the split outputs of EPI_RELU_SPLIT / EPI_PV
    int qkv_fmt;              // SplitFmt of q, k, v^T written by EPI_QKV (SPLIT_TF32 for attention.cu, SPLIT_BF16 for attention16.cu)
    int v_rowmajor;           // EPI_QKV: v is written like q and k ([seg][head][t][d_k], into vt_hi / vt_lo) instead of transposed
    int b_shared;             // batched A against one B (weights): every batch entry reads B[0]
    int vec8;                 // set by the tensor-core launcher: N, leading dimensions and bases allow 8-column vector epilogues
    float acc_scale;          // accumulator scale applied before the bias (undoes the SPLIT_F16 operand scales; 1 otherwise)
    int n_valid;              // columns >= n_valid are padding (weights padded with zero rows)
    const float* bias;        // [n_valid] or nullptr
    int epi;
    float alpha;
    float* out0; float* out1; int64_t ldo; int64_t o_batch_stride;
    // geometry for the scatter epilogues
    int T, Tp, n_heads, d_k, d_model;
    float* q_hi; float* q_lo; float* k_hi; float* k_lo; float* vt_hi; float* vt_lo;
    // LayerNorm folded into the GEMMs on either side of it (tensor-core engines, vector epilogue; conformer.cu):
    //   producer (EPI_RESID): the updated residual rows also go out as raw split planes (ln_hi, ln_lo; pitch ldo, format
    //   out_fmt) -- the A operand of the next GEMM -- with per-row partial (sum, sum of squares) in ln_part[m][ln_slots]
    //   (slot 2 * n_blk + epilogue half);
    //   consumer (EPI_QKV, EPI_RELU_SPLIT): v = rstd[m] (acc - mean[m] csum[n]) + bias[n] with (mean, rstd) = ln_stats[m],
    //   csum[n] = sum_k of the gamma-scaled weights as stored, bias[n] = b[n] + sum_k beta[k] W[n][k].
    float* ln_hi; float* ln_lo; float2* ln_part; int ln_slots;
    const float2* ln_stats; const float* ln_csum;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// One output element.  (b, m, n) are in range: m < M, n < n_valid.
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, int b, int m, int n, float acc) {
    const float v = acc * p.acc_scale + (p.bias ? __ldg(p.bias + n) : 0.f);
    switch (p.epi) {
        case EPI_STORE:
            p.out0[(size_t)b * p.o_batch_stride + (size_t)m * p.ldo + n] = v;
            break;
        case EPI_RELU_SPLIT: {
            float hi, lo;
            split_tf32(fmaxf(v, 0.f), hi, lo);
            const size_t o = (size_t)m * p.ldo + n;
            p.out0[o] = hi;
            p.out1[o] = lo;
            break;
        }
        case EPI_RESID: {
            const size_t o = (size_t)m * p.ldo + n;
            p.out0[o] = p.out0[o] + p.alpha * v;
            break;
        }
        case EPI_QKV: {
            const int which = n / p.d_model, c = n - which * p.d_model;
            const int h = c / p.d_k, d = c - h * p.d_k;
            const int seg = m / p.T, t = m - seg * p.T;
            if (which < 2 || p.v_rowmajor) {
                const size_t o = (((size_t)seg * p.n_heads + h) * p.T + t) * p.d_k + d;
                split_store(p.qkv_fmt, which == 0 ? p.q_hi : which == 1 ? p.k_hi : p.vt_hi, which == 0 ? p.q_lo : which == 1 ? p.k_lo : p.vt_lo, o, v);
            } else {
                const size_t o = (((size_t)seg * p.n_heads + h) * p.d_k + d) * p.Tp + t;
                split_store(p.qkv_fmt, p.vt_hi, p.vt_lo, o, v);
            }
            break;
        }
        case EPI_PV: {
            const int seg = b / p.n_heads, h = b - seg * p.n_heads;
            float hi, lo;
            split_tf32(v, hi, lo);
            const size_t o = ((size_t)seg * p.T + m) * p.ldo + (size_t)h * p.d_k + n;
            p.out0[o] = hi;
            p.out1[o] = lo;
            break;
        }
        case EPI_GELU_SPLIT:
            split_store(p.out_fmt, p.out0, p.out1, (size_t)b * p.o_batch_stride + (size_t)m * p.ldo + n, gelu_erf(v));
            break;
        case EPI_GELU_POS:
            p.out0[((size_t)b * p.M + m) * p.ldo + n] = gelu_erf(v) + __ldg(p.out1 + (size_t)m * p.ldo + n);
            break;
        case EPI_MASK: {
            const int seg = m / p.T, t = m - seg * p.T;
            const int k = n / kBins, f = n - k * kBins;
            const int n_masks = p.n_valid / kBins;
            // torch.sigmoid(m), conformer.py:304
            p.out0[(((size_t)seg * n_masks + k) * kBins + f) * p.T + t] = 1.f / (1.f + expf(-v));
            break;
        }
        default: break;
    }
}

int gemm_simt_launch(const GemmParams& p, cudaStream_t stream);
// mode: 3 -> 3xTF32, 1 -> single TF32 pass, 16 -> three kind::f16 MMAs on 16-bit pairs (p.op_fmt = SPLIT_BF16 / SPLIT_F16),
//       116 -> one kind::f16 MMA on plain bf16 operands (p.op_fmt = SPLIT_BF16_1)
int gemm_tc_launch(const GemmParams& p, int mode, cudaStream_t stream);

// fused relative-position attention (attention.cu); q, k: [n_seg*heads][T][64], vt: [n_seg*heads][64][Tp], all split
int attn_fused_launch(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo, const float* vt_hi,
                      const float* vt_lo, const float* pe_hi, const float* pe_lo, int maxlen, int n_seg, int n_heads, int T, int Tp,
                      float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream);
// the same on bf16 head + remainder pairs (attention16.cu); the float pointers are reinterpreted as uint16_t arrays
int attn16_launch(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo, const float* vt_hi,
                  const float* vt_lo, const float* pe_hi, const float* pe_lo, int maxlen, int n_seg, int n_heads, int T, int Tp,
                  float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream);
// non-causal attention with online softmax on plain bf16 operands (flash_attn.cu); q, k, v [n_bh][T][64]
int flash_attn_launch(const void* q, const void* k, const void* vt, int n_batch, int n_heads, int T,
                      float* out_hi, float* out_lo, int64_t ldo, int out_fmt, cudaStream_t stream);
// y1 = LN(x; g1, b1) [relu]; optional fp32 store to out_x; y2 = LN(y1; g2, b2) if g2; optional split store in fmt
// (conformer.cu; one warp per row, d a multiple of 32, vector path for d = 128 * {1,2,3,4,5,6,8,10})
int ln_launch(const float* x, int M, int d, const float* g1, const float* b1, int relu1, float* out_x, const float* g2,
              const float* b2, float* out_hi, float* out_lo, int fmt, cudaStream_t s);
inline bool attn_fused_supported(int T, int d_k) { return T >= 2 && T <= 192 && d_k == 64; }

inline int gemm_launch(int engine, const GemmParams& p, cudaStream_t stream) {
    // algorithmic flops: 2 M N K per batch entry (the three TF32 passes of 3xTF32 count once)
    ProfScope prof(engine == NSF_GEMM_SIMT_FP32 ? PROF_GEMM_SIMT : PROF_GEMM_TC, 2.0 * p.M * p.N * p.K * p.batch, stream);
    if (engine == NSF_GEMM_SIMT_FP32) return gemm_simt_launch(p, stream);
    if (engine == NSF_GEMM_TC_2XBF16 || engine == NSF_GEMM_TC_2XF16) return gemm_tc_launch(p, 16, stream);
    if (engine == NSF_GEMM_TC_BF16) return gemm_tc_launch(p, 116, stream);
    return gemm_tc_launch(p, engine == NSF_GEMM_TC_3XTF32 ? 3 : 1, stream);
}

}  // namespace nsf
