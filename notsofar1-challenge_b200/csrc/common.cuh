// Shared helpers for the libnsf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/nsf_b200.h"

namespace nsf {

constexpr int kBins = NSF_NUM_BINS;     // 257
constexpr int kFrame = NSF_FRAME_LEN;   // 512
constexpr int kHop = NSF_FRAME_HOP;     // 256

void set_error(const char* fmt, ...);
void count_launch();

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return NSF_ERR_CUDA;
    }
    count_launch();
    return NSF_OK;
}

// Optional per-kernel-class timing with CUDA events on the launching stream (nsf_prof_* in the C ABI).
// A ProfScope brackets the launches of one class; `work` is the algorithmic work of the bracket
// (bytes for the HBM-bound classes, flops for the GEMM classes).
enum ProfClass : int {
    PROF_STFT = 0, PROF_FEATURES, PROF_GEMM_TC, PROF_GEMM_SIMT, PROF_ATTN, PROF_NET_OTHER, PROF_MVDR, PROF_PIT, PROF_STITCH,
    PROF_ACTIVITY, PROF_ISTFT, PROF_PCM16, PROF_NUM_CLASSES
};
bool prof_enabled();
void prof_begin(int cls, double work, cudaStream_t stream);
void prof_end(int cls, cudaStream_t stream);
struct ProfScope {
    int cls; cudaStream_t stream; bool on;
    ProfScope(int c, double work, cudaStream_t s) : cls(c), stream(s), on(prof_enabled()) { if (on) prof_begin(cls, work, stream); }
    ~ProfScope() { if (on) prof_end(cls, stream); }
};

#define NSF_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::nsf::set_error(__VA_ARGS__);                       \
            return NSF_ERR_INVALID_ARG;                          \
        }                                                        \
    } while (0)

#define NSF_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            ::nsf::set_error("%s: %s", #call, cudaGetErrorString(e__));         \
            return NSF_ERR_CUDA;                                                \
        }                                                                       \
    } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Split an fp32 value into its TF32-representable head (low 13 mantissa bits cleared, so the
// tensor core's own truncation is a no-op) and the exact remainder.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    lo = v - hi;
}

// Storage formats of a "split" activation / weight pair (X_hi, X_lo) that feeds a tensor-core GEMM.
//   SPLIT_TF32  two fp32 arrays: TF32 head + exact remainder                       (3xTF32 / TF32 engines)
//   SPLIT_BF16  two bf16 arrays: hi = bf16(x), lo = bf16(x - hi)                    (~2^-17 relative, fp32 range)
//   SPLIT_F16   two fp16 arrays of the scaled value s x: hi = f16(s x), lo = f16(s x - hi)  (~2^-22 relative; the
//               power-of-two scale keeps the remainder out of the fp16 subnormals: activations s = 2^4, weights
//               s = 2^8, undone exactly in the GEMM epilogue; |s x| saturates at 65504)
// The 16-bit arrays live in the same buffers as the fp32 ones (element index unchanged, half the bytes used).
//   SPLIT_BF16_1 one bf16 array (head only): the plain bf16 operand of the single-pass NSF_GEMM_TC_BF16 engine
//   SPLIT_FP32   one fp32 array (no split): plain outputs written through the same store helpers
//   SPLIT_F16_1  one fp16 array (unscaled, saturating): the operand of a single-pass fp16 GEMM (TitaNet under autocast semantics)
enum SplitFmt : int { SPLIT_TF32 = 0, SPLIT_BF16 = 1, SPLIT_F16 = 2, SPLIT_BF16_1 = 3, SPLIT_FP32 = 4, SPLIT_F16_1 = 5 };
constexpr float kF16ActScale = 16.f, kF16WeightScale = 256.f;

inline int split_fmt_of_engine(int engine) {
    return engine == NSF_GEMM_TC_2XBF16 ? SPLIT_BF16 : engine == NSF_GEMM_TC_2XF16 ? SPLIT_F16
         : engine == NSF_GEMM_TC_BF16 ? SPLIT_BF16_1 : SPLIT_TF32;
}

__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}
__device__ __forceinline__ void split_f16(float v, float scale, uint16_t& hi, uint16_t& lo) {
    const float s = fminf(fmaxf(v * scale, -65504.f), 65504.f);
    const __half h = __float2half_rn(s);
    const __half l = __float2half_rn(s - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}
// one value into a split pair of format `fmt` (warp-uniform); hi/lo are the buffer bases, idx the element index
__device__ __forceinline__ void split_store(int fmt, float* hi, float* lo, size_t idx, float v) {
    if (fmt == SPLIT_TF32) {
        float h, l;
        split_tf32(v, h, l);
        hi[idx] = h;
        lo[idx] = l;
    } else if (fmt == SPLIT_BF16_1) {
        reinterpret_cast<uint16_t*>(hi)[idx] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    } else if (fmt == SPLIT_F16_1) {
        reinterpret_cast<uint16_t*>(hi)[idx] = __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
    } else if (fmt == SPLIT_FP32) {
        hi[idx] = v;
    } else {
        uint16_t h, l;
        if (fmt == SPLIT_BF16) split_bf16(v, h, l); else split_f16(v, kF16ActScale, h, l);
        reinterpret_cast<uint16_t*>(hi)[idx] = h;
        reinterpret_cast<uint16_t*>(lo)[idx] = l;
    }
}
// four consecutive values (idx a multiple of 4, bases 16-byte aligned)
__device__ __forceinline__ void split_store4(int fmt, float* hi, float* lo, size_t idx, float4 v) {
    if (fmt == SPLIT_TF32) {
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi + idx) = h;
        *reinterpret_cast<float4*>(lo + idx) = l;
    } else if (fmt == SPLIT_BF16_1) {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + idx) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    } else if (fmt == SPLIT_F16_1) {
        const __half2 h01 = __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
        const __half2 h23 = __floats2half2_rn(fminf(fmaxf(v.z, -65504.f), 65504.f), fminf(fmaxf(v.w, -65504.f), 65504.f));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + idx) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    } else if (fmt == SPLIT_FP32) {
        *reinterpret_cast<float4*>(hi + idx) = v;
    } else {
        uint16_t h[4], l[4];
        if (fmt == SPLIT_BF16) {
            // packed converts (F2FP) instead of four scalar F2F pairs
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + idx) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo) + idx) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            return;
        } else {
            split_f16(v.x, kF16ActScale, h[0], l[0]); split_f16(v.y, kF16ActScale, h[1], l[1]);
            split_f16(v.z, kF16ActScale, h[2], l[2]); split_f16(v.w, kF16ActScale, h[3], l[3]);
        }
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + idx) =
            make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo) + idx) =
            make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

// eight consecutive values (idx a multiple of 8, bases 16-byte aligned): one 16-byte store per plane for the 16-bit formats
__device__ __forceinline__ void split_store8(int fmt, float* hi, float* lo, size_t idx, const float (&v)[8]) {
    if (fmt == SPLIT_TF32) {
        split_store4(fmt, hi, lo, idx, make_float4(v[0], v[1], v[2], v[3]));
        split_store4(fmt, hi, lo, idx + 4, make_float4(v[4], v[5], v[6], v[7]));
    } else if (fmt == SPLIT_FP32) {
        *reinterpret_cast<float4*>(hi + idx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(hi + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (fmt == SPLIT_BF16_1) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            h[e] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hi) + idx) = make_uint4(h[0], h[1], h[2], h[3]);
    } else if (fmt == SPLIT_F16_1) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __half2 hp = __floats2half2_rn(fminf(fmaxf(v[2 * e], -65504.f), 65504.f), fminf(fmaxf(v[2 * e + 1], -65504.f), 65504.f));
            h[e] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hi) + idx) = make_uint4(h[0], h[1], h[2], h[3]);
    } else if (fmt == SPLIT_BF16) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * e] - __low2float(hp), v[2 * e + 1] - __high2float(hp));
            h[e] = *reinterpret_cast<const uint32_t*>(&hp);
            l[e] = *reinterpret_cast<const uint32_t*>(&lp);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hi) + idx) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(lo) + idx) = make_uint4(l[0], l[1], l[2], l[3]);
    } else {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = fminf(fmaxf(v[2 * e] * kF16ActScale, -65504.f), 65504.f);
            const float b = fminf(fmaxf(v[2 * e + 1] * kF16ActScale, -65504.f), 65504.f);
            const __half2 hp = __floats2half2_rn(a, b);
            const __half2 lp = __floats2half2_rn(a - __low2float(hp), b - __high2float(hp));
            h[e] = *reinterpret_cast<const uint32_t*>(&hp);
            l[e] = *reinterpret_cast<const uint32_t*>(&lp);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hi) + idx) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(lo) + idx) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// twiddle table W512^j = (cos(2 pi j/512), sin(2 pi j/512)), j < 512; filled once per process
// from float64 on the host (exact to the last bit of fp32).
const float2* twiddle_table_device();      // returns device pointer (initialises on first use), nullptr on error
const float* hann_table_device();          // periodic hann(512)
const float* sqrt_hann16_table_device();   // sqrt(hann)/16  (feature.py:30-36 with the default 'sqrt_hann' window)

}  // namespace nsf
