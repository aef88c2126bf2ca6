// Shared helpers for the libnsf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
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

// twiddle table W512^j = (cos(2 pi j/512), sin(2 pi j/512)), j < 512; filled once per process
// from float64 on the host (exact to the last bit of fp32).
const float2* twiddle_table_device();      // returns device pointer (initialises on first use), nullptr on error
const float* hann_table_device();          // periodic hann(512)
const float* sqrt_hann16_table_device();   // sqrt(hann)/16  (feature.py:30-36 with the default 'sqrt_hann' window)

}  // namespace nsf
