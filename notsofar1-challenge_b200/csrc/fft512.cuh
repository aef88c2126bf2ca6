// 512-point complex FFT on 64 threads (8 values per thread, radix 8x8x8, two shared-memory
// exchanges).  Used by the STFT (two real channels packed per transform) and the iSTFT (two
// Hermitian-symmetrised spectra packed per transform).
//
// Index algebra (n = 64a + 8b + c, k = k0 + 8k1 + 64k2, W = exp(SIGN * 2 pi i / 512)):
//   pass 1, thread (b,c):   A1[k0]  = sum_a x[64a+8b+c] W8^(a k0)           then *= W64^(b k0)
//   pass 2, thread (k0,c):  A2[k1]  = sum_b A1[k0;b,c]  W8^(b k1)           then *= W512^(c (k0+8k1))
//   pass 3, thread (k0,k1): X[k0+8k1+64k2] = sum_c A2[k0,k1;c] W8^(c k2)
// Shared-memory exchange layouts (separate re/im float planes, 72 words per k0 group) are
// bank-conflict free for both the stores and the loads:
//   exchange 1: idx = 72 k0 + 8 b + c       exchange 2: idx = 72 k0 + 9 k1 + c
#pragma once
#include <cuda_runtime.h>

namespace nsf {

constexpr int kFftScratchFloats = 2 * 8 * 72;   // re plane + im plane per transform

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by (SIGN * i)
template <int SIGN>
__host__ __device__ __forceinline__ float2 cmul_i(float2 a) {
    return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <int SIGN>
__host__ __device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_i<SIGN>(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

// in-place: v[k] <- sum_a v[a] exp(SIGN * 2 pi i a k / 8)
template <int SIGN>
__host__ __device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float c = 0.70710678118654752440f;
    float2 s0 = cadd(v[0], v[4]), d0 = csub(v[0], v[4]);
    float2 s1 = cadd(v[1], v[5]), d1 = csub(v[1], v[5]);
    float2 s2 = cadd(v[2], v[6]), d2 = csub(v[2], v[6]);
    float2 s3 = cadd(v[3], v[7]), d3 = csub(v[3], v[7]);
    d1 = cmul(d1, make_float2(c, SIGN * c));
    d2 = cmul_i<SIGN>(d2);
    d3 = cmul(d3, make_float2(-c, SIGN * c));
    dft4<SIGN>(s0, s1, s2, s3);
    dft4<SIGN>(d0, d1, d2, d3);
    v[0] = s0; v[2] = s1; v[4] = s2; v[6] = s3;
    v[1] = d0; v[3] = d1; v[5] = d2; v[7] = d3;
}

#ifdef __CUDACC__
// tw: shared-memory copy of the 512-entry table (cos, sin)(2 pi j / 512).
template <int SIGN>
__device__ __forceinline__ float2 twiddle(const float2* tw, int j) {
    float2 w = tw[j & 511];
    if (SIGN < 0) w.y = -w.y;
    return w;
}

// One transform by a group of 64 threads (lane64 = thread index inside the group, 0..63).
// in:  v[a] = x[64 a + lane64]      out: v[k2] = X[k0 + 8 k1 + 64 k2] with k0 = lane64 / 8, k1 = lane64 % 8.
// scratch: kFftScratchFloats floats private to the group.  sync(): barrier over the 64 threads.
template <int SIGN, typename Sync>
__device__ __forceinline__ void fft512_group(float2 (&v)[8], int lane64, float* scratch, const float2* tw, Sync sync) {
    float* sre = scratch;
    float* sim = scratch + 8 * 72;
    {
        const int b = lane64 >> 3;
        dft8<SIGN>(v);
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) v[k0] = cmul(v[k0], twiddle<SIGN>(tw, 8 * b * k0));   // W64^(b k0) = W512^(8 b k0)
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) {
            sre[72 * k0 + lane64] = v[k0].x;
            sim[72 * k0 + lane64] = v[k0].y;
        }
    }
    sync();
    {
        const int k0 = lane64 >> 3, c = lane64 & 7;
#pragma unroll
        for (int b = 0; b < 8; ++b) v[b] = make_float2(sre[72 * k0 + 8 * b + c], sim[72 * k0 + 8 * b + c]);
        dft8<SIGN>(v);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) v[k1] = cmul(v[k1], twiddle<SIGN>(tw, c * (k0 + 8 * k1)));
        sync();     // all loads of exchange 1 done before it is overwritten
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            sre[72 * k0 + 9 * k1 + c] = v[k1].x;
            sim[72 * k0 + 9 * k1 + c] = v[k1].y;
        }
    }
    sync();
    {
        const int k0 = lane64 >> 3, k1 = lane64 & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = make_float2(sre[72 * k0 + 9 * k1 + c], sim[72 * k0 + 9 * k1 + c]);
        dft8<SIGN>(v);
    }
    sync();         // scratch free for the caller / next transform
}
#endif

}  // namespace nsf
