// Mask-weighted MVDR beamformer, one warp per (segment, frequency bin).
//
// Reference: css/css_with_conformer/utils/mvdr_util.py
//   make_wta   :50-55   winner-take-all over {speaker masks, sum of noise masks}; losers -> 1e-10
//   get_mask_scm :58-66 R_j[f] = sum_t m_j[f,t] x[f,t] x[f,t]^H + 1e-15 I          (7x7 Hermitian, 4 of them)
//   make_mvdr  :36-41   N_i = R_noise + sum_{j != i} R_j
//   calc_bfcoeffs :69-75  G = solve(N_i, R_i);  W = G[:,0] / trace(G)   (den[bin 0] += 1e-15)
//   get_bf     :78-80   y_i[f,t] = sum_c conj(W[f,c]) x[c,f,t]
// and the floored-mask multiply of css/css.py:223-227.
//
// The [T, 7] complex slab of a bin streams through shared memory 32 frames at a time, converted to fp64 on the
// way in (fp32 -> fp64 conversions are a slow pipe, so they are done once per sample instead of once per
// covariance entry), the four covariance matrices are accumulated in fp64 by 28 lanes (one per upper-triangle
// entry; only the winner-take-all winner of a frame gets its own update), the three 7x7 complex systems are solved by Gauss-Jordan
// elimination without pivoting (they are Hermitian positive definite) on 21 lanes (one matrix row per lane, pivot rows
// broadcast through shared memory), and the beamformer is applied
// from the staged slab.  fp64 because the noise covariances have condition numbers of 1e5..1e7
// (the reference's own complex64 result is only ~1e-2 accurate there): parity is
// checked against the reference evaluated in complex128.
// Algorithmic HBM bytes per (bin, frame): 7*8 (mix) + 4*4 (masks) + 3*8 (out) = 96 B.
#include "common.cuh"

namespace nsf {

constexpr int kMvdrWarps = 4;
constexpr int kMvdrC = 7;
constexpr int kMvdrS = 3;

__device__ __forceinline__ double2 shfl_d2(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// 1 / d: approximate fp32 reciprocal (one MUFU) + two Newton steps (2^-22 -> 2^-44 -> below fp64 rounding); the IEEE division (a long
// dependent sequence with special-case handling) only outside the fp32 range
__device__ __forceinline__ double rcp_fast(double d) {
    const float df = (float)d;
    if (!(df >= 1e-35f && df <= 1e35f)) return 1.0 / d;
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(df));
    double r = (double)rf;
    r = r * fma(-d, r, 2.0);
    r = r * fma(-d, r, 2.0);
    return r;
}
__device__ __forceinline__ double2 zinv(double2 a) {
    const double d = rcp_fast(a.x * a.x + a.y * a.y);
    return make_double2(a.x * d, -a.y * d);
}

// per-warp shared memory: one 32-frame chunk of the slab in fp64 + its winner-take-all weights, the four covariance
// matrices and the beamformer coefficients (8.8 KB, independent of T: occupancy is set by registers, not by T)
constexpr int kMvdrChunk = 32;
struct MvdrWarpSmem {
    double2 xs[kMvdrChunk * kMvdrC];                     // chunk of the slab, converted to fp64 once per sample
    double wext[kMvdrChunk];                             // winner weight - 1e-10 per frame
    int wmsk[kMvdrChunk];                                // winner bit mask per frame
    uint8_t wlist[kMvdrS + 1][kMvdrChunk];               // per mask: the chunk's frames it wins (bit 7: this list also adds the frame to the total)
    double2 Rm[(kMvdrS + 1) * kMvdrC * kMvdrC];          // covariance matrices
    double2 Wc[kMvdrS * 8];                              // beamformer coefficients
    double2 prow[2][kMvdrS][2 * kMvdrC];                 // pivot rows of the elimination (double buffered)
};

__global__ void __launch_bounds__(kMvdrWarps * 32, 5)
mvdr_kernel(const float* __restrict__ masks, int n_noise, const float2* __restrict__ X, int64_t T_long,
            int64_t T_valid, int64_t seg_first, int T, int hop, int n_bins, float mask_floor,
            float2* __restrict__ Y) {
    constexpr int C = kMvdrC, S = kMvdrS;
    __shared__ __align__(16) MvdrWarpSmem smem_all[kMvdrWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    const int seg = blockIdx.y;
    if (f >= n_bins) return;                 // warp-uniform; no block-level barriers below
    MvdrWarpSmem& sm = smem_all[warp];

    const int64_t st = (seg_first + seg) * (int64_t)hop;
    const int n_ch_total = S + n_noise;
    const float* mseg = masks + ((size_t)seg * n_ch_total * n_bins + f) * T;        // + k * n_bins * T
    const size_t mstride = (size_t)n_bins * T;
    const float2* Xf = X + ((size_t)f * T_long + st) * C;                            // [T][C] slab of this (segment, bin)
    int64_t n_valid64 = (T_valid - st) * C;                                          // samples beyond are the zero padding
    const int n_valid = (int)(n_valid64 < 0 ? 0 : (n_valid64 > (int64_t)T * C ? (int64_t)T * C : n_valid64));

    // ---- A + B. covariance accumulation, 32 frames at a time.
    // make_wta keeps a mask where it equals the maximum over {speakers, summed noise} and puts 1e-10 elsewhere, so
    //   R_k = sum_t w_k(t) P(t) = 1e-10 * sum_t P(t) + sum_{t: k wins} (m_k(t) - 1e-10) P(t),   P(t) = x(t) x(t)^H:
    // per frame only the winner (ties: every mask equal to the maximum) needs its own accumulation.
    // Lane l < 28 owns upper-triangle entry (i, j), i <= j.  The next chunk's samples and masks are fetched into
    // registers while the current chunk is being accumulated.
    int ei = 0, ej = 0;
    {
        int l = lane < 28 ? lane : 0, rowlen = C;
        while (l >= rowlen) { l -= rowlen; ++ei; --rowlen; }
        ej = ei + l;
    }
    double ar[S + 1], ai[S + 1], tr = 0.0, ti = 0.0;
    int wcnt[S + 1] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k <= S; ++k) { ar[k] = 0.0; ai[k] = 0.0; }

    float2 px[C];          // prefetched samples: element lane + 32 q of the chunk's [32][C] block
    float pm[S + 1];       // prefetched masks of frame t0 + lane (noise already summed)
    auto prefetch = [&](int t0) {
#pragma unroll
        for (int q = 0; q < C; ++q) {
            const int j = t0 * C + q * 32 + lane;
            px[q] = (j < n_valid && (q * 32 + lane) < kMvdrChunk * C) ? __ldg(Xf + j) : make_float2(0.f, 0.f);
        }
        const int t = t0 + lane;
#pragma unroll
        for (int k = 0; k <= S; ++k) pm[k] = 0.f;
        if (t < T) {
#pragma unroll
            for (int k = 0; k < S; ++k) pm[k] = __ldg(mseg + k * mstride + t);
            float nz = 0.f;
            for (int k = 0; k < n_noise; ++k) nz += __ldg(mseg + (S + k) * mstride + t);   // noise_masks.sum(axis=0)
            pm[S] = nz;
        }
    };
    prefetch(0);
    for (int t0 = 0; t0 < T; t0 += kMvdrChunk) {
        // stage the prefetched chunk
#pragma unroll
        for (int q = 0; q < C; ++q) sm.xs[q * 32 + lane] = make_double2((double)px[q].x, (double)px[q].y);
        {
            float mx = pm[0];
#pragma unroll
            for (int k = 1; k <= S; ++k) mx = fmaxf(mx, pm[k]);
            int bits = 0;
#pragma unroll
            for (int k = 0; k <= S; ++k) bits |= (pm[k] == mx) ? (1 << k) : 0;               // np.where(mask == mask_max, mask, 1e-10)
            sm.wext[lane] = (double)mx - 1e-10;
            sm.wmsk[lane] = bits;
            // frames sorted by winner: the accumulation below runs one branch-free loop per mask over the frames it wins
            // (a predicated update of all four masks per frame would issue 8 fp64 instructions of which 2 do work)
            if (t0 + lane >= T) bits = 0;
            const int lowest = bits & -bits;
#pragma unroll
            for (int k = 0; k <= S; ++k) {
                const unsigned m = __ballot_sync(0xffffffffu, (bits >> k) & 1);
                if ((bits >> k) & 1) sm.wlist[k][__popc(m & ((1u << lane) - 1u))] = (uint8_t)(lane | ((lowest == (1 << k)) ? 0x80 : 0));
                wcnt[k] = __popc(m);
            }
        }
        __syncwarp();
        if (t0 + kMvdrChunk < T) prefetch(t0 + kMvdrChunk);
        const int nt = min(kMvdrChunk, T - t0);
        (void)nt;
        if (lane < 28) {
#pragma unroll
            for (int k = 0; k <= S; ++k) {
                const int cnt = wcnt[k];                              // warp-uniform trip count
#pragma unroll 4
                for (int n = 0; n < cnt; ++n) {
                    const int e = sm.wlist[k][n];
                    const int t = e & 31;
                    const double2 xi = sm.xs[t * C + ei], xj = sm.xs[t * C + ej];
                    const double pr = xi.x * xj.x + xi.y * xj.y;     // x_i conj(x_j)
                    const double pi = xi.y * xj.x - xi.x * xj.y;
                    const double we = sm.wext[t];
                    ar[k] += we * pr; ai[k] += we * pi;
                    const double first = (e & 0x80) ? 1.0 : 0.0;     // exact ties: the frame enters the total once
                    tr += first * pr; ti += first * pi;
                }
            }
        }
        __syncwarp();
    }
    if (lane < 28) {
#pragma unroll
        for (int k = 0; k <= S; ++k) {
            double rr = ar[k] + 1e-10 * tr, ri = ai[k] + 1e-10 * ti;
            if (ei == ej) { rr += 1e-15; ri = 0.0; }             // Ri += 1e-15 * I
            sm.Rm[(k * C + ei) * C + ej] = make_double2(rr, ri);
            if (ei != ej) sm.Rm[(k * C + ej) * C + ei] = make_double2(rr, -ri);
        }
    }
    __syncwarp();
    double2* Rm = sm.Rm;
    double2* Wc = sm.Wc;

    // ---- C. three Gauss-Jordan solves, lane = (speaker s = lane / 8, row r = lane % 8).  The matrices are Hermitian
    // positive definite (non-negative combinations of x x^H plus 1e-15 I): elimination is stable without pivoting, the
    // pivot of step k is row k and sits in a lane known at compile time.
    {
        const int s = lane >> 3, r = lane & 7;
        const bool act = (s < S) && (r < C);
        double2 row[2 * C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double2 other = make_double2(0.0, 0.0), tgt = make_double2(0.0, 0.0), noi = make_double2(0.0, 0.0);
            if (act) {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const double2 v = Rm[(j * C + r) * C + c];
                    if (j == s) tgt = v; else { other.x += v.x; other.y += v.y; }
                }
                noi = Rm[(S * C + r) * C + c];
            } else if (r == c) {
                noi = make_double2(1.0, 0.0);                            // idle lanes carry an identity row
            }
            row[c] = make_double2(noi.x + other.x, noi.y + other.y);     // noise_scm + other_spks_scm
            row[C + c] = tgt;
        }
        double2 mypiv = make_double2(1.0, 0.0);
        const int sb = s < S ? s : 0;                                // the idle group follows speaker 0 (finite values, unused)
#pragma unroll
        for (int k = 0; k < C; ++k) {
            // Gauss-Jordan without normalising the pivot row: row_i -= (row_i[k] / piv) row_k for every other row (rows that
            // were pivots before included); the division by the pivots happens once at the end.  The pivot row goes
            // through shared memory (one 16-byte broadcast load per column instead of four shuffles).
            const bool is_p = (r == k);
            double2 (&prow)[S][2 * C] = sm.prow[k & 1];
            if (is_p && act) {
#pragma unroll
                for (int c = k; c < 2 * C; ++c) prow[s][c] = row[c];
            }
            __syncwarp();
            const double2 g = zmul(row[k], zinv(prow[sb][k]));
            if (is_p) mypiv = row[k];
#pragma unroll
            for (int c = k + 1; c < 2 * C; ++c) {
                const double2 prc = prow[sb][c];
                if (!is_p) { row[c].x -= g.x * prc.x - g.y * prc.y; row[c].y -= g.x * prc.y + g.y * prc.x; }
            }
        }
        // lane r holds row r of N^-1 R, still scaled by its pivot, in row[C..2C)
        const double2 minv = zinv(mypiv);
        double2 gd = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < C; ++c) if (act && r == c) gd = zmul(row[C + c], minv);
        double2 tr = gd;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            tr.x += __shfl_xor_sync(0xffffffffu, tr.x, o);
            tr.y += __shfl_xor_sync(0xffffffffu, tr.y, o);
        }
        if (f == 0) tr.x += 1e-15;                               // den[0] += 1e-15, mvdr_util.py:73
        if (act) Wc[s * 8 + r] = zmul(zmul(row[C], minv), zinv(tr));     // W[c] = G[c][0] / trace(G)
    }
    __syncwarp();

    // ---- D. apply: y_s[t] = sum_c conj(W_s[c]) x_c[t], then the floored-mask multiply (css.py:223-227).
    // The coefficients come out of the fp64 solve and are rounded to fp32 once; the 7-term sums run on the fp32 pipe
    // (the reference applies in complex64 too, mvdr_util.py:78-80), which keeps the fp64 pipe -- the kernel's bottleneck --
    // for the covariances and the solves.  The slab is read a second time (L2-resident), one frame per lane.
    float2 wc[S][C];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < C; ++c) wc[s][c] = make_float2((float)Wc[s * 8 + c].x, (float)Wc[s * 8 + c].y);
    for (int t = lane; t < T; t += 32) {
        float yr[S], yi[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { yr[s] = 0.f; yi[s] = 0.f; }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float2 x = (t * C + c < n_valid) ? __ldg(Xf + t * C + c) : make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                yr[s] = fmaf(wc[s][c].x, x.x, fmaf(wc[s][c].y, x.y, yr[s]));
                yi[s] = fmaf(wc[s][c].x, x.y, fmaf(-wc[s][c].y, x.x, yi[s]));
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float mk = fmaxf(__ldg(mseg + s * mstride + t), mask_floor);      // torch.clip(mask, min=floor)
            Y[(((size_t)seg * S + s) * n_bins + f) * T + t] = make_float2(yr[s] * mk, yi[s] * mk);
        }
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int nsf_mvdr(const float* masks, int n_spk, int n_noise, const float* X, int64_t T_long, int64_t T_valid,
                        int n_ch, int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor, float* Y,
                        void* stream) {
    NSF_REQUIRE(masks && X && Y, "nsf_mvdr: null pointer");
    if (n_spk != kMvdrS || n_ch != kMvdrC) {
        set_error("nsf_mvdr: only n_spk=3, n_ch=7 are built (got %d, %d)", n_spk, n_ch);
        return NSF_ERR_UNSUPPORTED;
    }
    NSF_REQUIRE(n_noise >= 1 && n_noise <= 4, "nsf_mvdr: n_noise=%d", n_noise);
    NSF_REQUIRE(T >= 1 && n_bins >= 1 && hop >= 1 && T_valid <= T_long, "nsf_mvdr: bad sizes");
    if (n_seg <= 0) return NSF_OK;
    dim3 grid(ceil_div(n_bins, kMvdrWarps), n_seg);
    // algorithmic bytes: 7*8 mix + (S+Nn)*4 masks + S*8 out per (bin, frame)  (96 B for S = 3, Nn = 1)
    ProfScope prof(PROF_MVDR, (double)n_seg * n_bins * T * (kMvdrC * 8.0 + (kMvdrS + n_noise) * 4.0 + kMvdrS * 8.0), (cudaStream_t)stream);
    mvdr_kernel<<<grid, kMvdrWarps * 32, 0, (cudaStream_t)stream>>>(masks, n_noise, reinterpret_cast<const float2*>(X), T_long,
                                                                      T_valid, seg_first, T, hop, n_bins, mask_floor,
                                                                      reinterpret_cast<float2*>(Y));
    return check_launch("mvdr_kernel");
}
