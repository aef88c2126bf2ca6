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
// The [T, 7] complex slab of a bin and its mask rows are copied into shared memory with cp.async as soon as the warp
// starts (13 KB in flight per warp: the path is HBM-latency bound at one warp per slab unless every byte is requested
// up front).  The covariances are accumulated in fp64 with ONE FRAME PER LANE and all 49 real entries of the Hermitian
// outer product in that lane's registers (98 fused multiply-adds per frame, no idle lanes, no per-entry operand
// traffic); lanes are reduced through shared memory once per matrix.  The three 7x7 complex systems are solved by
// Gauss-Jordan elimination without pivoting (the matrices are Hermitian positive definite) on 21 lanes (one matrix row
// per lane, pivot rows broadcast by warp shuffles), and the beamformer is applied from the staged slab.
// fp64 because the noise covariances have condition numbers of 1e5..1e7 (the reference's own complex64 result is only
// ~1e-2 accurate there): parity is checked against the reference evaluated in complex128.
// Algorithmic HBM bytes per (bin, frame): 7*8 (mix) + 4*4 (masks) + 3*8 (out) = 96 B.
#include "common.cuh"
#include <stdlib.h>

namespace nsf {

constexpr int kMvdrWarps = 2;          // 2 x 21.5 KB of shared memory per CTA: 5 CTAs = 10 warps per SM
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
__device__ __forceinline__ void cp_async8(void* smem, const void* g, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;                        // 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// A Hermitian 7x7 matrix is 49 real numbers: row i holds its diagonal entry and then (re, im) of the entries (i, j > i).
__host__ __device__ constexpr int mvdr_diag(int i) { return 14 * i - i * i; }
__host__ __device__ constexpr int mvdr_re(int i, int j) { return mvdr_diag(i) + 1 + 2 * (j - i - 1); }
constexpr int kMvdrE = 49;
constexpr int kMvdrChunk = 192;          // frames staged / accumulated per pass (T = 186: one pass)
constexpr int kRedPitch = 9;             // cross-lane reduction staging: [entry][8 partial sums], pitch 9 (conflict-free both ways)

// per-warp shared memory (21.2 KB, independent of T)
struct MvdrWarpSmem {
    float2 xs[kMvdrChunk * kMvdrC];                      // the chunk of the slab, [frame][mic]
    float ms[kMvdrS + 1][kMvdrChunk];                    // speaker masks and the summed noise mask of the chunk; after the
                                                         // sort row S holds the winning mask value of each frame
    union {
        double red[kMvdrE * kRedPitch];                                // reduction staging (also used as float[])
        struct {
            double2 Rm[(kMvdrS + 1) * kMvdrC * kMvdrC];                // covariance matrices
            double2 Wc[kMvdrS * 8];                                    // beamformer coefficients
        } s;
    } u;
    double2 prow[2][kMvdrS][2 * kMvdrC];                 // pivot rows of the elimination (double buffered)
    double accS[kMvdrS + 2][kMvdrE];                     // [k <= S]: sum over the frames mask k wins of (m_k - 1e-10) x x^H; [S + 1]: sum over all frames of x x^H
    uint8_t order[kMvdrS + 1][kMvdrChunk];               // per mask: the frames of the chunk it wins, ascending
    int cnt[8];                                          // list lengths ([S + 1]: frames with data)
};

// one frame's outer product into 49 accumulators: acc += (w x) x^H   (98 fused multiply-adds)
template <typename T>
__device__ __forceinline__ void mvdr_outer(T (&acc)[kMvdrE], const T (&xr)[kMvdrC], const T (&xi)[kMvdrC], T w) {
    constexpr int C = kMvdrC;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const T yr = w * xr[i], yi = w * xi[i];
        acc[mvdr_diag(i)] = fma(yr, xr[i], fma(yi, xi[i], acc[mvdr_diag(i)]));
#pragma unroll
        for (int j = i + 1; j < C; ++j) {                 // x_i conj(x_j)
            acc[mvdr_re(i, j)] = fma(yr, xr[j], fma(yi, xi[j], acc[mvdr_re(i, j)]));
            acc[mvdr_re(i, j) + 1] = fma(yi, xr[j], fma(-yr, xi[j], acc[mvdr_re(i, j) + 1]));
        }
    }
}

// out[e] += sum over the 32 lanes of acc[e]: two butterfly levels in registers (49 independent chains), the remaining
// 8 partial sums of every entry through shared memory, added as a tree by lane e
template <typename T>
__device__ __forceinline__ void mvdr_reduce(T (&acc)[kMvdrE], T* red, double* out, int lane) {
    constexpr int E = kMvdrE;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    if (lane < 8) {
#pragma unroll
        for (int e = 0; e < E; ++e) red[e * kRedPitch + lane] = acc[e];
    }
    __syncwarp();
    for (int e = lane; e < E; e += 32) {
        T v[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) v[l] = red[e * kRedPitch + l];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
#pragma unroll
            for (int l = 0; l < o; ++l) v[l] += v[l + o];
        out[e] += (double)v[0];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kMvdrWarps * 32, 5)
mvdr_kernel(const float* __restrict__ masks, int n_noise, const float2* __restrict__ X, int64_t T_long,
            int64_t T_valid, int64_t seg_first, int T, int hop, int n_bins, float mask_floor,
            float2* __restrict__ Y, int phases) {
    constexpr int C = kMvdrC, S = kMvdrS, E = kMvdrE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    const int seg = blockIdx.y;
    if (f >= n_bins) return;                 // warp-uniform; no block-level barriers below
    MvdrWarpSmem& sm = reinterpret_cast<MvdrWarpSmem*>(smem_raw)[warp];

    const int64_t st = (seg_first + seg) * (int64_t)hop;
    const int n_ch_total = S + n_noise;
    const float* mseg = masks + ((size_t)seg * n_ch_total * n_bins + f) * T;        // + k * n_bins * T
    const size_t mstride = (size_t)n_bins * T;
    const float2* Xf = X + ((size_t)f * T_long + st) * C;                            // [T][C] slab of this (segment, bin)
    const int64_t tv64 = T_valid - st;                                               // frames beyond are the zero padding
    const int tv = (int)(tv64 < 0 ? 0 : (tv64 > T ? T : tv64));

    // chunk [c0, c0 + nt) of the slab and of the mask rows -> shared memory; everything requested before anything is awaited
    auto stage = [&](int c0, int nt) {
        const int nx = nt * C, nxv = max(0, min(nt, tv - c0)) * C;
        const float2* src = Xf + (size_t)c0 * C;
        for (int j = lane; j < nx; j += 32) cp_async8(&sm.xs[j], j < nxv ? (const void*)(src + j) : (const void*)X, j < nxv);
        const int rows = n_noise == 1 ? S + 1 : S;
        for (int k = 0; k < rows; ++k)
            for (int tl = lane; tl < nt; tl += 32) cp_async4(&sm.ms[k][tl], mseg + k * mstride + c0 + tl);
        if (n_noise != 1)
            for (int tl = lane; tl < nt; tl += 32) {
                float nz = 0.f;
                for (int k = 0; k < n_noise; ++k) nz += __ldg(mseg + (S + k) * mstride + c0 + tl);   // noise_masks.sum(axis=0)
                sm.ms[S][tl] = nz;
            }
        cp_async_wait_all();
        __syncwarp();
    };

    // ---- A + B. covariance accumulation.
    // make_wta keeps a mask where it equals the maximum over {speakers, summed noise} and puts 1e-10 elsewhere, so
    //   R_k = sum_t w_k(t) P(t) = 1e-10 * sum_t P(t) + sum_{t: k wins} (m_k(t) - 1e-10) P(t),   P(t) = x(t) x(t)^H.
    // Both sums are fp64 (a mask that never wins in a bin has R_k = 1e-10 * total, and the solve amplifies its rounding
    // by the condition number).  For the second sum the frames of a chunk are sorted by winning mask and the 32 lanes
    // are divided among the masks in proportion to the list lengths, so a lane only ever accumulates frames of one mask
    // and one cross-lane reduction per chunk yields all four matrices.
    for (int i = lane; i < (S + 2) * E; i += 32) (&sm.accS[0][0])[i] = 0.0;

    for (int c0 = 0; c0 < T; c0 += kMvdrChunk) {
        const int nt = min(kMvdrChunk, T - c0);
        stage(c0, nt);
        // -- winners of the chunk's frames, sorted by mask
        {
            int cnt[S + 1] = {0, 0, 0, 0};
            for (int r0 = 0; r0 < nt; r0 += 32) {
                const int tl = r0 + lane;
                int bits = 0;
                if (tl < nt) {
                    float pm[S + 1];
#pragma unroll
                    for (int k = 0; k <= S; ++k) pm[k] = sm.ms[k][tl];
                    float mx = pm[0];
#pragma unroll
                    for (int k = 1; k <= S; ++k) mx = fmaxf(mx, pm[k]);
#pragma unroll
                    for (int k = 0; k <= S; ++k) bits |= (pm[k] == mx) ? (1 << k) : 0;           // np.where(mask == mask_max, mask, 1e-10)
                    sm.ms[S][tl] = mx;                       // the noise row is not needed again
                }
#pragma unroll
                for (int k = 0; k <= S; ++k) {
                    const unsigned m = __ballot_sync(0xffffffffu, (bits >> k) & 1);
                    if ((bits >> k) & 1) sm.order[k][cnt[k] + __popc(m & ((1u << lane) - 1u))] = (uint8_t)tl;   // exact ties: every winner
                    cnt[k] += __popc(m);
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k <= S; ++k) sm.cnt[k] = (phases & 2) ? cnt[k] : 0;
                sm.cnt[S + 1] = (phases & 1) ? max(0, min(nt, tv - c0)) : 0;
            }
        }
        __syncwarp();
        // -- list k <= S holds the frames mask k wins (weight m_k - 1e-10), list S + 1 every frame with data (weight 1);
        // 32 frames per round, all lanes reduced after the last round of a list.  The total enters R_k at 1e-10 relative
        // weight: it is summed in fp32 first (fp32 pipe, 1e-7 relative) and only redone in fp64 -- same loop body as the
        // weighted lists -- if some mask's own sum is so small (trace below 1e-2 of the total's) that the total's
        // rounding, amplified by the solve, could reach 1e-6 of the output (a mask that never wins has R_k = 1e-10 total).
#pragma unroll 1
        for (int k = 0; k <= S + 1; ++k) {
            const int nk = sm.cnt[k];
            if (nk == 0) continue;
            if (k == S + 1 && T <= kMvdrChunk) {
                float ta[E];
#pragma unroll
                for (int e = 0; e < E; ++e) ta[e] = 0.f;
#pragma unroll 1
                for (int tl = lane; tl < nk; tl += 32) {
                    float xr[C], xi[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) { const float2 v = sm.xs[tl * C + c]; xr[c] = v.x; xi[c] = v.y; }
                    mvdr_outer<float>(ta, xr, xi, 1.f);
                }
                mvdr_reduce<float>(ta, reinterpret_cast<float*>(sm.u.red), sm.accS[k], lane);
                double trT = 0.0;
                bool need64 = false;
#pragma unroll
                for (int i = 0; i < C; ++i) trT += sm.accS[S + 1][mvdr_diag(i)];
#pragma unroll
                for (int j = 0; j <= S; ++j) {
                    double tj = 0.0;
#pragma unroll
                    for (int i = 0; i < C; ++i) tj += sm.accS[j][mvdr_diag(i)];
                    need64 |= !(tj >= 1e-2 * trT);
                }
                if (!need64) break;
                __syncwarp();
                for (int e = lane; e < E; e += 32) sm.accS[S + 1][e] = 0.0;
                __syncwarp();
            }
            double acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.0;
#pragma unroll 1
            for (int r0 = 0; r0 < nk; r0 += 32) {
                const int idx = r0 + lane;
                const bool pa = idx < nk;
                const int tl = !pa ? 0 : (k <= S ? (int)sm.order[k][idx] : idx);
                const double w = !pa ? 0.0 : (k <= S ? (double)sm.ms[S][tl] - 1e-10 : 1.0);
                double xr[C], xi[C];
#pragma unroll
                for (int c = 0; c < C; ++c) { const float2 v = sm.xs[tl * C + c]; xr[c] = (double)v.x; xi[c] = (double)v.y; }
                mvdr_outer<double>(acc, xr, xi, w);
            }
            mvdr_reduce<double>(acc, sm.u.red, sm.accS[k], lane);
        }
    }
    {
        int ei = 0, ej = 0;                  // lane l < 28 owns upper-triangle entry (i, j), i <= j
        int l = lane < 28 ? lane : 0, rowlen = C;
        while (l >= rowlen) { l -= rowlen; ++ei; --rowlen; }
        ej = ei + l;
        if (lane < 28) {
            const int ire = ei == ej ? mvdr_diag(ei) : mvdr_re(ei, ej);
            const double tr = sm.accS[S + 1][ire], ti = ei == ej ? 0.0 : sm.accS[S + 1][ire + 1];
#pragma unroll
            for (int k = 0; k <= S; ++k) {
                double rr = sm.accS[k][ire] + 1e-10 * tr;
                const double ri = ei == ej ? 0.0 : sm.accS[k][ire + 1] + 1e-10 * ti;
                if (ei == ej) rr += 1e-15;                            // Ri += 1e-15 * I
                sm.u.s.Rm[(k * C + ei) * C + ej] = make_double2(rr, ri);
                if (ei != ej) sm.u.s.Rm[(k * C + ej) * C + ei] = make_double2(rr, -ri);
            }
        }
    }
    __syncwarp();
    double2* Rm = sm.u.s.Rm;
    double2* Wc = sm.u.s.Wc;

    // ---- C. three Gauss-Jordan solves, lane = (speaker s = lane / 8, row r = lane % 8).  The matrices are Hermitian
    // positive definite (non-negative combinations of x x^H plus 1e-15 I): elimination is stable without pivoting, the
    // pivot of step k is row k and sits in a lane known at compile time.
    if (phases & 4) {
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
    // (the reference applies in complex64 too, mvdr_util.py:78-80).  One frame per lane, from the staged slab.
    float2 wc[S][C];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < C; ++c) wc[s][c] = make_float2((float)Wc[s * 8 + c].x, (float)Wc[s * 8 + c].y);
    for (int c0 = 0; c0 < ((phases & 8) ? T : 0); c0 += kMvdrChunk) {
        const int nt = min(kMvdrChunk, T - c0);
        if (T > kMvdrChunk) { __syncwarp(); stage(c0, nt); }          // a single chunk is still resident
        for (int tl = lane; tl < nt; tl += 32) {
            float yr[S], yi[S];
#pragma unroll
            for (int s = 0; s < S; ++s) { yr[s] = 0.f; yi[s] = 0.f; }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float2 x = sm.xs[tl * C + c];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    yr[s] = fmaf(wc[s][c].x, x.x, fmaf(wc[s][c].y, x.y, yr[s]));
                    yi[s] = fmaf(wc[s][c].x, x.y, fmaf(-wc[s][c].y, x.x, yi[s]));
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float mk = fmaxf(sm.ms[s][tl], mask_floor);          // torch.clip(mask, min=floor)
                Y[(((size_t)seg * S + s) * n_bins + f) * T + c0 + tl] = make_float2(yr[s] * mk, yi[s] * mk);
            }
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
    // tools/bench_mvdr.py --phases: time the kernel with phases switched off (1 total, 2 covariances, 4 solves, 8 apply);
    // results are then meaningless.  Never set in production.
    static const int phases = [] { const char* e = getenv("NSF_MVDR_PHASES"); return e ? atoi(e) : 15; }();
    // algorithmic bytes: 7*8 mix + (S+Nn)*4 masks + S*8 out per (bin, frame)  (96 B for S = 3, Nn = 1)
    ProfScope prof(PROF_MVDR, (double)n_seg * n_bins * T * (kMvdrC * 8.0 + (kMvdrS + n_noise) * 4.0 + kMvdrS * 8.0), (cudaStream_t)stream);
    constexpr int smem_bytes = kMvdrWarps * (int)sizeof(MvdrWarpSmem);
    NSF_CUDA(cudaFuncSetAttribute(mvdr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    mvdr_kernel<<<grid, kMvdrWarps * 32, smem_bytes, (cudaStream_t)stream>>>(masks, n_noise, reinterpret_cast<const float2*>(X), T_long,
                                                                      T_valid, seg_first, T, hop, n_bins, mask_floor,
                                                                      reinterpret_cast<float2*>(Y), phases);
    return check_launch("mvdr_kernel");
}
