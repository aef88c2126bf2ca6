// Mask-weighted MVDR beamformer: one warp per frequency bin, covariances / solves in fp64 on chip.
//
// Reference: css/css_with_conformer/utils/mvdr_util.py
//   make_wta   :50-55   winner-take-all over {speaker masks, sum of noise masks}; losers -> 1e-10
//   get_mask_scm :58-66 R_j[f] = sum_t m_j[f,t] x[f,t] x[f,t]^H + 1e-15 I          (7x7 Hermitian, S + 1 of them)
//   make_mvdr  :36-41   N_i = R_noise + sum_{j != i} R_j
//   calc_bfcoeffs :69-75  G = solve(N_i, R_i);  W = G[:,0] / trace(G)   (den[bin 0] += 1e-15)
//   get_bf     :78-80   y_i[f,t] = sum_c conj(W[f,c]) x[c,f,t]
// and the floored-mask multiply of css/css.py:223-227.
//
// make_wta keeps a mask where it equals the maximum over {speakers, summed noise} and puts 1e-10 elsewhere, so with
// P(t) = x(t) x(t)^H
//   R_k = sum_t w_k(t) P(t) = 1e-10 * sum_t P(t) + sum_{t: k wins} (m_k(t) - 1e-10) P(t):
// per frame only the winner (exact ties: every mask equal to the maximum) needs its own accumulation, plus the total.
//
// One kernel template, two work decompositions (see mvdr_kernel below): STREAM (T == 2 hop: a warp walks a run of consecutive
// segments of one bin, every outer product is formed once for the two segments that share the frame) and per-(segment, bin)
// (any T / hop).  Frames are staged in shared memory in fp64, *sorted by winner class* (a counting sort with warp match), so the
// accumulation is one branch-free loop per class with compile-time accumulator registers and sequential shared-memory rows.
// The S 7x7 complex systems are solved by Gauss-Jordan elimination without pivoting (Hermitian positive definite) on S x 7
// lanes, pivot rows broadcast through shared memory; the beamformer is applied in fp32 from L2.  fp64 because the noise
// covariances have condition numbers of 1e5..1e7 (the reference's own complex64 result is only ~1e-2 accurate there; SURVEY.md
// 7.3-1): parity is checked against the reference evaluated in complex128, and against the reference's actual complex64 output
// where that is trustworthy (tests/golden/make_golden_t186.py).
// Algorithmic HBM bytes per (bin, frame): 7*8 (mix) + (S + Nn)*4 (masks) + S*8 (out) = 96 B for S = 3, Nn = 1.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace nsf {

constexpr int kMvdrWarps = 4;
constexpr int kMvdrC = 7;
constexpr int kMvdrHopMax = 96;              // frames per block the streaming kernel stages (hop of 3-s segments: 93)

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

// winner-take-all of one frame (mvdr_util.py:50-55): masks m[k * mstride], k < S, and the summed noise masks; returns the
// bit mask of the masks equal to the maximum (several bits: an exact tie, every tied mask keeps its value)
template <int S>
__device__ __forceinline__ int wta_bits(const float* __restrict__ m, size_t mstride, int n_noise, float& mx) {
    float v[S + 1];
#pragma unroll
    for (int k = 0; k < S; ++k) v[k] = __ldg(m + k * mstride);
    float nz = 0.f;
    for (int k = 0; k < n_noise; ++k) nz += __ldg(m + (S + k) * mstride);          // noise_masks.sum(axis=0)
    v[S] = nz;
    mx = v[0];
#pragma unroll
    for (int k = 1; k <= S; ++k) mx = fmaxf(mx, v[k]);
    int bits = 0;
#pragma unroll
    for (int k = 0; k <= S; ++k) bits |= (v[k] == mx) ? (1 << k) : 0;               // np.where(mask == mask_max, mask, 1e-10)
    return bits;
}

template <int S>
struct MvdrSolveSmem {
    double2 Rm[(S + 1) * kMvdrC * kMvdrC];           // covariance matrices
    double2 Wc[S * 8];                               // beamformer coefficients
    double2 prow[2][S][2 * kMvdrC];                  // pivot rows of the elimination (double buffered)
};

// S Gauss-Jordan solves, lane = (speaker s = lane / 8, row r = lane % 8).  The matrices are Hermitian positive definite
// (non-negative combinations of x x^H plus 1e-15 I): elimination is stable without pivoting, the pivot of step k is row k
// and sits in a lane known at compile time.  Reads sv.Rm, writes sv.Wc; the caller brackets the call with __syncwarp().
template <int S>
__device__ __forceinline__ void mvdr_solve(MvdrSolveSmem<S>& sv, int lane, int f) {
    constexpr int C = kMvdrC;
    const double2* Rm = sv.Rm;
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
#pragma unroll
    for (int k = 0; k < C; ++k) {
        // Gauss-Jordan without normalising the pivot row: row_i -= (row_i[k] / piv) row_k for every other row (rows that
        // were pivots before included); the division by the pivots happens once at the end.  The pivot row goes
        // through shared memory (one 16-byte broadcast load per column instead of four shuffles).
        const bool is_p = (r == k);
        double2 (&prow)[S][2 * C] = sv.prow[k & 1];
        if (is_p && act) {
#pragma unroll
            for (int c = k; c < 2 * C; ++c) prow[s][c] = row[c];
        }
        __syncwarp();
        if (s < S) {                                                 // whole quarter-warps without a system issue no shared-memory loads
            const double2 g = zmul(row[k], zinv(prow[s][k]));
            if (is_p) mypiv = row[k];
#pragma unroll
            for (int c = k + 1; c < 2 * C; ++c) {
                const double2 prc = prow[s][c];
                if (!is_p) { row[c].x -= g.x * prc.x - g.y * prc.y; row[c].y -= g.x * prc.y + g.y * prc.x; }
            }
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
    if (f == 0) tr.x += 1e-15;                                       // den[0] += 1e-15, mvdr_util.py:73
    if (act) sv.Wc[s * 8 + r] = zmul(zmul(row[C], minv), zinv(tr));  // W[c] = G[c][0] / trace(G)
}

// y_s[t] = sum_c conj(W_s[c]) x_c[t], then the floored-mask multiply (css.py:223-227).  The coefficients come out of the
// fp64 solve and are rounded to fp32 once; the 7-term sums run on the fp32 pipe (the reference applies in complex64 too,
// mvdr_util.py:78-80), which keeps the fp64 pipe -- the kernel's bottleneck -- for the covariances and the solves.  The
// slab is read from L2, one frame per lane.  mseg / Yseg point at (segment, mask 0, bin f, frame 0); stride per mask: mstride.
template <int S>
__device__ __forceinline__ void mvdr_apply(const double2* Wc, const float2* __restrict__ Xf, int n_valid, const float* __restrict__ mseg,
                                           size_t mstride, int T, float mask_floor, float2* __restrict__ Yseg, int lane) {
    constexpr int C = kMvdrC;
    float2 wc[S][C];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < C; ++c) wc[s][c] = make_float2((float)Wc[s * 8 + c].x, (float)Wc[s * 8 + c].y);
    for (int t0 = lane; t0 < T; t0 += 64) {                          // two frames per lane in flight (loads of both before the sums)
        float2 x[2][C];
        float mk[2][S];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = t0 + 32 * u;
#pragma unroll
            for (int c = 0; c < C; ++c) x[u][c] = (t < T && t * C + c < n_valid) ? __ldg(Xf + t * C + c) : make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 0; s < S; ++s) mk[u][s] = t < T ? fmaxf(__ldg(mseg + s * mstride + t), mask_floor) : 0.f;      // torch.clip(mask, min=floor)
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = t0 + 32 * u;
            if (t >= T) break;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float yr = 0.f, yi = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    yr = fmaf(wc[s][c].x, x[u][c].x, fmaf(wc[s][c].y, x[u][c].y, yr));
                    yi = fmaf(wc[s][c].x, x[u][c].y, fmaf(-wc[s][c].y, x[u][c].x, yi));
                }
                Yseg[s * mstride + t] = make_float2(yr * mk[u][s], yi * mk[u][s]);
            }
        }
    }
}

// ================================================================================================ the kernel
// Covariance accumulation.  The profile of the round-1 kernel (one covariance entry per lane: two 16-byte shared-memory loads
// per lane and frame for 8 fp64 operations) showed the shared-memory pipe at 88 % of its wavefront peak with the fp64 pipe at
// 24 %: the bound is bytes moved from shared memory into registers, so the entries are register-blocked.  The 21 off-diagonal
// entries of a 7x7 Hermitian matrix are the edges of K7, which the Fano plane splits into 7 triangles {c, c+1, c+3} (mod 7):
// lane c of a 7-lane group loads x_c, x_{c+1}, x_{c+3} (3 loads) and owns the three products between them plus the diagonal
// entry |x_c|^2 -- 7 real accumulators per matrix for 3 loads instead of 2 loads per complex entry -- and the four 7-lane
// groups of a warp (one per quarter-warp, so that every 16-byte shared-memory wavefront serves one row) work on four different
// frames of the same winner class (their partial sums meet in a shuffle reduction when a segment is complete).
template <int S, int NST = kMvdrHopMax>                // NST: frames staged at a time (a multiple of 32)
struct MvdrSmem {
    union {
        double2 xs[NST * kMvdrC];                    // the block, fp64, frames in class-sorted order
        MvdrSolveSmem<S> sv;                         // reused by the solve once the block has been accumulated
    } u;
    double2 w2[NST];                                 // sorted frame -> (winner weight - 1e-10 in the older segment, in the newer one)
    uint8_t rank[NST];                               // frame of the block -> sorted position
    uint8_t tbA[NST], tbB[NST];                      // sorted frame -> winner bit masks (read for the tie class only)
    uint8_t cnt[32];                                 // frames per class
    uint8_t off[32];                                 // exclusive prefix of cnt
};

// winners of the block's frames in the older (A) / newer (B) segment -> class = (winner in A, winner in B) (exact ties: class
// NK*NK), counting sort by class with warp match, then the block's [n_fr][7] samples -> fp64 rows in sorted order (fp32 -> fp64
// conversions are a slow pipe: once per sample).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// asks L2 for the [n_fr][7] samples at Xb and the (S + n_noise) mask rows at m (n_fr floats each) that a later
// mvdr_stage_block call of this warp will read, so that its loads find them on chip
__device__ __forceinline__ void mvdr_prefetch_block(int lane, const float2* Xb, int n_valid, const float* m, size_t mstride, int n_masks, int n_fr) {
    const char* xb = reinterpret_cast<const char*>(Xb);
    for (int o = lane * 128; o < n_valid * (int)sizeof(float2); o += 32 * 128) prefetch_l2(xb + o);
    if (m != nullptr) {
        const int lines = (n_fr * 4 + 127) / 128 + 1;
        for (int i = lane; i < n_masks * lines; i += 32) {
            const int k = i / lines, l = i - k * lines;
            prefetch_l2(reinterpret_cast<const char*>(m + k * mstride) + l * 128);
        }
    }
}

template <int S, int NST>
__device__ __forceinline__ void mvdr_stage_block(MvdrSmem<S, NST>& sm, int lane, bool hasA, bool hasB, const float* __restrict__ mA,
                                                 const float* __restrict__ mB, size_t mstride, int n_noise, int n_fr,
                                                 const float2* __restrict__ Xb, int n_valid) {
    constexpr int C = kMvdrC, NK = S + 1, NCLS = NK * NK;
    constexpr int kRounds = NST / 32;
    sm.cnt[lane] = 0;
    // winner-take-all of every frame in both segments: all loads first, then the sort
    double weA[kRounds], weB[kRounds];
    int cls[kRounds], pic[kRounds], btA[kRounds], btB[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int t = r * 32 + lane;
        cls[r] = 31; btA[r] = 1; btB[r] = 1;
        weA[r] = 0.0; weB[r] = 0.0;
        if (t < n_fr) {
            float mx;
            if (hasA) { btA[r] = wta_bits<S>(mA + t, mstride, n_noise, mx); weA[r] = (double)mx - 1e-10; }
            if (hasB) { btB[r] = wta_bits<S>(mB + t, mstride, n_noise, mx); weB[r] = (double)mx - 1e-10; }
            const bool tie = ((btA[r] & (btA[r] - 1)) | (btB[r] & (btB[r] - 1))) != 0;
            cls[r] = tie ? NCLS : (__ffs(btA[r]) - 1) * NK + (__ffs(btB[r]) - 1);
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int c = cls[r];
        const unsigned grp = __match_any_sync(0xffffffffu, c);
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if (lane == leader && c != 31) { base = sm.cnt[c]; sm.cnt[c] = (uint8_t)(base + __popc(grp)); }
        base = __shfl_sync(0xffffffffu, base, leader);
        pic[r] = base + __popc(grp & ((1u << lane) - 1u));
        __syncwarp();
    }
    {
        const int cv = lane <= NCLS ? sm.cnt[lane] : 0;
        int inc = cv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        sm.off[lane] = (uint8_t)(inc - cv);                          // lanes > NCLS: the number of frames in the block
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int t = r * 32 + lane;
        if (t < n_fr) {
            const int rk = sm.off[cls[r]] + pic[r];
            sm.rank[t] = (uint8_t)rk;
            sm.w2[rk] = make_double2(weA[r], weB[r]);
            sm.tbA[rk] = (uint8_t)btA[r];
            sm.tbB[rk] = (uint8_t)btB[r];
        }
    }
    __syncwarp();
    // samples: a quarter-warp moves one frame (7 channels + one idle lane), i.e. one sorted row of 7 consecutive 16-byte
    // words per shared-memory wavefront: no bank conflicts whatever the permutation
    const int q = lane >> 3, c = lane & 7;
    constexpr int kIters = NST / 4;                                  // 24 for 96 frames
#pragma unroll
    for (int i0 = 0; i0 < kIters; i0 += 8) {
        float2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = (i0 + i) * 4 + q;
            v[i] = (c < C && t * C + c < n_valid) ? __ldg(Xb + t * C + c) : make_float2(0.f, 0.f);   // beyond n_valid: the zero padding
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = (i0 + i) * 4 + q;
            if (c < C && t < n_fr) sm.u.xs[sm.rank[t] * C + c] = make_double2((double)v[i].x, (double)v[i].y);
        }
    }
    __syncwarp();
}

// the 7 real products lane c of a group owns for one frame: |x_a|^2, Re / Im of x_a conj(x_b), x_a conj(x_d), x_b conj(x_d)
__device__ __forceinline__ void fano_products(const double2 xa, const double2 xb, const double2 xd, double (&p)[7]) {
    p[0] = xa.x * xa.x + xa.y * xa.y;
    p[1] = xa.x * xb.x + xa.y * xb.y;  p[2] = xa.y * xb.x - xa.x * xb.y;
    p[3] = xa.x * xd.x + xa.y * xd.y;  p[4] = xa.y * xd.x - xa.x * xd.y;
    p[5] = xb.x * xd.x + xb.y * xd.y;  p[6] = xb.y * xd.x - xb.x * xd.y;
}

template <int S, bool HAS_A, bool HAS_B>
__device__ __forceinline__ void mvdr_accumulate(const MvdrSmem<S>& sm, int lane, double (&aA)[S + 1][7], double (&aB)[S + 1][7], double (&h)[7]) {
    constexpr int C = kMvdrC, NK = S + 1, NCLS = NK * NK;
    const int g = lane >> 3, c = lane & 7;                           // frame slot 0..3 = quarter-warp, Fano line (lane 7 of a quarter: none)
    if (c >= 7) return;
    const int b = c + 1 >= 7 ? c - 6 : c + 1, d = c + 3 >= 7 ? c - 4 : c + 3;
    const double2* xa_p = sm.u.xs + c;
    const double2* xb_p = sm.u.xs + b;
    const double2* xd_p = sm.u.xs + d;
#pragma unroll
    for (int ka = 0; ka < (HAS_A ? NK : 1); ++ka) {
#pragma unroll
        for (int kb = 0; kb < (HAS_B ? NK : 1); ++kb) {
            const int cl = ka * NK + kb;
            const int end = sm.off[cl + 1];
#pragma unroll 2
            for (int n = sm.off[cl] + g; n < end; n += 4) {
                double p[7];
                fano_products(xa_p[n * C], xb_p[n * C], xd_p[n * C], p);
                const double2 w = sm.w2[n];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    h[i] += p[i];
                    if (HAS_A) aA[ka][i] += w.x * p[i];
                    if (HAS_B) aB[kb][i] += w.y * p[i];
                }
            }
        }
    }
    // frames with an exact tie in either segment: every tied mask gets the update (rare; predicated)
    const int end = sm.off[NCLS + 1];
    for (int n = sm.off[NCLS] + g; n < end; n += 4) {
        double p[7];
        fano_products(xa_p[n * C], xb_p[n * C], xd_p[n * C], p);
        const double2 w = sm.w2[n];
        const int bA = sm.tbA[n], bB = sm.tbB[n];
#pragma unroll
        for (int i = 0; i < 7; ++i) h[i] += p[i];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                if (HAS_A && ((bA >> k) & 1)) aA[k][i] += w.x * p[i];
                if (HAS_B && ((bB >> k) & 1)) aB[k][i] += w.y * p[i];
            }
        }
    }
}

// sum of the four frame-slot groups' partials (lane = 8 g + c): every lane ends with the total of its c
__device__ __forceinline__ double group_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}

// a complete segment: acc[k][.] = sum_{t: k wins} (m_k - 1e-10) P(t) (per-group partials), tot[.] = sum_t P(t) (already summed
// over the groups, valid in lanes 0..6) -> covariance matrices in shared memory -> S solves -> beamformer applied
template <int S>
__device__ __forceinline__ void mvdr_finish_segment(MvdrSolveSmem<S>& sv, int lane, int f, const double (&acc)[S + 1][7], const double (&tot)[7],
                                                    const float2* __restrict__ Xf, int n_valid, const float* __restrict__ mseg, size_t mstride,
                                                    int T, float mask_floor, float2* __restrict__ Yseg) {
    constexpr int C = kMvdrC;
    const int c = lane & 7, b = (c + 1) % 7, d = (c + 3) % 7;
#pragma unroll
    for (int k = 0; k <= S; ++k) {
        double r[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) r[i] = group_sum(acc[k][i]) + 1e-10 * tot[i];
        if (lane < 7) {
            double2* R = sv.Rm + k * C * C;
            R[c * C + c] = make_double2(r[0] + 1e-15, 0.0);                           // Ri += 1e-15 * I
            R[c * C + b] = make_double2(r[1], r[2]);  R[b * C + c] = make_double2(r[1], -r[2]);
            R[c * C + d] = make_double2(r[3], r[4]);  R[d * C + c] = make_double2(r[3], -r[4]);
            R[b * C + d] = make_double2(r[5], r[6]);  R[d * C + b] = make_double2(r[5], -r[6]);
        }
    }
    __syncwarp();
    mvdr_solve<S>(sv, lane, f);
    __syncwarp();
    mvdr_apply<S>(sv.Wc, Xf, n_valid, mseg, mstride, T, mask_floor, Yseg, lane);
    __syncwarp();                                                    // Wc is read before the next block is staged over it
}

__device__ __forceinline__ int clamp_valid(int64_t frames_left, int n_fr) {
    const int64_t v = frames_left * kMvdrC;
    return (int)(v < 0 ? 0 : (v > (int64_t)n_fr * kMvdrC ? (int64_t)n_fr * kMvdrC : v));
}

// STREAM (T == 2 hop, the 50 % overlap every shipped configuration uses): neighbouring segments share their frames -- block b =
//   frames [b hop, (b+1) hop) is the second half of segment b-1 (role A) and the first half of segment b (role B).  A warp owns
//   one bin and a run of `run_len` consecutive segments and streams the blocks once: every product is formed once and added to
//   the block total, to the winner's matrix of the older segment and to the winner's matrix of the newer one; a segment's total
//   is the sum of its two block totals.  Loads, conversions, staging and products are shared by the two segments of a frame.
// !STREAM (any T, hop): a warp owns one (segment, bin) and walks the segment in chunks of <= 96 frames (single role).
template <int S, bool STREAM>
__global__ void __launch_bounds__(kMvdrWarps * 32, STREAM ? 2 : 4)
mvdr_kernel(const float* __restrict__ masks, int n_noise, const float2* __restrict__ X, int64_t T_long, int64_t T_valid,
            int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor, float2* __restrict__ Y, int run_len) {
    constexpr int C = kMvdrC, NK = S + 1;
    extern __shared__ __align__(16) unsigned char mvdr_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    if (f >= n_bins) return;                 // warp-uniform; no block-level barriers below
    MvdrSmem<S>& sm = reinterpret_cast<MvdrSmem<S>*>(mvdr_smem_raw)[warp];
    const int n_ch_total = S + n_noise;
    const size_t mstride = (size_t)n_bins * T;
    const float2* Xbin = X + (size_t)f * T_long * C;

    double aA[NK][7], aB[NK][7], h[7], hp[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        h[i] = 0.0; hp[i] = 0.0;
#pragma unroll
        for (int k = 0; k < NK; ++k) { aA[k][i] = 0.0; aB[k][i] = 0.0; }
    }

    if (STREAM) {
        const int s0 = blockIdx.y * run_len, s1 = min(n_seg, s0 + run_len);
        for (int j = 0; j <= s1 - s0; ++j) {
            const bool hasA = j > 0, hasB = s0 + j < s1;
            const int64_t frame0 = (seg_first + s0 + j) * (int64_t)hop;                             // first frame of the block
            const float* mB = masks + ((size_t)(s0 + j) * n_ch_total * n_bins + f) * T;             // newer segment, frames [0, hop)
            const float* mA = mB - (size_t)n_ch_total * n_bins * T + hop;                           // older segment, frames [hop, 2 hop)
            mvdr_stage_block<S, kMvdrHopMax>(sm, lane, hasA, hasB, mA, mB, mstride, n_noise, hop, Xbin + frame0 * C, clamp_valid(T_valid - frame0, hop));
            if (j < s1 - s0)                                         // the next block's samples and the next segment's masks -> L2
                mvdr_prefetch_block(lane, Xbin + (frame0 + hop) * C, clamp_valid(T_valid - frame0 - hop, hop),
                                    s0 + j + 1 < s1 ? mB + (size_t)n_ch_total * n_bins * T : nullptr, mstride, n_ch_total, T);
            // an absent role (first / last block of a run) has weight 0 and winner 0: the same code serves all blocks
            mvdr_accumulate<S, true, true>(sm, lane, aA, aB, h);
            __syncwarp();                                            // xs is dead from here: the solve scratch shares its memory
            double tot[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) { const double hs = group_sum(h[i]); tot[i] = hp[i] + hs; hp[i] = hs; h[i] = 0.0; }
            if (hasA) {
                const int sA = s0 + j - 1;
                const int64_t st = (seg_first + sA) * (int64_t)hop;
                mvdr_finish_segment<S>(sm.u.sv, lane, f, aA, tot, Xbin + st * C, clamp_valid(T_valid - st, T),
                                       masks + ((size_t)sA * n_ch_total * n_bins + f) * T, mstride, T, mask_floor,
                                       Y + ((size_t)sA * S * n_bins + f) * T);
            }
#pragma unroll
            for (int k = 0; k < NK; ++k)
#pragma unroll
                for (int i = 0; i < 7; ++i) { aA[k][i] = aB[k][i]; aB[k][i] = 0.0; }
        }
    } else {
        const int seg = blockIdx.y;
        const int64_t st = (seg_first + seg) * (int64_t)hop;
        const float* mseg = masks + ((size_t)seg * n_ch_total * n_bins + f) * T;
        for (int t0 = 0; t0 < T; t0 += kMvdrHopMax) {
            const int n_fr = min(kMvdrHopMax, T - t0);
            mvdr_stage_block<S, kMvdrHopMax>(sm, lane, false, true, nullptr, mseg + t0, mstride, n_noise, n_fr, Xbin + (st + t0) * C,
                                clamp_valid(T_valid - st - t0, n_fr));
            if (t0 + kMvdrHopMax < T)
                mvdr_prefetch_block(lane, Xbin + (st + t0 + kMvdrHopMax) * C, clamp_valid(T_valid - st - t0 - kMvdrHopMax, min(kMvdrHopMax, T - t0 - kMvdrHopMax)),
                                    nullptr, mstride, 0, 0);
            mvdr_accumulate<S, false, true>(sm, lane, aA, aB, h);
            __syncwarp();
        }
        double tot[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) tot[i] = group_sum(h[i]);
        mvdr_finish_segment<S>(sm.u.sv, lane, f, aB, tot, Xbin + st * C, clamp_valid(T_valid - st, T), mseg, mstride, T, mask_floor,
                               Y + ((size_t)seg * S * n_bins + f) * T);
    }
}

// ------------------------------------------------------------------------------------------------ one long utterance (split-T)
// make_mvdr accepts any T (mvdr_util.py:5-47): one covariance set per bin over the whole utterance.  A warp per (bin, T-chunk)
// accumulates partial covariances (phase 1), a warp per bin adds the partials in chunk order, solves and stores the coefficients
// (phase 2), a warp per (bin, T-chunk) applies them (phase 3).  partial: [n_chunks][n_bins][S + 2][7][7] f64 (lane c's 7 reals of
// every mask + the total), coef: [n_bins][S][8] complex f64.
template <int S>
__global__ void __launch_bounds__(kMvdrWarps * 32, 4)
mvdr_partial_kernel(const float* __restrict__ masks, int n_noise, const float2* __restrict__ X, int64_t T, int n_bins, int chunk,
                    double* __restrict__ partial) {
    constexpr int C = kMvdrC, NK = S + 1;
    extern __shared__ __align__(16) unsigned char mvdr_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    if (f >= n_bins) return;
    MvdrSmem<S>& sm = reinterpret_cast<MvdrSmem<S>*>(mvdr_smem_raw)[warp];
    const int64_t t_lo = (int64_t)blockIdx.y * chunk, t_hi = min(T, t_lo + chunk);
    const size_t mstride = (size_t)n_bins * T;
    const float* mrow = masks + (size_t)f * T;
    const float2* Xbin = X + (size_t)f * T * C;
    double aA[NK][7], aB[NK][7], h[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        h[i] = 0.0;
#pragma unroll
        for (int k = 0; k < NK; ++k) { aA[k][i] = 0.0; aB[k][i] = 0.0; }
    }
    for (int64_t t0 = t_lo; t0 < t_hi; t0 += kMvdrHopMax) {
        const int n_fr = (int)min((int64_t)kMvdrHopMax, t_hi - t0);
        mvdr_stage_block<S, kMvdrHopMax>(sm, lane, false, true, nullptr, mrow + t0, mstride, n_noise, n_fr, Xbin + t0 * C, n_fr * C);
        if (t0 + kMvdrHopMax < t_hi) mvdr_prefetch_block(lane, Xbin + (t0 + kMvdrHopMax) * C, (int)min((int64_t)kMvdrHopMax, t_hi - t0 - kMvdrHopMax) * C, nullptr, mstride, 0, 0);
        mvdr_accumulate<S, false, true>(sm, lane, aA, aB, h);
        __syncwarp();
    }
    double* out = partial + (((size_t)blockIdx.y * n_bins + f) * (S + 2)) * 49;
#pragma unroll
    for (int k = 0; k <= S + 1; ++k) {
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double v = group_sum(k <= S ? aB[k < NK ? k : 0][i] : h[i]);
            if (lane < 7) out[(k * 7 + lane) * 7 + i] = v;
        }
    }
}

template <int S>
__global__ void __launch_bounds__(kMvdrWarps * 32, 4)
mvdr_coef_kernel(const double* __restrict__ partial, int n_chunks, int n_bins, double2* __restrict__ coef) {
    constexpr int C = kMvdrC;
    __shared__ __align__(16) MvdrSolveSmem<S> sv_all[kMvdrWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    if (f >= n_bins) return;
    MvdrSolveSmem<S>& sv = sv_all[warp];
    if (lane < 7) {
        const int c = lane, b = (c + 1) % 7, d = (c + 3) % 7;
        double tot[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) tot[i] = 0.0;
        for (int ch = 0; ch < n_chunks; ++ch) {
            const double* p = partial + ((((size_t)ch * n_bins + f) * (S + 2) + (S + 1)) * 7 + c) * 7;
#pragma unroll
            for (int i = 0; i < 7; ++i) tot[i] += p[i];
        }
        for (int k = 0; k <= S; ++k) {
            double r[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) r[i] = 0.0;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const double* p = partial + ((((size_t)ch * n_bins + f) * (S + 2) + k) * 7 + c) * 7;
#pragma unroll
                for (int i = 0; i < 7; ++i) r[i] += p[i];
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) r[i] += 1e-10 * tot[i];
            double2* R = sv.Rm + k * C * C;
            R[c * C + c] = make_double2(r[0] + 1e-15, 0.0);
            R[c * C + b] = make_double2(r[1], r[2]);  R[b * C + c] = make_double2(r[1], -r[2]);
            R[c * C + d] = make_double2(r[3], r[4]);  R[d * C + c] = make_double2(r[3], -r[4]);
            R[b * C + d] = make_double2(r[5], r[6]);  R[d * C + b] = make_double2(r[5], -r[6]);
        }
    }
    __syncwarp();
    mvdr_solve<S>(sv, lane, f);
    __syncwarp();
    if (lane < S * 8) coef[(size_t)f * S * 8 + lane] = sv.Wc[lane];
}

template <int S>
__global__ void __launch_bounds__(kMvdrWarps * 32, 8)
mvdr_apply_kernel(const double2* __restrict__ coef, const float* __restrict__ masks, const float2* __restrict__ X, int64_t T, int n_bins,
                  int chunk, float mask_floor, float2* __restrict__ Y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    if (f >= n_bins) return;
    const int64_t t_lo = (int64_t)blockIdx.y * chunk;
    const int n = (int)min((int64_t)chunk, T - t_lo);
    const size_t mstride = (size_t)n_bins * T;
    mvdr_apply<S>(coef + (size_t)f * S * 8, X + ((size_t)f * T + t_lo) * kMvdrC, n * kMvdrC, masks + (size_t)f * T + t_lo, mstride, n, mask_floor,
                  Y + (size_t)f * T + t_lo, lane);
}

template <int S>
static int launch_mvdr_utterance(const float* masks, int n_noise, const float2* X, int64_t T, int n_bins, float mask_floor, float2* Y,
                                 void* workspace, cudaStream_t stream) {
    const int chunk = 4 * kMvdrHopMax;                                   // 384 frames per warp
    const int n_chunks = (int)ceil_div64(T, chunk);
    double* partial = reinterpret_cast<double*>(workspace);
    const size_t n_partial = ((size_t)n_chunks * n_bins * (S + 2) * 49 + 1) & ~(size_t)1;          // coef (double2) stays 16-byte aligned
    double2* coef = reinterpret_cast<double2*>(partial + n_partial);
    const size_t smem = sizeof(MvdrSmem<S>) * kMvdrWarps;
    NSF_CUDA(cudaFuncSetAttribute(mvdr_partial_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n_bins, kMvdrWarps), n_chunks);
    mvdr_partial_kernel<S><<<grid, kMvdrWarps * 32, smem, stream>>>(masks, n_noise, X, T, n_bins, chunk, partial);
    int rc = check_launch("mvdr_partial_kernel");
    if (rc) return rc;
    mvdr_coef_kernel<S><<<ceil_div(n_bins, kMvdrWarps), kMvdrWarps * 32, 0, stream>>>(partial, n_chunks, n_bins, coef);
    if ((rc = check_launch("mvdr_coef_kernel"))) return rc;
    mvdr_apply_kernel<S><<<grid, kMvdrWarps * 32, 0, stream>>>(coef, masks, X, T, n_bins, chunk, mask_floor, Y);
    return check_launch("mvdr_apply_kernel");
}

// ------------------------------------------------------------------------------------------------ streaming, one entry per lane
// The same streaming decomposition with the round-1 register layout: lane l < 28 owns upper-triangle entry (i, j) and loads x_i,
// x_j per frame.  Twice the shared-memory bytes per product of the Fano layout, but 20 instead of 63 fp64 accumulators per lane:
// 16..24 resident warps per SM instead of 8, which is what hides the latency of the sort / solve / apply chains of the other warps.
__device__ __forceinline__ void entry_of_lane(int lane, int& ei, int& ej) {
    int l = lane < 28 ? lane : 0, rowlen = kMvdrC;
    ei = 0;
    while (l >= rowlen) { l -= rowlen; ++ei; --rowlen; }
    ej = ei + l;
}

template <int S, int NST>
__device__ __forceinline__ void mvdr_accumulate_entry(const MvdrSmem<S, NST>& sm, int ei, int ej, double (&aAr)[S + 1], double (&aAi)[S + 1],
                                                      double (&aBr)[S + 1], double (&aBi)[S + 1], double& hr, double& hi) {
    constexpr int C = kMvdrC, NK = S + 1, NCLS = NK * NK;
    const double2* xi_p = sm.u.xs + ei;
    const double2* xj_p = sm.u.xs + ej;
#pragma unroll
    for (int ka = 0; ka < NK; ++ka) {
#pragma unroll
        for (int kb = 0; kb < NK; ++kb) {
            const int c = ka * NK + kb;
            const int beg = sm.off[c], end = sm.off[c + 1];          // warp-uniform
#pragma unroll 2
            for (int n = beg; n < end; ++n) {
                const double2 xi = xi_p[n * C], xj = xj_p[n * C];
                const double2 w = sm.w2[n];
                const double pr = xi.x * xj.x + xi.y * xj.y;         // x_i conj(x_j)
                const double pi = xi.y * xj.x - xi.x * xj.y;
                hr += pr; hi += pi;
                aAr[ka] += w.x * pr; aAi[ka] += w.x * pi;
                aBr[kb] += w.y * pr; aBi[kb] += w.y * pi;
            }
        }
    }
    for (int n = sm.off[NCLS]; n < sm.off[NCLS + 1]; ++n) {          // exact ties (rare; predicated)
        const double2 xi = xi_p[n * C], xj = xj_p[n * C];
        const double2 w = sm.w2[n];
        const int bA = sm.tbA[n], bB = sm.tbB[n];
        const double pr = xi.x * xj.x + xi.y * xj.y;
        const double pi = xi.y * xj.x - xi.x * xj.y;
        hr += pr; hi += pi;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            if ((bA >> k) & 1) { aAr[k] += w.x * pr; aAi[k] += w.x * pi; }
            if ((bB >> k) & 1) { aBr[k] += w.y * pr; aBi[k] += w.y * pi; }
        }
    }
}

template <int S, int NST>
__global__ void __launch_bounds__(kMvdrWarps * 32, NST <= 64 ? 5 : 4)
mvdr_stream_entry_kernel(const float* __restrict__ masks, int n_noise, const float2* __restrict__ X, int64_t T_long, int64_t T_valid,
                         int64_t seg_first, int n_seg, int hop, int n_bins, float mask_floor, float2* __restrict__ Y, int run_len, int skip) {
    constexpr int C = kMvdrC, NK = S + 1;
    extern __shared__ __align__(16) unsigned char mvdr_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMvdrWarps + warp;
    if (f >= n_bins) return;                 // warp-uniform; no block-level barriers below
    MvdrSmem<S, NST>& sm = reinterpret_cast<MvdrSmem<S, NST>*>(mvdr_smem_raw)[warp];
    const int s0 = blockIdx.y * run_len, s1 = min(n_seg, s0 + run_len);
    const int T = 2 * hop;
    const int n_ch_total = S + n_noise;
    const size_t mstride = (size_t)n_bins * T;
    const float2* Xbin = X + (size_t)f * T_long * C;
    int ei, ej;
    entry_of_lane(lane, ei, ej);
    double aAr[NK], aAi[NK], aBr[NK], aBi[NK], hr = 0.0, hi = 0.0, hpr = 0.0, hpi = 0.0;
#pragma unroll
    for (int k = 0; k < NK; ++k) { aAr[k] = aAi[k] = aBr[k] = aBi[k] = 0.0; }

    for (int j = 0; j <= s1 - s0; ++j) {
        const bool hasA = j > 0, hasB = s0 + j < s1;
        const int64_t frame0 = (seg_first + s0 + j) * (int64_t)hop;                                 // first frame of the block
        const float* mB = masks + ((size_t)(s0 + j) * n_ch_total * n_bins + f) * T;                 // newer segment, frames [0, hop)
        const float* mA = mB - (size_t)n_ch_total * n_bins * T + hop;                               // older segment, frames [hop, 2 hop)
        for (int h0 = 0; h0 < hop; h0 += NST) {
            const int n_fr = min(NST, hop - h0);
            mvdr_stage_block<S, NST>(sm, lane, hasA, hasB, mA + h0, mB + h0, mstride, n_noise, n_fr, Xbin + (frame0 + h0) * C,
                                     clamp_valid(T_valid - frame0 - h0, n_fr));
            if (h0 + NST < hop)
                mvdr_prefetch_block(lane, Xbin + (frame0 + h0 + NST) * C, clamp_valid(T_valid - frame0 - h0 - NST, min(NST, hop - h0 - NST)), nullptr, mstride, 0, 0);
            else if (j < s1 - s0)                                    // the next block's samples and the next segment's masks -> L2
                mvdr_prefetch_block(lane, Xbin + (frame0 + hop) * C, clamp_valid(T_valid - frame0 - hop, min(NST, hop)),
                                    s0 + j + 1 < s1 ? mB + (size_t)n_ch_total * n_bins * T : nullptr, mstride, n_ch_total, T);
            if (lane < 28 && !(skip & 1)) mvdr_accumulate_entry<S, NST>(sm, ei, ej, aAr, aAi, aBr, aBi, hr, hi);
            __syncwarp();                                            // xs is dead from here: the solve scratch shares its memory
        }
        if (hasA) {
            if (lane < 28) {
                const double tr = hpr + hr, ti = hpi + hi;           // sum_t P(t) over the segment = its two block totals
#pragma unroll
                for (int k = 0; k <= S; ++k) {
                    double rr = aAr[k] + 1e-10 * tr, ri = aAi[k] + 1e-10 * ti;
                    if (ei == ej) { rr += 1e-15; ri = 0.0; }         // Ri += 1e-15 * I
                    sm.u.sv.Rm[(k * C + ei) * C + ej] = make_double2(rr, ri);
                    if (ei != ej) sm.u.sv.Rm[(k * C + ej) * C + ei] = make_double2(rr, -ri);
                }
            }
            __syncwarp();
            if (!(skip & 2)) mvdr_solve<S>(sm.u.sv, lane, f);
            __syncwarp();
            const int sA = s0 + j - 1;
            const int64_t st = (seg_first + sA) * (int64_t)hop;
            if (!(skip & 4))
            mvdr_apply<S>(sm.u.sv.Wc, Xbin + st * C, clamp_valid(T_valid - st, T), masks + ((size_t)sA * n_ch_total * n_bins + f) * T,
                          mstride, T, mask_floor, Y + ((size_t)sA * S * n_bins + f) * T, lane);
            __syncwarp();                                            // Wc is read before the next block is staged over it
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) { aAr[k] = aBr[k]; aAi[k] = aBi[k]; aBr[k] = 0.0; aBi[k] = 0.0; }
        hpr = hr; hpi = hi; hr = 0.0; hi = 0.0;
    }
}

template <int S, int NST>
static int launch_mvdr_entry(const float* masks, int n_noise, const float2* X, int64_t T_long, int64_t T_valid, int64_t seg_first,
                             int n_seg, int hop, int n_bins, float mask_floor, float2* Y, int run_len, cudaStream_t stream) {
    const size_t smem = sizeof(MvdrSmem<S, NST>) * kMvdrWarps;
    NSF_CUDA(cudaFuncSetAttribute(mvdr_stream_entry_kernel<S, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n_bins, kMvdrWarps), ceil_div(n_seg, run_len));
    // NSF_MVDR_SKIP: timing-only switch (bit 0: covariances, 1: solves, 2: apply are skipped; results are then meaningless)
    const char* sk = getenv("NSF_MVDR_SKIP");
    mvdr_stream_entry_kernel<S, NST><<<grid, kMvdrWarps * 32, smem, stream>>>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, hop, n_bins,
                                                                               mask_floor, Y, run_len, sk ? atoi(sk) : 0);
    return check_launch("mvdr_stream_entry_kernel");
}

template <int S, bool STREAM>
static int launch_mvdr_impl(const float* masks, int n_noise, const float2* X, int64_t T_long, int64_t T_valid, int64_t seg_first,
                            int n_seg, int T, int hop, int n_bins, float mask_floor, float2* Y, int run_len, cudaStream_t stream) {
    const size_t smem = sizeof(MvdrSmem<S>) * kMvdrWarps;
    NSF_CUDA(cudaFuncSetAttribute(mvdr_kernel<S, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every launch
    dim3 grid(ceil_div(n_bins, kMvdrWarps), STREAM ? ceil_div(n_seg, run_len) : n_seg);
    mvdr_kernel<S, STREAM><<<grid, kMvdrWarps * 32, smem, stream>>>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, T, hop, n_bins,
                                                                     mask_floor, Y, run_len);
    return check_launch("mvdr_kernel");
}

enum MvdrImpl { MVDR_GENERIC = 0, MVDR_STREAM_FANO, MVDR_STREAM_ENTRY96, MVDR_STREAM_ENTRY64, MVDR_STREAM_ENTRY32 };

template <int S>
static int launch_mvdr(int impl, const float* masks, int n_noise, const float2* X, int64_t T_long, int64_t T_valid, int64_t seg_first,
                       int n_seg, int T, int hop, int n_bins, float mask_floor, float2* Y, int run_len, cudaStream_t stream) {
    switch (impl) {
        case MVDR_STREAM_FANO: return launch_mvdr_impl<S, true>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, T, hop, n_bins, mask_floor, Y, run_len, stream);
        case MVDR_STREAM_ENTRY96: return launch_mvdr_entry<S, 96>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, hop, n_bins, mask_floor, Y, run_len, stream);
        case MVDR_STREAM_ENTRY64: return launch_mvdr_entry<S, 64>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, hop, n_bins, mask_floor, Y, run_len, stream);
        case MVDR_STREAM_ENTRY32: return launch_mvdr_entry<S, 32>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, hop, n_bins, mask_floor, Y, run_len, stream);
        default: return launch_mvdr_impl<S, false>(masks, n_noise, X, T_long, T_valid, seg_first, n_seg, T, hop, n_bins, mask_floor, Y, run_len, stream);
    }
}

}  // namespace nsf

using namespace nsf;

extern "C" int64_t nsf_mvdr_utterance_workspace_bytes(int n_spk, int64_t T, int n_bins) {
    if (n_spk < 2 || n_spk > 4 || T < 1 || n_bins < 1) return 0;
    const int64_t n_chunks = ceil_div64(T, 4 * kMvdrHopMax);
    return ((n_chunks * n_bins * (n_spk + 2) * 49 + 1) & ~(int64_t)1) * 8 + (int64_t)n_bins * n_spk * 8 * 16 + 256;
}

extern "C" int nsf_mvdr_utterance(const float* masks, int n_spk, int n_noise, const float* X, int64_t T, int n_ch, int n_bins, float mask_floor,
                                  float* Y, void* workspace, int64_t workspace_bytes, void* stream) {
    NSF_REQUIRE(masks && X && Y && workspace, "nsf_mvdr_utterance: null pointer");
    if (n_spk < 2 || n_spk > 4 || n_ch != kMvdrC) {
        set_error("nsf_mvdr_utterance: built for 2..4 speaker masks and 7 microphones (got n_spk=%d, n_ch=%d)", n_spk, n_ch);
        return NSF_ERR_UNSUPPORTED;
    }
    NSF_REQUIRE(n_noise >= 1 && n_noise <= 4 && T >= 1 && n_bins >= 1, "nsf_mvdr_utterance: bad sizes");
    NSF_REQUIRE(workspace_bytes >= nsf_mvdr_utterance_workspace_bytes(n_spk, T, n_bins) && ((uintptr_t)workspace & 15) == 0,
                "nsf_mvdr_utterance: workspace too small or misaligned");
    ProfScope prof(PROF_MVDR, (double)n_bins * T * (kMvdrC * 8.0 + (n_spk + n_noise) * 4.0 + n_spk * 8.0), (cudaStream_t)stream);
    const float2* Xc = reinterpret_cast<const float2*>(X);
    float2* Yc = reinterpret_cast<float2*>(Y);
    switch (n_spk) {
        case 2: return launch_mvdr_utterance<2>(masks, n_noise, Xc, T, n_bins, mask_floor, Yc, workspace, (cudaStream_t)stream);
        case 3: return launch_mvdr_utterance<3>(masks, n_noise, Xc, T, n_bins, mask_floor, Yc, workspace, (cudaStream_t)stream);
        default: return launch_mvdr_utterance<4>(masks, n_noise, Xc, T, n_bins, mask_floor, Yc, workspace, (cudaStream_t)stream);
    }
}

extern "C" int nsf_mvdr(const float* masks, int n_spk, int n_noise, const float* X, int64_t T_long, int64_t T_valid,
                        int n_ch, int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor, float* Y,
                        void* stream) {
    NSF_REQUIRE(masks && X && Y, "nsf_mvdr: null pointer");
    if (n_spk < 2 || n_spk > 4 || n_ch != kMvdrC) {
        // the reference's own SCM code is written for 7 microphones (np.eye(7), mvdr_util.py:63)
        set_error("nsf_mvdr: built for 2..4 speaker masks and 7 microphones (got n_spk=%d, n_ch=%d)", n_spk, n_ch);
        return NSF_ERR_UNSUPPORTED;
    }
    NSF_REQUIRE(n_noise >= 1 && n_noise <= 4, "nsf_mvdr: n_noise=%d", n_noise);
    NSF_REQUIRE(T >= 1 && n_bins >= 1 && hop >= 1 && T_valid <= T_long, "nsf_mvdr: bad sizes");
    if (n_seg <= 0) return NSF_OK;
    // NSF_MVDR_IMPL=generic|stream|fano|entry96|entry64|entry32 and NSF_MVDR_RUN=<segments per warp run>: test / tuning switches, read per call
    const char* env = getenv("NSF_MVDR_IMPL");
    const bool can_stream = (T == 2 * hop) && hop <= kMvdrHopMax;
    int impl = can_stream ? MVDR_STREAM_ENTRY96 : MVDR_GENERIC;
    if (env && strcmp(env, "generic") == 0) impl = MVDR_GENERIC;
    else if (env && strcmp(env, "generic") != 0) {
        if (!can_stream) {
            set_error("nsf_mvdr: NSF_MVDR_IMPL=%s needs T == 2 hop and hop <= %d (got T=%d, hop=%d)", env, kMvdrHopMax, T, hop);
            return NSF_ERR_UNSUPPORTED;
        }
        impl = strcmp(env, "fano") == 0 ? MVDR_STREAM_FANO : strcmp(env, "entry64") == 0 ? MVDR_STREAM_ENTRY64
             : strcmp(env, "entry32") == 0 ? MVDR_STREAM_ENTRY32 : MVDR_STREAM_ENTRY96;      // "stream" / "entry96": the default
    }
    int run_len = 8;
    if (const char* r = getenv("NSF_MVDR_RUN")) run_len = atoi(r) > 0 ? atoi(r) : run_len;
    // algorithmic bytes: 7*8 mix + (S+Nn)*4 masks + S*8 out per (bin, frame)  (96 B for S = 3, Nn = 1)
    ProfScope prof(PROF_MVDR, (double)n_seg * n_bins * T * (kMvdrC * 8.0 + (n_spk + n_noise) * 4.0 + n_spk * 8.0), (cudaStream_t)stream);
    const float2* Xc = reinterpret_cast<const float2*>(X);
    float2* Yc = reinterpret_cast<float2*>(Y);
    switch (n_spk) {
        case 2: return launch_mvdr<2>(impl, masks, n_noise, Xc, T_long, T_valid, seg_first, n_seg, T, hop, n_bins, mask_floor, Yc, run_len, (cudaStream_t)stream);
        case 3: return launch_mvdr<3>(impl, masks, n_noise, Xc, T_long, T_valid, seg_first, n_seg, T, hop, n_bins, mask_floor, Yc, run_len, (cudaStream_t)stream);
        default: return launch_mvdr<4>(impl, masks, n_noise, Xc, T_long, T_valid, seg_first, n_seg, T, hop, n_bins, mask_floor, Yc, run_len, (cudaStream_t)stream);
    }
}
