// FP64 scores of the quartet-resolution hot path: singular values of the three 16 x 16 flattenings of a quartet's
// count tensor, numpy's rank rule, minrank, tail 2-norms, first argmin (reference tetrad/src/resolve_quartets.py:241-251).
//
// ONE implementation, used by every resolve kernel (the fused one-warp-per-quartet kernel, the tensor-core count
// kernel's score warps, the score pass behind the tip-pair GEMM): svd_scores4 takes up to four quartets, runs their 12
// flattenings through the same pivoted-QR + one-sided Jacobi in three passes of four matrices (8 lanes each), and every
// matrix stops rotating after its own last sweep -- so a quartet's scores are the same bits whatever kernel produced its
// counts and whatever quartets shared its pass.
//
// One-sided (Hestenes) Jacobi with the columns held in REGISTERS.  Lane (m, s) = (lane / 8, lane % 8) owns two columns
// P and Q of matrix m.  A sweep is a recursive-halving tournament of 8 + 4 + 2 + 1 = 15 steps that meets every pair of
// the 16 columns once: inside a group of g lanes the Q columns circulate by one lane per step (warp shuffles), and
// between phases the lower half of each group trades its Q for its partner's P through a small shared-memory buffer.
// The rotation tangent uses approximate reciprocal / rsqrt (a slightly-off angle only slows convergence); cos is refined
// to full FP64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tq {

static constexpr unsigned kFullMask = 0xFFFFFFFFu;
#define kFull kFullMask
static constexpr int kLdm = 18;               // padded column length of the smem matrices (doubles)
static constexpr double kEps = 2.220446049250313e-16;
static constexpr int kMaxSweeps = 30;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

#ifndef TQ_TAU2
#define TQ_TAU2 1e-18
#endif
static constexpr double kTau2 = TQ_TAU2;  // rotate when gamma^2 > tau^2 * |x|^2 |y|^2  (tau = 1e-9)
#ifndef TQ_STOP2
#define TQ_STOP2 1e-8
#endif
static constexpr double kStop2 = TQ_STOP2;    // a sweep whose largest relative gamma was < 1e-4 is the last

// Self-scaling ("fast") rotations: a column is stored as s * v with a scalar s, its reciprocal r = 1 / s travels along.
// Rotating x = sP P, y = sQ Q by (cos, sin) is exactly
//     x' = (cos sP) [P - tan (sQ / sP) Q],     y' = (cos sQ) [Q + tan (sP / sQ) P],
// so the 16 rows cost two FMAs each instead of two multiplies and two FMAs, and the factor cos goes into the scales
// (FP64 instructions are what the score warps wait for: profiles/r01_v7_resolve_ncu.md).  The scales are folded
// back into the columns once per sweep (fold_scales), so they stay within a few orders of magnitude of 1.
__device__ __forceinline__ void fold_scales(double (&P)[16], double (&Q)[16], double &sP, double &rP, double &sQ, double &rQ)
{
#pragma unroll
    for (int r = 0; r < 16; ++r) { P[r] *= sP; Q[r] *= sQ; }
    sP = 1.0; rP = 1.0; sQ = 1.0; rQ = 1.0;
}

__device__ __forceinline__ void jacobi_rotate(double (&P)[16], double (&Q)[16], double &sP, double &rP, double &sQ, double &rQ,
                                              double &nP, double &nQ, double tiny, bool &big, bool force = false,
                                              bool enable = true)
{
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
    for (int r = 0; r < 16; r += 4) {
        g0 = fma(P[r], Q[r], g0); g1 = fma(P[r + 1], Q[r + 1], g1);
        g2 = fma(P[r + 2], Q[r + 2], g2); g3 = fma(P[r + 3], Q[r + 3], g3);
    }
    const double ga = (sP * sQ) * ((g0 + g1) + (g2 + g3));
    const double ab = nP * nQ, gg = ga * ga;
    if (force || (enable && nP > tiny && nQ > tiny && gg > kTau2 * ab)) {
        big |= gg > kStop2 * ab;
        // Rotation that zeroes gamma:  with a = |Q|^2 - |P|^2, b = 2 gamma, r = sqrt(a^2 + b^2):
        //   cos^2 = (1 + |a| / r) / 2,   sin = sign(a) * b / (2 r cos),   tan = sin / cos.
        // Only two special-function ops (approximate FP64 rsqrt) sit on the dependency chain; the
        // pair (cos, sin) is then renormalised so that cos^2 + sin^2 = 1 to FP64 accuracy (the
        // angle itself may be ~1e-6 off, which only slows convergence).
        const double a = nQ - nP, b2 = ga + ga;
        const double q = fma(a, a, b2 * b2);
        double ir;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ir) : "d"(q));       // ~ 1 / r
        const double c2 = fma(0.5 * fabs(a), ir, 0.5);                   // cos^2 in [0.5, 1]
        double rc;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(c2));      // ~ 1 / cos
        double cs = c2 * rc;
        double sn = (a < 0.0 ? -0.5 : 0.5) * b2 * ir * rc;
        const double e = fma(cs, cs, fma(sn, sn, -1.0));                 // cos^2 + sin^2 - 1, ~1e-6
        const double f = fma(e, fma(0.375, e, -0.5), 1.0);               // (1 + e)^(-1/2) to O(e^3)
        cs *= f;
        sn *= f;
        double ic = rc;                                                  // 1 / cos: two Newton steps from the approximation
        ic = ic * fma(-cs, ic, 2.0);
        ic = ic * fma(-cs, ic, 2.0);
        const double t = sn * ic;                                        // tan
        const double al = t * (sQ * rP), be = t * (sP * rQ);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double p = P[r], qq = Q[r];
            P[r] = fma(-al, qq, p);
            Q[r] = fma(be, p, qq);
        }
        sP *= cs; sQ *= cs; rP *= ic; rQ *= ic;
        const double tg = t * ga;
        nP -= tg;
        nQ += tg;
    }
}

// ------------------------------------------------------------------------------------------
// Scores of up to FOUR quartets at a time (tensor-core kernel).  svd_scores above keeps lanes 24-31 idle: three
// matrices of 8 lanes each.  Here the 12 flattenings of four quartets go through the same QR + Jacobi code in three
// passes of four matrices, all 32 lanes busy: a quarter fewer passes per quartet.  Pass r, lane group m = lane / 8
// works on matrix 4r + m = flattening (4r + m) % 3 of quartet (4r + m) / 3; the sorted singular values of the 12
// matrices are parked in shared memory and lanes 0..3 finish one quartet each (rank, minrank, tail norms, argmin).
// ------------------------------------------------------------------------------------------
struct __align__(16) SvdScratch4 {
    double M[4][16][kLdm];                      // flattenings M[m][col][row]; later W (R factors) / exchange buffer
    double piv[4][16];                          // pivot-column buffer during the QR; unsorted singular values
    double sig[12][16];                         // sorted singular values of the 12 matrices
};

// Tbase: the CTA's slot array (256 counts per slot); slots: slot index of quartet j in byte j; actmask bit j: quartet j
// is present and has SNPs.  Lane j < 4 receives the result of quartet j.
// tptr(j): count tensor of quartet j (shared or global memory).
template <class TPtr>
__device__ __noinline__ void svd_scores4_impl(TPtr tptr, uint32_t actmask, SvdScratch4 &sv, int lane,
                                              double (&score_out)[3], uint32_t &best_out, uint32_t &minrank_out)
{
    const int m = lane >> 3, s = lane & 7;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int mi = 4 * pass + m, qi = mi / 3, f = mi - 3 * qi;
        const bool act = (actmask >> qi) & 1u;
        if (!__any_sync(kFull, act)) continue;      // none of this pass's four matrices exists (fewer than four quartets)
        const uint32_t *T = tptr(qi);
        __syncwarp();
        // this group's flattening (resolve_quartets.py:63,67-72), stored column-major; 32 of its 256 entries per lane
        double fro2 = 0.0;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int p = s + 8 * i;
            const int a = p >> 6, b = (p >> 4) & 3, c = (p >> 2) & 3, d = p & 3;
            const double v = act ? (double)T[p] : 0.0;
            const int col = (f == 0) ? 4 * c + d : (f == 1) ? 4 * b + d : 4 * b + c;
            const int row = (f == 0) ? 4 * a + b : (f == 1) ? 4 * a + c : 4 * a + d;
            sv.M[m][col][row] = v;
            fro2 = fma(v, v, fro2);
        }
        fro2 += __shfl_xor_sync(kFull, fro2, 4);
        fro2 += __shfl_xor_sync(kFull, fro2, 2);
        fro2 += __shfl_xor_sync(kFull, fro2, 1);
        const double tiny = (kEps * kEps) * fro2;       // columns this small are null directions
        __syncwarp();
        double P[16], Q[16];
        {
            const double2 *cp = reinterpret_cast<const double2 *>(&sv.M[m][s][0]);
            const double2 *cq = reinterpret_cast<const double2 *>(&sv.M[m][s + 8][0]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 x = cp[r], y = cq[r];
                P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
            }
        }
        __syncwarp();                                    // M is free from here on: reuse it below

        // QR preconditioner, as in svd_scores: pivoted modified Gram-Schmidt inside the 8-lane group, then the Jacobi
        // runs on the rows of R.
        {
            double *W = &sv.M[0][0][0];                  // [4][16 rows k][16 cols j], 8 KB
            bool doneP = !act, doneQ = !act;
            double mP = 0.0, mQ = 0.0;                   // squared norms of the remaining parts of P, Q
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {
                if ((k & 3) == 0) {                      // exact every 4th step, downdated in between
                    mP = 0.0; mQ = 0.0;
#pragma unroll
                    for (int r = 0; r < 16; ++r) { mP = fma(P[r], P[r], mP); mQ = fma(Q[r], Q[r], mQ); }
                }
                const double nP = doneP ? -1.0 : mP, nQ = doneQ ? -1.0 : mQ;
                double best = (nP >= nQ) ? nP : nQ;
                int idx = (nP >= nQ) ? s : s + 8;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    const double ob = __shfl_xor_sync(kFull, best, off);
                    const int oi = __shfl_xor_sync(kFull, idx, off);
                    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
                }
                const bool any = best >= 0.0;            // a column of this matrix is still unprocessed
                double2 *pv = reinterpret_cast<double2 *>(&sv.piv[m][0]);
                if (idx == s) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) pv[r] = make_double2(P[2 * r], P[2 * r + 1]);
                } else if (idx == s + 8) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) pv[r] = make_double2(Q[2 * r], Q[2 * r + 1]);
                }
                __syncwarp();
                double dP = 0.0, dQ = 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = pv[r];
                    dP = fma(x.x, P[2 * r], dP); dP = fma(x.y, P[2 * r + 1], dP);
                    dQ = fma(x.x, Q[2 * r], dQ); dQ = fma(x.y, Q[2 * r + 1], dQ);
                }
                const double pp = __shfl_sync(kFull, (idx == s) ? dP : dQ, (lane & ~7) | (idx & 7));
                const bool live = any && pp > tiny;      // else the pivot is a null direction: retire it with a zero row
                const double safe = live ? pp : 1.0;
                double rinv;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(safe));
                rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
                rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
                const double inv = rinv * rinv;
                const bool useP = live && !doneP, useQ = live && !doneQ;
                const double rP = useP ? dP * rinv : 0.0, rQ = useQ ? dQ * rinv : 0.0;
                W[(m * 16 + k) * 16 + s] = rP;
                W[(m * 16 + k) * 16 + s + 8] = rQ;
                mP = fmax(fma(-rP, rP, mP), 0.0);
                mQ = fmax(fma(-rQ, rQ, mQ), 0.0);
                const double cP = useP ? -dP * inv : 0.0, cQ = useQ ? -dQ * inv : 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = pv[r];
                    P[2 * r] = fma(cP, x.x, P[2 * r]); P[2 * r + 1] = fma(cP, x.y, P[2 * r + 1]);
                    Q[2 * r] = fma(cQ, x.x, Q[2 * r]); Q[2 * r + 1] = fma(cQ, x.y, Q[2 * r + 1]);
                }
                __syncwarp();
                if (any && idx == s) doneP = true;
                if (any && idx == s + 8) doneQ = true;
            }
            __syncwarp();
            {                                            // rows s and s+8 of R become this lane's columns
                const double2 *cp = reinterpret_cast<const double2 *>(W + (m * 16 + s) * 16);
                const double2 *cq = reinterpret_cast<const double2 *>(W + (m * 16 + s + 8) * 16);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = cp[r], y = cq[r];
                    P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
                }
            }
            __syncwarp();
        }
        double2 *xbuf = reinterpret_cast<double2 *>(&sv.M[0][0][0]);   // [32 lanes][11 double2] (176 B stride: conflict-free)

        // A matrix stops rotating after ITS last sweep, whatever the other three of the pass still need: the result of a
        // quartet must not depend on which quartets happen to share its batch (run-to-run reproducibility).
        bool gdone = !act;
        double sP = 1.0, rP = 1.0, sQ = 1.0, rQ = 1.0;   // column scales of the self-scaling rotations and their reciprocals
        for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
            fold_scales(P, Q, sP, rP, sQ, rQ);
            double nP = 0.0, nQ = 0.0;                   // refresh the cached squared norms
#pragma unroll
            for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
            bool big = false;
#pragma unroll 1
            for (int g = 8; g >= 1; g >>= 1) {
                for (int r = 0; r < g; ++r) {
                    jacobi_rotate(P, Q, sP, rP, sQ, rQ, nP, nQ, tiny, big, false, !gdone);
                    if (r + 1 < g) {                     // Q moves one lane down inside its group of g
                        const int src = (lane & ~(g - 1)) | ((lane + 1) & (g - 1));
#pragma unroll
                        for (int k = 0; k < 16; ++k) Q[k] = __shfl_sync(kFull, Q[k], src);
                        nQ = __shfl_sync(kFull, nQ, src); sQ = __shfl_sync(kFull, sQ, src); rQ = __shfl_sync(kFull, rQ, src);
                    }
                }
                if (g > 1) {
                    const int h = g >> 1;
                    const bool lower = (lane & (g - 1)) < h;
                    double2 *mine = xbuf + lane * 11;
                    if (lower) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) mine[k] = make_double2(Q[2 * k], Q[2 * k + 1]);
                        mine[8] = make_double2(nQ, sQ); mine[9] = make_double2(rQ, 0.0);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) mine[k] = make_double2(P[2 * k], P[2 * k + 1]);
                        mine[8] = make_double2(nP, sP); mine[9] = make_double2(rP, 0.0);
                    }
                    __syncwarp();
                    const double2 *theirs = xbuf + (lower ? lane + h : lane - h) * 11;
                    if (lower) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; Q[2 * k] = x.x; Q[2 * k + 1] = x.y; }
                        nQ = theirs[8].x; sQ = theirs[8].y; rQ = theirs[9].x;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; P[2 * k] = x.x; P[2 * k + 1] = x.y; }
                        nP = theirs[8].x; sP = theirs[8].y; rP = theirs[9].x;
                    }
                    __syncwarp();
                }
            }
            const uint32_t bigs = __ballot_sync(kFull, big);
            if (((bigs >> (lane & ~7)) & 0xFFu) == 0u) gdone = true;
            if (__all_sync(kFull, gdone)) break;
        }

        // singular values = column norms; sort descending per matrix by rank counting
        {
            fold_scales(P, Q, sP, rP, sQ, rQ);
            double nP = 0.0, nQ = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
            __syncwarp();
            sv.piv[m][s] = sqrt(nP); sv.piv[m][s + 8] = sqrt(nQ);
        }
        __syncwarp();
        for (int c = lane; c < 64; c += 32) {
            const int mm = c >> 4, col = c & 15;
            const double val = sv.piv[mm][col];
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double u = sv.piv[mm][j];
                pos += (u > val || (u == val && j < col)) ? 1 : 0;
            }
            sv.sig[4 * pass + mm][pos] = val;
        }
        __syncwarp();
    }
    __syncwarp();
    // lane j < 4 finishes quartet j: rank = #{sigma > sigma_max * 16 * eps} (numpy.linalg.matrix_rank, :243),
    // minrank (:246), tail norms (:248), first argmin (:251)
    if (lane < 4 && ((actmask >> lane) & 1u)) {
        int mr = 16;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double *sg = sv.sig[3 * lane + f];
            const double tol = sg[0] * (16.0 * kEps);
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) rank += (sg[j] > tol) ? 1 : 0;
            mr = min(mr, rank);
        }
        mr = min(mr, 10);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double *sg = sv.sig[3 * lane + f];
            double sc = 0.0;
            for (int j = mr; j < 16; ++j) sc = fma(sg[j], sg[j], sc);
            score_out[f] = sqrt(sc);
        }
        uint32_t best = 0;
        double lo = score_out[0];
        if (score_out[1] < lo) { best = 1; lo = score_out[1]; }
        if (score_out[2] < lo) { best = 2; }
        best_out = best;
        minrank_out = (uint32_t)mr;
    }
    __syncwarp();
}

// Tbase: the CTA's slot array (256 counts per slot); slots: slot index of quartet j in byte j.
__device__ __forceinline__ void svd_scores4(const uint32_t *Tbase, uint32_t slots, uint32_t actmask, SvdScratch4 &sv, int lane,
                                            double (&score_out)[3], uint32_t &best_out, uint32_t &minrank_out)
{
    svd_scores4_impl([=](int qi) { return Tbase + ((slots >> (8 * qi)) & 0xFFu) * 256u; }, actmask, sv, lane, score_out,
                     best_out, minrank_out);
}

// Zero-SNP quartets (reference :230-232): scores 0.001 and a split from a counter-based hash of (seed, tips) that the
// oracle shares, in place of the reference's unseeded np.random.randint(3).
__device__ __forceinline__ uint32_t tiebreak_split(uint64_t seed, uint4 t)
{
    uint64_t h = seed;
    h = splitmix64(h ^ t.x); h = splitmix64(h ^ t.y);
    h = splitmix64(h ^ t.z); h = splitmix64(h ^ t.w);
    return (uint32_t)(h % 3ull);
}

#undef kFull
}  // namespace tq
