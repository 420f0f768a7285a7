// Register-resident fused closed loop for small systems whose whole batch
// shares ONE equality-only controller (BASELINE configs 1-3: four-tank).
//
// One thread = one closed loop.  The plant state, the n-step measurement window
// and the planned inputs live in registers; the gain rows acting on the window
// (Kw) and the plant matrices arrive as a __grid_constant__ kernel parameter, so
// every DFMA takes its coefficient straight from the constant bank (no load
// instruction, warp-uniform).  The set-point part of the gain is folded into a
// per-loop constant once.  Measurement noise is drawn in-kernel (Philox4x32-10)
// or read from the caller's array (parity mode).  Trajectories are written in
// the reference layout (B, n_steps, m|p) as full 32-byte sectors: consecutive
// steps of one loop are paired (STG.256), and the thread->loop map puts loops of
// equal sector parity in the same warp so the stores stay warp-uniform.
//
// HBM traffic per loop-step: (m + p) * 8 B written, nothing read in Philox mode:
// this kernel is bound by the trajectory write (DESIGN.md "Roofline").
//
// Replaces the same reference code as k_closed_loop in solve.cu.
#include <type_traits>

#include "common.cuh"
#include "plan.cuh"

namespace ddmpc {

template <int N, int M, int P, int NX, int NMPC>
struct FastCoef {
    double Kt[N * (M + P)][NMPC * M];  // Kt[j][k] = Ku[k][j]: gain of window entry j on planned input k
    double A[NX][NX], B[NX][M], C[P][NX], D[P][M];
};

struct FastArgs {
    int B, n_steps;
    const double *Ksp;   // (NMPC*M, M+P) rows of Ku acting on [u_s; y_s]   (device)
    const double *x0, *u_past0, *y_past0, *u_s, *y_s, *w;
    unsigned long long seed, id0;
    double eps;
    double *u_sys, *y_sys, *x_final;
    int *status, *iters;
    uint32_t rk[20];     // Philox round keys (key + r * Weyl), filled on the host
    int zmask;           // always 0: defeats loop-invariant hoisting of coefficient loads
    int dbg_nostore;     // measurement aid (DDMPC_DEBUG_NOSTORE=1): k_closed_loop_ws skips the trajectory stores
    // CONVEX slack bound (device operators of controller 0, see plan.cuh)
    const double *Ks, *Phi, *Psi;   // (nb, nth), (nb, nb), (L*m, nb)
    double bound, tol;
    int nb, nth, max_iter;
};

// Warp-cooperative ADMM on the box rows of ONE loop, called only for the rare solves
// whose unconstrained slack violates the bound.  Lane i owns rows i and i + 32 (nb <= 64).
//   th(e)  : theta entry e of the loop (window then set-points)        up(k) : planned input k (corrected in place)
//   scr    : 2 x 64 doubles of shared scratch (d, t)
// Returns the number of iterations; *inaccurate is set when max_iter was reached.
template <typename ThetaF, typename UpF>
__device__ __noinline__ int admm_box_warp(const FastArgs &a, int R, ThetaF th, UpF up, double *scr, bool *inaccurate) {
    const int lane = threadIdx.x & 31, nb = a.nb, nth = a.nth;
    double *dsh = scr, *tsh = scr + 64;
    double sun[2], z[2], w[2];
    double smax = 0.0;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = lane + 32 * rr;
        double acc = 0.0;
        if (i < nb)
            for (int e = 0; e < nth; ++e) acc = fma(__ldg(a.Ks + (size_t)i * nth + e), th(e), acc);
        sun[rr] = acc;
        z[rr] = fmin(fmax(acc, -a.bound), a.bound);
        w[rr] = 0.0;
        smax = fmax(smax, fabs(acc));
    }
    for (int o = 16; o > 0; o >>= 1) smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    const double thr = a.tol * fmax(a.bound, smax);
    int it = 0;
    bool conv = false;
    while (it < a.max_iter && !conv) {
        ++it;
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = lane + 32 * rr;
            if (i < nb) dsh[i] = sun[rr] - z[rr] + w[rr];
        }
        __syncwarp();
        double res = 0.0;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = lane + 32 * rr;
            if (i < nb) {
                // Phi is symmetric: read column i (= row i) with the lanes along the unit stride
                double acc0 = 0.0, acc1 = 0.0;
                const double *col = a.Phi + i;
                int j = 0;
                for (; j + 7 < nb; j += 8) {       // 8 independent loads in flight (the loop is latency-bound)
                    double pv[8];
#pragma unroll
                    for (int u8 = 0; u8 < 8; ++u8) pv[u8] = __ldg(col + (size_t)(j + u8) * nb);
#pragma unroll
                    for (int u8 = 0; u8 < 8; u8 += 2) {
                        acc0 = fma(pv[u8], dsh[j + u8], acc0);
                        acc1 = fma(pv[u8 + 1], dsh[j + u8 + 1], acc1);
                    }
                }
                for (; j < nb; ++j) acc0 = fma(__ldg(col + (size_t)j * nb), dsh[j], acc0);
                const double si = (z[rr] - w[rr]) + (acc0 + acc1);
                const double sr = DDMPC_ADMM_RELAX * si + (1.0 - DDMPC_ADMM_RELAX) * z[rr];   // over-relaxation (solve.cu)
                const double zn = fmin(fmax(sr + w[rr], -a.bound), a.bound);
                res = fmax(res, fmax(fabs(si - zn), fabs(zn - z[rr])));
                w[rr] = w[rr] + sr - zn;
                z[rr] = zn;
            }
        }
        for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
        conv = res <= thr;
    }
    *inaccurate = !conv;
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = lane + 32 * rr;
        if (i < nb) dsh[i] = sun[rr] - z[rr] + w[rr];
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = lane + 32 * rr;
        if (i < nb) {
            double acc = 0.0;
            for (int j = 0; j < nb; ++j) acc = fma(__ldg(a.Phi + (size_t)j * nb + i), dsh[j], acc);
            tsh[i] = acc;                      // t = Phi d
        }
    }
    __syncwarp();
    if (lane < R) {                            // u = u0 - Psi t on the applied rows
        double acc = 0.0;
        const double *row = a.Psi + (size_t)lane * nb;
        for (int j = 0; j < nb; ++j) acc = fma(__ldg(row + j), tsh[j], acc);
        up(lane, acc);
    }
    __syncwarp();
    return it;
}

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                             uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}

__device__ __forceinline__ double unit32_fast(uint32_t x) {
    return __hiloint2double((int)(0x3FF00000u | (x >> 12)), (int)(x << 20));
}

// One trajectory element (EL doubles) per step.  With EL == 2 an element is 16 B and the
// elements f-1, f (f odd) fill one 32 B sector, so they leave as a single STG.256; the even
// element is not stored on its own - at the next step it is still the newest entry of the
// measurement window (`prev`).
template <int EL, bool PAIR>
__device__ __forceinline__ void emit(double *__restrict__ base, size_t f, bool odd, bool first,
                                     const double (&prev)[EL], const double (&cur)[EL]) {
    if constexpr (EL == 2 && PAIR) {
        if (odd) {
            if (first) {
                *reinterpret_cast<double2 *>(base + f * 2) = make_double2(cur[0], cur[1]);
            } else {
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(base + (f - 1) * 2), "d"(prev[0]),
                             "d"(prev[1]), "d"(cur[0]), "d"(cur[1])
                             : "memory");
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < EL; ++i) base[f * EL + i] = cur[i];
    }
}

// LPT = closed loops per thread.  Every coefficient fetched from the constant bank feeds LPT
// DFMAs (one per loop), so LPT = 2 halves the pressure on the indexed-constant (ADU/IDC) path and
// doubles the independent work between dependent instructions, at the price of half the warps.
template <int N, int M, int P, int NX, int NMPC, int LPT, bool PHILOX, bool PAIR, bool CVX = false>
__global__ void __launch_bounds__(LPT == 1 ? 64 : 32, 8)
k_closed_loop_fast(const __grid_constant__ FastCoef<N, M, P, NX, NMPC> cfp, const FastArgs a) {
    // Coefficients stay in the kernel-parameter constant bank and are fetched with indexed LDC
    // (address = parameter base + a run-time zero).  Two alternatives were measured and lost:
    //  * compile-time constant-bank operands: ptxas stages FP64 constant operands through the 63
    //    uniform registers, treats all 164 loads as loop invariant and spills them (R2UR/UMOV storm);
    //  * a shared-memory copy read with LDS.128: correct, but every DFMA then waits on the MIO pipe.
    using Coef = FastCoef<N, M, P, NX, NMPC>;
    constexpr int R = NMPC * M;                  // planned-input rows per solve
    constexpr bool ALIGNED = (NMPC % N) == 0;    // a block starts with the ring at slot 0
    constexpr int FAST_TPB = LPT == 1 ? 64 : 32; // one warp per block when a thread carries two loops
    // Tensor-core solve: with 8 planned-input rows the gain application U(8 x loops) = Ku(8 x 16) W(16 x loops)
    // is exactly an m8n8k4 FP64 MMA shape: the warp's 64 loops are 8 n-tiles, the window is read from
    // shared memory in B-fragment order and Ku lives in registers as A fragments.  This moves 128 of the
    // 272 FMAs per loop and n-step block from the FP64 pipe to the (otherwise idle) tensor pipe.
    constexpr bool USE_MMA = (R == 8) && (LPT == 2) && ((N * M) % 4 == 0) && ((N * P) % 4 == 0) && ALIGNED;
    constexpr int TP = FAST_TPB + (USE_MMA ? 2 : 0);  // +2 doubles: row stride 68 = 4 (mod 16) -> B-fragment
                                                      // loads of 4 rows x 8 loops spread over all banks
    __shared__ __align__(16) double csp_s[R][LPT][TP];     // per-loop set-point term of the planned inputs
    __shared__ __align__(16) double up_s[R][LPT][TP];      // planned inputs of the current n-step block
    __shared__ __align__(16) double wu_s[N * M][LPT][TP];  // measurement window, ring over N time slots
    __shared__ __align__(16) double wy_s[N * P][LPT][TP];
    static_assert(!CVX || USE_MMA, "the fused CONVEX path needs the tensor-core solve (LPT = 2, 8 planned-input rows)");
    __shared__ double sp_s[CVX ? M + P : 1][LPT][TP];      // set-points (theta tail) for the slack rows
    __shared__ int extra_s[CVX ? LPT : 1][TP];             // extra ADMM iterations per loop
    __shared__ int stat_s[CVX ? LPT : 1][TP];              // worst ADMM status per loop
    __shared__ double adm_s[CVX ? 128 : 1];                // scratch of the warp-cooperative ADMM
    // thread -> loop map (64 loops per block): LPT = 1: the first warp takes the even loops and the
    // second warp the odd ones; LPT = 2: thread t carries loops 2t (l = 0) and 2t + 1 (l = 1).
    // Either way the sector parity of a step is uniform across a warp for a given l.
    const int tl = threadIdx.x;
    int b[LPT];
    bool live[LPT];
    size_t f0[LPT];
    uint32_t sid_lo[LPT], sid_hi[LPT];
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        b[l] = LPT == 1 ? blockIdx.x * 64 + 2 * (tl % 32) + (tl / 32) : blockIdx.x * 64 + 2 * tl + l;
        live[l] = b[l] < a.B;
        if (!live[l]) b[l] = 0;                  // dead slots replay loop 0 and never store
        f0[l] = (size_t)b[l] * a.n_steps;
        const unsigned long long sid = a.id0 + (unsigned long long)b[l];
        sid_lo[l] = (uint32_t)sid;
        sid_hi[l] = (uint32_t)(sid >> 32);
    }
    if (!USE_MMA && !live[0]) return;             // (mma.sync needs the whole warp)
    double x[LPT][NX];
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[l][i] = a.x0[(size_t)b[l] * NX + i];
#pragma unroll
        for (int i = 0; i < N * M; ++i) wu_s[i][l][tl] = a.u_past0[(size_t)b[l] * N * M + i];
#pragma unroll
        for (int i = 0; i < N * P; ++i) wy_s[i][l][tl] = a.y_past0[(size_t)b[l] * N * P + i];
        double sp[M + P];
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = a.u_s[(size_t)b[l] * M + i];
#pragma unroll
        for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[(size_t)b[l] * P + i];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < M + P; ++j) acc = fma(__ldg(a.Ksp + k * (M + P) + j), sp[j], acc);
            csp_s[k][l][tl] = acc;
        }
        if constexpr (CVX) {
#pragma unroll
            for (int j = 0; j < M + P; ++j) sp_s[j][l][tl] = sp[j];
            extra_s[l][tl] = 0;
            stat_s[l][tl] = DDMPC_SOLVE_OPTIMAL;
        }
    }
    uint32_t nw[LPT][4];                         // words of the current Philox call
#pragma unroll
    for (int l = 0; l < LPT; ++l)
#pragma unroll
        for (int i = 0; i < 4; ++i) nw[l][i] = 0u;

    // ---- QP solve (equality-only => affine in the window): planned inputs = csp + Kw * window.
    // Window entry jj (0 = oldest) sits in ring slot (t0 + jj) % N.
    // `cfz` is &cfp plus a run-time zero that changes (formally) every block: without it ptxas
    // treats the coefficient loads as loop invariant, hoists all of them and spills.
    const Coef *cfz = &cfp;
    double afrag[USE_MMA ? (N * (M + P)) / 4 : 1];   // A fragments: Ku[lane/4][4*ks + lane%4]
    if constexpr (USE_MMA) {
#pragma unroll
        for (int ks = 0; ks < (N * (M + P)) / 4; ++ks) afrag[ks] = cfp.Kt[4 * ks + (tl & 3)][tl >> 2];
        __syncwarp();                            // window / csp written above by their owner lanes
    }
    auto solve = [&](const int t0) {
        const int zuni = t0 & a.zmask;           // a.zmask is 0 at run time
        cfz = &cfp + zuni;
        if constexpr (USE_MMA) {
            __syncwarp();                        // window entries of the previous block are visible
            const int g = tl >> 2, q = tl & 3;   // fragment coordinates of this lane
            constexpr int NT = FAST_TPB / 8;     // n-tiles (8 loops each) per l
            // C fragments: rows g, loops 8*t8 + 2q, +1  <- set-point term.  k-steps outermost so that
            // consecutive MMAs belong to different n-tiles (independent accumulators).
            double2 c[LPT][NT];
#pragma unroll
            for (int l = 0; l < LPT; ++l)
#pragma unroll
                for (int t8 = 0; t8 < NT; ++t8) c[l][t8] = *reinterpret_cast<const double2 *>(&csp_s[g][l][8 * t8 + 2 * q]);
#pragma unroll
            for (int ks = 0; ks < (N * (M + P)) / 4; ++ks) {
                // B fragment: window entry 4*ks + q of loop 8*t8 + g (entries 0..N*M-1 are u, then y)
                const int e = 4 * ks + q;
                double bv[LPT][NT];
#pragma unroll
                for (int l = 0; l < LPT; ++l)
#pragma unroll
                    for (int t8 = 0; t8 < NT; ++t8)
                        bv[l][t8] = (4 * ks < N * M) ? wu_s[e < N * M ? e : 0][l][8 * t8 + g]
                                                     : wy_s[e >= N * M ? e - N * M : 0][l][8 * t8 + g];
#pragma unroll
                for (int l = 0; l < LPT; ++l)
#pragma unroll
                    for (int t8 = 0; t8 < NT; ++t8)
                        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                            : "+d"(c[l][t8].x), "+d"(c[l][t8].y)
                            : "d"(afrag[ks]), "d"(bv[l][t8]));
            }
#pragma unroll
            for (int l = 0; l < LPT; ++l)
#pragma unroll
                for (int t8 = 0; t8 < NT; ++t8) *reinterpret_cast<double2 *>(&up_s[g][l][8 * t8 + 2 * q]) = c[l][t8];
            __syncwarp();                        // planned inputs visible to their owner lanes
            if constexpr (CVX) {
                // ---- CONVEX slack bound: s_unc (nb x loops) = Ks (nb x n_theta) [window; set-points], again as
                // m8n8k4 MMAs, two row tiles at a time; only max |s_unc| per loop is kept.
                constexpr int NKS = (N * (M + P) + M + P) / 4;
                static_assert((M + P) % 4 == 0, "set-point block must fill whole k-steps");
                const int nrt = (a.nb + 7) / 8;
                double2 smax[LPT][NT];
#pragma unroll
                for (int l = 0; l < LPT; ++l)
#pragma unroll
                    for (int t8 = 0; t8 < NT; ++t8) smax[l][t8] = make_double2(0.0, 0.0);
                for (int rt = 0; rt < nrt; rt += 2) {
                    double2 cc[2][LPT][NT];
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                        for (int l = 0; l < LPT; ++l)
#pragma unroll
                            for (int t8 = 0; t8 < NT; ++t8) cc[r2][l][t8] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int ks = 0; ks < NKS; ++ks) {
                        const int e = 4 * ks + q;
                        double bv[LPT][NT];
#pragma unroll
                        for (int l = 0; l < LPT; ++l)
#pragma unroll
                            for (int t8 = 0; t8 < NT; ++t8)
                                bv[l][t8] = (4 * ks < N * M)        ? wu_s[e < N * M ? e : 0][l][8 * t8 + g]
                                            : (4 * ks < N * (M + P)) ? wy_s[(e >= N * M && e < N * (M + P)) ? e - N * M : 0][l][8 * t8 + g]
                                                                     : sp_s[e >= N * (M + P) ? e - N * (M + P) : 0][l][8 * t8 + g];
#pragma unroll
                        for (int r2 = 0; r2 < 2; ++r2) {
                            const int row = 8 * (rt + r2) + g;
                            const double av = row < a.nb ? __ldg(a.Ks + (size_t)row * a.nth + e) : 0.0;
#pragma unroll
                            for (int l = 0; l < LPT; ++l)
#pragma unroll
                                for (int t8 = 0; t8 < NT; ++t8)
                                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                        : "+d"(cc[r2][l][t8].x), "+d"(cc[r2][l][t8].y)
                                        : "d"(av), "d"(bv[l][t8]));
                        }
                    }
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                        for (int l = 0; l < LPT; ++l)
#pragma unroll
                            for (int t8 = 0; t8 < NT; ++t8) {
                                smax[l][t8].x = fmax(smax[l][t8].x, fabs(cc[r2][l][t8].x));
                                smax[l][t8].y = fmax(smax[l][t8].y, fabs(cc[r2][l][t8].y));
                            }
                }
                // max over the 8 rows of a tile (lanes with equal q); bit (l, t8, hh) of vmask = "loop
                // slot 8*t8 + 2q + hh of group l violates the bound" (static indices only: smax stays in registers)
                unsigned vmask = 0u;
#pragma unroll
                for (int l = 0; l < LPT; ++l)
#pragma unroll
                    for (int t8 = 0; t8 < NT; ++t8) {
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            smax[l][t8].x = fmax(smax[l][t8].x, __shfl_xor_sync(0xffffffffu, smax[l][t8].x, o));
                            smax[l][t8].y = fmax(smax[l][t8].y, __shfl_xor_sync(0xffffffffu, smax[l][t8].y, o));
                        }
                        if (smax[l][t8].x > a.bound) vmask |= 1u << (2 * (l * NT + t8));
                        if (smax[l][t8].y > a.bound) vmask |= 1u << (2 * (l * NT + t8) + 1);
                    }
                if (g != 0) vmask = 0u;          // lanes 0..3 (g = 0, q = lane) speak for their loops
                if (__any_sync(0xffffffffu, vmask != 0u)) {
                    // rare: run the ADMM for each violating loop, the whole warp on one loop at a time
#pragma unroll 1
                    for (int src = 0; src < 4; ++src) {
                        unsigned mbits = __shfl_sync(0xffffffffu, vmask, src);
                        while (mbits) {
                            const int bit = __ffs(mbits) - 1;
                            mbits &= mbits - 1;
                            const int l = bit / (2 * NT), t8 = (bit >> 1) % NT, hh = bit & 1;
                            const int slot = 8 * t8 + 2 * src + hh;    // loop slot (l, slot)
                            auto th = [&](int e) -> double {
                                return e < N * M ? wu_s[e][l][slot]
                                                 : (e < N * (M + P) ? wy_s[e - N * M][l][slot] : sp_s[e - N * (M + P)][l][slot]);
                            };
                            auto upf = [&](int k, double corr) { up_s[k][l][slot] -= corr; };
                            bool inacc = false;
                            const int it = admm_box_warp(a, R, th, upf, adm_s, &inacc);
                            if (tl == 0) {
                                extra_s[l][slot] += it - 1;
                                if (inacc) stat_s[l][slot] = DDMPC_SOLVE_OPTIMAL_INACCURATE;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            return;
        }
        constexpr int SPLIT = (R * LPT >= 8) ? 1 : (R * LPT >= 4 ? 2 : 4);   // partial sums when rows are few
        double acc[SPLIT][LPT][R];
#pragma unroll
        for (int l = 0; l < LPT; ++l)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                acc[0][l][k] = csp_s[k + zuni][l][tl];
#pragma unroll
                for (int q = 1; q < SPLIT; ++q) acc[q][l][k] = 0.0;
            }
        const int base = ALIGNED ? 0 : (t0 % N);
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
            const int slot = ALIGNED ? jj : ((base + jj) % N);
#pragma unroll
            for (int i = 0; i < M + P; ++i) {
                double wj[LPT];
#pragma unroll
                for (int l = 0; l < LPT; ++l)
                    wj[l] = (i < M) ? wu_s[slot * M + (i < M ? i : 0)][l][tl] : wy_s[slot * P + (i >= M ? i - M : 0)][l][tl];
                const int col = (i < M) ? (jj * M + i) : (N * M + jj * P + (i - M));
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const double c = cfz->Kt[col][k];
#pragma unroll
                    for (int l = 0; l < LPT; ++l) acc[jj % SPLIT][l][k] = fma(c, wj[l], acc[jj % SPLIT][l][k]);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < LPT; ++l)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double v = acc[0][l][k];
#pragma unroll
                for (int q = 1; q < SPLIT; ++q) v += acc[q][l][k];
                up_s[k][l][tl] = v;
            }
    };
    // ---- one plant step with the s-th planned input: noise, y, x, record, window update
    auto step = [&](const int s, const int k) {
        double u[LPT][M], y[LPT][P];
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
#pragma unroll
            for (int i = 0; i < M; ++i) u[l][i] = up_s[s * M + i][l][tl];
            if constexpr (!PHILOX) {
#pragma unroll
                for (int i = 0; i < P; ++i) y[l][i] = __ldg(a.w + (f0[l] + k) * P + i);
            } else if constexpr ((NMPC * P) % 4 == 0) {
                // noise word q = k*P + i is word (q & 3) of Philox call (q >> 2); a block starts on a
                // call boundary, so call index and word are compile-time offsets from the block base
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int qs = s * P + i;
                    if ((qs & 3) == 0) {
                        uint32_t c0 = (uint32_t)(((unsigned)(k - s) * (unsigned)P) >> 2) + (uint32_t)(qs >> 2),
                                 c1 = 0u, c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                        for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                        nw[l][0] = c0; nw[l][1] = c1; nw[l][2] = c2; nw[l][3] = c3;
                    }
                    y[l][i] = a.eps * (2.0 * unit32_fast(nw[l][qs & 3]) - 3.0);
                }
            } else {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const unsigned q = (unsigned)k * (unsigned)P + (unsigned)i;
                    if (i == 0 || (q & 3u) == 0u) {
                        uint32_t c0 = q >> 2, c1 = 0u, c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                        for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                        nw[l][0] = c0; nw[l][1] = c1; nw[l][2] = c2; nw[l][3] = c3;
                    }
                    const unsigned w4 = q & 3u;
                    const uint32_t word = w4 == 0 ? nw[l][0] : (w4 == 1 ? nw[l][1] : (w4 == 2 ? nw[l][2] : nw[l][3]));
                    y[l][i] = a.eps * (2.0 * unit32_fast(word) - 3.0);
                }
            }
        }
        // y = C x + D u + w   (pre-update state; model_simulation.py:94)
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double acc[LPT], acd[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) acc[l] = acd[l] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const double c = cfz->C[i][j];
#pragma unroll
                for (int l = 0; l < LPT; ++l) acc[l] = fma(c, x[l][j], acc[l]);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const double c = cfz->D[i][j];
#pragma unroll
                for (int l = 0; l < LPT; ++l) acd[l] = fma(c, u[l][j], acd[l]);
            }
#pragma unroll
            for (int l = 0; l < LPT; ++l) y[l][i] = (acc[l] + acd[l]) + y[l][i];
        }
        // x <- A x + B u      (model_simulation.py:96)
        double xn[LPT][NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc[LPT], acb[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) acc[l] = acb[l] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const double c = cfz->A[i][j];
#pragma unroll
                for (int l = 0; l < LPT; ++l) acc[l] = fma(c, x[l][j], acc[l]);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const double c = cfz->B[i][j];
#pragma unroll
                for (int l = 0; l < LPT; ++l) acb[l] = fma(c, u[l][j], acb[l]);
            }
#pragma unroll
            for (int l = 0; l < LPT; ++l) xn[l][i] = acc[l] + acb[l];
        }
        const int slot = ALIGNED ? (s % N) : (k % N);            // oldest slot: overwritten below
        const int pslot = ALIGNED ? ((s + N - 1) % N) : ((k + N - 1) % N);
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
#pragma unroll
            for (int i = 0; i < NX; ++i) x[l][i] = xn[l][i];
            // record (full-sector stores); the previous element is the newest window entry
            const size_t f = f0[l] + k;
            const bool odd = (f & 1) != 0;
            if (live[l]) {
                if constexpr (PAIR) {
                    if (odd) {   // warp-uniform
                        double pvu[M], pvy[P];
#pragma unroll
                        for (int i = 0; i < M; ++i) pvu[i] = wu_s[pslot * M + i][l][tl];
#pragma unroll
                        for (int i = 0; i < P; ++i) pvy[i] = wy_s[pslot * P + i][l][tl];
                        emit<M, true>(a.u_sys, f, true, k == 0, pvu, u[l]);
                        emit<P, true>(a.y_sys, f, true, k == 0, pvy, y[l]);
                    }
                } else {
                    emit<M, false>(a.u_sys, f, odd, k == 0, u[l], u[l]);
                    emit<P, false>(a.y_sys, f, odd, k == 0, y[l], y[l]);
                }
            }
            // window update (controller.py:893-895): the oldest slot receives (u, y)
#pragma unroll
            for (int i = 0; i < M; ++i) wu_s[slot * M + i][l][tl] = u[l][i];
#pragma unroll
            for (int i = 0; i < P; ++i) wy_s[slot * P + i][l][tl] = y[l][i];
        }
    };

    int t0 = 0;
    for (; t0 + NMPC <= a.n_steps; t0 += NMPC) {   // full n-step blocks: no guards
        solve(t0);
#pragma unroll
        for (int s = 0; s < NMPC; ++s) step(s, t0 + s);
    }
    if (t0 < a.n_steps) {                          // last, partial block (controller_operation.py:278)
        solve(t0);
#pragma unroll
        for (int s = 0; s < NMPC; ++s)
            if (t0 + s < a.n_steps) step(s, t0 + s);
    }
    const int lslot = (a.n_steps + N - 1) % N;     // newest window entry
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        if (!live[l]) continue;
        if constexpr (PAIR) {   // an unpaired final element is still the newest window entry
            const size_t fl = f0[l] + a.n_steps - 1;
            if ((fl & 1) == 0) {
                if constexpr (M == 2)
                    *reinterpret_cast<double2 *>(a.u_sys + fl * 2) =
                        make_double2(wu_s[lslot * M][l][tl], wu_s[lslot * M + 1][l][tl]);
                if constexpr (P == 2)
                    *reinterpret_cast<double2 *>(a.y_sys + fl * 2) =
                        make_double2(wy_s[lslot * P][l][tl], wy_s[lslot * P + 1][l][tl]);
            }
        }
        bool finite = true;
#pragma unroll
        for (int i = 0; i < NX; ++i) finite = finite && isfinite(x[l][i]);
#pragma unroll
        for (int i = 0; i < N * P; ++i) finite = finite && isfinite(wy_s[i][l][tl]);
        int st_out = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE, it_out = (a.n_steps + NMPC - 1) / NMPC;
        if constexpr (CVX) {
            st_out = max(st_out, stat_s[l][tl]);
            it_out += extra_s[l][tl];
        }
        if (a.status) a.status[b[l]] = st_out;
        if (a.iters) a.iters[b[l]] = it_out;
        if (a.x_final) {
#pragma unroll
            for (int i = 0; i < NX; ++i) a.x_final[(size_t)b[l] * NX + i] = x[l][i];
        }
    }
}

// (A two-lanes-per-loop variant of the kernel above - each lane computing half of the rows of every product, halves
// exchanged through shuffles - was measured and removed: twice the warps but also twice the shared-memory
// instructions per loop, MIO-throttled at 0.49 ms against 0.34 ms on config 3.  Its ncu summary is kept in
// profiles/r1_k_closed_loop_pair_experiment_ncu_full_summary.txt.)

// ===========================================================================
// All-tensor-core variant (four-tank n-step shape: 8 planned-input rows, NMPC a multiple of N, M = P = 2).
// Per n-step block and warp (64 loops = 8 n-tiles of the m8n8k4 FP64 MMA):
//   solve :  U (8 x loops)         = Ku (8 x 16) [window_u; window_y] + csp          4 k-steps, 32 DMMA
//   plant :  [Y (8); x+ (4)] x loops = Mblk (12 x 12) [x (4); U (8)]                 3 k-steps x 2 row tiles, 48 DMMA
// where Mblk is the NMPC-step block map of the LTI plant (rows y_0..y_{NMPC-1}, x_NMPC; columns x_0, u_0..):
// model_simulation.py:93-98 unrolled NMPC times (measurement noise only enters y and is added afterwards).
// Everything lives in shared memory [value][loop]; the planned inputs of a block ARE the input half of the next
// measurement window (NMPC = N), so `up_s` doubles as the window.  What is left for the owner thread of a loop is
// the noise draw, y = Y + w, the sector-paired trajectory stores and the output half of the window.
// ===========================================================================
template <int N, int M, int P, int NX, int NMPC>
struct MmaCoef {
    double Ku[NMPC * M][N * (M + P)];               // gain rows on [u_past; y_past]
    double Mb[NMPC * P + NX][NX + NMPC * M];        // block map, full block
    double Mt[NMPC * P + NX][NX + NMPC * M];        // block map of the last, partial block (n_tail steps), zero padded
};

template <int N, int M, int P, int NX, int NMPC, bool PHILOX>
__global__ void __launch_bounds__(32, 8)
k_closed_loop_mma(const __grid_constant__ MmaCoef<N, M, P, NX, NMPC> cfp, const FastArgs a, const int n_tail) {
    constexpr int R = NMPC * M, NW = N * (M + P), LPT = 2, TPB = 32, TP = TPB + 2, NT = TPB / 8;
    constexpr int KB = NX + R, RB = NMPC * P + NX;
    static_assert(M == 2 && P == 2 && R == 8 && (NMPC % N) == 0 && NMPC == N, "shape not supported by the MMA kernel");
    static_assert(NW % 4 == 0 && (N * M) % 4 == 0 && KB % 4 == 0 && RB <= 16 && NMPC * P == 8, "fragment tiling");
    __shared__ __align__(16) double csp_s[R][LPT][TP];
    __shared__ __align__(16) double up_s[R][LPT][TP];          // planned inputs = input half of the window
    __shared__ __align__(16) double wy_s[N * P][LPT][TP];      // output half of the window
    __shared__ __align__(16) double x_s[NX][LPT][TP];          // plant state
    __shared__ __align__(16) double Y_s[NMPC * P][LPT][TP];    // noise-free outputs of the block
    const int tl = threadIdx.x, g = tl >> 2, q = tl & 3;
    int b[LPT];
    bool live[LPT];
    size_t f0[LPT];
    uint32_t sid_lo[LPT], sid_hi[LPT];
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        b[l] = blockIdx.x * 64 + 2 * tl + l;
        live[l] = b[l] < a.B;
        if (!live[l]) b[l] = 0;
        f0[l] = (size_t)b[l] * a.n_steps;
        const unsigned long long sid = a.id0 + (unsigned long long)b[l];
        sid_lo[l] = (uint32_t)sid;
        sid_hi[l] = (uint32_t)(sid >> 32);
#pragma unroll
        for (int i = 0; i < NX; ++i) x_s[i][l][tl] = a.x0[(size_t)b[l] * NX + i];
#pragma unroll
        for (int i = 0; i < N * M; ++i) up_s[i][l][tl] = a.u_past0[(size_t)b[l] * N * M + i];
#pragma unroll
        for (int i = 0; i < N * P; ++i) wy_s[i][l][tl] = a.y_past0[(size_t)b[l] * N * P + i];
        double sp[M + P];
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = a.u_s[(size_t)b[l] * M + i];
#pragma unroll
        for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[(size_t)b[l] * P + i];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < M + P; ++j) acc = fma(__ldg(a.Ksp + k * (M + P) + j), sp[j], acc);
            csp_s[k][l][tl] = acc;
        }
    }
    // A fragments (row g, column 4*ks + q of each k-step) stay in registers for the whole run
    double aK[NW / 4], aP[2][KB / 4];
#pragma unroll
    for (int ks = 0; ks < NW / 4; ++ks) aK[ks] = cfp.Ku[g][4 * ks + q];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ks = 0; ks < KB / 4; ++ks) aP[rt][ks] = (8 * rt + g < RB) ? cfp.Mb[(8 * rt + g) % RB][4 * ks + q] : 0.0;
    double pu[LPT][M], py[LPT][P];               // previous trajectory element (sector pairing)
    uint32_t nw[LPT][4];
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        pu[l][0] = pu[l][1] = py[l][0] = py[l][1] = 0.0;
        nw[l][0] = nw[l][1] = nw[l][2] = nw[l][3] = 0u;
    }
    __syncwarp();

    auto mma = [](double2 &c, double av, double bv) {
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c.x), "+d"(c.y)
            : "d"(av), "d"(bv));
    };
    auto block = [&](const int t0, const int steps) {
        // ---- solve: planned inputs (C fragments start from the set-point term)
        double2 c[LPT][NT];
#pragma unroll
        for (int l = 0; l < LPT; ++l)
#pragma unroll
            for (int t8 = 0; t8 < NT; ++t8) c[l][t8] = *reinterpret_cast<const double2 *>(&csp_s[g][l][8 * t8 + 2 * q]);
#pragma unroll
        for (int ks = 0; ks < NW / 4; ++ks) {
            const int e = 4 * ks + q;
#pragma unroll
            for (int l = 0; l < LPT; ++l)
#pragma unroll
                for (int t8 = 0; t8 < NT; ++t8) {
                    const double bv = (4 * ks < N * M) ? up_s[e < N * M ? e : 0][l][8 * t8 + g]
                                                       : wy_s[e >= N * M ? e - N * M : 0][l][8 * t8 + g];
                    mma(c[l][t8], aK[ks], bv);
                }
        }
        __syncwarp();                            // every lane has read the old window
#pragma unroll
        for (int l = 0; l < LPT; ++l)
#pragma unroll
            for (int t8 = 0; t8 < NT; ++t8) *reinterpret_cast<double2 *>(&up_s[g][l][8 * t8 + 2 * q]) = c[l][t8];
        __syncwarp();
        // ---- plant: NMPC steps at once through the block map
        double2 d[2][LPT][NT];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int l = 0; l < LPT; ++l)
#pragma unroll
                for (int t8 = 0; t8 < NT; ++t8) d[rt][l][t8] = make_double2(0.0, 0.0);
#pragma unroll
        for (int ks = 0; ks < KB / 4; ++ks) {
            const int e = 4 * ks + q;
#pragma unroll
            for (int l = 0; l < LPT; ++l)
#pragma unroll
                for (int t8 = 0; t8 < NT; ++t8) {
                    const double bv = (4 * ks < NX) ? x_s[e < NX ? e : 0][l][8 * t8 + g]
                                                    : up_s[e >= NX ? e - NX : 0][l][8 * t8 + g];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) mma(d[rt][l][t8], aP[rt][ks], bv);
                }
        }
        __syncwarp();                            // every lane has read the old state
#pragma unroll
        for (int l = 0; l < LPT; ++l)
#pragma unroll
            for (int t8 = 0; t8 < NT; ++t8) {
                *reinterpret_cast<double2 *>(&Y_s[g][l][8 * t8 + 2 * q]) = d[0][l][t8];
                if (g < NX) *reinterpret_cast<double2 *>(&x_s[g][l][8 * t8 + 2 * q]) = d[1][l][t8];
            }
        __syncwarp();
        // ---- owner thread: noise, outputs, trajectory stores, output half of the window
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
#pragma unroll
            for (int s = 0; s < NMPC; ++s) {
                if (s < steps) {
                    const int k = t0 + s;
                    double u[M], y[P];
#pragma unroll
                    for (int i = 0; i < M; ++i) u[i] = up_s[s * M + i][l][tl];
                    if constexpr (!PHILOX) {
#pragma unroll
                        for (int i = 0; i < P; ++i) y[i] = __ldg(a.w + (f0[l] + k) * P + i);
                    } else {
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            const int qs = s * P + i;          // word (qs & 3) of call t0*P/4 + (qs >> 2)
                            if ((qs & 3) == 0) {
                                uint32_t c0 = (uint32_t)(((unsigned)t0 * (unsigned)P) >> 2) + (uint32_t)(qs >> 2), c1 = 0u,
                                         c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                                for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                                nw[l][0] = c0; nw[l][1] = c1; nw[l][2] = c2; nw[l][3] = c3;
                            }
                            y[i] = a.eps * (2.0 * unit32_fast(nw[l][qs & 3]) - 3.0);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < P; ++i) y[i] = Y_s[s * P + i][l][tl] + y[i];
                    const size_t f = f0[l] + k;
                    if (live[l]) {
                        if (f & 1) {             // warp-uniform: completes the sector (f-1, f)
                            if (k == 0) {
                                *reinterpret_cast<double2 *>(a.u_sys + f * 2) = make_double2(u[0], u[1]);
                                *reinterpret_cast<double2 *>(a.y_sys + f * 2) = make_double2(y[0], y[1]);
                            } else {
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.u_sys + (f - 1) * 2),
                                             "d"(pu[l][0]), "d"(pu[l][1]), "d"(u[0]), "d"(u[1])
                                             : "memory");
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.y_sys + (f - 1) * 2),
                                             "d"(py[l][0]), "d"(py[l][1]), "d"(y[0]), "d"(y[1])
                                             : "memory");
                            }
                        }
                    }
                    pu[l][0] = u[0]; pu[l][1] = u[1]; py[l][0] = y[0]; py[l][1] = y[1];
#pragma unroll
                    for (int i = 0; i < P; ++i) wy_s[s * P + i][l][tl] = y[i];
                }
            }
        }
        __syncwarp();
    };

    int t0 = 0;
    for (; t0 + NMPC <= a.n_steps; t0 += NMPC) block(t0, NMPC);
    if (t0 < a.n_steps) {                        // last, partial block (controller_operation.py:278): its own block map
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ks = 0; ks < KB / 4; ++ks)
                aP[rt][ks] = (8 * rt + g < RB) ? cfp.Mt[(8 * rt + g) % RB][4 * ks + q] : 0.0;
        block(t0, n_tail);
    }
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        if (!live[l]) continue;
        const size_t fl = f0[l] + a.n_steps - 1;
        if ((fl & 1) == 0) {                     // an unpaired final element is still in (pu, py)
            *reinterpret_cast<double2 *>(a.u_sys + fl * 2) = make_double2(pu[l][0], pu[l][1]);
            *reinterpret_cast<double2 *>(a.y_sys + fl * 2) = make_double2(py[l][0], py[l][1]);
        }
        bool finite = isfinite(py[l][0]) && isfinite(py[l][1]);
#pragma unroll
        for (int i = 0; i < NX; ++i) finite = finite && isfinite(x_s[i][l][tl]);
        if (a.status) a.status[b[l]] = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
        if (a.iters) a.iters[b[l]] = (a.n_steps + NMPC - 1) / NMPC;
        if (a.x_final) {
#pragma unroll
            for (int i = 0; i < NX; ++i) a.x_final[(size_t)b[l] * NX + i] = x_s[i][l][tl];
        }
    }
}

// ===========================================================================
// Warp-specialised all-tensor-core variant.  The single-warp kernels above are latency-bound: a 65,536-loop
// batch leaves a B200 1.7 warps per scheduler, and each of them alternates between tensor-pipe phases (the two
// GEMMs), an ALU phase (Philox: ~27 % of all instructions) and the store phase.  Here a CTA is TWO warps working
// on the same 64 loops in lockstep, one block-iteration apart:
//   warp 0 (math): solve GEMM -> planned inputs -> plant GEMM, nothing else.  The measurement noise of the block
//                  is already sitting in the destination buffer of the outputs and INITIALISES the accumulators
//                  of the plant GEMM, so y = Y + w costs no instruction.
//   warp 1 (i/o) : during block t it records the trajectories of block t-1 (sector-paired stores) and draws the
//                  noise of block t+1 into the output buffer that block will use.
// One __syncthreads per block; buffers rotate (planned inputs x2, outputs x3) so the two warps never touch the
// same buffer in the same iteration except to read.  Same arithmetic as k_closed_loop_mma except that the noise
// is the first instead of the last summand of y.
// ===========================================================================
template <int N, int M, int P, int NX, int NMPC, bool PHILOX, int MW, bool MD = false>
__global__ void __launch_bounds__(32 * (MW + 1), 7)
k_closed_loop_ws(const __grid_constant__ MmaCoef<N, M, P, NX, NMPC> cfp, const FastArgs a, const int n_tail) {
    constexpr int R = NMPC * M, NW = N * (M + P), LPT = 2, TP = 36, NT = 4;
    constexpr int KB = NX + R, RB = NMPC * P + NX, RY = NMPC * P;
    constexpr int LM = MW == 1 ? LPT : 1;                      // loop groups (of 32 loops) per math warp
    constexpr int WPG = MW == 4 ? 2 : 1;                       // math warps per loop group (they split its n-tiles)
    constexpr int NTW = NT / WPG;                              // n-tiles (of 8 loops) per math warp and loop group
    static_assert(MW == 1 || MW == 2 || MW == 4, "one, two or four math warps");
    // MD (opt-in, DDMPC_WS_MATH_DRAWS=1): every math warp draws the Philox noise of its own loops and the i/o warp only
    // records.  Measured slower (0.263 vs 0.240 ms): the math warps, not the i/o warp, are the critical path.
    constexpr bool MATH_DRAWS = MD && PHILOX;
    static_assert(!MD || MW >= 2, "math warps draw their own noise only with one loop group per warp");
    static_assert(M == 2 && P == 2 && R == 8 && NMPC == N, "shape not supported by the warp-specialised kernel");
    static_assert(NW % 4 == 0 && (N * M) % 4 == 0 && KB % 4 == 0 && NX % 4 == 0 && RB <= 16 && RY == 8, "fragment tiling");
    __shared__ __align__(16) double csp_s[R][LPT][TP];         // set-point term of the planned inputs
    __shared__ __align__(16) double up_s[2][R][LPT][TP];       // planned inputs of block t (= input half of the next window)
    __shared__ __align__(16) double wy_s[3][RY][LPT][TP];      // noise, then outputs of block t (= output half of the window)
    __shared__ __align__(16) double x_s[NX][LPT][TP];          // plant state
    // Shared-memory layout [row][l][column], row stride 72 doubles (= 8 mod 16 words) and column ^= 4 on rows 2, 3
    // (mod 4): B-fragment reads (rows 4ks + q, columns g), C-fragment 128-bit accesses (rows g, columns 2q) and the
    // owner-thread accesses (row fixed, column = lane) are then all bank-conflict free.
    auto SW = [](int row, int col) { return col ^ (((row >> 1) & 1) << 2); };
    __shared__ int swap_s;
    const int tl = threadIdx.x & 31, g = tl >> 2, q = tl & 3;
    const int nblk = (a.n_steps + NMPC - 1) / NMPC;
    // Role of each warp: MW math warps (math warp j owns loop group l = j when MW == 2), then the i/o warp.
    // A CTA's warps get consecutive hardware warp slots and a slot's scheduler is slot % 4 (probed on B200:
    // scripts/probes/warp_slots*.cu).  Three-warp CTAs therefore rotate over the four schedulers by themselves;
    // two-warp CTAs would put every math warp on schedulers 0 and 2 and leave half of the SM's FP64 tensor pipes
    // idle, so bit 2 of the slot number swaps the roles of every other CTA pair.  Only performance depends on
    // this; any value of swap_s is correct.
    if (threadIdx.x == 0) {
        unsigned slot;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
        swap_s = MW == 1 ? (slot >> 2) & 1 : 0;
    }
    __syncthreads();
    const int warp = (threadIdx.x >> 5) ^ swap_s;
    const int l0 = MW == 1 ? 0 : warp / WPG;       // first loop group of this math warp
    const int t80 = (warp % WPG) * NTW;            // its first n-tile inside the group
    const int cb = g ^ (((q >> 1) & 1) << 2);      // swizzled column of a B-fragment element (row = 4ks + q)
    const int cc2 = (2 * q) ^ (((g >> 1) & 1) << 2);  // swizzled column of a C-fragment pair (row = g)

    if (warp == MW) {
        // ------------------------------------------------------------------ i/o warp: thread tl owns loops 2tl, 2tl+1
        int b[LPT];
        bool live[LPT];
        size_t f0[LPT];
        uint32_t sid_lo[LPT], sid_hi[LPT];
        double pu[LPT][M], py[LPT][P];           // previous trajectory element (sector pairing)
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            b[l] = blockIdx.x * 64 + 2 * tl + l;
            live[l] = b[l] < a.B;
            if (!live[l]) b[l] = 0;              // dead slots replay loop 0 and never store
            f0[l] = (size_t)b[l] * a.n_steps;
            const unsigned long long sid = a.id0 + (unsigned long long)b[l];
            sid_lo[l] = (uint32_t)sid;
            sid_hi[l] = (uint32_t)(sid >> 32);
            pu[l][0] = pu[l][1] = py[l][0] = py[l][1] = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) x_s[i][l][SW(i, tl)] = a.x0[(size_t)b[l] * NX + i];
#pragma unroll
            for (int i = 0; i < N * M; ++i) up_s[1][i][l][SW(i, tl)] = a.u_past0[(size_t)b[l] * N * M + i];
#pragma unroll
            for (int i = 0; i < N * P; ++i) wy_s[2][i][l][SW(i, tl)] = a.y_past0[(size_t)b[l] * N * P + i];
            double sp[M + P];
#pragma unroll
            for (int i = 0; i < M; ++i) sp[i] = a.u_s[(size_t)b[l] * M + i];
#pragma unroll
            for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[(size_t)b[l] * P + i];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < M + P; ++j) acc = fma(__ldg(a.Ksp + k * (M + P) + j), sp[j], acc);
                csp_s[k][l][SW(k, tl)] = acc;
            }
        }
        // noise of block tb into output buffer `buf` (word qs & 3 of Philox call tb*NMPC*P/4 + (qs >> 2), qs = s*P + i)
        auto draw = [&](const int tb, const int buf) {
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                if constexpr (PHILOX) {
#pragma unroll
                    for (int cc = 0; cc < RY / 4; ++cc) {
                        uint32_t c0 = (uint32_t)(((unsigned)tb * (unsigned)RY) >> 2) + (uint32_t)cc, c1 = 0u,
                                 c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                        for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                        wy_s[buf][4 * cc + 0][l][SW(4 * cc + 0, tl)] = a.eps * (2.0 * unit32_fast(c0) - 3.0);
                        wy_s[buf][4 * cc + 1][l][SW(4 * cc + 1, tl)] = a.eps * (2.0 * unit32_fast(c1) - 3.0);
                        wy_s[buf][4 * cc + 2][l][SW(4 * cc + 2, tl)] = a.eps * (2.0 * unit32_fast(c2) - 3.0);
                        wy_s[buf][4 * cc + 3][l][SW(4 * cc + 3, tl)] = a.eps * (2.0 * unit32_fast(c3) - 3.0);
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < NMPC; ++s) {
                        const int k = tb * NMPC + s;
#pragma unroll
                        for (int i = 0; i < P; ++i)
                            wy_s[buf][s * P + i][l][SW(s * P + i, tl)] = k < a.n_steps ? __ldg(a.w + (f0[l] + k) * P + i) : 0.0;
                    }
                }
            }
        };
        // record the `steps` trajectory elements of block tb (full 32-byte sectors, see emit())
        auto record = [&](const int tb, const int steps) {
            const int ub = tb & 1, yb = tb % 3;
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
#pragma unroll
                for (int s = 0; s < NMPC; ++s) {
                    if (s < steps) {
                        const int k = tb * NMPC + s;
                        double u[M], y[P];
#pragma unroll
                        for (int i = 0; i < M; ++i) u[i] = up_s[ub][s * M + i][l][SW(s * M + i, tl)];
#pragma unroll
                        for (int i = 0; i < P; ++i) y[i] = wy_s[yb][s * P + i][l][SW(s * P + i, tl)];
                        const size_t f = f0[l] + k;
                        if (live[l] && (f & 1) && !a.dbg_nostore) {   // warp-uniform: completes the sector (f-1, f)
                            if (k == 0) {
                                *reinterpret_cast<double2 *>(a.u_sys + f * 2) = make_double2(u[0], u[1]);
                                *reinterpret_cast<double2 *>(a.y_sys + f * 2) = make_double2(y[0], y[1]);
                            } else {
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.u_sys + (f - 1) * 2),
                                             "d"(pu[l][0]), "d"(pu[l][1]), "d"(u[0]), "d"(u[1])
                                             : "memory");
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.y_sys + (f - 1) * 2),
                                             "d"(py[l][0]), "d"(py[l][1]), "d"(y[0]), "d"(y[1])
                                             : "memory");
                            }
                        }
                        pu[l][0] = u[0]; pu[l][1] = u[1]; py[l][0] = y[0]; py[l][1] = y[1];
                    }
                }
            }
        };
        if (!MATH_DRAWS) draw(0, 0);
        __syncthreads();                                   // window, state and noise of block 0 are in place
        for (int t = 0; t < nblk; ++t) {
            if (t > 0) record(t - 1, NMPC);
            if (!MATH_DRAWS && t + 1 < nblk) draw(t + 1, (t + 1) % 3);
            __syncthreads();                               // block t is complete
        }
        record(nblk - 1, n_tail ? n_tail : NMPC);
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            if (!live[l]) continue;
            const size_t fl = f0[l] + a.n_steps - 1;
            if ((fl & 1) == 0) {                           // an unpaired final element is still in (pu, py)
                *reinterpret_cast<double2 *>(a.u_sys + fl * 2) = make_double2(pu[l][0], pu[l][1]);
                *reinterpret_cast<double2 *>(a.y_sys + fl * 2) = make_double2(py[l][0], py[l][1]);
            }
            bool finite = isfinite(py[l][0]) && isfinite(py[l][1]);
#pragma unroll
            for (int i = 0; i < NX; ++i) finite = finite && isfinite(x_s[i][l][SW(i, tl)]);
            if (a.status) a.status[b[l]] = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
            if (a.iters) a.iters[b[l]] = nblk;
            if (a.x_final) {
#pragma unroll
                for (int i = 0; i < NX; ++i) a.x_final[(size_t)b[l] * NX + i] = x_s[i][l][SW(i, tl)];
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- math warp
    // A fragments (row g, column 4*ks + q of each k-step) stay in registers for the whole run
    double aK[NW / 4], aP[2][KB / 4];
#pragma unroll
    for (int ks = 0; ks < NW / 4; ++ks) aK[ks] = cfp.Ku[g][4 * ks + q];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ks = 0; ks < KB / 4; ++ks) aP[rt][ks] = (8 * rt + g < RB) ? cfp.Mb[(8 * rt + g) % RB][4 * ks + q] : 0.0;
    auto mma = [](double2 &c, double av, double bv) {
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c.x), "+d"(c.y)
            : "d"(av), "d"(bv));
    };
    // noise of block tb for the warp's own loops: lane = (column, Philox call); see the i/o warp's draw()
    constexpr int CPL = RY / 4 / WPG;                      // Philox calls per lane and block
    const int ncol = WPG == 2 ? 8 * t80 + (tl >> 1) : tl, ncc0 = WPG == 2 ? (tl & 1) : 0;
    const unsigned long long nsid = a.id0 + (unsigned long long)min(blockIdx.x * 64 + 2 * ncol + l0, a.B - 1);
    auto mdraw = [&](const int tb, const int buf) {
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) {
            const int ncc = ncc0 + ci;
            uint32_t c0 = (uint32_t)(((unsigned)tb * (unsigned)RY) >> 2) + (uint32_t)ncc, c1 = 0u, c2 = (uint32_t)nsid,
                     c3 = (uint32_t)(nsid >> 32);
#pragma unroll
            for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
            wy_s[buf][4 * ncc + 0][l0][SW(4 * ncc + 0, ncol)] = a.eps * (2.0 * unit32_fast(c0) - 3.0);
            wy_s[buf][4 * ncc + 1][l0][SW(4 * ncc + 1, ncol)] = a.eps * (2.0 * unit32_fast(c1) - 3.0);
            wy_s[buf][4 * ncc + 2][l0][SW(4 * ncc + 2, ncol)] = a.eps * (2.0 * unit32_fast(c2) - 3.0);
            wy_s[buf][4 * ncc + 3][l0][SW(4 * ncc + 3, ncol)] = a.eps * (2.0 * unit32_fast(c3) - 3.0);
        }
    };
    if (MATH_DRAWS) mdraw(0, 0);
    __syncthreads();
    int cy = 0, py_ = 2;                                   // output buffers: current block, previous block
    for (int t = 0; t < nblk; ++t) {
        const int cu = t & 1, pu_ = cu ^ 1;
        if (MATH_DRAWS && t + 1 < nblk) mdraw(t + 1, cy == 2 ? 0 : cy + 1);   // that buffer was recorded during block t - 1
        if (t == nblk - 1 && n_tail != 0) {                // last, partial block (controller_operation.py:278)
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int ks = 0; ks < KB / 4; ++ks)
                    aP[rt][ks] = (8 * rt + g < RB) ? cfp.Mt[(8 * rt + g) % RB][4 * ks + q] : 0.0;
        }
        // ---- solve: U (8 x loops) = csp + Ku [window_u; window_y]
        double2 c[LM][NTW];
#pragma unroll
        for (int li = 0; li < LM; ++li)
#pragma unroll
            for (int t8 = 0; t8 < NTW; ++t8) c[li][t8] = *reinterpret_cast<const double2 *>(&csp_s[g][l0 + li][8 * (t80 + t8) + cc2]);
#pragma unroll
        for (int ks = 0; ks < NW / 4; ++ks) {
            const int e = 4 * ks + q;
#pragma unroll
            for (int li = 0; li < LM; ++li)
#pragma unroll
                for (int t8 = 0; t8 < NTW; ++t8) {
                    const double bv = (4 * ks < N * M) ? up_s[pu_][e < N * M ? e : 0][l0 + li][8 * (t80 + t8) + cb]
                                                       : wy_s[py_][e >= N * M ? e - N * M : 0][l0 + li][8 * (t80 + t8) + cb];
                    mma(c[li][t8], aK[ks], bv);
                }
        }
#pragma unroll
        for (int li = 0; li < LM; ++li)
#pragma unroll
            for (int t8 = 0; t8 < NTW; ++t8) *reinterpret_cast<double2 *>(&up_s[cu][g][l0 + li][8 * (t80 + t8) + cc2]) = c[li][t8];
        __syncwarp();
        // ---- plant: [Y; x+] = Mblk [x; U] (+ the noise waiting in the output buffer)
        double2 d[2][LM][NTW];
#pragma unroll
        for (int li = 0; li < LM; ++li)
#pragma unroll
            for (int t8 = 0; t8 < NTW; ++t8) {
                d[0][li][t8] = *reinterpret_cast<const double2 *>(&wy_s[cy][g][l0 + li][8 * (t80 + t8) + cc2]);
                d[1][li][t8] = make_double2(0.0, 0.0);
            }
#pragma unroll
        for (int ks = 0; ks < KB / 4; ++ks) {
            const int e = 4 * ks + q;
#pragma unroll
            for (int li = 0; li < LM; ++li)
#pragma unroll
                for (int t8 = 0; t8 < NTW; ++t8) {
                    const double bv = (4 * ks < NX) ? x_s[e < NX ? e : 0][l0 + li][8 * (t80 + t8) + cb]
                                                    : up_s[cu][e >= NX ? e - NX : 0][l0 + li][8 * (t80 + t8) + cb];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) mma(d[rt][li][t8], aP[rt][ks], bv);
                }
        }
        __syncwarp();                                      // every lane has read the old state
#pragma unroll
        for (int li = 0; li < LM; ++li)
#pragma unroll
            for (int t8 = 0; t8 < NTW; ++t8) {
                *reinterpret_cast<double2 *>(&wy_s[cy][g][l0 + li][8 * (t80 + t8) + cc2]) = d[0][li][t8];
                if (g < NX) *reinterpret_cast<double2 *>(&x_s[g][l0 + li][8 * (t80 + t8) + cc2]) = d[1][li][t8];
            }
        py_ = cy;
        cy = cy == 2 ? 0 : cy + 1;
        __syncthreads();                                   // block t is complete
    }
}

// ===========================================================================
// Register-chained all-tensor-core variant: no shared memory, no barriers.
//
// The products are computed TRANSPOSED, loops along M:   U^T (8 loops x 8) = W^T (8 x 16) Ku^T,
// [Y; x+]^T (8 loops x 12) = [x; U]^T (8 x 12) Mblk^T.   With the m8n8k4 fragment layouts (lane = 4g + q:
// A[g][q], B[q][g], C[g][2q..2q+1]) the C fragment of one product - lane (g, q) holds outputs 2q, 2q+1 of loop g -
// is directly a pair of A fragments of the next one, because the order in which a dot product visits its terms is
// free: k-step "0" takes entry 2q from lane q and k-step "1" entry 2q+1, and that permutation is folded into the
// constant B operands (the coefficient matrices, held in registers for the whole run).  So the planned inputs feed
// the plant product, and both feed the next solve, without ever leaving the register file; only the 4 plant
// states are re-spread over the lanes with two shuffles.  Lane (g, q) ends up holding step q of the block for loop
// g: it draws that step's noise (one Philox call, the accumulator of the plant product starts from it), and stores
// that step's u and y (16 B each; the four lanes of a loop write 64 contiguous bytes).
// A warp carries NT m-tiles (8 NT loops) as independent DMMA chains.
// ===========================================================================
// XP (needs NT = 4): the block's results are transposed through 4 KB of shared memory private to the warp, so that
// lane l records loop l of the warp's 32 as full 32-byte sectors (the pairing rule of emit()) instead of 16-byte pieces.
template <int N, int M, int P, int NX, int NMPC, bool PHILOX, int NT, bool XP = false>
__global__ void __launch_bounds__(32, NT >= 8 ? 7 : 14)
k_closed_loop_reg(const __grid_constant__ MmaCoef<N, M, P, NX, NMPC> cfp, const FastArgs a, const int n_tail) {
    constexpr int R = NMPC * M, RY = NMPC * P, NU = N * M;
    static_assert(!XP || NT == 4, "the transposing variant carries 32 loops per warp");
    __shared__ __align__(16) double2 xu_s[XP ? 32 * NMPC : 1], xy_s[XP ? 32 * NMPC : 1];
    // chunk (loop, step) -> 16-byte slot: conflict-free for the writers (lane (g, q): loop 8mt + g, step q) and the reader
    // (lane l: loop l, one step at a time); see k_closed_loop_rws
    auto CH = [](int loop, int s) { return ((loop ^ ((loop >> 3) & 1)) << 2) | (s ^ ((loop >> 1) & 3)); };
    static_assert(M == 2 && P == 2 && R == 8 && RY == 8 && NX == 4 && NMPC == N, "shape not supported by the register-chained kernel");
    const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
    const int nblk = (a.n_steps + NMPC - 1) / NMPC;
    // constant B fragments: lane (g, q) holds B[k = q][n = g] of every k-step
    double bK[4], bP[2][3];
    bK[0] = cfp.Ku[g][2 * q];          bK[1] = cfp.Ku[g][2 * q + 1];             // window inputs  2q, 2q+1
    bK[2] = cfp.Ku[g][NU + 2 * q];     bK[3] = cfp.Ku[g][NU + 2 * q + 1];        // window outputs 2q, 2q+1
    auto load_plant = [&](const double (&Mb)[NMPC * P + NX][NX + NMPC * M]) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
            const int row = 8 * tile + g;
            const bool valid = row < RY + NX;
            const int rr = valid ? row : 0;
            bP[tile][0] = valid ? Mb[rr][q] : 0.0;                               // state entry q
            bP[tile][1] = valid ? Mb[rr][NX + 2 * q] : 0.0;                      // planned inputs 2q, 2q+1
            bP[tile][2] = valid ? Mb[rr][NX + 2 * q + 1] : 0.0;
        }
    };
    load_plant(cfp.Mb);
    // Keep the coefficients in registers: left alone, the compiler re-fetches them inside the loop with LANE-INDEXED
    // constant loads (c[0x0][R + off]: 32 different addresses per warp = 32 serialised constant-cache accesses each).
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("" : "+d"(bK[i]));
#pragma unroll
    for (int i = 0; i < 6; ++i) asm volatile("" : "+d"(bP[i / 3][i % 3]));
    auto mma = [](double2 &c, double av, double bv) {
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c.x), "+d"(c.y)
            : "d"(av), "d"(bv));
    };
    // per m-tile state of loop g: lane (g, q) holds entries 2q, 2q+1 of the window halves, entry q of the state
    int b[NT];
    bool live[NT];
    double2 uC[NT], yC[NT], csp[NT], xC[NT];
    double xA[NT];
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
        b[mt] = (blockIdx.x * NT + mt) * 8 + g;
        live[mt] = b[mt] < a.B;
        if (!live[mt]) b[mt] = a.B - 1;                        // dead rows replay the last loop and never store
        const size_t bb = (size_t)b[mt];
        uC[mt] = *reinterpret_cast<const double2 *>(a.u_past0 + bb * NU + 2 * q);
        yC[mt] = *reinterpret_cast<const double2 *>(a.y_past0 + bb * (N * P) + 2 * q);
        xA[mt] = a.x0[bb * NX + q];
        xC[mt] = make_double2(0.0, 0.0);
        double sp[M + P];
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = a.u_s[bb * M + i];
#pragma unroll
        for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[bb * P + i];
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int j = 0; j < M + P; ++j) {
            c0 = fma(__ldg(a.Ksp + (2 * q) * (M + P) + j), sp[j], c0);
            c1 = fma(__ldg(a.Ksp + (2 * q + 1) * (M + P) + j), sp[j], c1);
        }
        csp[mt] = make_double2(c0, c1);
    }
    // transposing variant: lane records loop `bo`
    const int bo = blockIdx.x * NT * 8 + lane;
    const bool olive = bo < a.B;
    const size_t of0 = (size_t)(olive ? bo : 0) * a.n_steps;
    double2 pu = make_double2(0.0, 0.0), py = pu;              // previous trajectory element (sector pairing)
    for (int t = 0; t < nblk; ++t) {
        const int steps = (t == nblk - 1 && n_tail != 0) ? n_tail : NMPC;
        if (t == nblk - 1 && n_tail != 0) load_plant(cfp.Mt);   // last, partial block (controller_operation.py:278)
        // ---- solve: U^T = csp + W^T Ku^T   (k-steps outermost: consecutive DMMAs hit different accumulators)
        double2 nu[NT];
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) nu[mt] = csp[mt];
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) mma(nu[mt], uC[mt].x, bK[0]);
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) mma(nu[mt], uC[mt].y, bK[1]);
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) mma(nu[mt], yC[mt].x, bK[2]);
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) mma(nu[mt], yC[mt].y, bK[3]);
        // ---- measurement noise of step q of the block (the accumulator of the output product starts from it)
        const int k = t * NMPC + q;
        double2 d0[NT], d1[NT];
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) {
            if constexpr (PHILOX) {
                // noise word qs = q*P + i of the block is word (qs & 3) of Philox call t*RY/4 + (qs >> 2)
                const unsigned long long sid = a.id0 + (unsigned long long)b[mt];
                uint32_t c0 = (uint32_t)(((unsigned)t * (unsigned)RY) >> 2) + (uint32_t)(q >> 1), c1 = 0u, c2 = (uint32_t)sid,
                         c3 = (uint32_t)(sid >> 32);
#pragma unroll
                for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                const uint32_t w0 = (q & 1) ? c2 : c0, w1 = (q & 1) ? c3 : c1;
                d0[mt] = make_double2(a.eps * (2.0 * unit32_fast(w0) - 3.0), a.eps * (2.0 * unit32_fast(w1) - 3.0));
            } else {
                d0[mt] = k < a.n_steps ? *reinterpret_cast<const double2 *>(a.w + ((size_t)b[mt] * a.n_steps + k) * P)
                                       : make_double2(0.0, 0.0);
            }
            d1[mt] = make_double2(0.0, 0.0);
        }
        // ---- plant: [Y; x+]^T = w + [x; U]^T Mblk^T
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) { mma(d0[mt], xA[mt], bP[0][0]); mma(d1[mt], xA[mt], bP[1][0]); }
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) { mma(d0[mt], nu[mt].x, bP[0][1]); mma(d1[mt], nu[mt].x, bP[1][1]); }
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) { mma(d0[mt], nu[mt].y, bP[0][2]); mma(d1[mt], nu[mt].y, bP[1][2]); }
        // ---- record step q, hand the block over to the next one
        if constexpr (XP) {
            __syncwarp();                                      // the previous block has been read out
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) {
                xu_s[CH(8 * mt + g, q)] = nu[mt];
                xy_s[CH(8 * mt + g, q)] = d0[mt];
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < NMPC; ++s) {
                if (s < steps) {
                    const double2 u = xu_s[CH(lane, s)], y = xy_s[CH(lane, s)];
                    const size_t f = of0 + (size_t)(t * NMPC + s);
                    if (olive && (f & 1) && !a.dbg_nostore) {      // completes the sector (f - 1, f)
                        if (t == 0 && s == 0) {
                            *reinterpret_cast<double2 *>(a.u_sys + f * 2) = u;
                            *reinterpret_cast<double2 *>(a.y_sys + f * 2) = y;
                        } else {
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.u_sys + (f - 1) * 2), "d"(pu.x),
                                         "d"(pu.y), "d"(u.x), "d"(u.y)
                                         : "memory");
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.y_sys + (f - 1) * 2), "d"(py.x),
                                         "d"(py.y), "d"(y.x), "d"(y.y)
                                         : "memory");
                        }
                    }
                    pu = u;
                    py = y;
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) {
            if (!XP && live[mt] && q < steps && !a.dbg_nostore) {
                const size_t f = (size_t)b[mt] * a.n_steps + k;
                *reinterpret_cast<double2 *>(a.u_sys + f * M) = nu[mt];
                *reinterpret_cast<double2 *>(a.y_sys + f * P) = d0[mt];
            }
            uC[mt] = nu[mt];
            yC[mt] = d0[mt];
            xC[mt] = d1[mt];
            // state entry q of loop g sits in lane (g, q >> 1), component q & 1
            const int src = (lane & ~3) | (q >> 1);
            const double v0 = __shfl_sync(0xffffffffu, d1[mt].x, src), v1 = __shfl_sync(0xffffffffu, d1[mt].y, src);
            xA[mt] = (q & 1) ? v1 : v0;
        }
    }
    if constexpr (XP) {
        const size_t fl = of0 + a.n_steps - 1;
        if (olive && (fl & 1) == 0 && !a.dbg_nostore) {          // an unpaired final element is still in (pu, py)
            *reinterpret_cast<double2 *>(a.u_sys + fl * 2) = pu;
            *reinterpret_cast<double2 *>(a.y_sys + fl * 2) = py;
        }
    }
    // ---- per-loop results (loop g = lanes 4g .. 4g+3)
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
        const int sl = (n_tail ? n_tail : NMPC) - 1;             // last recorded step of the last block
        bool fin = isfinite(xA[mt]) && (q != sl || (isfinite(yC[mt].x) && isfinite(yC[mt].y)));
        const unsigned badm = __ballot_sync(0xffffffffu, !fin);
        const bool loop_bad = ((badm >> (4 * g)) & 0xfu) != 0u;
        if (live[mt] && q == 0) {
            if (a.status) a.status[b[mt]] = loop_bad ? DDMPC_SOLVE_NONFINITE : DDMPC_SOLVE_OPTIMAL;
            if (a.iters) a.iters[b[mt]] = nblk;
        }
        if (live[mt] && a.x_final) a.x_final[(size_t)b[mt] * NX + q] = xA[mt];
    }
}

// ===========================================================================
// Register-chained math warps + a DECOUPLED i/o warp (the two ideas above combined).
//
// Two math warps run the register-chained recurrence of k_closed_loop_reg on 4 m-tiles (32 loops) each; what
// they exchange with the i/o warp is only the block's results and its noise, as 16-byte chunks (one step of one
// loop: lane (g, q) of m-tile mt owns chunk (loop 8mt + g, step q)): per block and m-tile one LDS.128 (the noise,
// which initialises the output accumulator) and two STS.128 (planned inputs, outputs) instead of the 12 shared-
// memory accesses per m-tile of k_closed_loop_ws.  The i/o warp is that kernel's: lane tl owns loops 2tl, 2tl+1,
// draws noise and records results as full 32-byte sectors.
// Because the window lives in registers, the shared buffers are pure hand-over queues, so the warps need not run
// in lock-step: instead of one __syncthreads per block there are two rings of mbarriers,
//     full[t % 3]  "noise of block t is in wy_s[t % 3]"            i/o lane 0 arrives, math warps wait
//     done[t % 3]  "results of block t are in up_s[t % 3], wy_s[t % 3]"   lane 0 of each math warp arrives, i/o waits
// and the i/o warp's order   draw(j + 2); wait done[j]; record(j)   lets the math warps run up to two blocks ahead of
// the trajectory stores (full[t] is signalled after record(t - 3), which is what frees buffer t % 3), and the math
// warps never wait for each other.
// Chunk (loop, s) sits at 16-byte slot ((loop ^ bit3(loop)) << 2) | (s ^ ((loop >> 1) & 3)): a quarter-warp of a
// math warp (loops 2j, 2j+1, all four steps) and a quarter-warp of the i/o warp (8 consecutive even or odd loops,
// one step) then both touch eight different 16-byte bank groups.
// ===========================================================================
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!ok);
}

template <int N, int M, int P, int NX, int NMPC, bool PHILOX>
__global__ void __launch_bounds__(96, 7)   // 7 CTAs = 21 warps per SM = 6 per scheduler -> at most 80 registers (16K per scheduler)
k_closed_loop_rws(const __grid_constant__ MmaCoef<N, M, P, NX, NMPC> cfp, const FastArgs a, const int n_tail) {
    constexpr int R = NMPC * M, RY = NMPC * P, NU = N * M, NT = 4, LC = 64, LPT = 2;
    static_assert(M == 2 && P == 2 && R == 8 && RY == 8 && NX == 4 && NMPC == N && NMPC == 4, "shape not supported by the register-chained kernel");
    __shared__ __align__(16) double2 wy_s[3][LC * NMPC];       // noise, then outputs, of block t in buffer t % 3
    __shared__ __align__(16) double2 up_s[3][LC * NMPC];       // planned inputs of block t in buffer t % 3
    __shared__ __align__(16) double2 csp_s[LC * NMPC];         // set-point term of the planned inputs (same chunk layout)
    __shared__ __align__(8) uint64_t full_b[3], done_b[3];
    auto CH = [](int loop, int s) { return ((loop ^ ((loop >> 3) & 1)) << 2) | (s ^ ((loop >> 1) & 3)); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (a.n_steps + NMPC - 1) / NMPC;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            mbar_init(&full_b[i], 1);
            mbar_init(&done_b[i], 2);
        }
    }
    __syncthreads();

    if (warp == 2) {
        // ------------------------------------------------------------------ i/o warp: lane owns loops 2 lane, 2 lane + 1
        int b[LPT];
        bool live[LPT];
        size_t f0[LPT];
        uint32_t sid_lo[LPT], sid_hi[LPT];
        double2 pu[LPT], py[LPT];                // previous trajectory element (sector pairing)
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            b[l] = blockIdx.x * LC + 2 * lane + l;
            live[l] = b[l] < a.B;
            if (!live[l]) b[l] = a.B - 1;        // dead slots replay the last loop and never store
            f0[l] = (size_t)b[l] * a.n_steps;
            const unsigned long long sid = a.id0 + (unsigned long long)b[l];
            sid_lo[l] = (uint32_t)sid;
            sid_hi[l] = (uint32_t)(sid >> 32);
            pu[l] = py[l] = make_double2(0.0, 0.0);
        }
        // noise of block tb into wy_s[tb & 3] (word qs & 3 of Philox call tb*NMPC*P/4 + (qs >> 2), qs = s*P + i), then signal
        auto draw = [&](const int tb, const int buf) {
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                const int loop = 2 * lane + l;
                if constexpr (PHILOX) {
#pragma unroll
                    for (int cc = 0; cc < RY / 4; ++cc) {
                        uint32_t c0 = (uint32_t)(((unsigned)tb * (unsigned)RY) >> 2) + (uint32_t)cc, c1 = 0u,
                                 c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                        for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                        wy_s[buf][CH(loop, 2 * cc)] =
                            make_double2(a.eps * (2.0 * unit32_fast(c0) - 3.0), a.eps * (2.0 * unit32_fast(c1) - 3.0));
                        wy_s[buf][CH(loop, 2 * cc + 1)] =
                            make_double2(a.eps * (2.0 * unit32_fast(c2) - 3.0), a.eps * (2.0 * unit32_fast(c3) - 3.0));
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < NMPC; ++s) {
                        const int k = tb * NMPC + s;
                        wy_s[buf][CH(loop, s)] = k < a.n_steps ? __ldg(reinterpret_cast<const double2 *>(a.w + (f0[l] + k) * P))
                                                               : make_double2(0.0, 0.0);
                    }
                }
            }
            __syncwarp();                                  // every lane's chunks are written before lane 0 signals
            if (lane == 0) mbar_arrive(&full_b[buf]);
        };
        // record the `steps` trajectory elements of block tb (full 32-byte sectors, see emit())
        auto record = [&](const int tb, const int ub, const int steps) {
            const int yb = ub;
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                const int loop = 2 * lane + l;
#pragma unroll
                for (int s = 0; s < NMPC; ++s) {
                    if (s < steps) {
                        const int k = tb * NMPC + s;
                        const double2 u = up_s[ub][CH(loop, s)], y = wy_s[yb][CH(loop, s)];
                        const size_t f = f0[l] + k;
                        if (live[l] && (f & 1) && !a.dbg_nostore) {   // warp-uniform: completes the sector (f-1, f)
                            if (k == 0) {
                                *reinterpret_cast<double2 *>(a.u_sys + f * 2) = u;
                                *reinterpret_cast<double2 *>(a.y_sys + f * 2) = y;
                            } else {
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.u_sys + (f - 1) * 2),
                                             "d"(pu[l].x), "d"(pu[l].y), "d"(u.x), "d"(u.y)
                                             : "memory");
                                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.y_sys + (f - 1) * 2),
                                             "d"(py[l].x), "d"(py[l].y), "d"(y.x), "d"(y.y)
                                             : "memory");
                            }
                        }
                        pu[l] = u;
                        py[l] = y;
                    }
                }
            }
        };
        draw(0, 0);
        if (nblk > 1) draw(1, 1);
        int ub = 0;                                        // j % 3
        unsigned ph = 0u;                                  // (j / 3) & 1
        for (int j = 0; j < nblk; ++j) {
            if (j + 2 < nblk) draw(j + 2, ub == 0 ? 2 : ub - 1);   // (j + 2) % 3: that buffer was last read by record(j - 1)
            mbar_wait(&done_b[ub], ph);
            record(j, ub, (j == nblk - 1 && n_tail) ? n_tail : NMPC);
            if (ub == 2) { ub = 0; ph ^= 1u; } else ++ub;
        }
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            const size_t fl = f0[l] + a.n_steps - 1;
            if (live[l] && (fl & 1) == 0 && !a.dbg_nostore) {   // an unpaired final element is still in (pu, py)
                *reinterpret_cast<double2 *>(a.u_sys + fl * 2) = pu[l];
                *reinterpret_cast<double2 *>(a.y_sys + fl * 2) = py[l];
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- math warps (see k_closed_loop_reg)
    const int g = lane >> 2, q = lane & 3;
    double bK[4], bP[2][3];
    bK[0] = cfp.Ku[g][2 * q];          bK[1] = cfp.Ku[g][2 * q + 1];
    bK[2] = cfp.Ku[g][NU + 2 * q];     bK[3] = cfp.Ku[g][NU + 2 * q + 1];
    auto load_plant = [&](const double (&Mb)[NMPC * P + NX][NX + NMPC * M]) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
            const int row = 8 * tile + g;
            const bool valid = row < RY + NX;
            const int rr = valid ? row : 0;
            bP[tile][0] = valid ? Mb[rr][q] : 0.0;
            bP[tile][1] = valid ? Mb[rr][NX + 2 * q] : 0.0;
            bP[tile][2] = valid ? Mb[rr][NX + 2 * q + 1] : 0.0;
        }
    };
    load_plant(cfp.Mb);
    // Keep the coefficients in registers: left alone, the compiler re-fetches them inside the loop with LANE-INDEXED
    // constant loads (c[0x0][R + off], 32 different addresses per warp = 32 serialised constant-cache accesses each).
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("" : "+d"(bK[i]));
#pragma unroll
    for (int i = 0; i < 6; ++i) asm volatile("" : "+d"(bP[i / 3][i % 3]));
    auto mma = [](double2 &c, double av, double bv) {
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c.x), "+d"(c.y)
            : "d"(av), "d"(bv));
    };
    // chunk of m-tile mt: CH(32 warp + 8 mt + g, q) = ((mt & 1) ? chO : chE) + 32 mt
    const int chE = CH(32 * warp + g, q), chO = CH(32 * warp + (g ^ 1), q);
    const int b0 = blockIdx.x * LC + 32 * warp + g;
    double2 uC[NT], yC[NT];
    double xA[NT];
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
        const size_t bb = (size_t)min(b0 + 8 * mt, a.B - 1);    // dead rows replay the last loop and never store
        uC[mt] = *reinterpret_cast<const double2 *>(a.u_past0 + bb * NU + 2 * q);
        yC[mt] = *reinterpret_cast<const double2 *>(a.y_past0 + bb * (N * P) + 2 * q);
        xA[mt] = a.x0[bb * NX + q];
        double sp[M + P];
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = a.u_s[bb * M + i];
#pragma unroll
        for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[bb * P + i];
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int j = 0; j < M + P; ++j) {
            c0 = fma(__ldg(a.Ksp + (2 * q) * (M + P) + j), sp[j], c0);
            c1 = fma(__ldg(a.Ksp + (2 * q + 1) * (M + P) + j), sp[j], c1);
        }
        csp_s[((mt & 1) ? chO : chE) + 32 * mt] = make_double2(c0, c1);   // read back by the same lane only
    }
    int cu = 0;                                            // t % 3
    unsigned ph = 0u;                                      // (t / 3) & 1
    for (int t = 0; t < nblk; ++t) {
        if (t == nblk - 1 && n_tail != 0) {                // last, partial block (controller_operation.py:278)
            load_plant(cfp.Mt);
#pragma unroll
            for (int i = 0; i < 6; ++i) asm volatile("" : "+d"(bP[i / 3][i % 3]));
        }
        // two m-tiles at a time: 80 registers hold the windows of all four but the accumulators of only two
#pragma unroll
        for (int h = 0; h < NT; h += 2) {
            double2 nu[2], d0[2], d1[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                d1[i] = make_double2(0.0, 0.0);
                nu[i] = csp_s[(((h + i) & 1) ? chO : chE) + 32 * (h + i)];
            }
            // the input half of the solve does not need the noise: start it before waiting
#pragma unroll
            for (int i = 0; i < 2; ++i) { mma(nu[i], uC[h + i].x, bK[0]); mma(d1[i], xA[h + i], bP[1][0]); }
#pragma unroll
            for (int i = 0; i < 2; ++i) mma(nu[i], uC[h + i].y, bK[1]);
            if (h == 0) mbar_wait(&full_b[cu], ph);        // noise of block t is there, and buffer cu has been recorded
#pragma unroll
            for (int i = 0; i < 2; ++i) d0[i] = wy_s[cu][(((h + i) & 1) ? chO : chE) + 32 * (h + i)];   // accumulator starts from the noise
#pragma unroll
            for (int i = 0; i < 2; ++i) { mma(nu[i], yC[h + i].x, bK[2]); mma(d0[i], xA[h + i], bP[0][0]); }
#pragma unroll
            for (int i = 0; i < 2; ++i) mma(nu[i], yC[h + i].y, bK[3]);
#pragma unroll
            for (int i = 0; i < 2; ++i) { mma(d0[i], nu[i].x, bP[0][1]); mma(d1[i], nu[i].x, bP[1][1]); }
#pragma unroll
            for (int i = 0; i < 2; ++i) { mma(d0[i], nu[i].y, bP[0][2]); mma(d1[i], nu[i].y, bP[1][2]); }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int mt = h + i;
                up_s[cu][((mt & 1) ? chO : chE) + 32 * mt] = nu[i];
                wy_s[cu][((mt & 1) ? chO : chE) + 32 * mt] = d0[i];
                uC[mt] = nu[i];
                yC[mt] = d0[i];
                const int src = (lane & ~3) | (q >> 1);          // state entry q of loop g: lane (g, q >> 1), component q & 1
                const double v0 = __shfl_sync(0xffffffffu, d1[i].x, src), v1 = __shfl_sync(0xffffffffu, d1[i].y, src);
                xA[mt] = (q & 1) ? v1 : v0;
            }
        }
        __syncwarp();                                      // every lane's chunks are written before lane 0 signals
        if (lane == 0) mbar_arrive(&done_b[cu]);
        if (cu == 2) { cu = 0; ph ^= 1u; } else ++cu;
    }
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
        const int bm = b0 + 8 * mt;
        const bool live = bm < a.B;
        const int sl = (n_tail ? n_tail : NMPC) - 1;             // last recorded step of the last block
        const bool fin = isfinite(xA[mt]) && (q != sl || (isfinite(yC[mt].x) && isfinite(yC[mt].y)));
        const unsigned badm = __ballot_sync(0xffffffffu, !fin);
        const bool loop_bad = ((badm >> (4 * g)) & 0xfu) != 0u;
        if (live && q == 0) {
            if (a.status) a.status[bm] = loop_bad ? DDMPC_SOLVE_NONFINITE : DDMPC_SOLVE_OPTIMAL;
            if (a.iters) a.iters[bm] = nblk;
        }
        if (live && a.x_final) a.x_final[(size_t)bm * NX + q] = xA[mt];
    }
}

// s-step block map of the plant into Mout (rows y_0..y_{NMPC-1} (zero beyond s), then x_s; columns x_0, u_0..)
template <int M, int P, int NX, int NMPC>
static void host_block_map(const ddmpc_plant *pl, int s, double (&Mout)[NMPC * P + NX][NX + NMPC * M]) {
    double Ap[NMPC + 1][NX][NX] = {};
    for (int i = 0; i < NX; ++i) Ap[0][i][i] = 1.0;
    for (int k = 1; k <= s; ++k)
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) {
                double acc = 0.0;
                for (int l = 0; l < NX; ++l) acc += pl->A[i * NX + l] * Ap[k - 1][l][j];
                Ap[k][i][j] = acc;
            }
    double AB[NMPC][NX][M] = {};
    for (int k = 0; k < s; ++k)
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < M; ++j) {
                double acc = 0.0;
                for (int l = 0; l < NX; ++l) acc += Ap[k][i][l] * pl->B[l * M + j];
                AB[k][i][j] = acc;
            }
    for (auto &row : Mout)
        for (double &v : row) v = 0.0;
    for (int k = 0; k < s; ++k)
        for (int i = 0; i < P; ++i) {
            for (int c = 0; c < NX; ++c) {
                double acc = 0.0;
                for (int l = 0; l < NX; ++l) acc += pl->C[i * NX + l] * Ap[k][l][c];
                Mout[k * P + i][c] = acc;
            }
            for (int j = 0; j < k; ++j)
                for (int c = 0; c < M; ++c) {
                    double acc = 0.0;
                    for (int l = 0; l < NX; ++l) acc += pl->C[i * NX + l] * AB[k - 1 - j][l][c];
                    Mout[k * P + i][NX + j * M + c] = acc;
                }
            for (int c = 0; c < M; ++c) Mout[k * P + i][NX + k * M + c] = pl->D[i * M + c];
        }
    for (int i = 0; i < NX; ++i) {
        for (int c = 0; c < NX; ++c) Mout[NMPC * P + i][c] = Ap[s][i][c];
        for (int j = 0; j < s; ++j)
            for (int c = 0; c < M; ++c) Mout[NMPC * P + i][NX + j * M + c] = AB[s - 1 - j][i][c];
    }
}

// gather the Ksp block of Ku (rows 0..NMPC*M-1, columns n*(m+p)..nth-1) into a dense device array
__global__ void k_gather_ksp(const double *__restrict__ Ku, int nth, int nw, int rows, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsp = nth - nw;
    if (e < rows * nsp) out[e] = Ku[(size_t)(e / nsp) * nth + nw + (e % nsp)];
}

template <int N, int M, int P, int NX, int NMPC>
static int launch_fast(const ddmpc_set *set, const ddmpc_plant *plant, const FastArgs &fa, cudaStream_t st) {
    using Coef = FastCoef<N, M, P, NX, NMPC>;
    static_assert(sizeof(Coef) <= 3584, "coefficients must fit in the kernel parameter space");
    const Dims &d = set->plan.d;
    constexpr int NW = N * (M + P);
    // host copy of the gain rows (cached in the set after the first call)
    auto &cache = set->fast_host;
    const size_t need = (size_t)NMPC * M * d.nth;
    if (cache.size() != need) {
        cache.resize(need);
        DDMPC_CUDA(cudaMemcpy(cache.data(), set->plan.Ku.d(), sizeof(double) * need, cudaMemcpyDeviceToHost));
        DDMPC_CUDA(set->fast_ksp.alloc(sizeof(double) * NMPC * M * (M + P)));
        k_gather_ksp<<<ceil_div(NMPC * M * (M + P), 128), 128, 0, st>>>(set->plan.Ku.d(), d.nth, NW, NMPC * M,
                                                                          set->fast_ksp.d());
        DDMPC_LAUNCH_CHECK();
    }
    Coef cf;
    for (int k = 0; k < NMPC * M; ++k)
        for (int j = 0; j < NW; ++j) cf.Kt[j][k] = cache[(size_t)k * d.nth + j];
    for (int i = 0; i < NX; ++i) {
        for (int j = 0; j < NX; ++j) cf.A[i][j] = plant->A[i * NX + j];
        for (int j = 0; j < M; ++j) cf.B[i][j] = plant->B[i * M + j];
    }
    for (int i = 0; i < P; ++i) {
        for (int j = 0; j < NX; ++j) cf.C[i][j] = plant->C[i * NX + j];
        for (int j = 0; j < M; ++j) cf.D[i][j] = plant->D[i * M + j];
    }
    FastArgs a = fa;
    a.Ksp = set->fast_ksp.d();
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)a.seed + (uint32_t)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(a.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    // 32-byte pairing needs 16-byte elements (M == 2 and P == 2) and 32-byte aligned outputs
    const bool pair = (M == 2 && P == 2) && ((reinterpret_cast<uintptr_t>(a.u_sys) & 31) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.y_sys) & 31) == 0);
    // loops per thread: 2 once the batch is large enough to keep every SM busy with half the warps
    int lpt = a.B >= 16384 ? 2 : 1;
    if (const char *e = getenv("DDMPC_LPT")) lpt = (e[0] == '2') ? 2 : 1;
    if (d.convex) {
        // fused CONVEX path: slack rows through the tensor-core solve (needs LPT = 2 and 8 planned-input rows)
        constexpr bool CVX_OK = (NMPC * M == 8) && (NMPC % N == 0) && ((N * M) % 4 == 0) && ((N * P) % 4 == 0) &&
                                ((M + P) % 4 == 0);
        if constexpr (CVX_OK) {
            if (d.nb > 64) return -1;
            a.Ks = set->plan.Ks.d(); a.Phi = set->plan.Phi.d(); a.Psi = set->plan.Psi.d();
            a.bound = set->plan.bound; a.nb = d.nb; a.nth = d.nth;
            const dim3 grid(ceil_div(a.B, 64));
            if (a.w) {
                if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, 2, false, true, true><<<grid, 32, 0, st>>>(cf, a);
                else k_closed_loop_fast<N, M, P, NX, NMPC, 2, false, false, true><<<grid, 32, 0, st>>>(cf, a);
            } else {
                if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, 2, true, true, true><<<grid, 32, 0, st>>>(cf, a);
                else k_closed_loop_fast<N, M, P, NX, NMPC, 2, true, false, true><<<grid, 32, 0, st>>>(cf, a);
            }
            DDMPC_LAUNCH_CHECK();
            return DDMPC_OK;
        } else {
            return -1;
        }
    }
    {   // all-tensor-core variant (solve AND plant on the FP64 MMA pipe) for the four-tank n-step shape.
        // Measured equal to the hybrid kernel (0.255 vs 0.250 ms on config 3: both are latency-bound at 1.7
        // warps per scheduler, tensor pipe 47 % busy), so it is opt-in (DDMPC_PLANT_MMA=1): the hybrid kernel
        // steps the plant literally as the reference does instead of through a block map.
        constexpr bool MMA_OK = (M == 2 && P == 2 && NMPC * M == 8 && NMPC == N && (N * M) % 4 == 0 &&
                                 (NX + NMPC * M) % 4 == 0 && NMPC * P + NX <= 16 && NMPC * P == 8);
        const char *e = getenv("DDMPC_PLANT_MMA");
        const bool want = e && e[0] == '1';
        // warp-specialised variant (math warp + i/o warp per 64 loops): the default for large batches
        const char *ews = getenv("DDMPC_WS");
        const bool want_ws = ews ? ews[0] == '1' : true;
        if constexpr (MMA_OK) {
            // register-chained variant: opt-in (DDMPC_REG=1).  Its recurrence never leaves the register file (compute alone:
            // 0.19 ms on config 3 with NT = 2 against 0.213 ms for the warp-specialised kernel), but a lane then owns ONE
            // step of a loop, i.e. 16-byte stores, and those cost far more than the shared-memory hand-over they save
            // (0.33-0.44 ms in total); pairing them into sectors needs the same transposition again.
            const char *ereg = getenv("DDMPC_REG");
            const bool want_reg = ereg ? (ereg[0] >= '1' && ereg[0] <= '3') : false;
            if constexpr (NX == 4) {
                if (want_reg && want_ws && !want && pair && lpt == 2 && (reinterpret_cast<uintptr_t>(a.u_past0) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.y_past0) & 15) == 0 && (!a.w || (reinterpret_cast<uintptr_t>(a.w) & 15) == 0)) {
                    MmaCoef<N, M, P, NX, NMPC> mc;
                    for (int k = 0; k < NMPC * M; ++k)
                        for (int j = 0; j < NW; ++j) mc.Ku[k][j] = cache[(size_t)k * d.nth + j];
                    host_block_map<M, P, NX, NMPC>(plant, NMPC, mc.Mb);
                    const int n_tail = a.n_steps % NMPC;
                    host_block_map<M, P, NX, NMPC>(plant, n_tail ? n_tail : NMPC, mc.Mt);
                    if (const char *ens = getenv("DDMPC_DEBUG_NOSTORE")) a.dbg_nostore = ens[0] == '1';
                    const char *ent = getenv("DDMPC_REG_NT");
                    const int nt = ent ? atoi(ent) : 4;
                    if (ereg[0] == '2') {                            // register-chained math warps + decoupled i/o warp
                        auto go = [&](auto kern) -> int {
                            DDMPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                            kern<<<ceil_div(a.B, 64), 96, 0, st>>>(mc, a, n_tail);
                            return DDMPC_OK;
                        };
                        const int rc = a.w ? go(k_closed_loop_rws<N, M, P, NX, NMPC, false>) : go(k_closed_loop_rws<N, M, P, NX, NMPC, true>);
                        if (rc != DDMPC_OK) return rc;
                    } else if (ereg[0] == '3') {                     // register-chained, results transposed per warp (sector stores)
                        if (a.w) k_closed_loop_reg<N, M, P, NX, NMPC, false, 4, true><<<ceil_div(a.B, 32), 32, 0, st>>>(mc, a, n_tail);
                        else k_closed_loop_reg<N, M, P, NX, NMPC, true, 4, true><<<ceil_div(a.B, 32), 32, 0, st>>>(mc, a, n_tail);
                    } else if (nt == 2) {
                        if (a.w) k_closed_loop_reg<N, M, P, NX, NMPC, false, 2><<<ceil_div(a.B, 16), 32, 0, st>>>(mc, a, n_tail);
                        else k_closed_loop_reg<N, M, P, NX, NMPC, true, 2><<<ceil_div(a.B, 16), 32, 0, st>>>(mc, a, n_tail);
                    } else if (nt == 8) {
                        if (a.w) k_closed_loop_reg<N, M, P, NX, NMPC, false, 8><<<ceil_div(a.B, 64), 32, 0, st>>>(mc, a, n_tail);
                        else k_closed_loop_reg<N, M, P, NX, NMPC, true, 8><<<ceil_div(a.B, 64), 32, 0, st>>>(mc, a, n_tail);
                    } else {
                        if (a.w) k_closed_loop_reg<N, M, P, NX, NMPC, false, 4><<<ceil_div(a.B, 32), 32, 0, st>>>(mc, a, n_tail);
                        else k_closed_loop_reg<N, M, P, NX, NMPC, true, 4><<<ceil_div(a.B, 32), 32, 0, st>>>(mc, a, n_tail);
                    }
                    DDMPC_LAUNCH_CHECK();
                    return DDMPC_OK;
                }
            }
            if (want_ws && !want && pair && lpt == 2) {
                MmaCoef<N, M, P, NX, NMPC> mc;
                for (int k = 0; k < NMPC * M; ++k)
                    for (int j = 0; j < NW; ++j) mc.Ku[k][j] = cache[(size_t)k * d.nth + j];
                host_block_map<M, P, NX, NMPC>(plant, NMPC, mc.Mb);
                const int n_tail = a.n_steps % NMPC;
                host_block_map<M, P, NX, NMPC>(plant, n_tail ? n_tail : NMPC, mc.Mt);
                const dim3 gridw(ceil_div(a.B, 64));
                // 7 CTAs (30 KB of shared memory each) per SM put the 1024 CTAs of a 65,536-loop batch in ONE wave
                // on 148 SMs: ask for the largest shared-memory carveout instead of trusting the default
                if (const char *ens = getenv("DDMPC_DEBUG_NOSTORE")) a.dbg_nostore = ens[0] == '1';
                const char *emw = getenv("DDMPC_WS_MATH_WARPS");
                const int mw = (emw && emw[0] == '1') ? 1 : ((emw && emw[0] == '4') ? 4 : 2);
                static std::atomic<unsigned long long> carveout_done{0};
                if (first_time_on_device(carveout_done)) {
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                }
                if (mw == 4) {
                    static std::atomic<unsigned long long> carveout4_done{0};
                    if (first_time_on_device(carveout4_done)) {
                        DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                        DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, true, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    }
                    if (a.w) k_closed_loop_ws<N, M, P, NX, NMPC, false, 4><<<gridw, 160, 0, st>>>(mc, a, n_tail);
                    else k_closed_loop_ws<N, M, P, NX, NMPC, true, 4><<<gridw, 160, 0, st>>>(mc, a, n_tail);
                } else if (mw == 2) {
                    const char *emd = getenv("DDMPC_WS_MATH_DRAWS");
                    if (a.w) k_closed_loop_ws<N, M, P, NX, NMPC, false, 2><<<gridw, 96, 0, st>>>(mc, a, n_tail);
                    else if (emd && emd[0] == '1') {
                        static std::atomic<unsigned long long> carveout_md{0};
                        if (first_time_on_device(carveout_md))
                            DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                        k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, true><<<gridw, 96, 0, st>>>(mc, a, n_tail);
                    }
                    else k_closed_loop_ws<N, M, P, NX, NMPC, true, 2><<<gridw, 96, 0, st>>>(mc, a, n_tail);
                } else {
                    if (a.w) k_closed_loop_ws<N, M, P, NX, NMPC, false, 1><<<gridw, 64, 0, st>>>(mc, a, n_tail);
                    else k_closed_loop_ws<N, M, P, NX, NMPC, true, 1><<<gridw, 64, 0, st>>>(mc, a, n_tail);
                }
                DDMPC_LAUNCH_CHECK();
                return DDMPC_OK;
            }
            if (want && pair && lpt == 2) {
                MmaCoef<N, M, P, NX, NMPC> mc;
                for (int k = 0; k < NMPC * M; ++k)
                    for (int j = 0; j < NW; ++j) mc.Ku[k][j] = cache[(size_t)k * d.nth + j];
                host_block_map<M, P, NX, NMPC>(plant, NMPC, mc.Mb);
                const int n_tail = a.n_steps % NMPC;
                host_block_map<M, P, NX, NMPC>(plant, n_tail ? n_tail : NMPC, mc.Mt);
                const dim3 gridm(ceil_div(a.B, 64));
                if (a.w) k_closed_loop_mma<N, M, P, NX, NMPC, false><<<gridm, 32, 0, st>>>(mc, a, n_tail);
                else k_closed_loop_mma<N, M, P, NX, NMPC, true><<<gridm, 32, 0, st>>>(mc, a, n_tail);
                DDMPC_LAUNCH_CHECK();
                return DDMPC_OK;
            }
        }
    }
    const int tpb = lpt == 1 ? 64 : 32;
    const dim3 grid(ceil_div(a.B, 64));
#define DDMPC_LAUNCH_FAST(LPT_)                                                                              \
    do {                                                                                                     \
        if (a.w) {                                                                                           \
            if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, LPT_, false, true><<<grid, tpb, 0, st>>>(cf, a);   \
            else k_closed_loop_fast<N, M, P, NX, NMPC, LPT_, false, false><<<grid, tpb, 0, st>>>(cf, a);       \
        } else {                                                                                             \
            if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, LPT_, true, true><<<grid, tpb, 0, st>>>(cf, a);    \
            else k_closed_loop_fast<N, M, P, NX, NMPC, LPT_, true, false><<<grid, tpb, 0, st>>>(cf, a);        \
        }                                                                                                    \
    } while (0)
    if (lpt == 2) DDMPC_LAUNCH_FAST(2);
    else DDMPC_LAUNCH_FAST(1);
#undef DDMPC_LAUNCH_FAST
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}


// Returns DDMPC_OK when the fast kernel handled the call, -1 when it does not apply.
int closed_loop_fast_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                         cudaStream_t st) {
    const Dims &d = set->plan.d;
    if (ctrl_idx || set->plan.count != 1 || !d.robust) return -1;
    if (d.nbu > 0 || d.nby > 0) return -1;       // input / output box: per-row bounds, generic kernel
    const char *force = getenv("DDMPC_FORCE_GENERIC");
    if (force && force[0] == '1') return -1;
    FastArgs fa{};
    fa.tol = tol > 0.0 ? tol : 1e-8;
    fa.max_iter = max_iter > 0 ? max_iter : 1000;
    fa.B = B; fa.n_steps = n_steps;
    fa.x0 = x0; fa.u_past0 = u_past0; fa.y_past0 = y_past0; fa.u_s = u_s; fa.y_s = y_s; fa.w = w;
    fa.seed = seed; fa.id0 = id0; fa.eps = eps;
    fa.u_sys = u_sys; fa.y_sys = y_sys; fa.x_final = x_final; fa.status = status; fa.iters = iters;
    const int nmpc = set->prm.n_mpc_step;
#define DDMPC_FAST_CASE(N_, M_, P_, NX_, NMPC_)                                                    \
    if (d.n == N_ && d.m == M_ && d.p == P_ && plant->n_x == NX_ && nmpc == NMPC_)                 \
        return launch_fast<N_, M_, P_, NX_, NMPC_>(set, plant, fa, st);
    DDMPC_FAST_CASE(4, 2, 2, 4, 4)
    DDMPC_FAST_CASE(4, 2, 2, 4, 1)
    DDMPC_FAST_CASE(4, 2, 2, 4, 2)
    DDMPC_FAST_CASE(2, 1, 1, 2, 1)
    DDMPC_FAST_CASE(2, 1, 1, 2, 2)
#undef DDMPC_FAST_CASE
    return -1;
}

}  // namespace ddmpc
