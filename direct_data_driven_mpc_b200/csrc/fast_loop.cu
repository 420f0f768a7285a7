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
#include "fast_common.cuh"
#include "ws_kernel.cuh"

namespace ddmpc {

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

// gather the Ksp block of Ku (rows 0..NMPC*M-1, columns n*(m+p)..nth-1) into a dense device array
__global__ void k_gather_ksp(const double *__restrict__ Ku, int nth, int nw, int rows, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsp = nth - nw;
    if (e < rows * nsp) out[e] = Ku[(size_t)(e / nsp) * nth + nw + (e % nsp)];
}

// Called once at the end of ddmpc_set_create (which synchronises): host copy of the applied gain rows and device
// copy of their set-point block.  Building them lazily on the first caller's stream raced with a second caller on
// another non-blocking stream (chunked host API), so nothing here is created at launch time any more.
int closed_loop_fast_prepare(ddmpc_set *set, cudaStream_t st) {
    const Dims &d = set->plan.d;
    const int rows = set->prm.n_mpc_step * d.m;
    if (set->plan.count != 1 || !d.robust || rows < 1 || rows > d.Lm || (size_t)rows * d.nth > 1024) return DDMPC_OK;
    set->fast_host.resize((size_t)rows * d.nth);
    DDMPC_CUDA(cudaMemcpyAsync(set->fast_host.data(), set->plan.Ku.d(), sizeof(double) * set->fast_host.size(),
                               cudaMemcpyDeviceToHost, st));
    DDMPC_CUDA(set->fast_ksp.alloc(sizeof(double) * rows * (d.m + d.p)));
    k_gather_ksp<<<ceil_div(rows * (d.m + d.p), 128), 128, 0, st>>>(set->plan.Ku.d(), d.nth, d.n * (d.m + d.p), rows,
                                                                      set->fast_ksp.d());
    DDMPC_LAUNCH_CHECK();
    DDMPC_CUDA(cudaStreamSynchronize(st));
    return DDMPC_OK;
}

template <int N, int M, int P, int NX, int NMPC>
static int launch_fast(const ddmpc_set *set, const ddmpc_plant *plant, const FastArgs &fa, cudaStream_t st) {
    using Coef = FastCoef<N, M, P, NX, NMPC>;
    static_assert(sizeof(Coef) <= 3584, "coefficients must fit in the kernel parameter space");
    const Dims &d = set->plan.d;
    constexpr int NW = N * (M + P);
    const auto &cache = set->fast_host;                      // gain rows, copied at set creation
    if (cache.size() != (size_t)NMPC * M * d.nth || !set->fast_ksp.p) return -1;
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
    const int lpt = set->opt_lpt == 1 || set->opt_lpt == 2 ? set->opt_lpt : (a.B >= 16384 ? 2 : 1);
    if (d.convex && set->opt_layout == 1) return -1;          // step-major trajectories: k_closed_loop_ws only
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
    {   // warp-specialised all-tensor-core kernel (solve AND plant on the FP64 MMA pipe, two math warps + one i/o warp
        // per 64 loops) for the four-tank n-step shape: the default from 16,384 loops on
        constexpr bool MMA_OK = (M == 2 && P == 2 && NMPC * M == 8 && NMPC == N && (N * M) % 4 == 0 &&
                                 (NX + NMPC * M) % 4 == 0 && NMPC * P + NX <= 16 && NMPC * P == 8);
        if constexpr (MMA_OK) {
            // (its serial chain is 0.071 ms for any batch up to 8192 loops: the fastest kernel from ~6000 loops on)
            const bool want_ws = set->opt_path == DDMPC_PATH_WS || set->opt_path == DDMPC_PATH_AUTO;
            a.step_major = set->opt_layout;
            if (want_ws && pair) {
                MmaCoef<N, M, P, NX, NMPC> mc;
                for (int k = 0; k < NMPC * M; ++k)
                    for (int j = 0; j < NW; ++j) mc.Ku[k][j] = cache[(size_t)k * d.nth + j];
                host_block_map<M, P, NX, NMPC>(plant, NMPC, mc.Mb);
                const int n_tail = a.n_steps % NMPC;
                host_block_map<M, P, NX, NMPC>(plant, n_tail ? n_tail : NMPC, mc.Mt);
                const dim3 gridw(ceil_div(a.B, 64));
                // 7 CTAs (30 KB of shared memory each) per SM put the 1024 CTAs of a 65,536-loop batch in ONE wave
                // on 148 SMs: ask for the largest shared-memory carveout instead of trusting the default
                static std::atomic<unsigned long long> carveout_done{0};
                if (first_time_on_device(carveout_done)) {
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop_ws<N, M, P, NX, NMPC, true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                }
                if (a.w) k_closed_loop_ws<N, M, P, NX, NMPC, false, 2><<<gridw, 96, 0, st>>>(mc, a, n_tail);
                else k_closed_loop_ws<N, M, P, NX, NMPC, true, 2><<<gridw, 96, 0, st>>>(mc, a, n_tail);
                DDMPC_LAUNCH_CHECK();
                return DDMPC_OK;
            }
        }
        if (set->opt_path == DDMPC_PATH_WS || set->opt_layout == 1) return -1;   // asked for explicitly but not available for this shape
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
    const int path = set->opt_path;
    if (path != DDMPC_PATH_AUTO && path != DDMPC_PATH_FAST && path != DDMPC_PATH_WS) return -1;
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
