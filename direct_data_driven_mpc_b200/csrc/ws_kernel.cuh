// Warp-specialised all-tensor-core closed-loop kernel (the bench kernel of BASELINE config 3).  A header because the
// product (fast_loop.cu) instantiates it with two math warps only, and experiments/closed_loop_variants.cu instantiates
// the measured-and-dropped shapes (one or four math warps, math warps drawing their own noise).
#pragma once
#include "fast_common.cuh"

#ifndef WS_PHILOX_ROUNDS
#define WS_PHILOX_ROUNDS 10   // (experiments build this header with fewer rounds to measure what the drawing costs)
#endif

namespace ddmpc {

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
// NOSTORE (experiments only): the kernel without its trajectory stores, a measurement aid.
// IOW = 2 (experiments): the i/o work on TWO warps, one drawing, one recording (the review's lever: CTAs of 2 + 2 warps,
// roles swapped on bit 2 of the hardware warp slot so that the math warps of successive CTAs alternate between the
// scheduler pairs).
// LPT_ = 1 (experiments, with MW = 1): CTAs of 32 loops - one math warp with four n-tiles, one i/o warp whose lanes own ONE
// loop each - 14 CTAs per SM instead of 7, the i/o work of a block on twice as many warps.
template <int N, int M, int P, int NX, int NMPC, bool PHILOX, int MW, int MD = 0, bool NOSTORE = false, int IOW = 1, int LPT_ = 2>
__global__ void __launch_bounds__(32 * (MW + IOW), LPT_ == 1 ? 14 : 7)
k_closed_loop_ws(const __grid_constant__ MmaCoef<N, M, P, NX, NMPC> cfp, const FastArgs a, const int n_tail) {
    constexpr int R = NMPC * M, NW = N * (M + P), LPT = LPT_, TP = 36, NT = 4;
    static_assert(LPT_ == 2 || (LPT_ == 1 && MW == 1 && MD == 0 && IOW == 1), "one loop per i/o lane: 1 + 1 warps");
    constexpr int KB = NX + R, RB = NMPC * P + NX, RY = NMPC * P;
    constexpr int LM = MW == 1 ? LPT : 1;                      // loop groups (of 32 loops) per math warp
    constexpr int WPG = MW == 4 ? 2 : 1;                       // math warps per loop group (they split its n-tiles)
    constexpr int NTW = NT / WPG;                              // n-tiles (of 8 loops) per math warp and loop group
    static_assert(MW == 1 || MW == 2 || MW == 4, "one, two or four math warps");
    // MD = 1 (experiments): every math warp draws the Philox noise of its own loops and the i/o warp only records.
    // Measured slower (0.263 vs 0.240 ms): all of the drawing is too much for the math warps.
    // MD = 2: the drawing is SHARED - a math lane draws the first Philox call of its loop and block, the i/o lane the
    // second one of its two loops (source-level stall attribution showed the math warps waiting at the block barrier for
    // the i/o warp a third of the time).
    constexpr bool MATH_DRAWS = MD != 0 && PHILOX;
    constexpr bool HALF = MD == 2 && PHILOX;
    static_assert(MD == 0 || MW >= 2, "math warps draw their own noise only with one loop group per warp");
    static_assert(MD != 2 || (MW == 2 && (NMPC * P) / 4 == 2), "shared drawing: two Philox calls per loop and block");
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
        swap_s = (MW == 1 || IOW == 2) ? (slot >> 2) & 1 : 0;
    }
    __syncthreads();
    static_assert(IOW == 1 || (IOW == 2 && MW == 2 && MD == 0), "two i/o warps: with two math warps, i/o warps draw");
    const int warp = IOW == 2 ? (threadIdx.x >> 5) ^ (swap_s << 1) : (threadIdx.x >> 5) ^ swap_s;
    const int l0 = MW == 1 ? 0 : warp / WPG;       // first loop group of this math warp
    const int t80 = (warp % WPG) * NTW;            // its first n-tile inside the group
    const int cb = g ^ (((q >> 1) & 1) << 2);      // swizzled column of a B-fragment element (row = 4ks + q)
    const int cc2 = (2 * q) ^ (((g >> 1) & 1) << 2);  // swizzled column of a C-fragment pair (row = g)

    if (warp >= MW) {
        // ------------------------------------------------------------------ i/o warp: thread tl owns loops 2tl, 2tl+1
        const bool do_draw = IOW == 1 || warp == MW, do_rec = IOW == 1 || warp == MW + 1;
        int b[LPT];
        bool live[LPT];
        size_t f0[LPT];
        uint32_t sid_lo[LPT], sid_hi[LPT];
        double pu[LPT][M], py[LPT][P];           // previous trajectory element (sector pairing)
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            b[l] = blockIdx.x * (32 * LPT) + LPT * tl + l;
            live[l] = b[l] < a.B;
            if (!live[l]) b[l] = 0;              // dead slots replay loop 0 and never store
            f0[l] = (size_t)b[l] * a.n_steps;
            const unsigned long long sid = a.id0 + (unsigned long long)b[l];
            sid_lo[l] = (uint32_t)sid;
            sid_hi[l] = (uint32_t)(sid >> 32);
            pu[l][0] = pu[l][1] = py[l][0] = py[l][1] = 0.0;
            if (!do_draw) continue;                  // (the drawing warp sets the CTA's state up)
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
                    for (int cc = HALF ? 1 : 0; cc < RY / 4; ++cc) {
                        uint32_t c0 = (uint32_t)(((unsigned)tb * (unsigned)RY) >> 2) + (uint32_t)cc, c1 = 0u,
                                 c2 = sid_lo[l], c3 = sid_hi[l];
#pragma unroll
                        for (int r = 0; r < WS_PHILOX_ROUNDS; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
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
            if (a.step_major && LPT == 2) {
                // (n_steps, B, m): the two loops of a lane are neighbours, a warp writes 1024 contiguous bytes per step
                const size_t e0 = (size_t)blockIdx.x * 64 + 2 * tl;
                const bool l0 = e0 < (size_t)a.B, l1 = e0 + 1 < (size_t)a.B;
#pragma unroll
                for (int s = 0; s < NMPC; ++s) {
                    if (s < steps && l0 && !NOSTORE) {
                        const size_t e = (size_t)(tb * NMPC + s) * a.B + e0;
                        double u[2][M], y[2][P];
#pragma unroll
                        for (int l = 0; l < LPT; ++l) {
#pragma unroll
                            for (int i = 0; i < M; ++i) u[l][i] = up_s[ub][s * M + i][l][SW(s * M + i, tl)];
#pragma unroll
                            for (int i = 0; i < P; ++i) y[l][i] = wy_s[yb][s * P + i][l][SW(s * P + i, tl)];
                        }
                        if (l1 && (e & 1) == 0) {
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.u_sys + e * 2), "d"(u[0][0]),
                                         "d"(u[0][1]), "d"(u[1][0]), "d"(u[1][1])
                                         : "memory");
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.y_sys + e * 2), "d"(y[0][0]),
                                         "d"(y[0][1]), "d"(y[1][0]), "d"(y[1][1])
                                         : "memory");
                        } else {
                            *reinterpret_cast<double2 *>(a.u_sys + e * 2) = make_double2(u[0][0], u[0][1]);
                            *reinterpret_cast<double2 *>(a.y_sys + e * 2) = make_double2(y[0][0], y[0][1]);
                            if (l1) {
                                *reinterpret_cast<double2 *>(a.u_sys + (e + 1) * 2) = make_double2(u[1][0], u[1][1]);
                                *reinterpret_cast<double2 *>(a.y_sys + (e + 1) * 2) = make_double2(y[1][0], y[1][1]);
                            }
                        }
                    }
                }
                // (the verdict below reads the last outputs from (py): keep it current)
#pragma unroll
                for (int l = 0; l < LPT; ++l) {
                    const int sl = steps - 1;
#pragma unroll
                    for (int i = 0; i < P; ++i) py[l][i] = wy_s[yb][sl * P + i][l][SW(sl * P + i, tl)];
                }
                return;
            }
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
                        if (live[l] && (f & 1) && !NOSTORE) {   // warp-uniform: completes the sector (f-1, f)
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
        if ((!MATH_DRAWS || HALF) && do_draw) draw(0, 0);
        __syncthreads();                                   // window, state and noise of block 0 are in place
        for (int t = 0; t < nblk; ++t) {
            if (t > 0 && do_rec) record(t - 1, NMPC);
            if ((!MATH_DRAWS || HALF) && do_draw && t + 1 < nblk) draw(t + 1, (t + 1) % 3);
            __syncthreads();                               // block t is complete
        }
        if (!do_rec) return;
        record(nblk - 1, n_tail ? n_tail : NMPC);
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            if (!live[l]) continue;
            const size_t fl = f0[l] + a.n_steps - 1;
            if ((fl & 1) == 0 && !a.step_major) {          // an unpaired final element is still in (pu, py)
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
    constexpr int CPL = HALF ? 1 : RY / 4 / WPG;           // Philox calls per lane and block
    const int ncol = WPG == 2 ? 8 * t80 + (tl >> 1) : tl, ncc0 = WPG == 2 ? (tl & 1) : 0;
    const unsigned long long nsid = a.id0 + (unsigned long long)min(blockIdx.x * 64 + 2 * ncol + l0, a.B - 1);
    auto mdraw = [&](const int tb, const int buf) {
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) {
            const int ncc = ncc0 + ci;
            uint32_t c0 = (uint32_t)(((unsigned)tb * (unsigned)RY) >> 2) + (uint32_t)ncc, c1 = 0u, c2 = (uint32_t)nsid,
                     c3 = (uint32_t)(nsid >> 32);
#pragma unroll
            for (int r = 0; r < WS_PHILOX_ROUNDS; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
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

}  // namespace ddmpc
