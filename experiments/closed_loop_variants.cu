// Closed-loop kernel designs that were built, measured on B200 and LOST to k_closed_loop_ws / k_closed_loop_fast
// (profiles/r1_probes.txt items 6, 7).  They are kept as runnable experiments - bit-compatible with the
// product kernels to 1.8e-15 - but are NOT part of libddmpc.so: scripts/time_variants.py builds this file into
// experiments/libddmpc_experiments.so and times the variants through ddmpc_exp_closed_loop().
#include <cstdlib>
#include <string>

#include "../direct_data_driven_mpc_b200/csrc/fast_common.cuh"
#include "../direct_data_driven_mpc_b200/csrc/ws_kernel.cuh"

namespace ddmpc {

__constant__ int exp_nostore = 0;   // measurement aid: 1 = skip the trajectory stores (compute only)

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
                    if (olive && (f & 1) && !exp_nostore) {      // completes the sector (f - 1, f)
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
            if (!XP && live[mt] && q < steps && !exp_nostore) {
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
        if (olive && (fl & 1) == 0 && !exp_nostore) {          // an unpaired final element is still in (pu, py)
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
                        if (live[l] && (f & 1) && !exp_nostore) {   // warp-uniform: completes the sector (f-1, f)
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
            if (live[l] && (fl & 1) == 0 && !exp_nostore) {   // an unpaired final element is still in (pu, py)
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


}  // namespace ddmpc

using namespace ddmpc;

// Runs one variant on the four-tank n-step shape (n = 4, m = p = 2, n_x = 4, n_mpc_step = 4), shared ROBUST / NONE
// controller, device pointers.  variant: "mma", "reg" (NT = 4), "reg_nt2", "reg_nt8", "regx", "rws", "ws1", "ws4",
// "ws2md", "ws2" (the product kernel, here for the no-store timing).  nostore = 1 skips the trajectory stores where
// the variant supports it.
extern "C" int ddmpc_exp_closed_loop(const ddmpc_set *set, const ddmpc_plant *plant, const char *variant, int nostore,
                                     int B, const double *x0, const double *u_past0, const double *y_past0,
                                     const double *u_s, const double *y_s, const double *w, uint64_t seed, uint64_t id0,
                                     double eps, int n_steps, double *u_sys, double *y_sys, int32_t *status,
                                     int32_t *iters, double *x_final, void *stream) {
    constexpr int N = 4, M = 2, P = 2, NX = 4, NMPC = 4, NW = N * (M + P);
    if (!set || !plant || !variant) return fail(DDMPC_ERR_INVALID_ARG, "exp: null argument");
    const Dims &d = set->plan.d;
    if (d.n != N || d.m != M || d.p != P || plant->n_x != NX || set->prm.n_mpc_step != NMPC || set->plan.count != 1 ||
        !d.robust || d.nb > 0 || set->fast_host.size() != (size_t)NMPC * M * d.nth)
        return fail(DDMPC_ERR_INVALID_ARG, "exp: four-tank n-step shape with a shared ROBUST / NONE controller only");
    cudaStream_t st = (cudaStream_t)stream;
    FastArgs a{};
    a.B = B; a.n_steps = n_steps;
    a.Ksp = set->fast_ksp.d();
    a.x0 = x0; a.u_past0 = u_past0; a.y_past0 = y_past0; a.u_s = u_s; a.y_s = y_s; a.w = w;
    a.seed = seed; a.id0 = id0; a.eps = eps;
    a.u_sys = u_sys; a.y_sys = y_sys; a.x_final = x_final; a.status = status; a.iters = iters;
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    MmaCoef<N, M, P, NX, NMPC> mc;
    for (int k = 0; k < NMPC * M; ++k)
        for (int j = 0; j < NW; ++j) mc.Ku[k][j] = set->fast_host[(size_t)k * d.nth + j];
    host_block_map<M, P, NX, NMPC>(plant, NMPC, mc.Mb);
    const int n_tail = n_steps % NMPC;
    host_block_map<M, P, NX, NMPC>(plant, n_tail ? n_tail : NMPC, mc.Mt);
    static const int flags[2] = {0, 1};                      // static storage: the copy may be captured in a CUDA graph
    DDMPC_CUDA(cudaMemcpyToSymbolAsync(exp_nostore, &flags[nostore ? 1 : 0], sizeof(int), 0, cudaMemcpyHostToDevice, st));
    const std::string v(variant);
    const int g64 = ceil_div(B, 64);
    auto carve = [](auto kern) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
#define EXP_RUN(KERN_W, KERN_P, GRID, BLOCK)                  \
    do {                                                      \
        if (w) { DDMPC_CUDA(carve(KERN_W)); KERN_W<<<GRID, BLOCK, 0, st>>>(mc, a, n_tail); } \
        else { DDMPC_CUDA(carve(KERN_P)); KERN_P<<<GRID, BLOCK, 0, st>>>(mc, a, n_tail); }   \
    } while (0)
    if (v == "mma") EXP_RUN((k_closed_loop_mma<N, M, P, NX, NMPC, false>), (k_closed_loop_mma<N, M, P, NX, NMPC, true>), g64, 32);
    else if (v == "reg") EXP_RUN((k_closed_loop_reg<N, M, P, NX, NMPC, false, 4>), (k_closed_loop_reg<N, M, P, NX, NMPC, true, 4>), ceil_div(B, 32), 32);
    else if (v == "reg_nt2") EXP_RUN((k_closed_loop_reg<N, M, P, NX, NMPC, false, 2>), (k_closed_loop_reg<N, M, P, NX, NMPC, true, 2>), ceil_div(B, 16), 32);
    else if (v == "reg_nt8") EXP_RUN((k_closed_loop_reg<N, M, P, NX, NMPC, false, 8>), (k_closed_loop_reg<N, M, P, NX, NMPC, true, 8>), g64, 32);
    else if (v == "regx") EXP_RUN((k_closed_loop_reg<N, M, P, NX, NMPC, false, 4, true>), (k_closed_loop_reg<N, M, P, NX, NMPC, true, 4, true>), ceil_div(B, 32), 32);
    else if (v == "rws") EXP_RUN((k_closed_loop_rws<N, M, P, NX, NMPC, false>), (k_closed_loop_rws<N, M, P, NX, NMPC, true>), g64, 96);
    else if (v == "ws1") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 1>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 1>), g64, 64);
    else if (v == "ws4") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 4>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 4>), g64, 160);
    else if (v == "ws2md") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 1>), g64, 96);
    else if (v == "ws2mh" && nostore) EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2, 0, true>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 2, true>), g64, 96);
    else if (v == "ws2mh") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 2>), g64, 96);
    else if (v == "ws1l1" && nostore) EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 1, 0, true, 1, 1>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 1, 0, true, 1, 1>), ceil_div(B, 32), 64);
    else if (v == "ws1l1") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 1, 0, false, 1, 1>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 1, 0, false, 1, 1>), ceil_div(B, 32), 64);
    else if (v == "ws2io2" && nostore) EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2, 0, true, 2>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 0, true, 2>), g64, 128);
    else if (v == "ws2io2") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2, 0, false, 2>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 0, false, 2>), g64, 128);
    else if (v == "ws2" && nostore) EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2, 0, true>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2, 0, true>), g64, 96);
    else if (v == "ws2") EXP_RUN((k_closed_loop_ws<N, M, P, NX, NMPC, false, 2>), (k_closed_loop_ws<N, M, P, NX, NMPC, true, 2>), g64, 96);
    else return fail(DDMPC_ERR_INVALID_ARG, "exp: unknown variant '%s'", variant);
#undef EXP_RUN
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}
