// Fused closed loop for PER-LOOP controllers (BASELINE config 2: every closed loop has its own (u_d, y_d), hence
// its own gains; config 5: a grid of controllers x a few loops each) and for small batches of a shared controller.
//
// When every loop has its own gain block there is no shared matrix, hence no GEMM: a solve is a private
// (n_mpc*m x n_theta) matrix-vector product.  A batch of 4096 such loops offers a B200 far fewer independent chains
// than it has FP64 lanes, so the design goal is LATENCY of the serial chain solve -> plant -> window, not throughput:
//
//   * EIGHT LANES PER LOOP (4 loops per warp).  Lane k of a loop owns planned-input row k of the gain (its window
//     coefficients live in REGISTERS for the whole run; the set-point part is folded into one constant per row) and
//     rows k, k + 8 of the plant's n_mpc-step block map  [Y; x+] = Mblk [x; U]  (model_simulation.py:93-98 unrolled;
//     measurement noise enters y only).  A solve is therefore ONE 16-term dot product deep instead of 8 of them, and
//     the n_mpc plant steps of a block are ONE 12-term dot product deep instead of n_mpc * 2 matrix-vector products.
//   * the loop state (window, plant state) is replicated in the registers of the 8 lanes; what a lane produces (one
//     planned input, one or two outputs / next-state entries) reaches the other seven through 160 bytes of shared
//     memory per loop: two __syncwarp per MPC iteration, no block-level barrier, no atomics.
//   * measurement noise: Philox4x32-10 in-kernel (lane = one output of the block) or the caller's array (parity mode);
//   * trajectories in the reference layout (B, n_steps, m|p): the 8 lanes of a loop write 64 contiguous bytes per
//     block and array (the thread-per-loop kernel wrote 8-byte pieces 3.2 KB apart: 1.49x DRAM traffic).
//
// Replaces, per loop, the same reference code as k_closed_loop (solve.cu): update_and_solve_data_driven_mpc
// (direct_data_driven_mpc_controller.py:389-407), get_optimal_control_input_at_step (:810-842),
// store_input_output_measurement (:844-895), LTIModel.simulate_step (utilities/model_simulation.py:93-98) and the loop
// of simulate_data_driven_mpc_control_loop (utilities/controller/controller_operation.py:263-305), for equality-only
// controllers (ROBUST / slack NONE, NOMINAL) of the compiled shapes.
#include <vector>

#include "common.cuh"
#include "plan.cuh"

namespace ddmpc {

std::vector<double> block_map(const ddmpc_plant *pl, int s);   // gemm_loop.cu

template <int RB, int KB>
struct PlMaps {
    double Mb[RB][KB];   // block map of a full n_mpc-step block: rows y_0..y_{n_mpc-1}, x_{n_mpc}; columns x_0, u_0..
    double Mt[RB][KB];   // block map of the last, partial block (controller_operation.py:278), same layout, zero padded
};

struct PerLoopArgs {
    int B, n_steps, nth, Lm, nfix;
    const int *ctrl_idx;               // (B) controller of each loop, NULL = controller 0
    const double *Ku;                  // (count, Lm, nth)
    const double *F;                   // (count, nfix, nth) nominal feasibility map, NULL for ROBUST
    const int *Fnz;                    // (count) 1 when F of that controller is not identically zero
    const double *x0, *u_past0, *y_past0, *u_s, *y_s, *w;
    unsigned long long id0;
    double eps;
    double *u_sys, *y_sys, *x_final;
    int *status, *iters;
    uint32_t rk[20];                   // Philox round keys (key + r * Weyl), filled on the host
    // box rows (CONVEX slack bound), per controller: Ks (nb, nth), Phi (nb, nb), Psi (Lm, nb), scaled bounds, tolerance scale
    int nb, max_iter;
    double tol;
    const double *Ks, *Phi, *Psi, *blo, *bhi, *bmax;
};

__device__ __forceinline__ void pl_philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                                uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}

// BOX: the controllers carry box rows (CONVEX slack bound): after the gain product the 8 lanes of a loop compute its
// slack rows Ks theta (rows k, k + 8, ..) and, if one leaves the box, run the box-row ADMM of DESIGN.md 1.1 cooperatively:
// every operator is PRIVATE to the loop (no shared matrix, no GEMM), streamed from L2 with 16-byte loads.
template <int N, int M, int P, int NX, int NMPC, bool PHILOX, bool BOX = false>
__global__ void __launch_bounds__(128)
k_closed_loop_perloop(const __grid_constant__ PlMaps<NMPC * P + NX, NX + NMPC * M> maps, const PerLoopArgs a) {
    constexpr int G = 8;                                   // lanes per loop
    constexpr int R = NMPC * M, RY = NMPC * P, RB = RY + NX, KB = NX + R;
    constexpr int NWU = N * M, NWY = N * P, NW = NWU + NWY, NSP = M + P;
    constexpr int RU = (R + G - 1) / G, RP = (RB + G - 1) / G;   // solve rows / plant rows per lane
    constexpr int EX = R + RB;                             // exchange buffer of a loop: [U (R); Y (RY); x+ (NX)]
    constexpr int S = ((EX + 11) / 16) * 16 + 4;           // stride = 4 (mod 16) doubles: the 4 loops of a warp read
                                                           // their (broadcast) 16-byte pieces from disjoint banks
    __shared__ __align__(16) double ex_s[16][S];
    __shared__ __align__(16) double box_s[BOX ? 16 : 1][BOX ? 3 : 1][BOX ? 64 : 2];   // per loop: ADMM iterate d (then t), z, w
    __shared__ double lv0_s[BOX ? 16 : 1][BOX ? NW + 1 : 1];   // per loop: level-0 screen, cmax (NW) and max |s_ss|
    const int lane = threadIdx.x & 31, k = lane & (G - 1);
    const int slot = threadIdx.x >> 3;                     // loop slot in the CTA
    double *ex = ex_s[slot];
    int b = blockIdx.x * (blockDim.x >> 3) + slot;
    const bool live = b < a.B;
    if (!live) b = a.B - 1;                                // dead slots replay the last loop and never store
    const int c = a.ctrl_idx ? a.ctrl_idx[b] : 0;
    const size_t f0 = (size_t)b * a.n_steps;
    const unsigned long long sid = a.id0 + (unsigned long long)b;
    const uint32_t sid_lo = (uint32_t)sid, sid_hi = (uint32_t)(sid >> 32);

    // ---- per-loop constants: gain rows on the window (registers), set-point term, plant block-map rows
    double sp[NSP];
#pragma unroll
    for (int i = 0; i < M; ++i) sp[i] = a.u_s[(size_t)b * M + i];
#pragma unroll
    for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[(size_t)b * P + i];
    double Kw[RU][NW], csp[RU];
#pragma unroll
    for (int i = 0; i < RU; ++i) {
        const int r = k + G * i;
        const double *row = a.Ku + ((size_t)c * a.Lm + (r < R ? r : 0)) * a.nth;
        const double on = r < R ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < NW; ++j) Kw[i][j] = on * __ldg(row + j);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NSP; ++j) acc = fma(__ldg(row + NW + j), sp[j], acc);
        csp[i] = on * acc;
    }
    double Mc[RP][KB];
    auto load_map = [&](const double (&Mm)[RB][KB]) {
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            const int r = k + G * i;
#pragma unroll
            for (int j = 0; j < KB; ++j) Mc[i][j] = r < RB ? Mm[r < RB ? r : 0][j] : 0.0;
        }
    };
    load_map(maps.Mb);

    // ---- loop state, replicated in the 8 lanes: measurement window (oldest first) and plant state
    double win[NW], x[NX];
#pragma unroll
    for (int j = 0; j < NWU; ++j) win[j] = a.u_past0[(size_t)b * NWU + j];
#pragma unroll
    for (int j = 0; j < NWY; ++j) win[NWU + j] = a.y_past0[(size_t)b * NWY + j];
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = a.x0[(size_t)b * NX + j];

    // ---- BOX, level-0 screen (cvx_loop.cu has the same test for a shared controller): with theta_ss the window of the
    //      settled loop (its set-point repeated),  |s_j(theta)| <= max_j |s_ss,j| + sum_k cmax_k |theta_k - theta_ss,k|,
    //      cmax_k = max_j |Ks[j][k]|.  Both constants cost one pass over the loop's own Ks (what the exact check reads in
    //      EVERY block: 9.6 KB per four-tank loop); the test itself is NW multiply-adds.
    double thr0 = 0.0;
    if constexpr (BOX) {
        const int nb = a.nb;
        const double *Ksc = a.Ks + (size_t)c * nb * a.nth;
        double cm[NW], sss = 0.0;
#pragma unroll
        for (int j = 0; j < NW; ++j) cm[j] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = k + G * i;
            if (r < nb) {
                const double *row = Ksc + (size_t)r * a.nth;
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    const double kv = __ldg(row + j);
                    cm[j] = fmax(cm[j], fabs(kv));
                    acc = fma(kv, j < NWU ? sp[j % M] : sp[M + (j - NWU) % P], acc);
                }
#pragma unroll
                for (int j = 0; j < NSP; ++j) acc = fma(__ldg(row + NW + j), sp[j], acc);
                sss = fmax(sss, fabs(acc));
            }
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
#pragma unroll
            for (int j = 0; j < NW; ++j) cm[j] = fmax(cm[j], __shfl_xor_sync(0xffffffffu, cm[j], o));
            sss = fmax(sss, __shfl_xor_sync(0xffffffffu, sss, o));
        }
        if (k == 0) {
#pragma unroll
            for (int j = 0; j < NW; ++j) lv0_s[slot][j] = cm[j];
            lv0_s[slot][NW] = sss;
        }
        __syncwarp();
        thr0 = __ldg(a.bhi + (size_t)c * nb) * (1.0 - 1e-9);
    }

    const double *Fc = nullptr;
    int fnz = 0;
    if (a.F) {
        fnz = a.Fnz[c];
        Fc = a.F + (size_t)c * a.nfix * a.nth;
    }
    const bool any_f = a.F && __any_sync(0xffffffffu, fnz != 0);

    int status = DDMPC_SOLVE_OPTIMAL, extra = 0;
    const int nblk = (a.n_steps + NMPC - 1) / NMPC;
    // uploaded noise (parity mode) of output row k + G i of the block starting at step t0
    auto fetch_row = [&](int t0, int i) -> double {
        const int r = k + G * i;
        return (G * i < RY && r < RY && t0 + r / P < a.n_steps) ? __ldg(a.w + (f0 + t0) * P + r) : 0.0;
    };
    // noise of that row: the prefetched value, or word (q & 3) of Philox call (q >> 2) for noise word q = step * P + output
    auto noise_row = [&](int t0, int i, double fetched) -> double {
        if constexpr (PHILOX) {
            const int r = k + G * i;
            if (G * i >= RY) return 0.0;
            const unsigned q = (unsigned)t0 * (unsigned)P + (unsigned)(r < RY ? r : 0);
            uint32_t c0 = q >> 2, c1 = 0u, c2 = sid_lo, c3 = sid_hi;
#pragma unroll
            for (int rr = 0; rr < 10; ++rr) pl_philox_round(c0, c1, c2, c3, a.rk[2 * rr], a.rk[2 * rr + 1]);
            const unsigned w4 = q & 3u;
            const uint32_t word = w4 == 0 ? c0 : (w4 == 1 ? c1 : (w4 == 2 ? c2 : c3));
            const double v = a.eps * (2.0 * __hiloint2double((int)(0x3FF00000u | (word >> 12)), (int)(word << 20)) - 3.0);
            return r < RY ? v : 0.0;
        } else {
            return fetched;
        }
    };
    double nz_next[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) nz_next[i] = PHILOX ? 0.0 : fetch_row(0, i);
    double Y[RY];
#pragma unroll
    for (int j = 0; j < RY; ++j) Y[j] = 0.0;
    for (int blk = 0, t0 = 0; blk < nblk; ++blk, t0 += NMPC) {
        const int steps = min(NMPC, a.n_steps - t0);
        if (steps < NMPC) load_map(maps.Mt);
        // ---- measurement noise of this lane's output rows: Philox words are computed here (independent of the solve chain);
        //      uploaded noise was fetched one block ahead, so its global-memory latency is off the critical path
        double nz[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) nz[i] = noise_row(t0, i, nz_next[i]);
        if constexpr (!PHILOX) {
            if (blk + 1 < nblk) {
#pragma unroll
                for (int i = 0; i < RP; ++i) nz_next[i] = fetch_row(t0 + NMPC, i);
            }
        }
        // ---- NOMINAL with rank-deficient data: consistency of the window with range(H) (solve.cu, k_closed_loop)
        if (any_f) {
            double fe = 0.0, thmax = 0.0;
#pragma unroll
            for (int j = 0; j < NW; ++j) thmax = fmax(thmax, fabs(win[j]));
#pragma unroll
            for (int j = 0; j < NSP; ++j) thmax = fmax(thmax, fabs(sp[j]));
            if (fnz) {
                for (int i = k; i < a.nfix; i += G) {
                    const double *row = Fc + (size_t)i * a.nth;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NW; ++j) acc = fma(__ldg(row + j), win[j], acc);
#pragma unroll
                    for (int j = 0; j < NSP; ++j) acc = fma(__ldg(row + NW + j), sp[j], acc);
                    fe = fmax(fe, fabs(acc));
                }
            }
#pragma unroll
            for (int o = 1; o < G; o <<= 1) fe = fmax(fe, __shfl_xor_sync(0xffffffffu, fe, o));
            if (fnz && fe > 1e-6 * (1.0 + thmax)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
        }
        // ---- solve (equality-only => affine in the window): planned input r = csp_r + Kw_r . window
        double u_own[RU];
#pragma unroll
        for (int i = 0; i < RU; ++i) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < NW; ++j) acc[j & 3] = fma(Kw[i][j], win[j], acc[j & 3]);
            u_own[i] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + csp[i];
        }
        bool check_box = BOX;
        if constexpr (BOX) {
            double b0 = lv0_s[slot][NW];
#pragma unroll
            for (int j = 0; j < NW; ++j)
                b0 = fma(lv0_s[slot][j], fabs(win[j] - (j < NWU ? sp[j % M] : sp[M + (j - NWU) % P])), b0);
            check_box = !__all_sync(0xffffffffu, b0 <= thr0);        // (NaN fails the test) warp-uniform
        }
        if (check_box) {
            // ---- slack rows of this loop, rows k + 8 i: s_unc = Ks theta
            const int nb = a.nb;
            const double *Ksc = a.Ks + (size_t)c * nb * a.nth, *loc = a.blo + (size_t)c * nb, *hic = a.bhi + (size_t)c * nb;
            double su[8];
            bool viol = false, bad = false;
            double smax = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = k + G * i;
                su[i] = 0.0;
                if (r < nb) {
                    const double2 *row = reinterpret_cast<const double2 *>(Ksc + (size_t)r * a.nth);
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int j = 0; j < NW / 2; ++j) {
                        const double2 kv = __ldg(row + j);
                        a0 = fma(kv.x, win[2 * j], a0);
                        a1 = fma(kv.y, win[2 * j + 1], a1);
                    }
#pragma unroll
                    for (int j = 0; j < NSP / 2; ++j) {
                        const double2 kv = __ldg(row + NW / 2 + j);
                        a0 = fma(kv.x, sp[2 * j], a0);
                        a1 = fma(kv.y, sp[2 * j + 1], a1);
                    }
                    const double sv = a0 + a1;
                    su[i] = sv;
                    viol = viol || sv < __ldg(loc + r) || sv > __ldg(hic + r);
                    bad = bad || !isfinite(sv);
                    smax = fmax(smax, fabs(sv));
                }
            }
            const unsigned gmask = 0xffu << (lane & 24);             // the 8 lanes of this loop
            const unsigned vm = __ballot_sync(0xffffffffu, viol), bm = __ballot_sync(0xffffffffu, bad);
            const bool act = (vm & gmask) != 0u && (bm & gmask) == 0u;
            if (__any_sync(0xffffffffu, act)) {                      // warp-uniform: some loop of the warp violates its box
                const double *Phic = a.Phi + (size_t)c * nb * nb;
                double *bd = box_s[slot][0], *bz = box_s[slot][1], *bw = box_s[slot][2];
#pragma unroll
                for (int o = 1; o < G; o <<= 1) smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
                const double thr = a.tol * fmax(__ldg(a.bmax + c), smax);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = k + G * i;
                    if (r < nb) {
                        const double zz = fmin(fmax(su[i], __ldg(loc + r)), __ldg(hic + r));
                        bz[r] = zz;
                        bw[r] = 0.0;
                        bd[r] = su[i] - zz;
                    }
                }
                // row r of Phi times the iterate in bd
                auto phi_row = [&](int r) -> double {
                    const double2 *row = reinterpret_cast<const double2 *>(Phic + (size_t)r * nb);
                    const double2 *dv = reinterpret_cast<const double2 *>(bd);
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    int j = 0;
                    for (; j + 1 < nb / 2; j += 2) {
                        const double2 p0 = __ldg(row + j), p1 = __ldg(row + j + 1), d0 = dv[j], d1 = dv[j + 1];
                        a0 = fma(p0.x, d0.x, a0); a1 = fma(p0.y, d0.y, a1);
                        a2 = fma(p1.x, d1.x, a2); a3 = fma(p1.y, d1.y, a3);
                    }
                    if (j < nb / 2) {
                        const double2 p0 = __ldg(row + j), d0 = dv[j];
                        a0 = fma(p0.x, d0.x, a0); a1 = fma(p0.y, d0.y, a1);
                    }
                    return (a0 + a1) + (a2 + a3);
                };
                bool conv = !act;
                int it = 0, it_conv = 1;
                while (true) {
                    __syncwarp();                                    // the iterate of every loop is complete
                    if (!__any_sync(0xffffffffu, !conv) || it >= a.max_iter) break;
                    ++it;
                    double dn[8], res = 0.0;
                    if (!conv) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = k + G * i;
                            dn[i] = 0.0;
                            if (r < nb) {
                                const double zi = bz[r], wi = bw[r];
                                const double si = (zi - wi) + phi_row(r);
                                const double sr = DDMPC_ADMM_RELAX * si + (1.0 - DDMPC_ADMM_RELAX) * zi;   // over-relaxation (solve.cu)
                                const double zn = fmin(fmax(sr + wi, __ldg(loc + r)), __ldg(hic + r));
                                res = fmax(res, fmax(fabs(si - zn), fabs(zn - zi)));
                                const double wn = wi + sr - zn;
                                bz[r] = zn;
                                bw[r] = wn;
                                dn[i] = su[i] - zn + wn;
                            }
                        }
                    }
                    __syncwarp();                                    // every lane has read the old iterate
                    if (!conv) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = k + G * i;
                            if (r < nb) bd[r] = dn[i];
                        }
                    }
#pragma unroll
                    for (int o = 1; o < G; o <<= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
                    if (!conv) {
                        it_conv = it;
                        conv = res <= thr;
                    }
                }
                if (act) {
                    if (!conv) status = max(status, (int)DDMPC_SOLVE_OPTIMAL_INACCURATE);
                    extra += it_conv - 1;
                }
                // t = Phi d at the fixed point, then the correction of the planned inputs u -= Psi t
                double tv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = k + G * i;
                    tv[i] = (act && r < nb) ? phi_row(r) : 0.0;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = k + G * i;
                    if (r < nb) bd[r] = tv[i];
                }
                __syncwarp();
                if (act) {
                    const double *Psic = a.Psi + (size_t)c * a.Lm * nb;
#pragma unroll
                    for (int i = 0; i < RU; ++i) {
                        const int r = k + G * i;
                        if (r < R) {
                            const double2 *row = reinterpret_cast<const double2 *>(Psic + (size_t)r * nb);
                            const double2 *dv = reinterpret_cast<const double2 *>(bd);
                            double a0 = 0.0, a1 = 0.0;
                            for (int j = 0; j < nb / 2; ++j) {
                                const double2 pv = __ldg(row + j), d0 = dv[j];
                                a0 = fma(pv.x, d0.x, a0);
                                a1 = fma(pv.y, d0.y, a1);
                            }
                            u_own[i] -= a0 + a1;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RU; ++i) {
            const int r = k + G * i;
            if (r < R) ex[r] = u_own[i];
        }
        __syncwarp();
        double U[R];
        if constexpr (R % 2 == 0) {
#pragma unroll
            for (int j = 0; j < R; j += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(ex + j);
                U[j] = v.x;
                U[j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) U[j] = ex[j];
        }
        // ---- plant: the n_mpc steps of the block at once, row r of [Y; x+] = Mblk [x; U] (+ noise on the y rows)
        double v_own[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < NX; ++j) acc[j & 3] = fma(Mc[i][j], x[j], acc[j & 3]);
#pragma unroll
            for (int j = 0; j < R; ++j) acc[(NX + j) & 3] = fma(Mc[i][NX + j], U[j], acc[(NX + j) & 3]);
            v_own[i] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + nz[i];
            const int r = k + G * i;
            if (r < RB) ex[R + r] = v_own[i];
        }
        // ---- record (controller_operation.py:290-300): the lanes of a loop write one contiguous run per array
        if (live) {
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int r = k + G * i;
                if (r < steps * M) a.u_sys[(f0 + t0) * M + r] = u_own[i];
            }
#pragma unroll
            for (int i = 0; i < RP; ++i) {
                const int r = k + G * i;
                if (G * i < RY && r < steps * P) a.y_sys[(f0 + t0) * P + r] = v_own[i];
            }
        }
        __syncwarp();
        if constexpr (RY % 2 == 0 && R % 2 == 0) {
#pragma unroll
            for (int j = 0; j < RY; j += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(ex + R + j);
                Y[j] = v.x;
                Y[j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < RY; ++j) Y[j] = ex[R + j];
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = ex[R + RY + j];
        // ---- window update (controller.py:893-895), n_mpc steps at once; a partial block is the last one
        if constexpr (NMPC >= N) {
#pragma unroll
            for (int j = 0; j < NWU; ++j) win[j] = U[(NMPC - N) * M + j];
#pragma unroll
            for (int j = 0; j < NWY; ++j) win[NWU + j] = Y[(NMPC - N) * P + j];
        } else {
#pragma unroll
            for (int j = 0; j < NWU - R; ++j) win[j] = win[j + R];
#pragma unroll
            for (int j = 0; j < R; ++j) win[NWU - R + j] = U[j];
#pragma unroll
            for (int j = 0; j < NWY - RY; ++j) win[NWU + j] = win[NWU + j + RY];
#pragma unroll
            for (int j = 0; j < RY; ++j) win[NWU + NWY - RY + j] = Y[j];
        }
    }
    // ---- per-loop verdict: a non-finite input or output turns the state non-finite for good
    bool finite = true;
#pragma unroll
    for (int j = 0; j < NX; ++j) finite = finite && isfinite(x[j]);
    const int last_steps = a.n_steps - (nblk - 1) * NMPC;
#pragma unroll
    for (int j = 0; j < RY; ++j)
        if (j < last_steps * P) finite = finite && isfinite(Y[j]);
    if (!finite) status = max(status, (int)DDMPC_SOLVE_NONFINITE);
    if (live) {
        if (k == 0) {
            if (a.status) a.status[b] = status;
            if (a.iters) a.iters[b] = nblk + extra;
        }
        if (a.x_final) {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                if ((j & (G - 1)) == k) a.x_final[(size_t)b * NX + j] = x[j];
        }
    }
}

template <int N, int M, int P, int NX, int NMPC>
static int launch_perloop(const ddmpc_plant *plant, PerLoopArgs a, uint64_t seed, cudaStream_t st) {
    constexpr int R = NMPC * M, RY = NMPC * P, RB = RY + NX, KB = NX + R;
    using Maps = PlMaps<RB, KB>;
    static_assert(sizeof(Maps) + sizeof(PerLoopArgs) <= 4000, "block maps must fit in the kernel parameter space");
    Maps maps;
    const std::vector<double> Mb = block_map(plant, NMPC);
    for (int r = 0; r < RB; ++r)
        for (int j = 0; j < KB; ++j) maps.Mb[r][j] = maps.Mt[r][j] = Mb[(size_t)r * KB + j];
    const int rem = a.n_steps % NMPC;
    if (rem) {
        const std::vector<double> Mr = block_map(plant, rem);
        const int cr = NX + rem * M;
        for (int r = 0; r < RB; ++r)
            for (int j = 0; j < KB; ++j) maps.Mt[r][j] = 0.0;
        for (int r = 0; r < rem * P; ++r)
            for (int j = 0; j < cr; ++j) maps.Mt[r][j] = Mr[(size_t)r * cr + j];
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < cr; ++j) maps.Mt[RY + i][j] = Mr[(size_t)(rem * P + i) * cr + j];
    }
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    // one-warp CTAs deal a small batch to the SMs at the finest grain; larger CTAs once the batch fills the machine
    const int wpc = a.B > 65536 ? 4 : (a.B > 16384 ? 2 : 1);
    const int grid = ceil_div(a.B, 4 * wpc);
    if (a.nb > 0) {
        if constexpr ((N * (M + P)) % 2 == 0 && (M + P) % 2 == 0) {
            if (a.w) k_closed_loop_perloop<N, M, P, NX, NMPC, false, true><<<grid, 32 * wpc, 0, st>>>(maps, a);
            else k_closed_loop_perloop<N, M, P, NX, NMPC, true, true><<<grid, 32 * wpc, 0, st>>>(maps, a);
        } else {
            return -1;
        }
    } else if (a.w) k_closed_loop_perloop<N, M, P, NX, NMPC, false><<<grid, 32 * wpc, 0, st>>>(maps, a);
    else k_closed_loop_perloop<N, M, P, NX, NMPC, true><<<grid, 32 * wpc, 0, st>>>(maps, a);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// Largest batch of a SHARED equality-only controller that still takes this kernel.  Measured on B200 (config 3, ms per
// 401-step pass; scripts/time_paths.py, profiles/r2_closed_loop_paths.txt):
//      loops     1024     4096     8192    16384    65536
//      perloop  0.045    0.053    0.102    0.194    0.665       8 lanes per loop: shortest chain, 8x the instructions
//      ws       0.071    0.071    0.073    0.096    0.243       warp-specialised tensor-core kernel (n-step shape)
//      hybrid   0.141    0.143    0.145    0.151    0.249       thread per loop
// (n_mpc_step = 1: perloop 0.126 / 0.157 / - / 0.487 / 1.76, hybrid 0.174 / 0.174 / - / 0.227 / 0.355.)
static constexpr int kPerLoopSharedMaxB = 6143;

// Returns DDMPC_OK when handled, -1 when this path does not apply.
int closed_loop_perloop_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                            const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                            const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                            double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                            cudaStream_t st) {
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    // box rows: the CONVEX slack bound only (input / output boxes add an infeasibility check: generic kernel); 16-byte
    // loads of the per-loop operators need even row lengths
    if (d.nb > 0 && (d.nbu > 0 || d.nby > 0 || d.nb > 64 || (d.nb & 1) || (d.nth & 1))) return -1;
    const int path = set->opt_path;
    if (path != DDMPC_PATH_AUTO && path != DDMPC_PATH_PERLOOP) return -1;
    const bool per_loop = ctrl_idx != nullptr || pl.count != 1 || !d.robust;
    // shared equality-only controller: large batches belong to the thread-per-loop / tensor-core kernels; a shared
    // CONVEX controller of a shape k_closed_loop_cvx does not cover stays here whatever the batch
    if (path == DDMPC_PATH_AUTO && !per_loop && d.nb == 0 && B > kPerLoopSharedMaxB) return -1;
    PerLoopArgs a{};
    a.B = B; a.n_steps = n_steps; a.nth = d.nth; a.Lm = d.Lm; a.nfix = d.nfix;
    a.ctrl_idx = ctrl_idx;
    a.Ku = pl.Ku.d();
    a.F = d.robust ? nullptr : pl.F.d();
    a.Fnz = d.robust ? nullptr : pl.Fnz.i();
    a.x0 = x0; a.u_past0 = u_past0; a.y_past0 = y_past0; a.u_s = u_s; a.y_s = y_s; a.w = w;
    a.id0 = id0; a.eps = eps;
    a.u_sys = u_sys; a.y_sys = y_sys; a.x_final = x_final; a.status = status; a.iters = iters;
    a.nb = d.nb; a.max_iter = max_iter > 0 ? max_iter : 1000; a.tol = tol > 0.0 ? tol : 1e-8;
    a.Ks = pl.Ks.d(); a.Phi = pl.Phi.d(); a.Psi = pl.Psi.d(); a.blo = pl.lo.d(); a.bhi = pl.hi.d(); a.bmax = pl.bmax.d();
    const int nmpc = set->prm.n_mpc_step;
#define DDMPC_PERLOOP_CASE(N_, M_, P_, NX_, NMPC_)                                                   \
    if (d.n == N_ && d.m == M_ && d.p == P_ && plant->n_x == NX_ && nmpc == NMPC_)                   \
        return launch_perloop<N_, M_, P_, NX_, NMPC_>(plant, a, seed, st);
    DDMPC_PERLOOP_CASE(4, 2, 2, 4, 4)
    DDMPC_PERLOOP_CASE(4, 2, 2, 4, 1)
    DDMPC_PERLOOP_CASE(4, 2, 2, 4, 2)
    DDMPC_PERLOOP_CASE(2, 1, 1, 2, 1)
    DDMPC_PERLOOP_CASE(2, 1, 1, 2, 2)
#undef DDMPC_PERLOOP_CASE
    return -1;
}

}  // namespace ddmpc
