// Fused closed loops of LARGE systems (BASELINE config 4: n = 20, m = p = 4, n_x = 20, n_theta = 168) on the FP64
// tensor cores: ONE launch for the whole run, one warp per n-tile of 8 closed loops.
//
// Per MPC iteration (s = n_mpc steps) and warp, both products are chains of mma.sync.m8n8k4.f64 (DMMA):
//   solve :  U (n_mpc*m x 8)            = Ku[0:n_mpc*m, :] (x n_theta)  [window_u; window_y; u_s; y_s] (n_theta x 8)
//   plant :  [Y (n_mpc*p); x+ (n_x)] x 8 = Mblk (x (n_x + n_mpc*m))      [x; U]
// Mblk is the n_mpc-step block map of the LTI plant (utilities/model_simulation.py:93-98 unrolled; measurement
// noise only enters y and is added in the epilogue).  The warp's state - measurement window as a ring over the n
// time slots, set-points, plant state - lives in its own shared-memory tile [value][8 loops] (8-word rows with an
// XOR swizzle: B-fragment reads and 128-bit C-fragment writes are bank-conflict free), so there is no block-level
// synchronisation at all.  The planned inputs and the outputs of a block are written straight into the ring slots
// they occupy in the next window (controller.py:893-895).  A fragments are pre-packed in fragment order
// ([m-tile][k-step][lane]) and fetched with coalesced 256-byte loads that hit L1/L2; accumulators of every m-tile
// stay in registers (template parameters MTS / MTP), with split-K chains when there are fewer than four m-tiles.
//
// Replaces, for this shape class, closed_loop_gemm (three launches per iteration) and the thread-per-loop kernel.
// Requirements: one equality-only ROBUST controller shared by the batch, m, p, n_x multiples of 4, n_mpc <= n.
#include <vector>

#include "common.cuh"
#include "plan.cuh"

namespace ddmpc {

std::vector<double> block_map(const ddmpc_plant *pl, int s);   // gemm_loop.cu

struct DmmaArgs {
    int B, n_steps, n, m, p, nx, nmpc, nth;
    int ksS, ksP;                       // k-steps of the solve (n_theta / 4) and of the plant ((n_x + n_mpc*m) / 4)
    const double *KuF, *MbF, *MtF;      // packed A fragments: [m-tile][k-step][32]
    const double *x0, *u_past0, *y_past0, *u_s, *y_s, *w;
    unsigned long long seed, id0;
    double eps;
    double *u_sys, *y_sys, *x_final;
    int *status, *iters;
};

__device__ __forceinline__ void dl_philox(uint32_t c0, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void dl_mma(double2 &c, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c.x), "+d"(c.y)
        : "d"(a), "d"(b));
}

// C (8*MT x 8) = A (packed fragments) * B, all MT accumulators in registers.  offtab[ks] = offset (in doubles) of
// the tile row that holds the first of the 4 operand values of k-step ks; tb = tile + 8*q + swizzled column.
// With MT < 4 the k-steps are dealt round-robin to SK = 4 / MT partial accumulators (independent DMMA chains).
// KS > 0: compile-time k-step count, the whole product is straight-line code (small shapes: the scheduler hoists
// the loads of later k-steps over the DMMAs of earlier ones); KS == 0: run-time loop over ksN.
template <int MT, int KS>
__device__ __forceinline__ void dl_gemm(const double *__restrict__ AF, int ksN, const int *offtab, const double *tb,
                                        int lane, double2 (&out)[MT]) {
    constexpr int SK = MT >= 3 ? 1 : 4 / MT;     // partial accumulators per m-tile
    constexpr int U = SK > 2 ? 4 : 2;            // k-steps per trip (one 8- or 16-byte table load)
    double2 acc[MT][SK];
#pragma unroll
    for (int j = 0; j < MT; ++j)
#pragma unroll
        for (int s = 0; s < SK; ++s) acc[j][s] = make_double2(0.0, 0.0);
    if constexpr (KS > 0) {
        const double *ap = AF + lane;
        int ov[U];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            constexpr int KV = KS - KS % U;      // k-steps covered by whole vector loads of the table
            if (ks < KV && ks % U == 0) {
                if constexpr (U == 4) {
                    const int4 t = *reinterpret_cast<const int4 *>(offtab + ks);
                    ov[0] = t.x; ov[1] = t.y; ov[2] = t.z; ov[3] = t.w;
                } else {
                    const int2 t = *reinterpret_cast<const int2 *>(offtab + ks);
                    ov[0] = t.x; ov[1] = t.y;
                }
            }
            const double bv = tb[ks < KV ? ov[ks % U] : offtab[ks]];
#pragma unroll
            for (int j = 0; j < MT; ++j) dl_mma(acc[j][ks % SK], __ldg(ap + (j * KS + ks) * 32), bv);
        }
    } else {
        const double *ap = AF + lane;
        const size_t tstride = (size_t)ksN * 32;
        int ks = 0;
        for (; ks + U <= ksN; ks += U) {
            int o[U];
            if constexpr (U == 4) {
                const int4 t = *reinterpret_cast<const int4 *>(offtab + ks);
                o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
            } else {
                const int2 t = *reinterpret_cast<const int2 *>(offtab + ks);
                o[0] = t.x; o[1] = t.y;
            }
            double bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) bv[u] = tb[o[u]];
            // k-step outermost: consecutive DMMAs go to different accumulators
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < MT; ++j) dl_mma(acc[j][u % SK], __ldg(ap + j * tstride + 32 * u), bv[u]);
            ap += 32 * U;
        }
        for (; ks < ksN; ++ks) {
            const double b0 = tb[offtab[ks]];
#pragma unroll
            for (int j = 0; j < MT; ++j) dl_mma(acc[j][0], __ldg(ap + j * tstride), b0);
            ap += 32;
        }
    }
#pragma unroll
    for (int j = 0; j < MT; ++j) {
        double2 v = acc[j][0];
#pragma unroll
        for (int s = 1; s < SK; ++s) { v.x += acc[j][s].x; v.y += acc[j][s].y; }
        out[j] = v;
    }
}

// Same product with the A fragments already in registers (small shapes: the block map is loop invariant).
template <int MT, int KS>
__device__ __forceinline__ void dl_gemm_reg(const double (&areg)[MT * KS], const int *offtab, const double *tb,
                                            double2 (&out)[MT]) {
    constexpr int SK = MT >= 3 ? 1 : 4 / MT;
#pragma unroll
    for (int j = 0; j < MT; ++j) out[j] = make_double2(0.0, 0.0);
    double2 acc[MT][SK];
#pragma unroll
    for (int j = 0; j < MT; ++j)
#pragma unroll
        for (int s = 0; s < SK; ++s) acc[j][s] = make_double2(0.0, 0.0);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const double bv = tb[offtab[ks]];
#pragma unroll
        for (int j = 0; j < MT; ++j) dl_mma(acc[j][ks % SK], areg[j * KS + ks], bv);
    }
#pragma unroll
    for (int j = 0; j < MT; ++j) {
        double2 v = acc[j][0];
#pragma unroll
        for (int s = 1; s < SK; ++s) { v.x += acc[j][s].x; v.y += acc[j][s].y; }
        out[j] = v;
    }
}

// DL_WARPS warps (n-tiles of 8 loops) per CTA.  The warps are independent (no block-level barrier), so the CTA size only
// sets the granularity with which the batch is dealt to the SMs.
// KU, KY > 0 (small shapes, one m-tile of planned inputs): k-steps of the input and output halves of the window.  The solve
// then walks the ring in PHYSICAL slot order - the order of the terms of a dot product is free - so its B operands sit at
// compile-time offsets of the tile, and the rotation moves to the A side: the gain fragments of each half are stored
// twice in a row, and the fragment of physical k-step j is entry j + (K - base * k-steps-per-slot) of that doubled run.
// No row table, no table rotation, no address arithmetic per DMMA.
template <int MTS, int MTP, int KSS, int KSP, int DL_WARPS, int KU = 0, int KY = 0>
__global__ void __launch_bounds__(32 * DL_WARPS, 16 / DL_WARPS)
k_closed_loop_dmma(const DmmaArgs a) {
    constexpr bool PHYS = KU > 0 && KY > 0;
    static_assert(!PHYS || (MTS == 1 && KSS > KU + KY), "physical-order solve: one m-tile, compile-time k-step counts");
    extern __shared__ __align__(16) double dl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int n = a.n, m = a.m, p = a.p, nx = a.nx, nmpc = a.nmpc;
    const int nm = n * m, npp = n * p, R = nmpc * m, RY = nmpc * p;
    // tile rows: [window_u (n*m) | window_y (n*p) | u_s, y_s (m + p) | x (n_x)], then the two offset tables
    const int oWY = nm, oSP = nm + npp, oX = oSP + m + p, rows = oX + nx;
    const int ksSe = (a.ksS + 3) & ~3, ksPe = (a.ksP + 3) & ~3;   // tables start on 16-byte boundaries
    const int tile_doubles = rows * 8 + (((ksSe + ksPe) / 2 + 1) & ~1);   // even: tiles stay 16-byte aligned
    double *tile = dl_smem + (size_t)warp * tile_doubles;
    int *tabS = reinterpret_cast<int *>(tile + rows * 8), *tabP = tabS + ksSe;
    const int b0 = (blockIdx.x * DL_WARPS + warp) * 8;
    if (b0 >= a.B) return;                                   // whole warp idle (no block-level barriers anywhere)
    auto SW = [](int row, int col) { return col ^ (((row >> 1) & 1) << 2); };
    // B fragment: row (4 | table row) + q, column g -> constant per-lane offset
    const double *tb = tile + q * 8 + (g ^ (((q >> 1) & 1) << 2));

    // ---- load the state of the warp's 8 loops
    for (int e = lane; e < rows * 8; e += 32) {
        const int r = e >> 3, col = e & 7;
        const int b = min(b0 + col, a.B - 1);                // dead columns replay the last loop and never store
        double v;
        if (r < oWY) v = a.u_past0[(size_t)b * nm + r];
        else if (r < oSP) v = a.y_past0[(size_t)b * npp + (r - oWY)];
        else if (r < oSP + m) v = a.u_s[(size_t)b * m + (r - oSP)];
        else if (r < oX) v = a.y_s[(size_t)b * p + (r - oSP - m)];
        else v = a.x0[(size_t)b * nx + (r - oX)];
        tile[r * 8 + SW(r, col)] = v;
    }
    // ---- offset tables for ring base 0 (entries of ring rows rotate by `steps` slots after every block)
    if constexpr (!PHYS) {
        for (int ks = lane; ks < a.ksS; ks += 32) {
            const int e0 = 4 * ks;                            // theta order: window_u, window_y, set-points
            tabS[ks] = 8 * (e0 < nm + npp ? e0 : oSP + (e0 - nm - npp));
        }
    }
    for (int ks = lane; ks < a.ksP; ks += 32) {
        const int e0 = 4 * ks;                                // [x; U]: U of step s sits in ring slot (base + s) % n
        tabP[ks] = 8 * (e0 < nx ? oX + e0 : e0 - nx);
    }
    // (row, channel) of the C-fragment rows this lane holds, per m-tile
    int sU[MTS], iU[MTS], sY[MTP], iY[MTP];
#pragma unroll
    for (int j = 0; j < MTS; ++j) { sU[j] = (8 * j + g) / m; iU[j] = (8 * j + g) % m; }
#pragma unroll
    for (int j = 0; j < MTP; ++j) { sY[j] = (8 * j + g) / p; iY[j] = (8 * j + g) % p; }
    // the two loops whose C-fragment columns this lane holds
    const int bc[2] = {b0 + 2 * q, b0 + 2 * q + 1};
    const bool live[2] = {bc[0] < a.B, bc[1] < a.B};
    double *const urow[2] = {a.u_sys + (size_t)(live[0] ? bc[0] : 0) * a.n_steps * m,
                             a.u_sys + (size_t)(live[1] ? bc[1] : 0) * a.n_steps * m};
    double *const yrow[2] = {a.y_sys + (size_t)(live[0] ? bc[0] : 0) * a.n_steps * p,
                             a.y_sys + (size_t)(live[1] ? bc[1] : 0) * a.n_steps * p};
    // noise: lane handles (loop = lane % 8, Philox calls lane / 8, lane / 8 + 4, ...) of every block
    const int nb_ = min(b0 + (lane & 7), a.B - 1);
    const unsigned long long nsid = a.id0 + (unsigned long long)nb_;
    const int cps = p >> 2;                                   // Philox calls per step (4 noise words each)
    bool fin[2] = {true, true};
    int base = 0, n_iter = 0;                                 // ring slot of the oldest window entry
    // (slot within the block, channel) of the plant-table entry this lane rotates: no division inside the loop
    const bool pRot = lane < a.ksP && 4 * lane >= nx;
    const int pj = pRot ? (4 * lane - nx) / m : 0, pi = pRot ? (4 * lane - nx) % m : 0;
    // small shapes: the A fragments of the block map stay in registers for the whole run
    constexpr bool PLANT_REG = KSP > 0 && MTP * KSP <= 24;
    double aP[PLANT_REG ? MTP * KSP : 1];
    if constexpr (PLANT_REG) {
#pragma unroll
        for (int e = 0; e < MTP * KSP; ++e) aP[e] = __ldg(a.MbF + (size_t)e * 32 + lane);
    }
    __syncwarp();

    for (int t0 = 0; t0 < a.n_steps; t0 += nmpc, ++n_iter) {
        const int steps = min(nmpc, a.n_steps - t0);
        const double *MF = steps == nmpc ? a.MbF : a.MtF;
        // a non-finite input or output turns the state non-finite for good (x+ = A x + B u, the window feeds u), so
        // looking at the outputs and the state of the last block is enough
        const bool last = t0 + nmpc >= a.n_steps;
        // noise words 4c .. 4c+3 of the block (word qw = k*p + i is word (qw & 3) of Philox call (qw >> 2)) for loop lane % 8
        auto draw = [&](const int c, double (&nz)[4]) {
            const int s = cps == 1 ? c : c / cps, i0 = 4 * (c - s * cps), k = t0 + s;
            if (a.w) {
#pragma unroll
                for (int i = 0; i < 4; ++i) nz[i] = __ldg(a.w + ((size_t)nb_ * a.n_steps + k) * p + i0 + i);
            } else {
                uint32_t o[4];
                dl_philox((unsigned)k * (unsigned)cps + (unsigned)(c - s * cps), (uint32_t)(nsid & 0xffffffffu),
                          (uint32_t)(nsid >> 32), (uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), o);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    nz[i] = a.eps * (2.0 * __hiloint2double((int)(0x3FF00000u | (o[i] >> 12)), (int)(o[i] << 20)) - 3.0);
            }
        };
        // ---- solve: planned inputs of the block
        double2 uo[MTS];
        if constexpr (PHYS) {
            constexpr int KP = KSS - KU - KY;                  // k-steps of the set-point rows
            const double *apU = a.KuF + (size_t)(KU - base * (m >> 2)) * 32 + lane;
            const double *apY = a.KuF + (size_t)(2 * KU + KY - base * (p >> 2)) * 32 + lane;
            const double *apS = a.KuF + (size_t)(2 * KU + 2 * KY) * 32 + lane;
            constexpr int NA = 8;                              // independent accumulator chains (a 2048-warp batch gives a
            double2 acc[NA];                                   // scheduler only 3-4 warps to hide the DMMA latency with)
#pragma unroll
            for (int i = 0; i < NA; ++i) acc[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int ks = 0; ks < KU; ++ks) dl_mma(acc[ks % NA], __ldg(apU + 32 * ks), tb[32 * ks]);
#pragma unroll
            for (int ks = 0; ks < KY; ++ks) dl_mma(acc[(KU + ks) % NA], __ldg(apY + 32 * ks), tb[32 * (KU + ks)]);
#pragma unroll
            for (int ks = 0; ks < KP; ++ks) dl_mma(acc[(KU + KY + ks) % NA], __ldg(apS + 32 * ks), tb[32 * (KU + KY + ks)]);
#pragma unroll
            for (int w = NA / 2; w >= 1; w >>= 1)
#pragma unroll
                for (int i = 0; i < w; ++i) { acc[i].x += acc[i + w].x; acc[i].y += acc[i + w].y; }
            uo[0] = acc[0];
        } else {
            dl_gemm<MTS, KSS>(a.KuF, a.ksS, tabS, tb, lane, uo);
        }
        __syncwarp();                                         // every lane has read the old window
#pragma unroll
        for (int j = 0; j < MTS; ++j) {
            if (8 * j + g < R && sU[j] < steps) {
                int slot = base + sU[j];
                if (slot >= n) slot -= n;                     // the slot this input occupies in the next window
                const int r = slot * m + iU[j];
                *reinterpret_cast<double2 *>(&tile[r * 8 + SW(r, 2 * q)]) = uo[j];
                const size_t f = (size_t)(t0 + sU[j]) * m + iU[j];
                if (live[0]) urow[0][f] = uo[j].x;
                if (live[1]) urow[1][f] = uo[j].y;
            }
        }
        // ---- measurement noise of the block, parked in the output slots of the ring (the old outputs there have
        // been consumed by the solve)
        for (int c = lane >> 3; c < steps * cps; c += 4) {
            const int s = cps == 1 ? c : c / cps, i0 = 4 * (c - s * cps);
            double nz[4];
            draw(c, nz);
            int slot = base + s;
            if (slot >= n) slot -= n;
            const int r = oWY + slot * p + i0;
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[(r + i) * 8 + SW(r + i, lane & 7)] = nz[i];
        }
        // rotate the U entries of the plant table to this block's slots, then the plant GEMM
        if (pRot) {
            int slot = base + pj;
            if (slot >= n) slot -= n;
            tabP[lane] = 8 * (slot * m + pi);
        }
        for (int ks = lane + 32; ks < a.ksP; ks += 32) {
            const int e0 = 4 * ks;
            int slot = base + (e0 - nx) / m;
            if (slot >= n) slot -= n;
            tabP[ks] = 8 * (slot * m + (e0 - nx) % m);
        }
        __syncwarp();
        // ---- plant: outputs of the block and the state after it
        double2 yo[MTP];
        if constexpr (PLANT_REG) {
            if (steps != nmpc) {                              // last, partial block: its own block map
#pragma unroll
                for (int e = 0; e < MTP * KSP; ++e) aP[e] = __ldg(a.MtF + (size_t)e * 32 + lane);
            }
            dl_gemm_reg<MTP, KSP>(aP, tabP, tb, yo);
        } else {
            dl_gemm<MTP, KSP>(MF, a.ksP, tabP, tb, lane, yo);
        }
        __syncwarp();                                         // every lane has read the old state
#pragma unroll
        for (int j = 0; j < MTP; ++j) {
            const int row = 8 * j + g;
            if (row < RY) {
                if (sY[j] < steps) {
                    int slot = base + sY[j];
                    if (slot >= n) slot -= n;
                    const int r = oWY + slot * p + iY[j];
                    double2 *cell = reinterpret_cast<double2 *>(&tile[r * 8 + SW(r, 2 * q)]);
                    const double2 nz = *cell;
                    const double2 y = make_double2(yo[j].x + nz.x, yo[j].y + nz.y);
                    *cell = y;
                    const size_t f = (size_t)(t0 + sY[j]) * p + iY[j];
                    if (live[0]) yrow[0][f] = y.x;
                    if (live[1]) yrow[1][f] = y.y;
                    if (last) {
                        fin[0] = fin[0] && isfinite(y.x);
                        fin[1] = fin[1] && isfinite(y.y);
                    }
                }
            } else if (row < RY + nx) {
                const int r = oX + (row - RY);
                *reinterpret_cast<double2 *>(&tile[r * 8 + SW(r, 2 * q)]) = yo[j];
                if (last) {
                    fin[0] = fin[0] && isfinite(yo[j].x);
                    fin[1] = fin[1] && isfinite(yo[j].y);
                }
            }
        }
        // ---- rotate the ring entries of the solve table by `steps` slots
        base += steps;
        if (base >= n) base -= n;
        if constexpr (!PHYS) {
            for (int ks = lane; ks < a.ksS; ks += 32) {
                int o = tabS[ks];
                if (o < 8 * oWY) { o += 8 * steps * m; if (o >= 8 * oWY) o -= 8 * nm; }
                else if (o < 8 * oSP) { o += 8 * steps * p; if (o >= 8 * oSP) o -= 8 * npp; }
                tabS[ks] = o;
            }
        }
        __syncwarp();
    }
    // ---- per-loop results: a loop's columns are spread over the 8 lanes with the same q
    const unsigned bad0 = __ballot_sync(0xffffffffu, !fin[0]), bad1 = __ballot_sync(0xffffffffu, !fin[1]);
    unsigned qmask = 0u;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) qmask |= 1u << (4 * gg + q);
    const bool loop_bad[2] = {(bad0 & qmask) != 0u, (bad1 & qmask) != 0u};
    if (g == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!live[h]) continue;
            if (a.status) a.status[bc[h]] = loop_bad[h] ? DDMPC_SOLVE_NONFINITE : DDMPC_SOLVE_OPTIMAL;
            if (a.iters) a.iters[bc[h]] = n_iter;
        }
    }
    if (a.x_final) {
        for (int e = lane; e < nx * 8; e += 32) {
            const int i = e >> 3, col = e & 7;
            if (b0 + col < a.B) a.x_final[(size_t)(b0 + col) * nx + i] = tile[(oX + i) * 8 + SW(oX + i, col)];
        }
    }
}

// pack rows [0, rows) x cols [0, cols) of a row-major matrix (ld) into A-fragment order [m-tile][k-step][lane]
static void pack_fragments(const double *M, int ld, int rows, int cols, int mt, int ks, std::vector<double> &out) {
    out.assign((size_t)mt * ks * 32, 0.0);
    for (int j = 0; j < mt; ++j)
        for (int k = 0; k < ks; ++k)
            for (int lane = 0; lane < 32; ++lane) {
                const int r = 8 * j + (lane >> 2), c = 4 * k + (lane & 3);
                if (r < rows && c < cols) out[((size_t)j * ks + k) * 32 + lane] = M[(size_t)r * ld + c];
            }
}

template <int MTS, int MTP, int KSS, int KSP, int DL_WARPS, int KU, int KY>
static int launch_dmma_w(const DmmaArgs &a, size_t smem_per_warp, cudaStream_t st) {
    const size_t smem = smem_per_warp * DL_WARPS;
    auto kern = k_closed_loop_dmma<MTS, MTP, KSS, KSP, DL_WARPS, KU, KY>;
    DDMPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DDMPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<ceil_div(a.B, 8 * DL_WARPS), 32 * DL_WARPS, smem, st>>>(a);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// Warps per CTA.  One-warp CTAs deal the batch to the SMs at the finest grain (16,384 loops = 2048 warps on 148 SMs: 14 or
// 13 per SM instead of 16 or 12 with CTAs of four) and measured 6 % faster than CTAs of 2 or 4 on config 4
// (1.455 vs 1.541 / 1.553 ms per 401-step pass, 0.903 vs 0.953 / 0.951 ms with n_mpc_step = 20);
// ddmpc_set_option("dmma_warps", 2 | 4) selects the larger ones.
template <int MTS, int MTP, int KSS, int KSP, int KU = 0, int KY = 0>
static int launch_dmma(const DmmaArgs &a, size_t smem_per_warp, int best, cudaStream_t st) {
    if (best == 4) return launch_dmma_w<MTS, MTP, KSS, KSP, 4, KU, KY>(a, smem_per_warp, st);
    if (best == 2) return launch_dmma_w<MTS, MTP, KSS, KSP, 2, KU, KY>(a, smem_per_warp, st);
    return launch_dmma_w<MTS, MTP, KSS, KSP, 1, KU, KY>(a, smem_per_warp, st);
}

// Returns DDMPC_OK when handled, -1 when this path does not apply.
int closed_loop_dmma_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st) {
    const Dims &d = set->plan.d;
    if (ctrl_idx || set->plan.count != 1 || d.nb > 0 || !d.robust) return -1;
    if (set->opt_path != DDMPC_PATH_AUTO && set->opt_path != DDMPC_PATH_DMMA) return -1;
    const int n = d.n, m = d.m, p = d.p, nx = plant->n_x, nth = d.nth, nmpc = set->prm.n_mpc_step;
    if ((m % 4) || (p % 4) || (nx % 4) || nmpc > n || nth < 64 || B < 256) return -1;
    const int R = nmpc * m, rowsP = nmpc * p + nx, colsP = nx + R;
    const int mtS = ceil_div(R, 8), mtP = ceil_div(rowsP, 8), ksS = nth / 4, ksP = colsP / 4;
    // compiled shapes: config 4 (n = 20, m = p = 4, n_x = 20) with n_mpc = 1 (straight-line products) and n_mpc = 20
    const bool shape1 = mtS == 1 && mtP == 3 && ksS == 42 && ksP == 6 && n * m == 80 && n * p == 80,
               shape20 = mtS == 10 && mtP == 13;
    if (!shape1 && !shape20) return -1;
    const int rows = n * (m + p) + m + p + nx;
    const size_t smem = sizeof(double) * ((size_t)rows * 8 + (((((ksS + 3) & ~3) + ((ksP + 3) & ~3)) / 2 + 1) & ~1));   // per warp
    if (smem * 4 > 200 * 1024) return -1;

    // packed operands (cached in the set; rebuilt when the plant changes)
    const int rem = n_steps % nmpc;
    std::vector<double> key(plant->A, plant->A + (size_t)nx * nx);
    key.insert(key.end(), plant->B, plant->B + (size_t)nx * m);
    key.insert(key.end(), plant->C, plant->C + (size_t)p * nx);
    key.insert(key.end(), plant->D, plant->D + (size_t)p * m);
    key.push_back((double)rem);
    // shape 1 walks the window in physical slot order: each half of the gain is stored twice in a row (see the kernel)
    const int ksKu = shape1 ? (2 * n * m + 2 * n * p + m + p) / 4 : ksS;
    const size_t nKu = (size_t)mtS * ksKu * 32, nMb = (size_t)mtP * ksP * 32;
    if (set->dmma_key != key || set->dmma_ws.bytes < sizeof(double) * (nKu + 2 * nMb)) {
        DDMPC_CUDA(cudaDeviceSynchronize());                  // loops still reading the previous operands
        std::vector<double> hKu((size_t)R * nth), fKu, fMb, fMt;
        DDMPC_CUDA(cudaMemcpy(hKu.data(), set->plan.Ku.d(), sizeof(double) * hKu.size(), cudaMemcpyDeviceToHost));
        if (shape1) {
            const int nm = n * m, npp = n * p, c2 = 2 * nm + 2 * npp + m + p;
            std::vector<double> K2((size_t)R * c2);
            for (int r = 0; r < R; ++r) {
                const double *src = hKu.data() + (size_t)r * nth;
                double *dst = K2.data() + (size_t)r * c2;
                for (int c = 0; c < nm; ++c) dst[c] = dst[nm + c] = src[c];
                for (int c = 0; c < npp; ++c) dst[2 * nm + c] = dst[2 * nm + npp + c] = src[nm + c];
                for (int c = 0; c < m + p; ++c) dst[2 * nm + 2 * npp + c] = src[nm + npp + c];
            }
            pack_fragments(K2.data(), c2, R, c2, mtS, ksKu, fKu);
        } else {
            pack_fragments(hKu.data(), nth, R, nth, mtS, ksS, fKu);
        }
        const std::vector<double> Mb = block_map(plant, nmpc);
        pack_fragments(Mb.data(), colsP, rowsP, colsP, mtP, ksP, fMb);
        // last, partial block (controller_operation.py:278): its block map in the layout of the full one
        std::vector<double> Mt((size_t)rowsP * colsP, 0.0);
        if (rem) {
            const std::vector<double> Mr = block_map(plant, rem);
            const int cr = nx + rem * m;
            for (int r = 0; r < rem * p; ++r)
                for (int c = 0; c < cr; ++c) Mt[(size_t)r * colsP + c] = Mr[(size_t)r * cr + c];
            for (int i = 0; i < nx; ++i)
                for (int c = 0; c < cr; ++c) Mt[(size_t)(nmpc * p + i) * colsP + c] = Mr[(size_t)(rem * p + i) * cr + c];
        }
        pack_fragments(Mt.data(), colsP, rowsP, colsP, mtP, ksP, fMt);
        DDMPC_CUDA(set->dmma_ws.alloc(sizeof(double) * (nKu + 2 * nMb)));
        DDMPC_CUDA(cudaMemcpy(set->dmma_ws.d(), fKu.data(), sizeof(double) * nKu, cudaMemcpyHostToDevice));
        DDMPC_CUDA(cudaMemcpy(set->dmma_ws.d() + nKu, fMb.data(), sizeof(double) * nMb, cudaMemcpyHostToDevice));
        DDMPC_CUDA(cudaMemcpy(set->dmma_ws.d() + nKu + nMb, fMt.data(), sizeof(double) * nMb, cudaMemcpyHostToDevice));
        set->dmma_key = key;
    }
    DmmaArgs a{};
    a.B = B; a.n_steps = n_steps; a.n = n; a.m = m; a.p = p; a.nx = nx; a.nmpc = nmpc; a.nth = nth;
    a.ksS = ksS; a.ksP = ksP;
    a.KuF = set->dmma_ws.d(); a.MbF = a.KuF + nKu; a.MtF = a.MbF + nMb;
    a.x0 = x0; a.u_past0 = u_past0; a.y_past0 = y_past0; a.u_s = u_s; a.y_s = y_s; a.w = w;
    a.seed = seed; a.id0 = id0; a.eps = eps;
    a.u_sys = u_sys; a.y_sys = y_sys; a.x_final = x_final; a.status = status; a.iters = iters;
    if (shape1) return launch_dmma<1, 3, 42, 6, 20, 20>(a, smem, set->opt_dmma_warps, st);
    return launch_dmma<10, 13, 0, 0>(a, smem, set->opt_dmma_warps, st);
}

}  // namespace ddmpc
