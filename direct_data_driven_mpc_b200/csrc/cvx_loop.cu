// The ITERATIVE part of the QP solver on the FP64 tensor cores: box-row ADMM for a controller that is shared by the
// batch (ROBUST + CONVEX slack bound ||sigma_pred||_inf <= c eps_max, direct_data_driven_mpc_controller.py:659-675;
// cvxpy's problem.solve() at :753 is what it replaces).
//
// with the equalities kept inside the factorised system, the only split rows are the n_B box rows and one
// ADMM iteration is
//        d = s_unc - (z - w);   s = (z - w) + Phi d;   s_r = alpha s + (1 - alpha) z;   z+ = clip(s_r + w);   w+ = w + s_r - z+
// i.e. ONE shared n_B x n_B matrix times the batch of iterates, then an element-wise projection / dual update and two
// max-norm residuals per problem.  Here a CTA of four warps carries 32 problems:
//   * Phi d as mma.sync.m8n8k4.f64 (DMMA): warp j owns box rows 16j .. 16j+15 (two m-tiles) of all four n-tiles (8 problems
//     each).  Its A fragments of Phi (2 x 16 k-steps) stay in REGISTERS for the whole solve; the B operand d lives in shared
//     memory [row][problem] (row stride 36: conflict-free fragment reads), double-buffered, written by the epilogue in
//     C-fragment order.  120 DMMA per warp and iteration, 8 independent accumulator chains per warp.
//   * the epilogue works on the C fragments in place: projection, dual update, residuals; the two max-norms of a problem
//     are reduced over its rows with warp shuffles (the 8 lanes of a fragment column) and over the four warps with one
//     shared-memory atomicMax per problem;
//   * problems converge independently: a converged problem is frozen (its iterate no longer changes), so its result does
//     not depend on which other problems share its CTA; the CTA leaves the loop when all 32 are frozen.
// Two users:
//   k_admm_dmma          batched solve (ddmpc_solve_batch with a shared CONVEX controller): S = Theta Ks^T and the other
//                        products of the solve are k_gemm calls around it (linalg.cuh);
//   k_closed_loop_cvx    fused closed loop (four-tank n-step shape): per MPC iteration the slack check Ks theta (the same
//                        row split, 40 DMMA per warp), the gain product, the ADMM when any of the CTA's 32 loops violates
//                        the bound, the plant block map, noise and recording - one launch for the whole run.
// Replaces the thread-per-problem ADMM of solve.cu / the one-loop-at-a-time warp ADMM of fast_loop.cu for these shapes.
#include <type_traits>
#include <vector>

#include "linalg.cuh"
#include "plan.cuh"

namespace ddmpc {

std::vector<double> block_map(const ddmpc_plant *pl, int s);   // gemm_loop.cu

constexpr int CV_NL = 32;   // problems (closed loops) per CTA
constexpr int CV_NT = 4;    // n-tiles of 8 problems
constexpr int CV_DS = 36;   // row stride of the [value][problem] shared tiles (4 mod 16 doubles)
constexpr int CV_NR = 64;   // box rows, padded: 8 m-tiles, two per warp
constexpr int CV_KS = 16;   // k-steps of Phi d

struct AdmmSmem {
    double d[2][CV_NR][CV_DS];              // iterate d = s_unc - z + w (B operand), double-buffered; also t for the correction
    double2 su[2 * CV_NT][128];             // s_unc in C-fragment order, per thread
    unsigned long long red[3][CV_NL];       // per-problem residual max over the four warps (rotating slots)
    unsigned long long smax[CV_NL];         // per-problem max |s_unc|
    double thr[CV_NL];                      // per-problem stopping threshold
    unsigned vmask[2][4], bmask[2][4];      // per-warp violation / non-finite masks over the 32 problems (by block parity)
    int extra[CV_NL], stat[CV_NL];          // closed loop: ADMM iterations beyond the check, worst ADMM status
};

__device__ __forceinline__ void dmma884(double2 &c, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c.x), "+d"(c.y)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double colmax8(double v) {   // max over the 8 lanes (g = 0..7) that share a fragment column
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 16));
    return v;
}
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// acc (rows of this warp x 32 problems) = Phi[rows, :] * D, D = sm.d[buf].  (Skipping the n-tiles whose eight problems have
// all converged was measured and dropped: the predicates cost more than the DMMAs they save, c = 0.3 went 4.0 -> 4.4 ms.)
// A operand of the ADMM iteration: this warp's rows of Phi, either resident in registers (64 per thread) or streamed from
// the zero-padded 64 x 64 copy (L1-resident; the closed-loop kernel's register budget goes to its per-block path).
struct PhiRegs {
    double a[2][CV_KS];
    __device__ __forceinline__ void load(const double *Phi, int nb, int warp, int g, int q) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int r = 16 * warp + 8 * mt + g;
#pragma unroll
            for (int ks = 0; ks < CV_KS; ++ks) {
                const int c = 4 * ks + q;
                a[mt][ks] = (r < nb && c < nb) ? __ldg(Phi + (size_t)r * nb + c) : 0.0;
            }
        }
    }
    __device__ __forceinline__ double get(int mt, int ks) const { return a[mt][ks]; }
};
struct PhiStream {
    const double *row[2];
    __device__ __forceinline__ void load(const double *Phi64, int warp, int g, int q) {
        row[0] = Phi64 + (size_t)(16 * warp + g) * CV_NR + q;
        row[1] = row[0] + 8 * CV_NR;
    }
    __device__ __forceinline__ double get(int mt, int ks) const { return __ldg(row[mt] + 4 * ks); }
};

template <class PhiA>
__device__ __forceinline__ void phi_times_d(const PhiA &aPhi, const double (*D)[CV_DS], int ks_n, int g, int q,
                                            double2 (&acc)[2][CV_NT]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) acc[mt][nt] = make_double2(0.0, 0.0);
#pragma unroll
    for (int ks = 0; ks < CV_KS; ++ks) {
        if (ks < ks_n) {
            const double a0 = aPhi.get(0, ks), a1 = aPhi.get(1, ks);
            double b[CV_NT];
#pragma unroll
            for (int nt = 0; nt < CV_NT; ++nt) b[nt] = D[4 * ks + q][8 * nt + g];
#pragma unroll
            for (int nt = 0; nt < CV_NT; ++nt) {
                dmma884(acc[0][nt], a0, b[nt]);
                dmma884(acc[1][nt], a1, b[nt]);
            }
        }
    }
}

// CTA-cooperative ADMM on the box rows of 32 problems (all 128 threads must call it).
//   aPhi   A fragments of Phi for this warp's rows          lo, hi  bounds of this lane's two rows (16 warp + 8 mt + g)
//   act    bit (2 nt + h) set: problem 8 nt + 2 q + h of this lane's fragment columns is active (its box is violated)
//   sm.su  s_unc (C-fragment order), sm.thr thresholds: written by the caller, visible after the first barrier in here
// Out: t = Phi d at the fixed point (zero for inactive problems), iterations per problem (1 when inactive), bits of the
// problems that hit max_iter.  Ends with a barrier: sm.d may be reused by the caller.
template <class PhiA>
__device__ __forceinline__ void admm_cta(const PhiA &aPhi, const double (&lo)[2], const double (&hi)[2],
                                         unsigned act, int ks_n, int max_iter, AdmmSmem &sm, double2 (&t)[2][CV_NT],
                                         int (&iters)[CV_NT][2], unsigned &inacc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    double2 z[2][CV_NT], w[2][CV_NT];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) {
            const double2 s = sm.su[mt * CV_NT + nt][tid];
            z[mt][nt] = make_double2(clipd(s.x, lo[mt], hi[mt]), clipd(s.y, lo[mt], hi[mt]));
            w[mt][nt] = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(&sm.d[0][16 * warp + 8 * mt + g][8 * nt + 2 * q]) =
                make_double2(s.x - z[mt][nt].x, s.y - z[mt][nt].y);
        }
#pragma unroll
    for (int nt = 0; nt < CV_NT; ++nt) iters[nt][0] = iters[nt][1] = 1;
    if (tid < CV_NL) sm.red[0][tid] = sm.red[1][tid] = sm.red[2][tid] = 0ull;
    unsigned frozen = ~act & 0xffu;
    int cur = 0, it = 0;
    while (true) {
        const int all = __syncthreads_and(frozen == 0xffu);          // also: d[cur], su, thr, red slots are visible
        if (all || it >= max_iter) break;
        ++it;
        const int slot = it % 3;
        double2 acc[2][CV_NT];
        phi_times_d(aPhi, sm.d[cur], ks_n, g, q, acc);
        double res[CV_NT][2];
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) res[nt][0] = res[nt][1] = 0.0;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < CV_NT; ++nt) {
                const double2 s = sm.su[mt * CV_NT + nt][tid];
                double dn[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double &zz = h ? z[mt][nt].y : z[mt][nt].x;
                    double &ww = h ? w[mt][nt].y : w[mt][nt].x;
                    const double a = h ? acc[mt][nt].y : acc[mt][nt].x, sv = h ? s.y : s.x;
                    const double si = (zz - ww) + a;
                    const double sr = DDMPC_ADMM_RELAX * si + (1.0 - DDMPC_ADMM_RELAX) * zz;   // over-relaxation (solve.cu)
                    const double zn = clipd(sr + ww, lo[mt], hi[mt]);
                    if (!((frozen >> (2 * nt + h)) & 1u)) {
                        res[nt][h] = fmax(res[nt][h], fmax(fabs(si - zn), fabs(zn - zz)));
                        ww = ww + sr - zn;
                        zz = zn;
                    }
                    dn[h] = sv - zz + ww;
                }
                *reinterpret_cast<double2 *>(&sm.d[cur ^ 1][16 * warp + 8 * mt + g][8 * nt + 2 * q]) = make_double2(dn[0], dn[1]);
            }
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double r = colmax8(res[nt][h]);                // (every lane takes part in the shuffles: columns differ in q)
                if (g == 0 && !((frozen >> (2 * nt + h)) & 1u))
                    atomicMax(&sm.red[slot][8 * nt + 2 * q + h], (unsigned long long)__double_as_longlong(r));
            }
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if ((frozen >> (2 * nt + h)) & 1u) continue;
                const int p = 8 * nt + 2 * q + h;
                iters[nt][h] = it;
                if (__longlong_as_double((long long)sm.red[slot][p]) <= sm.thr[p]) frozen |= 1u << (2 * nt + h);
            }
        // the slot of iteration it + 1 was last read in iteration it - 2: clear it (visible after the next barrier)
        if (tid < CV_NL) sm.red[(it + 1) % 3][tid] = 0ull;
        cur ^= 1;
    }
    inacc = act & ~frozen & 0xffu;
    double2 acc[2][CV_NT];
    phi_times_d(aPhi, sm.d[cur], ks_n, g, q, acc);                   // t = Phi d at the fixed point
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt)
            t[mt][nt] = make_double2(((act >> (2 * nt)) & 1u) ? acc[mt][nt].x : 0.0, ((act >> (2 * nt + 1)) & 1u) ? acc[mt][nt].y : 0.0);
    __syncthreads();
}

// Per-problem max |s_unc| -> stopping thresholds tol * max(bmax, smax) in sm.thr.  All threads; ends with a barrier.
__device__ __forceinline__ void admm_thresholds(const double2 (&su)[2][CV_NT], double tol, double bmax, AdmmSmem &sm) {
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
    if (tid < CV_NL) sm.smax[tid] = 0ull;
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double m = 0.0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) m = fmax(m, fabs(h ? su[mt][nt].y : su[mt][nt].x));   // (fmax drops NaN)
            m = colmax8(m);
            if (g == 0) atomicMax(&sm.smax[8 * nt + 2 * q + h], (unsigned long long)__double_as_longlong(m));
        }
    __syncthreads();
    if (tid < CV_NL) sm.thr[tid] = tol * fmax(bmax, __longlong_as_double((long long)sm.smax[tid]));
    // (visible to everybody after the first barrier of admm_cta)
}

// Which of the 32 problems violate their box / hold a non-finite slack, from this warp's rows: bit p of the results.
__device__ __forceinline__ void box_flags(const double2 (&su)[2][CV_NT], const double (&lo)[2], const double (&hi)[2], int q,
                                          unsigned &viol, unsigned &bad) {
    viol = bad = 0u;
#pragma unroll
    for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            bool v = false, b = false;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const double s = h ? su[mt][nt].y : su[mt][nt].x;
                v = v || s < lo[mt] || s > hi[mt];
                b = b || !isfinite(s);
            }
            // lanes q, q + 4, .. hold the rows of column (nt, 2q + h)
            const unsigned bv = __ballot_sync(0xffffffffu, v), bb = __ballot_sync(0xffffffffu, b);
            if ((bv >> q) & 0x11111111u) viol |= 1u << (8 * nt + 2 * q + h);
            if ((bb >> q) & 0x11111111u) bad |= 1u << (8 * nt + 2 * q + h);
        }
    viol = __reduce_or_sync(0xffffffffu, viol);
    bad = __reduce_or_sync(0xffffffffu, bad);
}

// active problems of this lane's fragment columns: bit (2 nt + h) <- bit (8 nt + 2 q + h) of the CTA-wide mask
__device__ __forceinline__ unsigned lane_bits(unsigned mask32, int q) {
    unsigned a = 0u;
#pragma unroll
    for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) a |= ((mask32 >> (8 * nt + 2 * q + h)) & 1u) << (2 * nt + h);
    return a;
}

// ===========================================================================
// Batched ADMM: S (B x nb) unconstrained box rows -> T (B x nb) corrections, iterations, status.
// ===========================================================================
__global__ void __launch_bounds__(128, 2)
k_admm_dmma(int B, int nb, const double *__restrict__ S, const double *__restrict__ Phi, const double *__restrict__ blo,
            const double *__restrict__ bhi, const double *__restrict__ bmax, double tol, int max_iter,
            double *__restrict__ T, int *__restrict__ iters_out, int *__restrict__ status_out) {
    extern __shared__ __align__(16) unsigned char cv_raw[];
    AdmmSmem &sm = *reinterpret_cast<AdmmSmem *>(cv_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int b0 = blockIdx.x * CV_NL, ks_n = (nb + 3) / 4;
    PhiRegs aPhi;
    double lo[2], hi[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r = 16 * warp + 8 * mt + g;
        lo[mt] = r < nb ? blo[r] : -INFINITY;
        hi[mt] = r < nb ? bhi[r] : INFINITY;
    }
    aPhi.load(Phi, nb, warp, g, q);
    double2 su[2][CV_NT];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) {
            const int r = 16 * warp + 8 * mt + g, p0 = b0 + 8 * nt + 2 * q;
            su[mt][nt].x = (r < nb && p0 < B) ? S[(size_t)p0 * nb + r] : 0.0;
            su[mt][nt].y = (r < nb && p0 + 1 < B) ? S[(size_t)(p0 + 1) * nb + r] : 0.0;
            sm.su[mt * CV_NT + nt][tid] = su[mt][nt];
        }
    unsigned viol, bad;
    box_flags(su, lo, hi, q, viol, bad);
    if (lane == 0) { sm.vmask[0][warp] = viol; sm.bmask[0][warp] = bad; }
    admm_thresholds(su, tol, bmax[0], sm);                           // (two barriers: the masks are visible after them)
    const unsigned vall = sm.vmask[0][0] | sm.vmask[0][1] | sm.vmask[0][2] | sm.vmask[0][3];
    const unsigned ball = sm.bmask[0][0] | sm.bmask[0][1] | sm.bmask[0][2] | sm.bmask[0][3];
    const unsigned act = lane_bits(vall & ~ball, q);
    double2 t[2][CV_NT];
    int iters[CV_NT][2];
    unsigned inacc;
    admm_cta(aPhi, lo, hi, act, ks_n, max_iter, sm, t, iters, inacc);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) {
            const int r = 16 * warp + 8 * mt + g, p0 = b0 + 8 * nt + 2 * q;
            if (r < nb && p0 < B) T[(size_t)p0 * nb + r] = t[mt][nt].x;
            if (r < nb && p0 + 1 < B) T[(size_t)(p0 + 1) * nb + r] = t[mt][nt].y;
        }
    if (warp == 0 && g == 0) {
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = b0 + 8 * nt + 2 * q + h;
                if (p >= B) continue;
                if (iters_out) iters_out[p] = iters[nt][h];
                if (status_out && ((inacc >> (2 * nt + h)) & 1u))
                    status_out[p] = max(status_out[p], (int)DDMPC_SOLVE_OPTIMAL_INACCURATE);
            }
    }
}

// theta_b = [u_past; y_past; u_s; y_s] (B x nth), status <- finite ? optimal : non-finite, iters <- 1
__global__ void k_pack_theta(int B, int nm, int npp, int m, int p, const double *__restrict__ u_past,
                             const double *__restrict__ y_past, const double *__restrict__ u_s, const double *__restrict__ y_s,
                             double *__restrict__ Th, int *__restrict__ status, int *__restrict__ iters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nth = nm + npp + m + p;
    bool fin = true;
    double *row = Th + (size_t)b * nth;
    for (int i = 0; i < nth; ++i) {
        double v;
        if (i < nm) v = u_past[(size_t)b * nm + i];
        else if (i < nm + npp) v = y_past[(size_t)b * npp + (i - nm)];
        else if (i < nm + npp + m) v = u_s[(size_t)b * m + (i - nm - npp)];
        else v = y_s[(size_t)b * p + (i - nm - npp - m)];
        row[i] = v;
        fin = fin && isfinite(v);
    }
    status[b] = fin ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
    iters[b] = 1;
}

// cost_b = theta_b . (Z theta_b) + rho2^2 t_b . (Lam t_b)
__global__ void k_cost_rows(int B, int nth, int nb, const double *__restrict__ Th, const double *__restrict__ ZT,
                            const double *__restrict__ T, const double *__restrict__ LT, const double *__restrict__ rho2,
                            double *__restrict__ cost) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double j0 = 0.0, j1 = 0.0;
    for (int i = 0; i < nth; ++i) j0 = fma(Th[(size_t)b * nth + i], ZT[(size_t)b * nth + i], j0);
    for (int i = 0; i < nb; ++i) j1 = fma(T[(size_t)b * nb + i], LT[(size_t)b * nb + i], j1);
    const double r = rho2[0];
    cost[b] = j0 + r * r * j1;
}

static int admm_smem_attr(const void *kern, size_t bytes) {
    DDMPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    DDMPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return DDMPC_OK;
}

// Batched solve for ONE shared CONVEX controller on the tensor cores.  Returns -1 when this path does not apply.
int solve_batch_cvx_dmma(const ddmpc_set *set, int B, const int *ctrl_idx, const double *u_past, const double *y_past,
                         const double *u_s, const double *y_s, double tol, int max_iter, double *optimal_u, double *cost,
                         int *status, int *iters, double *t_out, cudaStream_t st) {
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    if (ctrl_idx || pl.count != 1 || !d.robust || d.nb <= 0 || d.nb > CV_NR || d.nbu > 0 || d.nby > 0 || B < 1024) return -1;
    ScratchStreamScope scope(st);
    DevBuf th, S, T, zt, lt, stb, itb;
    const int nth = d.nth, nb = d.nb, Lm = d.Lm;
    DDMPC_CUDA(th.alloc(sizeof(double) * (size_t)B * nth));
    DDMPC_CUDA(S.alloc(sizeof(double) * (size_t)B * nb));
    if (!t_out) { DDMPC_CUDA(T.alloc(sizeof(double) * (size_t)B * nb)); t_out = T.d(); }
    if (!status) { DDMPC_CUDA(stb.alloc(sizeof(int) * (size_t)B)); status = stb.i(); }
    if (!iters) { DDMPC_CUDA(itb.alloc(sizeof(int) * (size_t)B)); iters = itb.i(); }
    k_pack_theta<<<ceil_div(B, 256), 256, 0, st>>>(B, d.n * d.m, d.n * d.p, d.m, d.p, u_past, y_past, u_s, y_s, th.d(), status, iters);
    DDMPC_LAUNCH_CHECK();
    const Mat Th = mat(th.d(), nth, 1, 0);
    // u0 = Theta Ku^T,  s_unc = Theta Ks^T          (Ku (Lm x nth), Ks (nb x nth) row-major: their transposes have rs = 1)
    DDMPC_TRY(gemm(st, 1, B, Lm, nth, 1.0, Th, mat(pl.Ku.d(), 1, nth, 0), 0.0, optimal_u, Lm, 1, 0));
    DDMPC_TRY(gemm(st, 1, B, nb, nth, 1.0, Th, mat(pl.Ks.d(), 1, nth, 0), 0.0, S.d(), nb, 1, 0));
    static std::atomic<unsigned long long> attr_done{0};
    if (first_time_on_device(attr_done)) DDMPC_TRY(admm_smem_attr((const void *)k_admm_dmma, sizeof(AdmmSmem)));
    k_admm_dmma<<<ceil_div(B, CV_NL), 128, sizeof(AdmmSmem), st>>>(B, nb, S.d(), pl.Phi.d(), pl.lo.d(), pl.hi.d(), pl.bmax.d(),
                                                                    tol > 0.0 ? tol : 1e-8, max_iter > 0 ? max_iter : 1000,
                                                                    t_out, iters, status);
    DDMPC_LAUNCH_CHECK();
    // u = u0 - T Psi^T
    DDMPC_TRY(gemm(st, 1, B, Lm, nb, -1.0, mat(t_out, nb, 1, 0), mat(pl.Psi.d(), 1, nb, 0), 1.0, optimal_u, Lm, 1, 0));
    if (cost) {
        DDMPC_CUDA(zt.alloc(sizeof(double) * (size_t)B * nth));
        DDMPC_CUDA(lt.alloc(sizeof(double) * (size_t)B * nb));
        DDMPC_TRY(gemm(st, 1, B, nth, nth, 1.0, Th, mat(pl.Z.d(), 1, nth, 0), 0.0, zt.d(), nth, 1, 0));
        DDMPC_TRY(gemm(st, 1, B, nb, nb, 1.0, mat(t_out, nb, 1, 0), mat(pl.Lam.d(), 1, nb, 0), 0.0, lt.d(), nb, 1, 0));
        k_cost_rows<<<ceil_div(B, 256), 256, 0, st>>>(B, nth, nb, th.d(), zt.d(), t_out, lt.d(), pl.rho2.d(), cost);
        DDMPC_LAUNCH_CHECK();
    }
    return DDMPC_OK;
}

// ===========================================================================
// Fused closed loop, shared CONVEX controller, four-tank n-step shape (n = 4, m = p = 2, n_x = 4, n_mpc_step = 4).
// ===========================================================================
struct CvxMaps {
    double Mb[12][12];   // 4-step block map of the plant: rows y_0..y_3 (8), x_4 (4); columns x_0 (4), u_0..u_3 (8)
    double Mt[12][12];   // block map of the last, partial block
};

struct CvxArgs {
    int B, n_steps, nb, nth, n_tail, max_iter;
    const double *Ku, *Ks, *Phi, *Phi64, *Psi, *blo, *bhi, *bmax;
    const double *x0, *u_past0, *y_past0, *u_s, *y_s, *w;
    unsigned long long id0;
    double eps, tol;
    double *u_sys, *y_sys, *x_final;
    int *status, *iters;
    uint32_t rk[20];
};

constexpr int CV_FS = 40;      // row stride of the FP32 window copy (rows q = 0..3 land on disjoint banks)
struct CvxSmem {
    double th[2][16][CV_DS];   // measurement window [U (8 rows); Y (8 rows)] x 32 loops, by block parity
    double sp[4][CV_DS];       // set-points [u_s; y_s]
    double x[4][CV_DS];        // plant state
    float thf[2][16][CV_FS];   // FP32 copy of the windows: B operand of the TF32 screen
    double cmax[16];           // level-0 screen: max_j |Ks[j][k]| over the slack rows, per window entry k
    double gss[CV_NR][4];      // level-0 screen: slack rows at a set-point, s_ss = gss [u_s; y_s]
    double smax_ss[CV_NL];     // level-0 screen: max_j |s_ss,j| of each loop
    unsigned qflag[2][4];      // level-0 verdict per warp for the window of that parity (0 = provably inside the box)
    AdmmSmem admm;
};

__device__ __forceinline__ void cv_philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}
__device__ __forceinline__ double cv_unit32(uint32_t x) {
    return __hiloint2double((int)(0x3FF00000u | (x >> 12)), (int)(x << 20));
}

// TF32 tensor-core product of the slack-row screen (mma.sync m16n8k8; the operands' low 13 mantissa bits are ignored).
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
constexpr float kScreenRel = 1.0f / 256.0f;
constexpr uint32_t kAbs = 0x7fffffffu;

// MINB = CTAs per SM the register allocation is held to.  3 (168 registers, the default: option "cvx_ctas_per_sm"): the
// ADMM section streams Phi from L1 instead of holding it in 64 registers; 2 (255 registers): Phi in registers - the
// steady state is 12 % slower, runs dominated by ADMM iterations (c = 0.3) are 7 % faster.
template <bool PHILOX, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_closed_loop_cvx(const __grid_constant__ CvxMaps maps, const CvxArgs a) {
    constexpr int NW = 16, NSPK = 5;                                 // window rows; k-steps of the 20 theta entries
    extern __shared__ __align__(16) unsigned char cv_raw[];
    CvxSmem &sm = *reinterpret_cast<CvxSmem *>(cv_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int nb = a.nb, ks_n = (nb + 3) / 4;
    // ---- i/o role of a lane: loop 8 warp + g of the CTA, step q of every block
    const int lcol = 8 * warp + g;
    int b = blockIdx.x * CV_NL + lcol;
    const bool live = b < a.B;
    if (!live) b = a.B - 1;                                          // dead slots replay the last loop and never store
    const size_t f0 = (size_t)b * a.n_steps;
    const unsigned long long sid = a.id0 + (unsigned long long)b;
    const uint32_t sid_lo = (uint32_t)sid, sid_hi = (uint32_t)(sid >> 32);
    {
        const double2 up = *reinterpret_cast<const double2 *>(a.u_past0 + (size_t)b * 8 + 2 * q);
        const double2 yp = *reinterpret_cast<const double2 *>(a.y_past0 + (size_t)b * 8 + 2 * q);
        sm.th[0][2 * q][lcol] = up.x; sm.th[0][2 * q + 1][lcol] = up.y;
        sm.th[0][8 + 2 * q][lcol] = yp.x; sm.th[0][8 + 2 * q + 1][lcol] = yp.y;
        sm.thf[0][2 * q][lcol] = (float)up.x; sm.thf[0][2 * q + 1][lcol] = (float)up.y;
        sm.thf[0][8 + 2 * q][lcol] = (float)yp.x; sm.thf[0][8 + 2 * q + 1][lcol] = (float)yp.y;
        sm.x[q][lcol] = a.x0[(size_t)b * 4 + q];
        sm.sp[q][lcol] = q < 2 ? a.u_s[(size_t)b * 2 + q] : a.y_s[(size_t)b * 2 + (q - 2)];
    }
    if (tid < CV_NL) { sm.admm.extra[tid] = 0; sm.admm.stat[tid] = DDMPC_SOLVE_OPTIMAL; }
    // ---- level-0 screen (see the block loop): column maxima of |Ks| over the window entries, and the slack rows at a
    //      set-point as a 4-column map of [u_s; y_s] (the window of a settled loop is its set-point repeated)
    if (tid < 16) {
        double m = 0.0;
        for (int j = 0; j < a.nb; ++j) m = fmax(m, fabs(__ldg(a.Ks + (size_t)j * a.nth + tid)));
        sm.cmax[tid] = m;
    }
    for (int e = tid; e < CV_NR * 4; e += 128) {
        const int j = e >> 2, i = e & 3;
        double gsum = 0.0;
        if (j < a.nb) {
            const double *row = a.Ks + (size_t)j * a.nth;
            const int base = i < 2 ? i : 8 + (i - 2);                 // window entries of channel i: base + 2 t, t = 0..3
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) gsum += __ldg(row + base + 2 * tt);
            gsum += __ldg(row + 16 + i);                             // its set-point column
        }
        sm.gss[j][i] = gsum;
    }
    __syncthreads();
    {
        // the four lanes of a loop share its 64 rows
        const double u0 = sm.sp[0][lcol], u1 = sm.sp[1][lcol], y0 = sm.sp[2][lcol], y1 = sm.sp[3][lcol];
        double m = 0.0;
        for (int j = q; j < CV_NR; j += 4)
            m = fmax(m, fabs(sm.gss[j][0] * u0 + sm.gss[j][1] * u1 + sm.gss[j][2] * y0 + sm.gss[j][3] * y1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
        if (q == 0) sm.smax_ss[lcol] = m;
    }
    __syncwarp();
    // Level-0 screen of the window in buffer `par` (a compile-time constant at every call site).  For every slack row j:
    //   |s_j(theta)| = |s_ss,j + sum_k Ks[j][k] (theta_k - theta_ss,k)| <= max_j |s_ss,j| + sum_k cmax[k] |theta_k - theta_ss,k|
    // with theta_ss the window of the settled loop (its set-point repeated; the set-point entries of theta cancel).  Twenty
    // multiply-adds per loop prove "inside the box" for a settled loop, where the TF32 screen costs 24 tensor-core MMAs per
    // warp and the exact check 40 FP64 ones.  Lane (g, q) holds window step q of loop 8 warp + g; NaN fails the test.
    const double thr0 = a.bhi[0] * (1.0 - 1e-9);
    auto level0 = [&](auto par_c) {
        constexpr int par = decltype(par_c)::value;
        const double d = sm.cmax[2 * q] * fabs(sm.th[par][2 * q][lcol] - sm.sp[0][lcol]) +
                         sm.cmax[2 * q + 1] * fabs(sm.th[par][2 * q + 1][lcol] - sm.sp[1][lcol]) +
                         sm.cmax[8 + 2 * q] * fabs(sm.th[par][8 + 2 * q][lcol] - sm.sp[2][lcol]) +
                         sm.cmax[8 + 2 * q + 1] * fabs(sm.th[par][8 + 2 * q + 1][lcol] - sm.sp[3][lcol]);
        double tot = d + __shfl_xor_sync(0xffffffffu, d, 1);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        const bool fail = !(sm.smax_ss[lcol] + tot <= thr0);
        const unsigned any = __any_sync(0xffffffffu, fail) ? 1u : 0u;
        if (lane == 0) sm.qflag[par][warp] = any;
    };
    level0(std::integral_constant<int, 0>{});
    // ---- A fragments that stay in registers: gain rows (8 x 20), slack rows of this warp (16 x 20), plant block map
    double aKu[NSPK], aMb[2][3];
    uint32_t fKs[2][3][2], spb[CV_NT];                               // TF32 screen: slack rows, set-points
    // (a lambda: the fragments are loaded again after an ADMM solve, so that they need not stay in registers - or be
    //  spilled and re-read in every block - across the ADMM section, which takes all 255 registers)
    auto load_frags = [&]() {
#pragma unroll
        for (int ks = 0; ks < NSPK; ++ks) aKu[ks] = __ldg(a.Ku + (size_t)g * a.nth + 4 * ks + q);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int r = 16 * warp + 8 * mt + g;
#pragma unroll
            for (int s8 = 0; s8 < 3; ++s8)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = 8 * s8 + q + 4 * h;
                    fKs[mt][s8][h] = to_tf32((r < nb && c < a.nth) ? (float)__ldg(a.Ks + (size_t)r * a.nth + c) : 0.f);
                }
        }
#pragma unroll
        for (int nt = 0; nt < CV_NT; ++nt) {
            const int bb = min(blockIdx.x * CV_NL + 8 * nt + g, a.B - 1);
            spb[nt] = __float_as_uint((float)(q < 2 ? a.u_s[(size_t)bb * 2 + q] : a.y_s[(size_t)bb * 2 + (q - 2)]));
        }
    };
    load_frags();
    auto load_map = [&](const double (&Mm)[12][12]) {
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) aMb[rt][ks] = (8 * rt + g < 12) ? Mm[(8 * rt + g) % 12][4 * ks + q] : 0.0;
    };
    load_map(maps.Mb);
    const int nblk = (a.n_steps + 3) / 4;
    uint32_t pw0 = 0u, pw1 = 0u, pw2 = 0u, pw3 = 0u;                 // Philox words of the current pair of blocks
    // pure CONVEX: every box row is [-c eps_max, c eps_max]; the screen compares against the bound rounded towards zero
    const unsigned bound_bits = __float_as_uint(__double2float_rz(a.bhi[0]));
    // One MPC iteration.  The window buffers alternate with the block parity, which is a compile-time constant here (the
    // loop below is unrolled by two) so that every shared-memory address is a fixed offset.
    auto block = [&](auto parity, const int t) {
        constexpr int cur = decltype(parity)::value, nxt = cur ^ 1;
        if (t == nblk - 1 && a.n_tail != 0) load_map(maps.Mt);       // last, partial block (controller_operation.py:278)
        __syncthreads();                                             // window / state / set-points of all 32 loops are in place
        // level-0 verdicts of the four warps for this window (written at the end of the previous block)
        const bool quiet = (sm.qflag[cur][0] | sm.qflag[cur][1] | sm.qflag[cur][2] | sm.qflag[cur][3]) == 0u;
        // ---- measurement noise of (loop, step q): parked in the output rows of the next window, where the plant
        //      product's accumulators start from it
        {
            double n0, n1;
            if constexpr (PHILOX) {
                // noise word (4 t + q) * 2 + i is word 2 (q & 1) + i of Philox call 2 t + (q >> 1)   (solve.cu).  The four
                // lanes of a loop draw the four calls of a PAIR of blocks once (lane q: call 2 t + q, t even) and hand
                // the words round.
                if constexpr (cur == 0) {
                    pw0 = 2u * (unsigned)t + (unsigned)q; pw1 = 0u; pw2 = sid_lo; pw3 = sid_hi;
#pragma unroll
                    for (int r = 0; r < 10; ++r) cv_philox_round(pw0, pw1, pw2, pw3, a.rk[2 * r], a.rk[2 * r + 1]);
                }
                const int src = (lane & ~3) | (2 * cur + (q >> 1));
                const uint32_t v0 = __shfl_sync(0xffffffffu, pw0, src), v1 = __shfl_sync(0xffffffffu, pw1, src);
                const uint32_t v2 = __shfl_sync(0xffffffffu, pw2, src), v3 = __shfl_sync(0xffffffffu, pw3, src);
                n0 = a.eps * (2.0 * cv_unit32((q & 1) ? v2 : v0) - 3.0);
                n1 = a.eps * (2.0 * cv_unit32((q & 1) ? v3 : v1) - 3.0);
            } else {
                const int k = 4 * t + q;
                n0 = k < a.n_steps ? __ldg(a.w + (f0 + k) * 2) : 0.0;
                n1 = k < a.n_steps ? __ldg(a.w + (f0 + k) * 2 + 1) : 0.0;
            }
            sm.th[nxt][8 + 2 * q][lcol] = n0;
            sm.th[nxt][8 + 2 * q + 1][lcol] = n1;
        }
        // ---- gain product U = Ku theta (own n-tile), FP64
        double2 cu = make_double2(0.0, 0.0);
#pragma unroll
        for (int ks = 0; ks < NSPK; ++ks) {
            const double *row = ks < NW / 4 ? sm.th[cur][4 * ks + q] : sm.sp[q];
            dmma884(cu, aKu[ks], row[8 * warp + g]);
        }
        // ---- screening of the slack rows on the TF32 tensor cores (this warp's 16 rows x 32 loops): s~ = Ks theta and
        //      A = |Ks| |theta| in one pass each; |s - s~| <= kScreenRel A (operand rounding 2^-11 + truncation 2^-10, FP32
        //      accumulation; kScreenRel = 2^-8 leaves a factor two), so  |s~| + kScreenRel A <= bound  proves that no
        //      row of this fragment leaves the box.  NaN/Inf anywhere in a window compare as "suspicious" (integer
        //      compare of the bit patterns), and the exact FP64 check below decides.
        unsigned act32 = 0u;
        double2 su[2][CV_NT];
        double lo[2], hi[2];
        if (!quiet) {                                                // CTA-uniform
            float cs[CV_NT][4], ca[CV_NT][4];
#pragma unroll
            for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) cs[nt][i] = ca[nt][i] = 0.f;
#pragma unroll
            for (int s8 = 0; s8 < 3; ++s8)
#pragma unroll
                for (int nt = 0; nt < CV_NT; ++nt) {
                    uint32_t b0, b1;
                    if (s8 < 2) {
                        b0 = __float_as_uint(sm.thf[cur][8 * s8 + q][8 * nt + g]);
                        b1 = __float_as_uint(sm.thf[cur][8 * s8 + 4 + q][8 * nt + g]);
                    } else {
                        b0 = spb[nt];
                        b1 = 0u;
                    }
                    mma_tf32(cs[nt], fKs[0][s8][0], fKs[1][s8][0], fKs[0][s8][1], fKs[1][s8][1], b0, b1);
                    mma_tf32(ca[nt], fKs[0][s8][0] & kAbs, fKs[1][s8][0] & kAbs, fKs[0][s8][1] & kAbs, fKs[1][s8][1] & kAbs,
                             b0 & kAbs, b1 & kAbs);
                }
            unsigned worst = 0u;
#pragma unroll
            for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) worst = max(worst, __float_as_uint(fmaf(kScreenRel, ca[nt][i], fabsf(cs[nt][i]))));
            const bool susp = __any_sync(0xffffffffu, worst > bound_bits);
            if (lane == 0) sm.admm.vmask[cur][warp] = susp ? 1u : 0u;
            __syncthreads();
            if ((sm.admm.vmask[cur][0] | sm.admm.vmask[cur][1] | sm.admm.vmask[cur][2] | sm.admm.vmask[cur][3]) != 0u) {
                // rare (first blocks after a set-point change): the exact FP64 check - which loops violate, which hold
                // non-finite numbers
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < CV_NT; ++nt) su[mt][nt] = make_double2(0.0, 0.0);
#pragma unroll
                for (int ks = 0; ks < NSPK; ++ks) {
                    const double *row = ks < NW / 4 ? sm.th[cur][4 * ks + q] : sm.sp[q];
                    double bv[CV_NT], av[2];
#pragma unroll
                    for (int nt = 0; nt < CV_NT; ++nt) bv[nt] = row[8 * nt + g];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const int r = 16 * warp + 8 * mt + g;
                        av[mt] = r < nb ? __ldg(a.Ks + (size_t)r * a.nth + 4 * ks + q) : 0.0;
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < CV_NT; ++nt) dmma884(su[mt][nt], av[mt], bv[nt]);
                }
                for (int mt = 0; mt < 2; ++mt) {
                    const int r = 16 * warp + 8 * mt + g;
                    lo[mt] = r < nb ? __ldg(a.blo + r) : -INFINITY;
                    hi[mt] = r < nb ? __ldg(a.bhi + r) : INFINITY;
                }
                unsigned viol, bad;
                box_flags(su, lo, hi, q, viol, bad);
                if (lane == 0) { sm.admm.vmask[nxt][warp] = viol; sm.admm.bmask[cur][warp] = bad; }
                __syncthreads();
                const unsigned vall = sm.admm.vmask[nxt][0] | sm.admm.vmask[nxt][1] | sm.admm.vmask[nxt][2] | sm.admm.vmask[nxt][3];
                const unsigned ball = sm.admm.bmask[cur][0] | sm.admm.bmask[cur][1] | sm.admm.bmask[cur][2] | sm.admm.bmask[cur][3];
                act32 = vall & ~ball;
            }
        }
        if (act32 != 0u) {                                           // CTA-uniform: some loop violates its slack bound
            // ---- box-row ADMM on the tensor cores, then the correction of the planned inputs  U -= Psi[0:8, :] t
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < CV_NT; ++nt) sm.admm.su[mt * CV_NT + nt][tid] = su[mt][nt];
            admm_thresholds(su, a.tol, a.bmax[0], sm.admm);
            std::conditional_t<(MINB > 2), PhiStream, PhiRegs> aPhi;
            if constexpr (MINB > 2) aPhi.load(a.Phi64, warp, g, q);
            else aPhi.load(a.Phi, nb, warp, g, q);
            double2 tt[2][CV_NT];
            int its[CV_NT][2];
            unsigned inacc;
            admm_cta(aPhi, lo, hi, lane_bits(act32, q), ks_n, a.max_iter, sm.admm, tt, its, inacc);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < CV_NT; ++nt)
                    *reinterpret_cast<double2 *>(&sm.admm.d[0][16 * warp + 8 * mt + g][8 * nt + 2 * q]) = tt[mt][nt];
            if (warp == 0 && g == 0) {
#pragma unroll
                for (int nt = 0; nt < CV_NT; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int p = 8 * nt + 2 * q + h;
                        sm.admm.extra[p] += its[nt][h] - 1;
                        if ((inacc >> (2 * nt + h)) & 1u) sm.admm.stat[p] = DDMPC_SOLVE_OPTIMAL_INACCURATE;
                    }
            }
            __syncthreads();
            double2 corr = make_double2(0.0, 0.0);
            for (int ks = 0; ks < ks_n; ++ks) {
                const int c = 4 * ks + q;
                const double av = c < nb ? __ldg(a.Psi + (size_t)g * nb + c) : 0.0;
                dmma884(corr, av, sm.admm.d[0][4 * ks + q][8 * warp + g]);
            }
            cu.x -= corr.x;
            cu.y -= corr.y;
            load_frags();
            if (t == nblk - 1 && a.n_tail != 0) load_map(maps.Mt);
            else load_map(maps.Mb);
            // (sm.admm.d is next written after the barrier at the top of a later block)
        }
        // ---- planned inputs -> input rows of the next window (own n-tile)
        *reinterpret_cast<double2 *>(&sm.th[nxt][g][8 * warp + 2 * q]) = cu;
        *reinterpret_cast<float2 *>(&sm.thf[nxt][g][8 * warp + 2 * q]) = make_float2((float)cu.x, (float)cu.y);
        __syncwarp();
        // ---- plant: [Y; x+] = Mblk [x; U], the outputs on top of the noise
        double2 d0 = *reinterpret_cast<const double2 *>(&sm.th[nxt][8 + g][8 * warp + 2 * q]), d1 = make_double2(0.0, 0.0);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            const double bv = ks == 0 ? sm.x[q][8 * warp + g] : sm.th[nxt][4 * (ks - 1) + q][8 * warp + g];
            dmma884(d0, aMb[0][ks], bv);
            dmma884(d1, aMb[1][ks], bv);
        }
        __syncwarp();                                                // every lane has read the old state
        *reinterpret_cast<double2 *>(&sm.th[nxt][8 + g][8 * warp + 2 * q]) = d0;
        *reinterpret_cast<float2 *>(&sm.thf[nxt][8 + g][8 * warp + 2 * q]) = make_float2((float)d0.x, (float)d0.y);
        if (g < 4) *reinterpret_cast<double2 *>(&sm.x[g][8 * warp + 2 * q]) = d1;
        __syncwarp();
        // ---- record step q of loop 8 warp + g: 16 bytes per array; the four lanes of a loop write 64 contiguous bytes
        const int k = 4 * t + q;
        if (live && k < a.n_steps) {
            *reinterpret_cast<double2 *>(a.u_sys + (f0 + k) * 2) = make_double2(sm.th[nxt][2 * q][lcol], sm.th[nxt][2 * q + 1][lcol]);
            *reinterpret_cast<double2 *>(a.y_sys + (f0 + k) * 2) = make_double2(sm.th[nxt][8 + 2 * q][lcol], sm.th[nxt][8 + 2 * q + 1][lcol]);
        }
        level0(std::integral_constant<int, nxt>{});                  // verdict for the next block's window
    };
    {
        int t = 0;
        for (; t + 1 < nblk; t += 2) {
            block(std::integral_constant<int, 0>{}, t);
            block(std::integral_constant<int, 1>{}, t + 1);
        }
        if (t < nblk) block(std::integral_constant<int, 0>{}, t);
    }
    __syncthreads();
    // ---- verdict of loop 8 warp + g (lanes q = 0..3 hold one state entry and one step of the last block each)
    const int fin_buf = nblk & 1, last_steps = a.n_steps - 4 * (nblk - 1);
    bool fin = isfinite(sm.x[q][lcol]);
    if (q < last_steps) fin = fin && isfinite(sm.th[fin_buf][8 + 2 * q][lcol]) && isfinite(sm.th[fin_buf][8 + 2 * q + 1][lcol]);
    const unsigned badm = __ballot_sync(0xffffffffu, !fin);
    const bool loop_bad = ((badm >> (4 * g)) & 0xfu) != 0u;
    if (live) {
        if (q == 0) {
            if (a.status) a.status[b] = loop_bad ? (int)DDMPC_SOLVE_NONFINITE : sm.admm.stat[lcol];
            if (a.iters) a.iters[b] = nblk + sm.admm.extra[lcol];
        }
        if (a.x_final) a.x_final[(size_t)b * 4 + q] = sm.x[q][lcol];
    }
}

// Set creation: the zero-padded copy of Phi the 3-CTAs-per-SM build streams its A operand from.
int closed_loop_cvx_prepare(ddmpc_set *set, cudaStream_t st) {
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    if (pl.count != 1 || !d.robust || !d.convex || d.nb <= 0 || d.nb > CV_NR) return DDMPC_OK;
    ScratchStreamScope scope(st);
    DDMPC_CUDA(set->cvx_phi64.alloc(sizeof(double) * CV_NR * CV_NR));
    DDMPC_CUDA(cudaMemsetAsync(set->cvx_phi64.p, 0, sizeof(double) * CV_NR * CV_NR, st));
    DDMPC_CUDA(cudaMemcpy2DAsync(set->cvx_phi64.p, sizeof(double) * CV_NR, pl.Phi.p, sizeof(double) * d.nb,
                                 sizeof(double) * d.nb, d.nb, cudaMemcpyDeviceToDevice, st));
    return DDMPC_OK;
}

// Returns DDMPC_OK when handled, -1 when this path does not apply.
int closed_loop_cvx_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                        const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                        const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                        double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                        cudaStream_t st) {
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    if (ctrl_idx || pl.count != 1 || !d.robust || !d.convex || d.nbu > 0 || d.nby > 0 || d.nb <= 0 || d.nb > CV_NR) return -1;
    if (set->opt_path != DDMPC_PATH_AUTO && set->opt_path != DDMPC_PATH_CVX) return -1;
    if (!(d.n == 4 && d.m == 2 && d.p == 2 && plant->n_x == 4 && set->prm.n_mpc_step == 4 && d.nth == 20)) return -1;
    // 16-byte vector accesses on the windows and the trajectories
    if ((reinterpret_cast<uintptr_t>(u_past0) | reinterpret_cast<uintptr_t>(y_past0) | reinterpret_cast<uintptr_t>(u_sys) |
         reinterpret_cast<uintptr_t>(y_sys)) & 15) return -1;
    CvxMaps maps;
    const std::vector<double> Mb = block_map(plant, 4);
    for (int r = 0; r < 12; ++r)
        for (int j = 0; j < 12; ++j) maps.Mb[r][j] = maps.Mt[r][j] = Mb[(size_t)r * 12 + j];
    const int rem = n_steps % 4;
    if (rem) {
        const std::vector<double> Mr = block_map(plant, rem);
        const int cr = 4 + rem * 2;
        for (int r = 0; r < 12; ++r)
            for (int j = 0; j < 12; ++j) maps.Mt[r][j] = 0.0;
        for (int r = 0; r < rem * 2; ++r)
            for (int j = 0; j < cr; ++j) maps.Mt[r][j] = Mr[(size_t)r * cr + j];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < cr; ++j) maps.Mt[8 + i][j] = Mr[(size_t)(rem * 2 + i) * cr + j];
    }
    CvxArgs a{};
    a.B = B; a.n_steps = n_steps; a.nb = d.nb; a.nth = d.nth; a.n_tail = rem;
    a.max_iter = max_iter > 0 ? max_iter : 1000;
    a.tol = tol > 0.0 ? tol : 1e-8;
    a.Ku = pl.Ku.d(); a.Ks = pl.Ks.d(); a.Phi = pl.Phi.d(); a.Phi64 = set->cvx_phi64.d(); a.Psi = pl.Psi.d();
    a.blo = pl.lo.d(); a.bhi = pl.hi.d(); a.bmax = pl.bmax.d();
    a.x0 = x0; a.u_past0 = u_past0; a.y_past0 = y_past0; a.u_s = u_s; a.y_s = y_s; a.w = w;
    a.id0 = id0; a.eps = eps;
    a.u_sys = u_sys; a.y_sys = y_sys; a.x_final = x_final; a.status = status; a.iters = iters;
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    static std::atomic<unsigned long long> attr_done{0};
    if (first_time_on_device(attr_done)) {
        DDMPC_TRY(admm_smem_attr((const void *)k_closed_loop_cvx<true, 2>, sizeof(CvxSmem)));
        DDMPC_TRY(admm_smem_attr((const void *)k_closed_loop_cvx<false, 2>, sizeof(CvxSmem)));
        DDMPC_TRY(admm_smem_attr((const void *)k_closed_loop_cvx<true, 3>, sizeof(CvxSmem)));
        DDMPC_TRY(admm_smem_attr((const void *)k_closed_loop_cvx<false, 3>, sizeof(CvxSmem)));
    }
    const int grid = ceil_div(B, CV_NL);
    if (set->opt_cvx_ctas == 2) {
        if (w) k_closed_loop_cvx<false, 2><<<grid, 128, sizeof(CvxSmem), st>>>(maps, a);
        else k_closed_loop_cvx<true, 2><<<grid, 128, sizeof(CvxSmem), st>>>(maps, a);
    } else {
        if (w) k_closed_loop_cvx<false, 3><<<grid, 128, sizeof(CvxSmem), st>>>(maps, a);
        else k_closed_loop_cvx<true, 3><<<grid, 128, sizeof(CvxSmem), st>>>(maps, a);
    }
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

}  // namespace ddmpc
