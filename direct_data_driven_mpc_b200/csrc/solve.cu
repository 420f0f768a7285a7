// Batched DD-MPC QP solve and fused closed loop (generic, runtime-sized path).
//
// One thread owns one scenario; its state lives in shared memory laid out
// [variable][thread] (conflict-free), the per-controller operators are read
// through L1/L2 (warp-uniform addresses when a warp shares a controller).
//
// Replaces, per scenario: update_and_solve_data_driven_mpc
// (direct_data_driven_mpc_controller.py:389-407 -> cvxpy solve :753),
// get_optimal_control_input[_at_step] (:780-842), store_input_output_measurement
// (:844-895), LTIModel.simulate_step (utilities/model_simulation.py:93-98) and the
// loop of simulate_data_driven_mpc_control_loop
// (utilities/controller/controller_operation.py:263-305).
#include <algorithm>
#include <vector>

#include "linalg.cuh"
#include "plan.cuh"

namespace ddmpc {

struct KArgs {
    int n, m, p, L, nth, nb, Lm, nx, nfix, r, cols, ny, nu;
    int convex, robust;
    const double *Ku, *Z, *Ks, *Phi, *Psi, *Lam, *rho2, *F, *X0, *Yf;
    const int *Fnz;                  // nominal: per controller, 1 when F is not identically zero
    const double *lo, *hi, *bmax;    // per controller: scaled bounds of the box rows (nb), largest finite |bound|
    const double *umin, *umax;       // input box (m) or NULL
    const double *ymin, *ymax;       // output box (p) or NULL
    int terminal;
    double tol;
    int max_iter;
};

// Terminal equality + input / output box: the last n predicted inputs (outputs) are fixed to u_s (y_s), so the QP is
// infeasible when the set-point itself violates the box (cvxpy would report "infeasible").
__device__ __forceinline__ bool setpoint_outside_box(const KArgs &a, const double *th, int TS, int tid) {
    if (!a.terminal || (!a.umin && !a.ymin)) return false;
    const int o = a.n * (a.m + a.p);
    bool bad = false;
    for (int j = 0; j < a.m && a.umin; ++j) {
        const double v = th[(size_t)(o + j) * TS + tid];
        const double sl = 1e-9 * (1.0 + fabs(v));
        bad = bad || v < a.umin[j] - sl || v > a.umax[j] + sl;
    }
    for (int j = 0; j < a.p && a.ymin; ++j) {
        const double v = th[(size_t)(o + a.m + j) * TS + tid];
        const double sl = 1e-9 * (1.0 + fabs(v));
        bad = bad || v < a.ymin[j] - sl || v > a.ymax[j] + sl;
    }
    return bad;
}

#define SMV(arr, i) (arr)[(size_t)(i) * TS + tid]

// ---- ADMM on the box rows --------------------------------------------------
// In : s_unc (unconstrained slack rows), smax = max|s_unc| > bound.
// Out: s_unc overwritten by t = Phi d at the fixed point, so that
//      u = u0 - Psi t,  x = x0 - Yf t,  cost += rho2^2 t^T Lam t.
// Iteration (DESIGN.md "ADMM on the condensed box rows"), over-relaxed with alpha = DDMPC_ADMM_RELAX:
//      d = s_unc - (z - w);  s = (z - w) + Phi d;  sr = alpha s + (1 - alpha) z;  z+ = clip(sr + w);  w+ = w + sr - z+.
// alpha = 1.8 cuts the iteration count 2-3x on the slack box and 1.9x on the input box (1.9 and above degrade).
__device__ __forceinline__ int admm_box(const KArgs &a, const double *__restrict__ Phi, const double *__restrict__ lo,
                                        const double *__restrict__ hi, double bscale, double *s_unc, double *z,
                                        double *w, double *d, double smax, int TS, int tid, int *status) {
    const int nb = a.nb;
    const double thr = a.tol * fmax(bscale, smax);
    for (int j = 0; j < nb; ++j) {
        const double s = SMV(s_unc, j);
        SMV(z, j) = fmin(fmax(s, lo[j]), hi[j]);
        SMV(w, j) = 0.0;
    }
    int it = 0;
    bool conv = false;
    while (it < a.max_iter && !conv) {
        ++it;
        for (int j = 0; j < nb; ++j) SMV(d, j) = SMV(s_unc, j) - SMV(z, j) + SMV(w, j);
        double rp = 0.0, rd = 0.0;
        for (int i = 0; i < nb; ++i) {
            const double *row = Phi + (size_t)i * nb;
            double acc0 = 0.0, acc1 = 0.0;
            int j = 0;
            for (; j + 1 < nb; j += 2) {
                acc0 = fma(row[j], SMV(d, j), acc0);
                acc1 = fma(row[j + 1], SMV(d, j + 1), acc1);
            }
            if (j < nb) acc0 = fma(row[j], SMV(d, j), acc0);
            const double zi = SMV(z, i), wi = SMV(w, i);
            const double si = (zi - wi) + (acc0 + acc1);
            const double sr = DDMPC_ADMM_RELAX * si + (1.0 - DDMPC_ADMM_RELAX) * zi;
            const double zn = fmin(fmax(sr + wi, lo[i]), hi[i]);
            rp = fmax(rp, fabs(si - zn));
            rd = fmax(rd, fabs(zn - zi));
            SMV(w, i) = wi + sr - zn;
            SMV(z, i) = zn;
        }
        conv = fmax(rp, rd) <= thr;
    }
    if (!conv) *status = max(*status, (int)DDMPC_SOLVE_OPTIMAL_INACCURATE);
    for (int j = 0; j < nb; ++j) SMV(d, j) = SMV(s_unc, j) - SMV(z, j) + SMV(w, j);
    for (int i = 0; i < nb; ++i) {
        const double *row = Phi + (size_t)i * nb;
        double acc = 0.0;
        for (int j = 0; j < nb; ++j) acc = fma(row[j], SMV(d, j), acc);
        SMV(s_unc, i) = acc;
    }
    return it;
}

__device__ __forceinline__ double dot_theta(const double *__restrict__ row, const double *th, int nth, int TS, int tid) {
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
    for (; j + 1 < nth; j += 2) {
        a0 = fma(row[j], SMV(th, j), a0);
        a1 = fma(row[j + 1], SMV(th, j + 1), a1);
    }
    if (j < nth) a0 = fma(row[j], SMV(th, j), a0);
    return a0 + a1;
}

// ---------------------------------------------------------------------------
// Batched solve: thread b solves QP b.
// t_out (optional, B x nb): the box correction vector (zeros when inactive).
// ---------------------------------------------------------------------------
__global__ void k_solve_batch(KArgs a, int B, const int *__restrict__ ctrl_idx, const double *__restrict__ u_past,
                              const double *__restrict__ y_past, const double *__restrict__ u_s,
                              const double *__restrict__ y_s, double *__restrict__ optimal_u,
                              double *__restrict__ cost, int *__restrict__ status_out, int *__restrict__ iters_out,
                              double *__restrict__ t_out) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, TS = blockDim.x;
    const int b = blockIdx.x * blockDim.x + tid;
    if (b >= B) return;
    double *th = sm;
    double *s_unc = th + (size_t)a.nth * TS, *z = s_unc + (size_t)a.nb * TS, *w = z + (size_t)a.nb * TS,
           *d = w + (size_t)a.nb * TS;
    const int c = ctrl_idx ? ctrl_idx[b] : 0;
    const int nm = a.n * a.m, npp = a.n * a.p;
    double thmax = 0.0;
    bool finite = true;
    for (int i = 0; i < a.nth; ++i) {
        double v;
        if (i < nm) v = u_past[(size_t)b * nm + i];
        else if (i < nm + npp) v = y_past[(size_t)b * npp + (i - nm)];
        else if (i < nm + npp + a.m) v = u_s[(size_t)b * a.m + (i - nm - npp)];
        else v = y_s[(size_t)b * a.p + (i - nm - npp - a.m)];
        SMV(th, i) = v;
        thmax = fmax(thmax, fabs(v));
        finite = finite && isfinite(v);
    }
    int status = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
    int iters = 1;
    if (a.F && a.Fnz[c]) {  // nominal: consistency of the fixed coordinates with range(H)
        const double *F = a.F + (size_t)c * a.nfix * a.nth;
        double fe = 0.0;
        for (int i = 0; i < a.nfix; ++i) fe = fmax(fe, fabs(dot_theta(F + (size_t)i * a.nth, th, a.nth, TS, tid)));
        if (fe > 1e-6 * (1.0 + thmax)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
    }
    const double *Ku = a.Ku + (size_t)c * a.Lm * a.nth;
    bool ofinite = true;       // a controller whose setup failed has NaN gains (setup.cu, k_poison)
    for (int k = 0; k < a.Lm; ++k) {
        const double v = dot_theta(Ku + (size_t)k * a.nth, th, a.nth, TS, tid);
        ofinite = ofinite && isfinite(v);
        optimal_u[(size_t)b * a.Lm + k] = v;
    }
    if (finite && !ofinite) status = max(status, (int)DDMPC_SOLVE_NONFINITE);
    double J = 0.0;
    if (cost) {
        const double *Z = a.Z + (size_t)c * a.nth * a.nth;
        for (int i = 0; i < a.nth; ++i) J = fma(SMV(th, i), dot_theta(Z + (size_t)i * a.nth, th, a.nth, TS, tid), J);
    }
    bool active = false;
    if (setpoint_outside_box(a, th, TS, tid)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
    if (a.nb > 0) {
        const double *Ks = a.Ks + (size_t)c * a.nb * a.nth;
        const double *lo = a.lo + (size_t)c * a.nb, *hi = a.hi + (size_t)c * a.nb;
        double smax = 0.0;
        bool viol = false;
        for (int j = 0; j < a.nb; ++j) {
            const double s = dot_theta(Ks + (size_t)j * a.nth, th, a.nth, TS, tid);
            SMV(s_unc, j) = s;
            smax = fmax(smax, fabs(s));
            viol = viol || s < lo[j] || s > hi[j];
        }
        if (viol && finite && status != DDMPC_SOLVE_INFEASIBLE) {
            active = true;
            iters = admm_box(a, a.Phi + (size_t)c * a.nb * a.nb, lo, hi, a.bmax[c], s_unc, z, w, d, smax, TS, tid, &status);
            const double *Psi = a.Psi + (size_t)c * a.Lm * a.nb;
            for (int k = 0; k < a.Lm; ++k) {
                double acc = 0.0;
                for (int j = 0; j < a.nb; ++j) acc = fma(Psi[(size_t)k * a.nb + j], SMV(s_unc, j), acc);
                optimal_u[(size_t)b * a.Lm + k] -= acc;
            }
            if (cost) {
                const double *Lam = a.Lam + (size_t)c * a.nb * a.nb;
                const double rho = a.rho2[c];
                double q = 0.0;
                for (int i = 0; i < a.nb; ++i) {
                    double acc = 0.0;
                    for (int j = 0; j < a.nb; ++j) acc = fma(Lam[(size_t)i * a.nb + j], SMV(s_unc, j), acc);
                    q = fma(SMV(s_unc, i), acc, q);
                }
                J += rho * rho * q;
            }
        }
    }
    if (t_out)
        for (int j = 0; j < a.nb; ++j) t_out[(size_t)b * a.nb + j] = active ? SMV(s_unc, j) : 0.0;
    if (cost) cost[b] = J;
    if (status_out) status_out[b] = status;
    if (iters_out) iters_out[b] = iters;
}

// ---------------------------------------------------------------------------
// Small batches (the B = 1 controller object of the reference API, up to a few dozen solves):
// one CTA per solve, rows of every operator spread over the threads, block-cooperative ADMM.
// Same arithmetic as k_solve_batch; exists for latency (p50 single-loop step latency metric).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_max(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (blockDim.x + 31) / 32; ++w) r = fmax(r, red[w]);
    return r;
}
__device__ __forceinline__ double block_sum(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) r += red[w];
    return r;
}
__device__ __forceinline__ double dot_row(const double *__restrict__ row, const double *v, int n) {
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
    for (; j + 1 < n; j += 2) {
        a0 = fma(__ldg(row + j), v[j], a0);
        a1 = fma(__ldg(row + j + 1), v[j + 1], a1);
    }
    if (j < n) a0 = fma(__ldg(row + j), v[j], a0);
    return a0 + a1;
}

__global__ void __launch_bounds__(128)
k_solve_small(KArgs a, int B, const int *__restrict__ ctrl_idx, const double *__restrict__ u_past,
              const double *__restrict__ y_past, const double *__restrict__ u_s, const double *__restrict__ y_s,
              double *__restrict__ optimal_u, double *__restrict__ cost, int *__restrict__ status_out,
              int *__restrict__ iters_out, double *__restrict__ t_out) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    double *th = sm, *su = th + a.nth, *z = su + a.nb, *w = z + a.nb, *dd = w + a.nb, *red = dd + a.nb, *uo = red + 64;
    const int c = ctrl_idx ? ctrl_idx[b] : 0;
    const int nm = a.n * a.m, npp = a.n * a.p;
    double tmax = 0.0, bad = 0.0;
    for (int i = tid; i < a.nth; i += T) {
        double v;
        if (i < nm) v = u_past[(size_t)b * nm + i];
        else if (i < nm + npp) v = y_past[(size_t)b * npp + (i - nm)];
        else if (i < nm + npp + a.m) v = u_s[(size_t)b * a.m + (i - nm - npp)];
        else v = y_s[(size_t)b * a.p + (i - nm - npp - a.m)];
        th[i] = v;
        tmax = fmax(tmax, fabs(v));
        if (!isfinite(v)) bad = 1.0;
    }
    const double thmax = block_max(tmax, red);
    const bool finite = block_max(bad, red) == 0.0;
    int status = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
    int iters = 1;
    if (a.F && a.Fnz[c]) {
        const double *F = a.F + (size_t)c * a.nfix * a.nth;
        double fe = 0.0;
        for (int i = tid; i < a.nfix; i += T) fe = fmax(fe, fabs(dot_row(F + (size_t)i * a.nth, th, a.nth)));
        fe = block_max(fe, red);
        if (fe > 1e-6 * (1.0 + thmax)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
    }
    const double *Ku = a.Ku + (size_t)c * a.Lm * a.nth;
    // (the result is staged in shared memory and written once: optimal_u may be mapped host memory)
    double obad = 0.0;         // a controller whose setup failed has NaN gains (setup.cu, k_poison)
    for (int k = tid; k < a.Lm; k += T) {
        uo[k] = dot_row(Ku + (size_t)k * a.nth, th, a.nth);
        if (!isfinite(uo[k])) obad = 1.0;
    }
    if (block_max(obad, red) != 0.0) status = max(status, (int)DDMPC_SOLVE_NONFINITE);
    double J = 0.0;
    if (cost) {
        const double *Z = a.Z + (size_t)c * a.nth * a.nth;
        double part = 0.0;
        for (int i = tid; i < a.nth; i += T) part = fma(th[i], dot_row(Z + (size_t)i * a.nth, th, a.nth), part);
        J = block_sum(part, red);
    }
    bool active = false;
    if (setpoint_outside_box(a, th, 1, 0)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
    if (a.nb > 0) {
        const double *Ks = a.Ks + (size_t)c * a.nb * a.nth;
        const double *lo = a.lo + (size_t)c * a.nb, *hi = a.hi + (size_t)c * a.nb;
        double sm_l = 0.0, vi_l = 0.0;
        for (int j = tid; j < a.nb; j += T) {
            const double s = dot_row(Ks + (size_t)j * a.nth, th, a.nth);
            su[j] = s;
            sm_l = fmax(sm_l, fabs(s));
            if (s < lo[j] || s > hi[j]) vi_l = 1.0;
        }
        const double smax = block_max(sm_l, red);
        const bool viol = block_max(vi_l, red) != 0.0;
        if (viol && finite && status != DDMPC_SOLVE_INFEASIBLE) {
            active = true;
            const double *Phi = a.Phi + (size_t)c * a.nb * a.nb;
            const double thr = a.tol * fmax(a.bmax[c], smax);
            for (int j = tid; j < a.nb; j += T) {
                z[j] = fmin(fmax(su[j], lo[j]), hi[j]);
                w[j] = 0.0;
            }
            int it = 0;
            bool conv = false;
            while (it < a.max_iter && !conv) {
                ++it;
                __syncthreads();
                for (int j = tid; j < a.nb; j += T) dd[j] = su[j] - z[j] + w[j];
                __syncthreads();
                double res = 0.0;
                for (int j = tid; j < a.nb; j += T) {
                    const double acc = dot_row(Phi + (size_t)j * a.nb, dd, a.nb);
                    const double si = (z[j] - w[j]) + acc;
                    const double sr = DDMPC_ADMM_RELAX * si + (1.0 - DDMPC_ADMM_RELAX) * z[j];
                    const double zn = fmin(fmax(sr + w[j], lo[j]), hi[j]);
                    res = fmax(res, fmax(fabs(si - zn), fabs(zn - z[j])));
                    w[j] = w[j] + sr - zn;
                    z[j] = zn;
                }
                conv = block_max(res, red) <= thr;
            }
            if (!conv) status = max(status, (int)DDMPC_SOLVE_OPTIMAL_INACCURATE);
            iters = it;
            __syncthreads();
            for (int j = tid; j < a.nb; j += T) dd[j] = su[j] - z[j] + w[j];
            __syncthreads();
            for (int j = tid; j < a.nb; j += T) su[j] = dot_row(Phi + (size_t)j * a.nb, dd, a.nb);   // t = Phi d
            __syncthreads();
            const double *Psi = a.Psi + (size_t)c * a.Lm * a.nb;
            for (int k = tid; k < a.Lm; k += T) uo[k] -= dot_row(Psi + (size_t)k * a.nb, su, a.nb);
            if (cost) {
                const double *Lam = a.Lam + (size_t)c * a.nb * a.nb;
                const double rho = a.rho2[c];
                double part = 0.0;
                for (int i = tid; i < a.nb; i += T) part = fma(su[i], dot_row(Lam + (size_t)i * a.nb, su, a.nb), part);
                J += rho * rho * block_sum(part, red);
            }
        }
    }
    for (int k = tid; k < a.Lm; k += T) optimal_u[(size_t)b * a.Lm + k] = uo[k];   // same thread wrote uo[k]
    if (t_out)
        for (int j = tid; j < a.nb; j += T) t_out[(size_t)b * a.nb + j] = active ? su[j] : 0.0;
    if (tid == 0) {
        if (cost) cost[b] = J;
        if (status_out) status_out[b] = status;
        if (iters_out) iters_out[b] = iters;
    }
}

// full primal: x = X0 theta - Yf t      (one thread per (b, i))
__global__ void k_full_x(KArgs a, int B, const int *__restrict__ ctrl_idx, const double *__restrict__ u_past,
                         const double *__restrict__ y_past, const double *__restrict__ u_s,
                         const double *__restrict__ y_s, const double *__restrict__ t, double *__restrict__ x) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * a.nx) return;
    const int b = (int)(e / a.nx), i = (int)(e % a.nx);
    const int c = ctrl_idx ? ctrl_idx[b] : 0;
    const int nm = a.n * a.m, npp = a.n * a.p;
    const double *row = a.X0 + ((size_t)c * a.nx + i) * a.nth;
    double acc = 0.0;
    for (int j = 0; j < a.nth; ++j) {
        double v;
        if (j < nm) v = u_past[(size_t)b * nm + j];
        else if (j < nm + npp) v = y_past[(size_t)b * npp + (j - nm)];
        else if (j < nm + npp + a.m) v = u_s[(size_t)b * a.m + (j - nm - npp)];
        else v = y_s[(size_t)b * a.p + (j - nm - npp - a.m)];
        acc = fma(row[j], v, acc);
    }
    if (a.nb > 0) {
        const double *yr = a.Yf + ((size_t)c * a.nx + i) * a.nb;
        for (int j = 0; j < a.nb; ++j) acc = fma(-yr[j], t[(size_t)b * a.nb + j], acc);
    }
    x[e] = acc;
}

// g = Om (T x)   (B x r)   then   alpha = H^T g   (B x cols)
__global__ void k_full_g(KArgs a, int B, const int *__restrict__ ctrl_idx, int shared_data, const double *__restrict__ Om,
                         const double *__restrict__ x, double *__restrict__ g) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * a.r) return;
    const int b = (int)(e / a.r), i = (int)(e % a.r);
    const int c = (ctrl_idx && !shared_data) ? ctrl_idx[b] : 0;
    const double *row = Om + ((size_t)c * a.r + i) * a.r;
    const double *xb = x + (size_t)b * a.nx;
    double acc = 0.0;
    for (int k = 0; k < a.r; ++k) {
        double tk = xb[k];
        if (k >= a.nu) tk += xb[k + a.ny];   // t_y = ybar + sigma
        acc = fma(row[k], tk, acc);
    }
    g[e] = acc;
}
__global__ void k_full_alpha(KArgs a, int B, const int *__restrict__ ctrl_idx, int shared_data, const double *__restrict__ H,
                             const double *__restrict__ g, double *__restrict__ alpha) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * a.cols) return;
    const int b = (int)(e / a.cols), col = (int)(e % a.cols);
    const int c = (ctrl_idx && !shared_data) ? ctrl_idx[b] : 0;
    const double *Hc = H + (size_t)c * a.r * a.cols;
    double acc = 0.0;
    for (int k = 0; k < a.r; ++k) acc = fma(Hc[(size_t)k * a.cols + col], g[(size_t)b * a.r + k], acc);
    alpha[e] = acc;
}
// t = T x = [ubar; ybar + sigma]   (B x r), the right-hand side of alpha = H^T W^-1 t
__global__ void k_full_t(KArgs a, int B, const double *__restrict__ x, double *__restrict__ t) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * a.r) return;
    const int b = (int)(e / a.r), k = (int)(e % a.r);
    const double *xb = x + (size_t)b * a.nx;
    t[e] = k >= a.nu ? xb[k] + xb[k + a.ny] : xb[k];
}
__global__ void k_split_x(KArgs a, int B, const double *__restrict__ x, double *__restrict__ ubar,
                          double *__restrict__ ybar, double *__restrict__ sigma) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * a.nx) return;
    const int b = (int)(e / a.nx), i = (int)(e % a.nx);
    const double v = x[e];
    if (i < a.nu) { if (ubar) ubar[(size_t)b * a.nu + i] = v; }
    else if (i < a.nu + a.ny) { if (ybar) ybar[(size_t)b * a.ny + (i - a.nu)] = v; }
    else if (sigma) sigma[(size_t)b * a.ny + (i - a.nu - a.ny)] = v;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (throughput-mode measurement noise; oracle/ddmpc_oracle.py
// philox_noise restates it bit-for-bit)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
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
// 32 random mantissa bits -> v in [1, 2); the noise sample is eps * (2 v - 3) in [-eps, eps)
__device__ __forceinline__ double unit32(uint32_t x) {
    return __hiloint2double((int)(0x3FF00000u | (x >> 12)), (int)(x << 20));
}

struct LoopArgs {
    int n_x, n_steps, n_mpc;
    const double *A, *Bm, *C, *D;   // device
    const double *w;                // (B, n_steps, p) or NULL
    uint64_t seed, id0;
    double eps;
};

// ---------------------------------------------------------------------------
// Fused closed loop: thread b runs closed loop b for n_steps steps.
// ---------------------------------------------------------------------------
__global__ void k_closed_loop(KArgs a, LoopArgs la, int B, const int *__restrict__ ctrl_idx,
                              const double *__restrict__ x0, const double *__restrict__ u_past0,
                              const double *__restrict__ y_past0, const double *__restrict__ u_s,
                              const double *__restrict__ y_s, double *__restrict__ u_sys,
                              double *__restrict__ y_sys, int *__restrict__ status_out,
                              int *__restrict__ iters_out, double *__restrict__ x_final) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, TS = blockDim.x;
    const int b = blockIdx.x * blockDim.x + tid;
    if (b >= B) return;
    const int nm = a.n * a.m, npp = a.n * a.p, m = a.m, p = a.p, nxp = la.n_x;
    const int nplan = la.n_mpc * m;
    double *th = sm;                                   // theta (nth): [u_past; y_past; u_s; y_s]
    double *xs = th + (size_t)a.nth * TS;              // plant state (n_x)
    double *xn = xs + (size_t)nxp * TS;                // next state (n_x)
    double *up = xn + (size_t)nxp * TS;                // planned inputs (n_mpc*m)
    double *yk = up + (size_t)nplan * TS;              // current output (p)
    double *s_unc = yk + (size_t)p * TS, *z = s_unc + (size_t)a.nb * TS, *w = z + (size_t)a.nb * TS,
           *d = w + (size_t)a.nb * TS;
    const int c = ctrl_idx ? ctrl_idx[b] : 0;
    const double *Ku = a.Ku + (size_t)c * a.Lm * a.nth;
    const bool boxed = a.nb > 0;
    const double *Ks = boxed ? a.Ks + (size_t)c * a.nb * a.nth : nullptr;
    const double *Phi = boxed ? a.Phi + (size_t)c * a.nb * a.nb : nullptr;
    const double *Psi = boxed ? a.Psi + (size_t)c * a.Lm * a.nb : nullptr;
    const double *lo = boxed ? a.lo + (size_t)c * a.nb : nullptr, *hi = boxed ? a.hi + (size_t)c * a.nb : nullptr;
    const double *F = (a.F && a.Fnz[c]) ? a.F + (size_t)c * a.nfix * a.nth : nullptr;

    for (int i = 0; i < nm; ++i) SMV(th, i) = u_past0[(size_t)b * nm + i];
    for (int i = 0; i < npp; ++i) SMV(th, nm + i) = y_past0[(size_t)b * npp + i];
    for (int i = 0; i < m; ++i) SMV(th, nm + npp + i) = u_s[(size_t)b * m + i];
    for (int i = 0; i < p; ++i) SMV(th, nm + npp + m + i) = y_s[(size_t)b * p + i];
    for (int i = 0; i < nxp; ++i) SMV(xs, i) = x0[(size_t)b * nxp + i];

    int status = DDMPC_SOLVE_OPTIMAL, iters = 0;
    if (setpoint_outside_box(a, th, TS, tid)) status = DDMPC_SOLVE_INFEASIBLE;
    const uint64_t sid = la.id0 + (uint64_t)b;
    for (int t = 0; t < la.n_steps; t += la.n_mpc) {
        // ---- solve: planned inputs = first n_mpc*m rows of Ku theta (+ box correction)
        if (F) {
            double fe = 0.0, thmax = 0.0;
            for (int i = 0; i < a.nth; ++i) thmax = fmax(thmax, fabs(SMV(th, i)));
            for (int i = 0; i < a.nfix; ++i) fe = fmax(fe, fabs(dot_theta(F + (size_t)i * a.nth, th, a.nth, TS, tid)));
            if (fe > 1e-6 * (1.0 + thmax)) status = max(status, (int)DDMPC_SOLVE_INFEASIBLE);
        }
        for (int k = 0; k < nplan; ++k) SMV(up, k) = dot_theta(Ku + (size_t)k * a.nth, th, a.nth, TS, tid);
        int it = 1;
        if (boxed) {
            double smax = 0.0;
            bool viol = false;
            for (int j = 0; j < a.nb; ++j) {
                const double s = dot_theta(Ks + (size_t)j * a.nth, th, a.nth, TS, tid);
                SMV(s_unc, j) = s;
                smax = fmax(smax, fabs(s));
                viol = viol || s < lo[j] || s > hi[j];
            }
            if (viol && isfinite(smax)) {
                it = admm_box(a, Phi, lo, hi, a.bmax[c], s_unc, z, w, d, smax, TS, tid, &status);
                for (int k = 0; k < nplan; ++k) {
                    double acc = 0.0;
                    for (int j = 0; j < a.nb; ++j) acc = fma(Psi[(size_t)k * a.nb + j], SMV(s_unc, j), acc);
                    SMV(up, k) -= acc;
                }
            }
        }
        iters += it;
        // ---- apply n_mpc inputs: plant step, record, shift the measurement window
        const int kend = min(t + la.n_mpc, la.n_steps);
        for (int k = t; k < kend; ++k) {
            const double *uk = up + (size_t)(k - t) * m * TS;
            // measurement noise
            if (la.w) {
                for (int i = 0; i < p; ++i) SMV(yk, i) = la.w[((size_t)b * la.n_steps + k) * p + i];
            } else {
                // noise word q = k*p + i is word (q & 3) of Philox call (q >> 2)
                uint32_t o[4];
                unsigned last = 0xffffffffu;
                for (int i = 0; i < p; ++i) {
                    const unsigned q = (unsigned)k * (unsigned)p + (unsigned)i;
                    if ((q >> 2) != last) {
                        last = q >> 2;
                        philox4x32_10(last, 0u, (uint32_t)(sid & 0xffffffffu), (uint32_t)(sid >> 32),
                                      (uint32_t)(la.seed & 0xffffffffu), (uint32_t)(la.seed >> 32), o);
                    }
                    SMV(yk, i) = la.eps * (2.0 * unit32(o[q & 3]) - 3.0);
                }
            }
            // y = C x + D u + w   (uses the pre-update state; model_simulation.py:94)
            for (int i = 0; i < p; ++i) {
                double acc = 0.0;
                for (int j = 0; j < nxp; ++j) acc = fma(la.C[i * nxp + j], SMV(xs, j), acc);
                double acd = 0.0;
                for (int j = 0; j < m; ++j) acd = fma(la.D[i * m + j], SMV(uk, j), acd);
                SMV(yk, i) = (acc + acd) + SMV(yk, i);
            }
            // x <- A x + B u      (model_simulation.py:96)
            for (int i = 0; i < nxp; ++i) {
                double acc = 0.0;
                for (int j = 0; j < nxp; ++j) acc = fma(la.A[i * nxp + j], SMV(xs, j), acc);
                double acb = 0.0;
                for (int j = 0; j < m; ++j) acb = fma(la.Bm[i * m + j], SMV(uk, j), acb);
                SMV(xn, i) = acc + acb;
            }
            for (int i = 0; i < nxp; ++i) SMV(xs, i) = SMV(xn, i);
            // record
            bool fin = true;
            for (int i = 0; i < m; ++i) u_sys[((size_t)b * la.n_steps + k) * m + i] = SMV(uk, i);
            for (int i = 0; i < p; ++i) {
                const double y = SMV(yk, i);
                fin = fin && isfinite(y);
                y_sys[((size_t)b * la.n_steps + k) * p + i] = y;
            }
            if (!fin) status = max(status, (int)DDMPC_SOLVE_NONFINITE);
            // window shift (controller.py:893-895)
            for (int i = 0; i < nm - m; ++i) SMV(th, i) = SMV(th, i + m);
            for (int i = 0; i < m; ++i) SMV(th, nm - m + i) = SMV(uk, i);
            for (int i = 0; i < npp - p; ++i) SMV(th, nm + i) = SMV(th, nm + i + p);
            for (int i = 0; i < p; ++i) SMV(th, nm + npp - p + i) = SMV(yk, i);
        }
    }
    if (status_out) status_out[b] = status;
    if (iters_out) iters_out[b] = iters;
    if (x_final)
        for (int i = 0; i < nxp; ++i) x_final[(size_t)b * nxp + i] = SMV(xs, i);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int closed_loop_fast_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                         cudaStream_t st);

int closed_loop_cvx_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                        const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                        const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                        double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                        cudaStream_t st);

int closed_loop_tc_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                       const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                       const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                       double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st);

int solve_batch_cvx_dmma(const ddmpc_set *set, int B, const int *ctrl_idx, const double *u_past, const double *y_past,
                         const double *u_s, const double *y_s, double tol, int max_iter, double *optimal_u, double *cost,
                         int *status, int *iters, double *t_out, cudaStream_t st);

int closed_loop_perloop_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                            const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                            const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                            double *y_sys, int *status, int *iters, double *x_final, double tol, int max_iter,
                            cudaStream_t st);

int closed_loop_gemm_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st);

int closed_loop_dmma_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st);

static KArgs make_kargs(const ddmpc_set *set, double tol, int max_iter) {
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    KArgs a{};
    a.n = d.n; a.m = d.m; a.p = d.p; a.L = d.L; a.nth = d.nth; a.nb = d.nb; a.Lm = d.Lm; a.nx = d.nx; a.nfix = d.nfix;
    a.r = d.r; a.cols = d.cols; a.ny = d.ny; a.nu = d.nu;
    a.convex = d.convex; a.robust = d.robust;
    a.Ku = pl.Ku.d(); a.Z = pl.Z.d(); a.X0 = pl.X0.d();
    a.Ks = pl.Ks.d(); a.Phi = pl.Phi.d(); a.Psi = pl.Psi.d(); a.Lam = pl.Lam.d(); a.Yf = pl.Yf.d();
    a.rho2 = pl.rho2.d();
    a.F = d.robust ? nullptr : pl.F.d();
    a.Fnz = d.robust ? nullptr : pl.Fnz.i();
    a.lo = pl.lo.d(); a.hi = pl.hi.d(); a.bmax = pl.bmax.d();
    a.umin = d.nbu > 0 ? pl.umin.d() : nullptr; a.umax = d.nbu > 0 ? pl.umax.d() : nullptr;
    a.ymin = d.nby > 0 ? pl.ymin.d() : nullptr; a.ymax = d.nby > 0 ? pl.ymax.d() : nullptr;
    a.terminal = d.terminal;
    a.tol = tol > 0.0 ? tol : 1e-8;
    a.max_iter = max_iter > 0 ? max_iter : 1000;
    return a;
}

// threads per block such that `per_thread` doubles of state fit in shared memory
static int pick_tpb(size_t per_thread_doubles, size_t *smem_bytes) {
    const size_t budget = 200 * 1024;
    int tpb = 128;
    while (tpb > 32 && (size_t)tpb * per_thread_doubles * 8 > budget) tpb -= 32;
    *smem_bytes = (size_t)tpb * per_thread_doubles * 8;
    return (*smem_bytes > budget) ? 0 : tpb;
}

int solve_batch_device(const ddmpc_set *set, int B, const int *ctrl_idx, const double *u_past, const double *y_past,
                       const double *u_s, const double *y_s, double tol, int max_iter, double *optimal_u, double *cost,
                       int *status, int *iters, double *t_out, cudaStream_t st) {
    if (!set || B < 0) return fail(DDMPC_ERR_INVALID_ARG, "solve_batch: bad set / batch size");
    if (B == 0) return DDMPC_OK;
    if (!u_past || !y_past || !u_s || !y_s || !optimal_u)
        return fail(DDMPC_ERR_INVALID_ARG, "solve_batch: null argument");
    if (set->opt_solve != 1) {   // shared CONVEX controller, large batch: GEMMs + box-row ADMM on the tensor cores
        const int rc = solve_batch_cvx_dmma(set, B, ctrl_idx, u_past, y_past, u_s, y_s, tol, max_iter, optimal_u, cost, status,
                                            iters, t_out, st);
        if (rc != -1) return rc;
    }
    KArgs a = make_kargs(set, tol, max_iter);
    if (B <= 64) {   // latency path: one CTA per solve
        const size_t sh = sizeof(double) * ((size_t)a.nth + 4 * (size_t)a.nb + 64 + (size_t)a.Lm);
        if (sh <= 200 * 1024) {
            if (sh > 48 * 1024)
                DDMPC_CUDA(cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
            k_solve_small<<<B, 128, sh, st>>>(a, B, ctrl_idx, u_past, y_past, u_s, y_s, optimal_u, cost, status, iters,
                                              t_out);
            DDMPC_LAUNCH_CHECK();
            return DDMPC_OK;
        }
    }
    size_t smem = 0;
    const int tpb = pick_tpb((size_t)a.nth + 4 * (size_t)a.nb, &smem);
    if (!tpb) return fail(DDMPC_ERR_INVALID_ARG, "solve_batch: problem too large for the generic kernel");
    DDMPC_CUDA(cudaFuncSetAttribute(k_solve_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_solve_batch<<<ceil_div(B, tpb), tpb, smem, st>>>(a, B, ctrl_idx, u_past, y_past, u_s, y_s, optimal_u, cost,
                                                        status, iters, t_out);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

}  // namespace ddmpc

using namespace ddmpc;

namespace {
// host -> device staging helper for the *_host entry points
struct Stage {
    std::vector<DevBuf *> bufs;
    ~Stage() { for (auto *b : bufs) delete b; }
    template <typename T> int up(const T *h, size_t n, T **d) {
        *d = nullptr;
        if (!h || n == 0) return DDMPC_OK;
        DevBuf *b = new DevBuf();
        bufs.push_back(b);
        DDMPC_CUDA(b->alloc(sizeof(T) * n));
        DDMPC_CUDA(cudaMemcpy(b->p, h, sizeof(T) * n, cudaMemcpyHostToDevice));
        *d = (T *)b->p;
        return DDMPC_OK;
    }
    template <typename T> int out(size_t n, T **d, bool wanted = true) {
        *d = nullptr;
        if (!wanted || n == 0) return DDMPC_OK;
        DevBuf *b = new DevBuf();
        bufs.push_back(b);
        DDMPC_CUDA(b->alloc(sizeof(T) * n));
        *d = (T *)b->p;
        return DDMPC_OK;
    }
};
}  // namespace

extern "C" {

int ddmpc_solve_batch(const ddmpc_set *set, int B, const int32_t *ctrl_idx, const double *u_past, const double *y_past,
                      const double *u_s, const double *y_s, double tol, int max_iter, double *optimal_u, double *cost,
                      int32_t *status, int32_t *iters, void *stream) {
    return solve_batch_device(set, B, ctrl_idx, u_past, y_past, u_s, y_s, tol, max_iter, optimal_u, cost, status, iters,
                              nullptr, (cudaStream_t)stream);
}

// B == 1 (the controller object of the reference API): one persistent pinned buffer that the kernel reads and
// writes directly over PCIe (pinned allocations are device-accessible under unified addressing), one kernel on a
// private stream, one synchronise - no allocation and no copy command on the per-step path.
static int solve_one_host_staged(const ddmpc_set *set, const int32_t *ctrl_idx, const double *u_past,
                                 const double *y_past, const double *u_s, const double *y_s, double tol, int max_iter,
                                 double *optimal_u, double *cost, int32_t *status, int32_t *iters) {
    const Dims &d = set->plan.d;
    const int nm = d.n * d.m, npp = d.n * d.p;
    const size_t n_in = (size_t)d.nth + 1, n_out = (size_t)d.Lm + 3;   // doubles (+ ctrl / cost, status, iters)
    if (!set->stage_host) {
        DDMPC_CUDA(cudaHostAlloc(&set->stage_host, sizeof(double) * (n_in + n_out), cudaHostAllocMapped));
        DDMPC_CUDA(cudaStreamCreateWithFlags(&set->stage_stream, cudaStreamNonBlocking));
    }
    double *h = (double *)set->stage_host;
    std::copy(u_past, u_past + nm, h);
    std::copy(y_past, y_past + npp, h + nm);
    std::copy(u_s, u_s + d.m, h + nm + npp);
    std::copy(y_s, y_s + d.p, h + nm + npp + d.m);
    int32_t *hci = reinterpret_cast<int32_t *>(h + d.nth);
    hci[0] = ctrl_idx ? ctrl_idx[0] : 0;
    double *dv = nullptr;
    DDMPC_CUDA(cudaHostGetDevicePointer((void **)&dv, h, 0));
    cudaStream_t st = set->stage_stream;
    double *dout = dv + n_in;
    int32_t *dints = reinterpret_cast<int32_t *>(dout + d.Lm + 1);
    DDMPC_TRY(solve_batch_device(set, 1, reinterpret_cast<const int *>(dv + d.nth), dv, dv + nm, dv + nm + npp,
                                 dv + nm + npp + d.m, tol, max_iter, dout, dout + d.Lm, dints, dints + 1, nullptr, st));
    DDMPC_CUDA(cudaStreamSynchronize(st));
    const double *hout = h + n_in;
    std::copy(hout, hout + d.Lm, optimal_u);
    if (cost) *cost = hout[d.Lm];
    const int32_t *hints = reinterpret_cast<const int32_t *>(hout + d.Lm + 1);
    if (status) *status = hints[0];
    if (iters) *iters = hints[1];
    return DDMPC_OK;
}

int ddmpc_solve_batch_host(const ddmpc_set *set, int B, const int32_t *ctrl_idx, const double *u_past,
                           const double *y_past, const double *u_s, const double *y_s, double tol, int max_iter,
                           double *optimal_u, double *cost, int32_t *status, int32_t *iters) {
    if (!set || B <= 0 || !optimal_u) return fail(DDMPC_ERR_INVALID_ARG, "solve_batch_host: bad argument");
    const Dims &d = set->plan.d;
    if (B == 1 && u_past && y_past && u_s && y_s)
        return solve_one_host_staged(set, ctrl_idx, u_past, y_past, u_s, y_s, tol, max_iter, optimal_u, cost, status, iters);
    Stage s;
    int32_t *dc, *dst, *dit;
    double *dup, *dyp, *dus, *dys, *dou, *dco;
    DDMPC_TRY(s.up(ctrl_idx, (size_t)B, &dc));
    DDMPC_TRY(s.up(u_past, (size_t)B * d.n * d.m, &dup));
    DDMPC_TRY(s.up(y_past, (size_t)B * d.n * d.p, &dyp));
    DDMPC_TRY(s.up(u_s, (size_t)B * d.m, &dus));
    DDMPC_TRY(s.up(y_s, (size_t)B * d.p, &dys));
    DDMPC_TRY(s.out((size_t)B * d.Lm, &dou));
    DDMPC_TRY(s.out((size_t)B, &dco, cost != nullptr));
    DDMPC_TRY(s.out((size_t)B, &dst, status != nullptr));
    DDMPC_TRY(s.out((size_t)B, &dit, iters != nullptr));
    DDMPC_TRY(solve_batch_device(set, B, dc, dup, dyp, dus, dys, tol, max_iter, dou, dco, dst, dit, nullptr, nullptr));
    DDMPC_CUDA(cudaStreamSynchronize(nullptr));
    DDMPC_CUDA(cudaMemcpy(optimal_u, dou, sizeof(double) * B * d.Lm, cudaMemcpyDeviceToHost));
    if (cost) DDMPC_CUDA(cudaMemcpy(cost, dco, sizeof(double) * B, cudaMemcpyDeviceToHost));
    if (status) DDMPC_CUDA(cudaMemcpy(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
    if (iters) DDMPC_CUDA(cudaMemcpy(iters, dit, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
    return DDMPC_OK;
}

int ddmpc_solve_full_batch(const ddmpc_set *set, int B, const int32_t *ctrl_idx, const double *u_past,
                           const double *y_past, const double *u_s, const double *y_s, double tol, int max_iter,
                           double *ubar, double *ybar, double *sigma, double *alpha, void *stream) {
    if (!set || B <= 0) return fail(DDMPC_ERR_INVALID_ARG, "solve_full_batch: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ScratchStreamScope scratch_scope(st);   // scratch below is allocated and freed in the order of `st`
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    if (alpha && !d.robust) return fail(DDMPC_ERR_INVALID_ARG, "alpha is only defined (unique) for ROBUST controllers");
    KArgs a = make_kargs(set, tol, max_iter);
    DevBuf ou, tb, xb, gb;
    DDMPC_CUDA(ou.alloc(sizeof(double) * (size_t)B * d.Lm));
    DDMPC_CUDA(tb.alloc(sizeof(double) * (size_t)B * std::max(d.nb, 1)));
    DDMPC_CUDA(xb.alloc(sizeof(double) * (size_t)B * d.nx));
    DDMPC_TRY(solve_batch_device(set, B, ctrl_idx, u_past, y_past, u_s, y_s, tol, max_iter, ou.d(), nullptr, nullptr,
                                 nullptr, d.nb > 0 ? tb.d() : nullptr, st));
    const int T = 256;
    k_full_x<<<ceil_div((long)B * d.nx, T), T, 0, st>>>(a, B, ctrl_idx, u_past, y_past, u_s, y_s, tb.d(), xb.d());
    DDMPC_LAUNCH_CHECK();
    k_split_x<<<ceil_div((long)B * d.nx, T), T, 0, st>>>(a, B, xb.d(), ubar, ybar, d.robust ? sigma : nullptr);
    DDMPC_LAUNCH_CHECK();
    if (alpha) {
        DDMPC_CUDA(gb.alloc(sizeof(double) * (size_t)B * d.r));
        const int shared_data = pl.data_count == 1 ? 1 : 0;
        if (shared_data && B >= 64) {
            // One data set for the whole batch: alpha = (T x) W^-1 H is the shared-Hankel x batch-of-iterates product
            // of the north star, two FP64 tensor-core GEMMs (k_gemm, linalg.cuh) over the batch:
            //   G (B x r) = T (B x r) Om       (Om = W^-1 symmetric),      alpha (B x cols) = G H
            DevBuf tv;
            DDMPC_CUDA(tv.alloc(sizeof(double) * (size_t)B * d.r));
            k_full_t<<<ceil_div((long)B * d.r, T), T, 0, st>>>(a, B, xb.d(), tv.d());
            DDMPC_LAUNCH_CHECK();
            DDMPC_TRY(gemm(st, 1, B, d.r, d.r, 1.0, mat(tv.d(), d.r, 1, 0), mat(pl.Om.d(), d.r, 1, 0), 0.0, gb.d(), d.r, 1, 0));
            DDMPC_TRY(gemm(st, 1, B, d.cols, d.r, 1.0, mat(gb.d(), d.r, 1, 0), mat(pl.H.d(), d.cols, 1, 0), 0.0, alpha, d.cols, 1, 0));
            return DDMPC_OK;
        }
        k_full_g<<<ceil_div((long)B * d.r, T), T, 0, st>>>(a, B, ctrl_idx, shared_data, pl.Om.d(), xb.d(), gb.d());
        DDMPC_LAUNCH_CHECK();
        k_full_alpha<<<ceil_div((long)B * d.cols, T), T, 0, st>>>(a, B, ctrl_idx, shared_data, pl.H.d(), gb.d(), alpha);
        DDMPC_LAUNCH_CHECK();
    }
    return DDMPC_OK;                        // (scratch is released in stream order, no synchronise needed)
}

int ddmpc_closed_loop_batch(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int32_t *ctrl_idx,
                            const double *x0, const double *u_past0, const double *y_past0, const double *u_s,
                            const double *y_s, const double *w, uint64_t noise_seed, uint64_t scenario_id0,
                            double noise_eps, int n_steps, double tol, int max_iter, double *u_sys, double *y_sys,
                            int32_t *status, int32_t *iters, double *x_final, void *stream) {
    if (!set || !plant || B < 0 || n_steps < 0) return fail(DDMPC_ERR_INVALID_ARG, "closed_loop_batch: bad argument");
    if (B == 0 || n_steps == 0) return DDMPC_OK;
    if (!x0 || !u_past0 || !y_past0 || !u_s || !y_s || !u_sys || !y_sys)
        return fail(DDMPC_ERR_INVALID_ARG, "closed_loop_batch: null argument");
    const Dims &d = set->plan.d;
    if (plant->m != d.m || plant->p != d.p || plant->n_x <= 0 || !plant->A || !plant->B || !plant->C || !plant->D)
        return fail(DDMPC_ERR_INVALID_ARG, "closed_loop_batch: plant does not match the controller (m=%d p=%d)", d.m, d.p);
    const int n_mpc = set->prm.n_mpc_step;
    if (n_mpc < 1 || n_mpc > d.L) return fail(DDMPC_ERR_INVALID_ARG, "n_mpc_step must be within [1, L]");
    if (B == 0 || n_steps == 0) return DDMPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nxp = plant->n_x;
    const int path = set->opt_path;
    if (set->opt_layout == 1) {
        // step-major trajectories (n_steps, B, m): an option for consumers that stay on the device; only the warp-specialised
        // kernel of the four-tank n-step shape writes them
        const int rc = (path == DDMPC_PATH_AUTO || path == DDMPC_PATH_WS)
                           ? closed_loop_fast_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                                  scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, tol,
                                                  max_iter, st)
                           : -1;
        if (rc != -1) return rc;
        return fail(DDMPC_ERR_NOT_IMPLEMENTED,
                    "trajectory_layout = step-major is implemented by k_closed_loop_ws only (one shared ROBUST controller without "
                    "slack bound, four-tank n-step shape, paths auto / ws)");
    }
    if (path != DDMPC_PATH_GENERIC) {
        if (path == DDMPC_PATH_TC) {   // opt-in: config-4 shape on tcgen05 / TMEM with TF32x3 arithmetic
            const int rc = closed_loop_tc_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                              scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, st);
            if (rc != -1) return rc;
        }
        {   // shared CONVEX controller, four-tank n-step shape: slack check and box-row ADMM on the FP64 tensor cores
            const int rc = closed_loop_cvx_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                               scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, tol,
                                               max_iter, st);
            if (rc != -1) return rc;
        }
        {   // 8 lanes per loop: per-loop controllers (any batch; also their CONVEX variant), NOMINAL, small batches of a
            // shared controller
            const int rc = closed_loop_perloop_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                                   scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, tol,
                                                   max_iter, st);
            if (rc != -1) return rc;
        }
        {   // register-resident / warp-specialised kernels (shared ROBUST controller, four-tank-size system)
            const int rc = closed_loop_fast_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                                scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, tol,
                                                max_iter, st);
            if (rc != -1) return rc;
        }
        {   // large systems, compiled shapes: one fused launch, a warp per 8 loops, both products on the FP64 tensor cores
            const int rc = closed_loop_dmma_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                                scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, st);
            if (rc != -1) return rc;
        }
        {   // large systems: the batch as the N dimension of FP64 tensor-core GEMMs
            const int rc = closed_loop_gemm_try(set, plant, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, w, noise_seed,
                                                scenario_id0, noise_eps, n_steps, u_sys, y_sys, status, iters, x_final, st);
            if (rc != -1) return rc;
        }
    }
    // plant matrices -> device (small; one staging buffer)
    const size_t nA = (size_t)nxp * nxp, nB = (size_t)nxp * d.m, nC = (size_t)d.p * nxp, nD = (size_t)d.p * d.m;
    std::vector<double> hp(nA + nB + nC + nD);
    std::copy(plant->A, plant->A + nA, hp.begin());
    std::copy(plant->B, plant->B + nB, hp.begin() + nA);
    std::copy(plant->C, plant->C + nC, hp.begin() + nA + nB);
    std::copy(plant->D, plant->D + nD, hp.begin() + nA + nB + nC);
    if (hp != set->plant_host) {
        // a different plant: wait for loops still reading the old one, then replace it
        DDMPC_CUDA(cudaDeviceSynchronize());
        DDMPC_CUDA(set->plant_dev.alloc(sizeof(double) * hp.size()));
        DDMPC_CUDA(cudaMemcpy(set->plant_dev.p, hp.data(), sizeof(double) * hp.size(), cudaMemcpyHostToDevice));
        set->plant_host = hp;
    }
    const DevBuf &dp = set->plant_dev;
    KArgs a = make_kargs(set, tol, max_iter);
    LoopArgs la{};
    la.n_x = nxp; la.n_steps = n_steps; la.n_mpc = n_mpc;
    la.A = dp.d(); la.Bm = dp.d() + nA; la.C = dp.d() + nA + nB; la.D = dp.d() + nA + nB + nC;
    la.w = w; la.seed = noise_seed; la.id0 = scenario_id0; la.eps = noise_eps;
    size_t smem = 0;
    const size_t per_thread = (size_t)a.nth + 2 * (size_t)nxp + (size_t)n_mpc * d.m + d.p + 4 * (size_t)a.nb;
    const int tpb = pick_tpb(per_thread, &smem);
    if (!tpb) return fail(DDMPC_ERR_INVALID_ARG, "closed_loop_batch: problem too large for the generic kernel");
    DDMPC_CUDA(cudaFuncSetAttribute(k_closed_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_closed_loop<<<ceil_div(B, tpb), tpb, smem, st>>>(a, la, B, ctrl_idx, x0, u_past0, y_past0, u_s, y_s, u_sys, y_sys,
                                                        status, iters, x_final);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

int ddmpc_closed_loop_batch_host(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int32_t *ctrl_idx,
                                 const double *x0, const double *u_past0, const double *y_past0, const double *u_s,
                                 const double *y_s, const double *w, uint64_t noise_seed, uint64_t scenario_id0,
                                 double noise_eps, int n_steps, double tol, int max_iter, double *u_sys, double *y_sys,
                                 int32_t *status, int32_t *iters, double *x_final) {
    if (!set || !plant || B <= 0 || n_steps <= 0 || !u_sys || !y_sys)
        return fail(DDMPC_ERR_INVALID_ARG, "closed_loop_batch_host: bad argument");
    const Dims &d = set->plan.d;
    Stage s;
    int32_t *dc, *dst, *dit;
    double *dx0, *dup, *dyp, *dus, *dys, *dw, *dU, *dY, *dxf;
    DDMPC_TRY(s.up(ctrl_idx, (size_t)B, &dc));
    DDMPC_TRY(s.up(x0, (size_t)B * plant->n_x, &dx0));
    DDMPC_TRY(s.up(u_past0, (size_t)B * d.n * d.m, &dup));
    DDMPC_TRY(s.up(y_past0, (size_t)B * d.n * d.p, &dyp));
    DDMPC_TRY(s.up(u_s, (size_t)B * d.m, &dus));
    DDMPC_TRY(s.up(y_s, (size_t)B * d.p, &dys));
    DDMPC_TRY(s.up(w, (size_t)B * n_steps * d.p, &dw));
    DDMPC_TRY(s.out((size_t)B * n_steps * d.m, &dU));
    DDMPC_TRY(s.out((size_t)B * n_steps * d.p, &dY));
    DDMPC_TRY(s.out((size_t)B, &dst, status != nullptr));
    DDMPC_TRY(s.out((size_t)B, &dit, iters != nullptr));
    DDMPC_TRY(s.out((size_t)B * plant->n_x, &dxf, x_final != nullptr));
    DDMPC_TRY(ddmpc_closed_loop_batch(set, plant, B, dc, dx0, dup, dyp, dus, dys, dw, noise_seed, scenario_id0,
                                      noise_eps, n_steps, tol, max_iter, dU, dY, dst, dit, dxf, nullptr));
    DDMPC_CUDA(cudaStreamSynchronize(nullptr));
    DDMPC_CUDA(cudaMemcpy(u_sys, dU, sizeof(double) * B * n_steps * d.m, cudaMemcpyDeviceToHost));
    DDMPC_CUDA(cudaMemcpy(y_sys, dY, sizeof(double) * B * n_steps * d.p, cudaMemcpyDeviceToHost));
    if (status) DDMPC_CUDA(cudaMemcpy(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
    if (iters) DDMPC_CUDA(cudaMemcpy(iters, dit, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
    if (x_final) DDMPC_CUDA(cudaMemcpy(x_final, dxf, sizeof(double) * B * plant->n_x, cudaMemcpyDeviceToHost));
    return DDMPC_OK;
}

}  // extern "C"
