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
#include "common.cuh"
#include "plan.cuh"

namespace ddmpc {

template <int N, int M, int P, int NX, int NMPC>
struct FastCoef {
    double Kw[NMPC * M][N * (M + P)];  // rows of Ku acting on [u_past; y_past]
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
};

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                             uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}

__device__ __forceinline__ double unit12_fast(uint32_t hi, uint32_t lo) {
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}

// One trajectory element (EL doubles) per step.  With EL == 2 an element is 16 B and the
// elements f-1, f (f odd) fill one 32 B sector: `pend` carries the even element until its
// partner arrives and the pair leaves as a single STG.256.
template <int EL, bool PAIR>
__device__ __forceinline__ void emit(double *__restrict__ base, size_t f, bool odd, bool first, double (&pend)[EL],
                                     const double (&cur)[EL]) {
    if constexpr (EL == 2 && PAIR) {
        if (odd) {
            if (first) {
                *reinterpret_cast<double2 *>(base + f * 2) = make_double2(cur[0], cur[1]);
            } else {
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(base + (f - 1) * 2), "d"(pend[0]),
                             "d"(pend[1]), "d"(cur[0]), "d"(cur[1])
                             : "memory");
            }
        }
        pend[0] = cur[0];
        pend[1] = cur[1];
    } else {
#pragma unroll
        for (int i = 0; i < EL; ++i) base[f * EL + i] = cur[i];
    }
}

template <int N, int M, int P, int NX, int NMPC, bool PHILOX, bool PAIR>
__global__ void __maxnreg__(144)
k_closed_loop_fast(const __grid_constant__ FastCoef<N, M, P, NX, NMPC> cf, const FastArgs a) {
    // thread -> loop map: first half of the block takes even loops, second half odd loops, so
    // the sector parity of a step is uniform across a warp
    const int half = blockDim.x >> 1;
    const int tl = threadIdx.x;
    const int b = blockIdx.x * blockDim.x + 2 * (tl % half) + (tl / half);
    if (b >= a.B) return;
    double x[NX], wu[N * M], wy[N * P], csp[NMPC * M];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = a.x0[(size_t)b * NX + i];
#pragma unroll
    for (int i = 0; i < N * M; ++i) wu[i] = a.u_past0[(size_t)b * N * M + i];
#pragma unroll
    for (int i = 0; i < N * P; ++i) wy[i] = a.y_past0[(size_t)b * N * P + i];
    {
        double sp[M + P];
#pragma unroll
        for (int i = 0; i < M; ++i) sp[i] = a.u_s[(size_t)b * M + i];
#pragma unroll
        for (int i = 0; i < P; ++i) sp[M + i] = a.y_s[(size_t)b * P + i];
#pragma unroll
        for (int k = 0; k < NMPC * M; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < M + P; ++j) acc = fma(__ldg(a.Ksp + k * (M + P) + j), sp[j], acc);
            csp[k] = acc;
        }
    }
    const size_t f0 = (size_t)b * a.n_steps;
    const unsigned long long sid = a.id0 + (unsigned long long)b;
    const uint32_t sid_lo = (uint32_t)sid, sid_hi = (uint32_t)(sid >> 32);
    double pu[M], py[P], up[NMPC * M];
#pragma unroll
    for (int i = 0; i < M; ++i) pu[i] = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) py[i] = 0.0;

    // ---- QP solve (equality-only => affine in the window): planned inputs
    auto solve = [&]() {
#pragma unroll
        for (int k = 0; k < NMPC * M; ++k) {
            double acc0 = csp[k], acc1 = 0.0;
#pragma unroll
            for (int j = 0; j < N * M; ++j) {
                if (j & 1) acc1 = fma(cf.Kw[k][j], wu[j], acc1);
                else acc0 = fma(cf.Kw[k][j], wu[j], acc0);
            }
#pragma unroll
            for (int j = 0; j < N * P; ++j) {
                if (j & 1) acc1 = fma(cf.Kw[k][N * M + j], wy[j], acc1);
                else acc0 = fma(cf.Kw[k][N * M + j], wy[j], acc0);
            }
            up[k] = acc0 + acc1;
        }
    };
    // ---- one plant step with the s-th planned input: noise, y, x, record, window shift
    auto step = [&](const int s, const int k) {
        double u[M], y[P];
#pragma unroll
        for (int i = 0; i < M; ++i) u[i] = up[s * M + i];
        if constexpr (!PHILOX) {
#pragma unroll
            for (int i = 0; i < P; ++i) y[i] = __ldg(a.w + (f0 + k) * P + i);
        } else {
#pragma unroll
            for (int ch = 0; ch < (P + 1) / 2; ++ch) {
                uint32_t c0 = (uint32_t)k, c1 = (uint32_t)ch, c2 = sid_lo, c3 = sid_hi;
#pragma unroll
                for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
                // same order of operations as the oracle: eps * (2 v - 3)
                y[2 * ch] = a.eps * (2.0 * unit12_fast(c0, c1) - 3.0);
                if (2 * ch + 1 < P) y[2 * ch + 1] = a.eps * (2.0 * unit12_fast(c2, c3) - 3.0);
            }
        }
        // y = C x + D u + w   (pre-update state; model_simulation.py:94)
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double acc = 0.0, acd = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(cf.C[i][j], x[j], acc);
#pragma unroll
            for (int j = 0; j < M; ++j) acd = fma(cf.D[i][j], u[j], acd);
            y[i] = (acc + acd) + y[i];
        }
        // x <- A x + B u      (model_simulation.py:96)
        double xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = 0.0, acb = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(cf.A[i][j], x[j], acc);
#pragma unroll
            for (int j = 0; j < M; ++j) acb = fma(cf.B[i][j], u[j], acb);
            xn[i] = acc + acb;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
        // record (full-sector stores)
        const size_t f = f0 + k;
        const bool odd = (f & 1) != 0;
        emit<M, PAIR>(a.u_sys, f, odd, k == 0, pu, u);
        emit<P, PAIR>(a.y_sys, f, odd, k == 0, py, y);
        // window shift (controller.py:893-895)
#pragma unroll
        for (int i = 0; i < N * M - M; ++i) wu[i] = wu[i + M];
#pragma unroll
        for (int i = 0; i < M; ++i) wu[N * M - M + i] = u[i];
#pragma unroll
        for (int i = 0; i < N * P - P; ++i) wy[i] = wy[i + P];
#pragma unroll
        for (int i = 0; i < P; ++i) wy[N * P - P + i] = y[i];
    };

    int solves = 0, t0 = 0;
    for (; t0 + NMPC <= a.n_steps; t0 += NMPC) {   // full n-step blocks: no guards
        solve();
        ++solves;
#pragma unroll
        for (int s = 0; s < NMPC; ++s) step(s, t0 + s);
    }
    if (t0 < a.n_steps) {                          // last, partial block (controller_operation.py:278)
        solve();
        ++solves;
#pragma unroll
        for (int s = 0; s < NMPC; ++s)
            if (t0 + s < a.n_steps) step(s, t0 + s);
    }
    if constexpr (PAIR) {   // an unpaired final element still sits in pend
        const size_t fl = f0 + a.n_steps - 1;
        if ((fl & 1) == 0) {
            if constexpr (M == 2) *reinterpret_cast<double2 *>(a.u_sys + fl * 2) = make_double2(pu[0], pu[1]);
            if constexpr (P == 2) *reinterpret_cast<double2 *>(a.y_sys + fl * 2) = make_double2(py[0], py[1]);
        }
    }
    bool finite = true;
#pragma unroll
    for (int i = 0; i < NX; ++i) finite = finite && isfinite(x[i]);
#pragma unroll
    for (int i = 0; i < N * P; ++i) finite = finite && isfinite(wy[i]);
    if (a.status) a.status[b] = finite ? DDMPC_SOLVE_OPTIMAL : DDMPC_SOLVE_NONFINITE;
    if (a.iters) a.iters[b] = solves;
    if (a.x_final) {
#pragma unroll
        for (int i = 0; i < NX; ++i) a.x_final[(size_t)b * NX + i] = x[i];
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
        for (int j = 0; j < NW; ++j) cf.Kw[k][j] = cache[(size_t)k * d.nth + j];
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
    const int tpb = 64;
    const dim3 grid(ceil_div(a.B, tpb));
    // 32-byte pairing needs 16-byte elements (M == 2 and P == 2) and 32-byte aligned outputs
    const bool pair = (M == 2 && P == 2) && ((reinterpret_cast<uintptr_t>(a.u_sys) & 31) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.y_sys) & 31) == 0);
    if (a.w) {
        if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, false, true><<<grid, tpb, 0, st>>>(cf, a);
        else k_closed_loop_fast<N, M, P, NX, NMPC, false, false><<<grid, tpb, 0, st>>>(cf, a);
    } else {
        if (pair) k_closed_loop_fast<N, M, P, NX, NMPC, true, true><<<grid, tpb, 0, st>>>(cf, a);
        else k_closed_loop_fast<N, M, P, NX, NMPC, true, false><<<grid, tpb, 0, st>>>(cf, a);
    }
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// Returns DDMPC_OK when the fast kernel handled the call, -1 when it does not apply.
int closed_loop_fast_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st) {
    const Dims &d = set->plan.d;
    if (ctrl_idx || set->plan.count != 1 || d.convex || !d.robust) return -1;
    const char *force = getenv("DDMPC_FORCE_GENERIC");
    if (force && force[0] == '1') return -1;
    FastArgs fa{};
    fa.B = B; fa.n_steps = n_steps;
    fa.x0 = x0; fa.u_past0 = u_past0; fa.y_past0 = y_past0; fa.u_s = u_s; fa.y_s = y_s; fa.w = w;
    fa.seed = seed; fa.id0 = id0; fa.eps = eps;
    fa.u_sys = u_sys; fa.y_sys = y_sys; fa.x_final = x_final; fa.status = status; fa.iters = iters;
    const int nmpc = set->prm.n_mpc_step;
#define DDMPC_FAST_CASE(N_, M_, P_, NX_, NMPC_)                                                    \
    if (d.n == N_ && d.m == M_ && d.p == P_ && plant->n_x == NX_ && nmpc == NMPC_)                 \
        return launch_fast<N_, M_, P_, NX_, NMPC_>(set, plant, fa, st);
    DDMPC_FAST_CASE(4, 2, 2, 4, 4)   // four-tank, n-step scheme (configs 1, 3; TEC-n-step)
    DDMPC_FAST_CASE(4, 2, 2, 4, 1)   // four-tank, 1-step schemes (TEC, UCON)
    DDMPC_FAST_CASE(4, 2, 2, 4, 2)
    DDMPC_FAST_CASE(2, 1, 1, 2, 1)
    DDMPC_FAST_CASE(2, 1, 1, 2, 2)
#undef DDMPC_FAST_CASE
    return -1;
}

}  // namespace ddmpc
