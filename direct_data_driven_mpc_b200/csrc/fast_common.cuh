// Shared pieces of the fused four-tank-size closed-loop kernels (fast_loop.cu, and the measured-and-dropped
// variants kept under experiments/): coefficient / argument structs, Philox round, sector-paired stores, the
// s-step block map of the plant.
#pragma once
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
    // CONVEX slack bound (device operators of controller 0, see plan.cuh)
    const double *Ks, *Phi, *Psi;   // (nb, nth), (nb, nb), (L*m, nb)
    double bound, tol;
    int nb, nth, max_iter;
    int step_major;      // 1: trajectories stored (n_steps, B, m) instead of (B, n_steps, m); k_closed_loop_ws only
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

template <int N, int M, int P, int NX, int NMPC>
struct MmaCoef {
    double Ku[NMPC * M][N * (M + P)];               // gain rows on [u_past; y_past]
    double Mb[NMPC * P + NX][NX + NMPC * M];        // block map, full block
    double Mt[NMPC * P + NX][NX + NMPC * M];        // block map of the last, partial block (n_tail steps), zero padded
};

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

}  // namespace ddmpc
