// Closed loops of LARGE systems (BASELINE config 4: n = 20, m = p = 4, L = 40) as a sequence of
// FP64 tensor-core GEMMs over the whole batch.
//
// The per-loop state of such a system (n_theta = 168 window numbers + 20 plant states) no longer
// fits in registers, and the gain block that is applied each solve (n_mpc*m x n_theta, up to 80 x 168)
// no longer fits in a kernel parameter.  Here the batch is the GEMM N dimension:
//
//   per MPC iteration (s = n_mpc steps, or the remainder in the last iteration)
//     1. Uplan (s*m x B)      = Ku[0:s*m, :] (s*m x n_theta)  *  ThetaT (n_theta x B)        k_gemm (DMMA)
//     2. O ((s*p + n_x) x B)  = Mblk(s)                       *  V = [X; Uplan]              k_gemm (DMMA)
//        Mblk(s) is the s-step block map of the LTI plant: rows y_0..y_{s-1} then x_s,
//        columns x_0 then u_0..u_{s-1}  (utilities/model_simulation.py:93-98 unrolled s times;
//        measurement noise only enters y, so it is added afterwards).
//     3. k_block_finish: y += noise, record (u, y), slide the measurement window inside ThetaT,
//        X <- x_s.
//
// State is stored value-major ([value][loop]) so that both GEMM operands are read with unit stride
// along the batch.  Equality-only controllers shared by the whole batch only (ROBUST / slack NONE);
// everything else takes the generic thread-per-loop kernel in solve.cu.
//
// Replaces the same reference code as k_closed_loop (solve.cu).
#include <vector>

#include "linalg.cuh"
#include "plan.cuh"

namespace ddmpc {

// ThetaT[i][b] <- theta_b[i];  V[j][b] <- x0_b[j]
__global__ void k_gl_init(int B, int nm, int npp, int m, int p, int nxp, const double *__restrict__ u_past0,
                          const double *__restrict__ y_past0, const double *__restrict__ u_s,
                          const double *__restrict__ y_s, const double *__restrict__ x0, double *__restrict__ ThetaT,
                          double *__restrict__ V) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int r = 0;
    for (int i = 0; i < nm; ++i) ThetaT[(size_t)(r++) * B + b] = u_past0[(size_t)b * nm + i];
    for (int i = 0; i < npp; ++i) ThetaT[(size_t)(r++) * B + b] = y_past0[(size_t)b * npp + i];
    for (int i = 0; i < m; ++i) ThetaT[(size_t)(r++) * B + b] = u_s[(size_t)b * m + i];
    for (int i = 0; i < p; ++i) ThetaT[(size_t)(r++) * B + b] = y_s[(size_t)b * p + i];
    for (int i = 0; i < nxp; ++i) V[(size_t)i * B + b] = x0[(size_t)b * nxp + i];
}

__device__ __forceinline__ void gl_philox(uint32_t c0, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                          uint32_t out[4]) {
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

// One thread per loop: finish an s-step block.
//   O rows: [y_0 (p) .. y_{s-1} (p); x_s (n_x)],  V rows: [x (n_x); u_0 (m) .. u_{s-1} (m)]
__global__ void k_block_finish(int B, int n, int m, int p, int nxp, int s, int t0, int n_steps,
                               const double *__restrict__ w, unsigned long long seed, unsigned long long id0,
                               double eps, double *__restrict__ O, double *__restrict__ V,
                               double *__restrict__ ThetaT, double *__restrict__ u_sys, double *__restrict__ y_sys,
                               int *__restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long sid = id0 + (unsigned long long)b;
    const int nm = n * m, npp = n * p;
    bool fin = true;
    // outputs: add the measurement noise, record
    uint32_t o[4];
    unsigned last = 0xffffffffu;
    for (int k = 0; k < s; ++k) {
        const size_t f = (size_t)b * n_steps + (t0 + k);
        for (int i = 0; i < p; ++i) {
            double nz;
            if (w) {
                nz = w[f * p + i];
            } else {
                const unsigned q = (unsigned)(t0 + k) * (unsigned)p + (unsigned)i;
                if ((q >> 2) != last) {
                    last = q >> 2;
                    gl_philox(last, (uint32_t)(sid & 0xffffffffu), (uint32_t)(sid >> 32), (uint32_t)(seed & 0xffffffffu),
                              (uint32_t)(seed >> 32), o);
                }
                const unsigned l = q & 3u;
                const uint32_t word = l == 0 ? o[0] : (l == 1 ? o[1] : (l == 2 ? o[2] : o[3]));
                nz = eps * (2.0 * __hiloint2double((int)(0x3FF00000u | (word >> 12)), (int)(word << 20)) - 3.0);
            }
            const double y = O[(size_t)(k * p + i) * B + b] + nz;
            O[(size_t)(k * p + i) * B + b] = y;
            y_sys[f * p + i] = y;
            fin = fin && isfinite(y);
        }
        for (int i = 0; i < m; ++i) u_sys[f * m + i] = V[(size_t)(nxp + k * m + i) * B + b];
    }
    // slide the measurement window (controller.py:893-895) by s steps
    if (s >= n) {
        for (int j = 0; j < n; ++j) {
            const int k = s - n + j;
            for (int i = 0; i < m; ++i) ThetaT[(size_t)(j * m + i) * B + b] = V[(size_t)(nxp + k * m + i) * B + b];
            for (int i = 0; i < p; ++i) ThetaT[(size_t)(nm + j * p + i) * B + b] = O[(size_t)(k * p + i) * B + b];
        }
    } else {
        for (int j = 0; j < n - s; ++j) {
            for (int i = 0; i < m; ++i) ThetaT[(size_t)(j * m + i) * B + b] = ThetaT[(size_t)((j + s) * m + i) * B + b];
            for (int i = 0; i < p; ++i)
                ThetaT[(size_t)(nm + j * p + i) * B + b] = ThetaT[(size_t)(nm + (j + s) * p + i) * B + b];
        }
        for (int k = 0; k < s; ++k) {
            const int j = n - s + k;
            for (int i = 0; i < m; ++i) ThetaT[(size_t)(j * m + i) * B + b] = V[(size_t)(nxp + k * m + i) * B + b];
            for (int i = 0; i < p; ++i) ThetaT[(size_t)(nm + j * p + i) * B + b] = O[(size_t)(k * p + i) * B + b];
        }
    }
    // plant state for the next block
    for (int i = 0; i < nxp; ++i) {
        const double xv = O[(size_t)(s * p + i) * B + b];
        V[(size_t)i * B + b] = xv;
        fin = fin && isfinite(xv);
    }
    if (!fin && status) status[b] = DDMPC_SOLVE_NONFINITE;
    (void)npp;
}

__global__ void k_gl_final(int B, int nxp, int iters_val, const double *__restrict__ V, int *__restrict__ status_init,
                           int *__restrict__ iters, double *__restrict__ x_final, int phase) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (phase == 0) {
        if (status_init) status_init[b] = DDMPC_SOLVE_OPTIMAL;
        return;
    }
    if (iters) iters[b] = iters_val;
    if (x_final)
        for (int i = 0; i < nxp; ++i) x_final[(size_t)b * nxp + i] = V[(size_t)i * B + b];
}

// s-step block map of the plant, row-major ((s*p + n_x) x (n_x + s*m)), built on the host
std::vector<double> block_map(const ddmpc_plant *pl, int s) {
    const int nx = pl->n_x, m = pl->m, p = pl->p;
    const int rows = s * p + nx, cols = nx + s * m;
    std::vector<double> M((size_t)rows * cols, 0.0);
    // Apow[k] = A^k
    std::vector<std::vector<double>> Apow(s + 1, std::vector<double>((size_t)nx * nx, 0.0));
    for (int i = 0; i < nx; ++i) Apow[0][(size_t)i * nx + i] = 1.0;
    for (int k = 1; k <= s; ++k)
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) {
                double acc = 0.0;
                for (int l = 0; l < nx; ++l) acc += pl->A[i * nx + l] * Apow[k - 1][(size_t)l * nx + j];
                Apow[k][(size_t)i * nx + j] = acc;
            }
    // AB[k] = A^k B  (nx x m)
    std::vector<std::vector<double>> AB(s, std::vector<double>((size_t)nx * m, 0.0));
    for (int k = 0; k < s; ++k)
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < m; ++j) {
                double acc = 0.0;
                for (int l = 0; l < nx; ++l) acc += Apow[k][(size_t)i * nx + l] * pl->B[l * m + j];
                AB[k][(size_t)i * m + j] = acc;
            }
    auto at = [&](int r, int c) -> double & { return M[(size_t)r * cols + c]; };
    for (int k = 0; k < s; ++k) {
        for (int i = 0; i < p; ++i) {
            // y_k = C A^k x_0 + sum_{j<k} C A^{k-1-j} B u_j + D u_k
            for (int c = 0; c < nx; ++c) {
                double acc = 0.0;
                for (int l = 0; l < nx; ++l) acc += pl->C[i * nx + l] * Apow[k][(size_t)l * nx + c];
                at(k * p + i, c) = acc;
            }
            for (int j = 0; j < k; ++j)
                for (int c = 0; c < m; ++c) {
                    double acc = 0.0;
                    for (int l = 0; l < nx; ++l) acc += pl->C[i * nx + l] * AB[k - 1 - j][(size_t)l * m + c];
                    at(k * p + i, nx + j * m + c) = acc;
                }
            for (int c = 0; c < m; ++c) at(k * p + i, nx + k * m + c) = pl->D[i * m + c];
        }
    }
    for (int i = 0; i < nx; ++i) {
        for (int c = 0; c < nx; ++c) at(s * p + i, c) = Apow[s][(size_t)i * nx + c];
        for (int j = 0; j < s; ++j)
            for (int c = 0; c < m; ++c) at(s * p + i, nx + j * m + c) = AB[s - 1 - j][(size_t)i * m + c];
    }
    return M;
}

// Returns DDMPC_OK when handled, -1 when this path does not apply.
int closed_loop_gemm_try(const ddmpc_set *set, const ddmpc_plant *plant, int B, const int *ctrl_idx, const double *x0,
                         const double *u_past0, const double *y_past0, const double *u_s, const double *y_s,
                         const double *w, uint64_t seed, uint64_t id0, double eps, int n_steps, double *u_sys,
                         double *y_sys, int *status, int *iters, double *x_final, cudaStream_t st) {
    const Dims &d = set->plan.d;
    if (ctrl_idx || set->plan.count != 1 || d.nb > 0 || !d.robust) return -1;
    if (set->opt_path != DDMPC_PATH_AUTO && set->opt_path != DDMPC_PATH_GEMM) return -1;
    const int n = d.n, m = d.m, p = d.p, nxp = plant->n_x, nth = d.nth;
    const int nmpc = set->prm.n_mpc_step;
    // worth it only when the per-loop state is too large for the thread-per-loop kernels and the
    // applied gain block has enough rows to fill a tensor-core tile (n-step schemes)
    if (!(nth >= 64 && B >= 512 && nmpc * m >= 32)) return -1;
    const int rem = n_steps % nmpc;
    const int rows_full = nmpc * p + nxp, cols_full = nxp + nmpc * m;

    // block maps of the plant: constant per (plant, n_mpc), cached in the set.  The loop state (ThetaT, V, O) is
    // per-call scratch, allocated and freed in stream order on the caller's stream: closed loops of one set may run
    // on several streams at once (ControllerSet.closed_loop_host alternates its chunks between two streams).
    const size_t nMf = (size_t)rows_full * cols_full;
    const size_t nMr = rem ? (size_t)(rem * p + nxp) * (nxp + rem * m) : 0;
    std::vector<double> hM = block_map(plant, nmpc);
    if (rem) {
        std::vector<double> hr = block_map(plant, rem);
        hM.insert(hM.end(), hr.begin(), hr.end());
    }
    if (set->gemm_host != hM) {
        DDMPC_CUDA(cudaDeviceSynchronize());   // loops still reading the previous maps
        DDMPC_CUDA(set->gemm_ws.alloc((nMf + nMr) * sizeof(double)));
        DDMPC_CUDA(cudaMemcpy(set->gemm_ws.p, hM.data(), sizeof(double) * hM.size(), cudaMemcpyHostToDevice));
        set->gemm_host = hM;
    }
    const double *Mf = set->gemm_ws.d(), *Mr = Mf + nMf;
    DDMPC_CUDA(pool_ready());
    double *scratch = nullptr;
    DDMPC_CUDA(cudaMallocAsync((void **)&scratch, sizeof(double) * (size_t)B * (nth + cols_full + rows_full), st));
    struct ScratchFree {
        double *p;
        cudaStream_t s;
        ~ScratchFree() { cudaFreeAsync(p, s); }           // stream-ordered: after the last kernel enqueued below
    } scratch_free{scratch, st};
    double *ThetaT = scratch, *V = ThetaT + (size_t)B * nth, *O = V + (size_t)B * cols_full;

    const int T = 128, G = ceil_div(B, T);
    k_gl_final<<<G, T, 0, st>>>(B, nxp, 0, V, status, nullptr, nullptr, 0);
    DDMPC_LAUNCH_CHECK();
    k_gl_init<<<G, T, 0, st>>>(B, n * m, n * p, m, p, nxp, u_past0, y_past0, u_s, y_s, x0, ThetaT, V);
    DDMPC_LAUNCH_CHECK();
    const double *Ku = set->plan.Ku.d();
    int n_iter = 0;
    for (int t0 = 0; t0 < n_steps; t0 += nmpc, ++n_iter) {
        const int s = std::min(nmpc, n_steps - t0);
        const double *Mb = (s == nmpc) ? Mf : Mr;
        const int rows = s * p + nxp, cols = nxp + s * m;
        // 1. planned inputs of the block: V[nxp : nxp + s*m, :] = Ku[0 : s*m, :] * ThetaT
        DDMPC_TRY(gemm(st, 1, s * m, B, nth, 1.0, mat(Ku, nth, 1, 0), mat(ThetaT, B, 1, 0), 0.0,
                       V + (size_t)nxp * B, B, 1, 0));
        // 2. s plant steps at once: O = Mblk(s) * V
        DDMPC_TRY(gemm(st, 1, rows, B, cols, 1.0, mat(Mb, cols, 1, 0), mat(V, B, 1, 0), 0.0, O, B, 1, 0));
        // 3. noise, record, window, state
        k_block_finish<<<G, T, 0, st>>>(B, n, m, p, nxp, s, t0, n_steps, w, seed, id0, eps, O, V, ThetaT, u_sys, y_sys,
                                        status);
        DDMPC_LAUNCH_CHECK();
    }
    k_gl_final<<<G, T, 0, st>>>(B, nxp, n_iter, V, nullptr, iters, x_final, 1);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

}  // namespace ddmpc
