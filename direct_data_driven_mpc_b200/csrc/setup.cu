// Per-controller setup on the GPU: Hankel matrices (bit-exact index kernel),
// persistency-of-excitation rank test, Gram matrix, factorisations and the
// condensed solve operators (DESIGN.md "Condensed formulation").
//
// Replaces direct_data_driven_mpc_controller.py:95-387 (constructor,
// evaluate_input_persistent_excitation, initialize_data_driven_mpc) and
// utilities/hankel_matrix.py:5-87 of the reference.
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "linalg.cuh"
#include "plan.cuh"

namespace ddmpc {

thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local cudaStream_t g_scratch_stream = nullptr;

// ---------------------------------------------------------------------------
// K1: Hankel gather.  H[row0 + r, col] = X[r + col * nch]   (hankel_matrix.py:47-51:
// column i is X[i:i+L,:].flatten()).  Pure FP64 copy -> bit-exact.
// One CTA writes HK_ROWS consecutive rows of one matrix: the (tiny) data sequence is staged in shared memory once,
// every output row is then written with unit-stride 8-byte stores, 256 B per warp instruction.  (The first version
// gave every output element its own thread in 128-thread CTAs: 835,584 CTAs for 4096 four-tank controllers and
// 1.9 TB/s; this one writes the same 0.76 GB at the HBM write rate.)  rowmap (optional) permutes output rows.
// ---------------------------------------------------------------------------
constexpr int HK_ROWS = 8;
__global__ void __launch_bounds__(256)
k_hankel(const double *__restrict__ X, long bsX, int nch, int rows, int cols, int row0,
         const int *__restrict__ rowmap, double *__restrict__ H, long ldH, long bsH, int in_smem) {
    extern __shared__ double xs[];
    const int nrb = (rows + HK_ROWS - 1) / HK_ROWS;          // 1-D grid: (matrix, row block)
    const long bz = blockIdx.x / nrb;
    const double *x = X + bz * bsX;
    const int r0 = (int)(blockIdx.x % nrb) * HK_ROWS, r1 = min(r0 + HK_ROWS, rows);
    // the rows r0..r1-1 read x[r0 .. r1-1 + (cols-1)*nch]
    const int lo = r0, hi = (r1 - 1) + (cols - 1) * nch + 1;
    if (in_smem) {
        for (int e = lo + threadIdx.x; e < hi; e += blockDim.x) xs[e - lo] = x[e];
        __syncthreads();
    }
    const double *src = in_smem ? xs - lo : x;
    double *Hb = H + bz * bsH;
    for (int r = r0; r < r1; ++r) {
        const int orow = rowmap ? rowmap[row0 + r] : row0 + r;
        double *dst = Hb + (long)orow * ldH;
        for (int col = threadIdx.x; col < cols; col += blockDim.x) dst[col] = src[(long)r + (long)col * nch];
    }
}

static int launch_hankel(cudaStream_t st, int batch, const double *X, long bsX, int N, int nch, int L,
                         int row0, const int *rowmap, double *H, long ldH, long bsH) {
    const int rows = L * nch, cols = N - L + 1;
    const size_t sh = sizeof(double) * ((size_t)(HK_ROWS - 1) + (size_t)(cols - 1) * nch + 1);
    const int in_smem = sh <= 96 * 1024;
    if (in_smem && sh > 48 * 1024) {
        static std::atomic<unsigned long long> attr_done{0};
        if (first_time_on_device(attr_done))
            DDMPC_CUDA(cudaFuncSetAttribute(k_hankel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    const long grid = (long)ceil_div(rows, HK_ROWS) * batch;
    k_hankel<<<(unsigned)grid, 256, in_smem ? sh : 0, st>>>(X, bsX, nch, rows, cols, row0, rowmap, H, ldH, bsH, in_smem);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// ---------------------------------------------------------------------------
// small fill / gather kernels
// ---------------------------------------------------------------------------
__global__ void k_copy_bcast(long n, const double *__restrict__ src, long bs_src, double *__restrict__ dst, long bs_dst) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[(long)blockIdx.y * bs_dst + e] = src[(long)blockIdx.y * bs_src + e];
}
static int copy_bcast(cudaStream_t st, int batch, long n, const double *src, long bs_src, double *dst, long bs_dst) {
    if (n <= 0) return DDMPC_OK;
    dim3 grid(ceil_div(n, 256), batch);
    k_copy_bcast<<<grid, 256, 0, st>>>(n, src, bs_src, dst, bs_dst);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// strided 2-D copy: dst[i*ldd + j] = src[i*lds + j]
__global__ void k_copy2d(int rows, int cols, const double *__restrict__ src, long lds, long bs_src,
                         double *__restrict__ dst, long ldd, long bs_dst) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)rows * cols) return;
    const int i = (int)(e / cols), j = (int)(e % cols);
    dst[(long)blockIdx.y * bs_dst + (long)i * ldd + j] = src[(long)blockIdx.y * bs_src + (long)i * lds + j];
}
static int copy2d(cudaStream_t st, int batch, int rows, int cols, const double *src, long lds, long bs_src,
                  double *dst, long ldd, long bs_dst) {
    if (rows <= 0 || cols <= 0) return DDMPC_OK;
    dim3 grid(ceil_div((long)rows * cols, 256), batch);
    k_copy2d<<<grid, 256, 0, st>>>(rows, cols, src, lds, bs_src, dst, ldd, bs_dst);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

__global__ void k_identity(int n, double *__restrict__ A, long ld, long bs, double diag) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    A[(long)blockIdx.y * bs + (long)i * ld + j] = (i == j) ? diag : 0.0;
}
static int set_identity(cudaStream_t st, int batch, int n, double *A, long ld, long bs, double diag = 1.0) {
    if (n <= 0) return DDMPC_OK;
    dim3 grid(ceil_div((long)n * n, 256), batch);
    k_identity<<<grid, 256, 0, st>>>(n, A, ld, bs, diag);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// A <- (A + A^T)/2 + shift*I
__global__ void k_symmetrize(int n, double *__restrict__ A, long ld, long bs, double shift) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    if (j > i) return;
    double *a = A + (long)blockIdx.y * bs;
    const double v = 0.5 * (a[(long)i * ld + j] + a[(long)j * ld + i]) + (i == j ? shift : 0.0);
    a[(long)i * ld + j] = v;
    a[(long)j * ld + i] = v;
}
static int symmetrize(cudaStream_t st, int batch, int n, double *A, long ld, long bs, double shift = 0.0) {
    if (n <= 0) return DDMPC_OK;
    dim3 grid(ceil_div((long)n * n, 256), batch);
    k_symmetrize<<<grid, 256, 0, st>>>(n, A, ld, bs, shift);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// Spectrum post-processing (one warp per batch entry):
//   mode 0: d[k] = lam[k] > rel*max ? 1 : 0        mode 1: d[k] = lam[k] > rel*max ? 1/lam[k] : 0
//   rank[b] = #{lam[k] > rel*max}
__global__ void k_spectrum(int n, const double *__restrict__ lam, long bsl, double rel, int mode,
                           double *__restrict__ d, long bsd, int *__restrict__ rank) {
    const int b = blockIdx.x, lane = threadIdx.x;
    lam += (long)b * bsl;
    double mx = 0.0;
    for (int k = lane; k < n; k += 32) mx = fmax(mx, lam[k]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const double thr = rel * mx;
    int cnt = 0;
    for (int k = lane; k < n; k += 32) {
        const bool keep = lam[k] > thr && mx > 0.0;
        cnt += keep ? 1 : 0;
        if (d) d[(long)b * bsd + k] = keep ? (mode ? 1.0 / lam[k] : 1.0) : 0.0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && rank) rank[b] = cnt;
}

struct FillArgs {
    int n, m, p, L, nu, ny, nx, nth, terminal, robust;
};

__device__ __forceinline__ double sym_at(const double *M, int ld, int a, int b) {
    return 0.5 * (M[(long)a * ld + b] + M[(long)b * ld + a]);
}

// cost weight D(i,j) of primal coordinates i,j (controller.py:703-719)
__device__ __forceinline__ double dval(const FillArgs &f, const double *R, const double *Q, double lamS, int i, int j) {
    const int nm = f.n * f.m, npp = f.n * f.p, Lm = f.L * f.m, Lpp = f.L * f.p;
    if (i < f.nu && j < f.nu) return (i >= nm && j >= nm) ? sym_at(R, Lm, i - nm, j - nm) : 0.0;
    if (i >= f.nu && i < f.nu + f.ny && j >= f.nu && j < f.nu + f.ny) {
        const int a = i - f.nu - npp, b = j - f.nu - npp;
        return (a >= 0 && b >= 0) ? sym_at(Q, Lpp, a, b) : 0.0;
    }
    if (i >= f.nu + f.ny && i == j) return lamS;
    return 0.0;
}

// Pperm[a][b] = D(perm a, perm b) + lamA * Om[tmap(perm a)][tmap(perm b)]
__global__ void k_fill_P(FillArgs f, const int *__restrict__ perm, const double *__restrict__ R,
                         const double *__restrict__ Q, const double *__restrict__ lamA,
                         const double *__restrict__ lamS, const double *__restrict__ Om, long bsOm,
                         double *__restrict__ P, long bsP) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)f.nx * f.nx) return;
    const int c = blockIdx.y;
    const int a = (int)(e / f.nx), b = (int)(e % f.nx);
    const int i = perm[a], j = perm[b];
    double v = dval(f, R, Q, lamS ? lamS[c] : 0.0, i, j);
    if (Om) {
        const int r = f.nu + f.ny;
        const int ti = i < r ? i : i - f.ny, tj = j < r ? j : j - f.ny;
        v = fma(lamA[c], Om[(long)c * bsOm + (long)ti * r + tj], v);
    }
    P[(long)c * bsP + e] = v;
}

// Cperm (selection of theta into fixed coordinates) and DTperm = D * Tile
__global__ void k_fill_CDT(FillArgs f, const int *__restrict__ perm, const double *__restrict__ R,
                           const double *__restrict__ Q, double *__restrict__ Cp, double *__restrict__ DTp) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)f.nx * f.nth) return;
    const int a = (int)(e / f.nth), t = (int)(e % f.nth);
    const int i = perm[a];
    const int nm = f.n * f.m, npp = f.n * f.p, Lm = f.L * f.m, Lpp = f.L * f.p;
    const int o_yp = nm, o_us = nm + npp, o_ys = nm + npp + f.m;
    int sel = -1;
    if (i < nm) sel = i;
    else if (i >= f.nu && i < f.nu + npp) sel = o_yp + (i - f.nu);
    else if (f.terminal && i >= Lm && i < f.nu) sel = o_us + (i - Lm) % f.m;
    else if (f.terminal && i >= f.nu + Lpp && i < f.nu + f.ny) sel = o_ys + (i - f.nu - Lpp) % f.p;
    Cp[e] = (sel == t) ? 1.0 : 0.0;
    double dt = 0.0;
    if (i >= nm && i < f.nu && t >= o_us && t < o_us + f.m) {
        const int j = t - o_us;
        for (int k = 0; k < f.L; ++k) dt += sym_at(R, Lm, i - nm, k * f.m + j);
    } else if (i >= f.nu + npp && i < f.nu + f.ny && t >= o_ys && t < o_ys + f.p) {
        const int j = t - o_ys;
        for (int k = 0; k < f.L; ++k) dt += sym_at(Q, Lpp, i - f.nu - npp, k * f.p + j);
    }
    DTp[e] = dt;
}

// ZT = Tile^T D Tile  (nth x nth)
__global__ void k_fill_ZT(FillArgs f, const int *__restrict__ invperm, const double *__restrict__ DTp,
                          double *__restrict__ ZT) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= f.nth * f.nth) return;
    const int t1 = e / f.nth, t2 = e % f.nth;
    const int nm = f.n * f.m, npp = f.n * f.p;
    const int o_us = nm + npp, o_ys = nm + npp + f.m;
    double v = 0.0;
    if (t1 >= o_us && t1 < o_us + f.m) {
        for (int k = 0; k < f.L; ++k) v += DTp[(long)invperm[(f.n + k) * f.m + (t1 - o_us)] * f.nth + t2];
    } else if (t1 >= o_ys) {
        for (int k = 0; k < f.L; ++k) v += DTp[(long)invperm[f.nu + (f.n + k) * f.p + (t1 - o_ys)] * f.nth + t2];
    }
    ZT[e] = v;
}

// Y (nf x nb) = B^T : unit columns at the free positions of sigma_pred
__global__ void k_fill_Bt(int nf, int nb, const int *__restrict__ bpos, double *__restrict__ Y, long bs) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)nf * nb) return;
    const int a = (int)(e / nb), j = (int)(e % nb);
    Y[(long)blockIdx.y * bs + e] = (bpos[j] == a) ? 1.0 : 0.0;
}

// Row scales S, Lam = sym(S B Y S), rho2 = nb / trace(Lam), M = I + rho2 Lam and the scaled bounds.  One CTA per
// controller.  The box rows come in up to three groups - sigma_pred (CONVEX), predicted inputs, predicted outputs -
// whose diagonal of Lam differs by orders of magnitude (1e-3 vs 4e2 for sigma vs inputs on the four-tank data), and a
// single ADMM penalty only suits all of them after equilibration: the first non-empty group keeps scale 1 (a
// CONVEX-only problem is untouched), every other group g gets sqrt(mean Lam_ref / mean Lam_gg).
__global__ void __launch_bounds__(256)
k_lam_rho(int nf, int nb, int nbs, int nbu, const int *__restrict__ bpos, const double *__restrict__ Y, long bsY,
          const double *__restrict__ blo, const double *__restrict__ bhi, double *__restrict__ Lam,
          double *__restrict__ Mm, double *__restrict__ rho2, double *__restrict__ rs, double *__restrict__ lo,
          double *__restrict__ hi, double *__restrict__ bmax) {
    const int c = blockIdx.x, tid = threadIdx.x;
    Y += (long)c * bsY;
    Lam += (long)c * nb * nb;
    Mm += (long)c * nb * nb;
    __shared__ double red[4][8];
    __shared__ double s_rho, s_sc[3];
    const int g1 = nbs, g2 = nbs + nbu;
    auto group = [&](int j) { return j < g1 ? 0 : (j < g2 ? 1 : 2); };
    double t[3] = {0.0, 0.0, 0.0}, bm = 0.0;
    for (int j = tid; j < nb; j += blockDim.x) t[group(j)] += Y[(long)bpos[j] * nb + j];
    for (int o = 16; o > 0; o >>= 1)
        for (int g = 0; g < 3; ++g) t[g] += __shfl_xor_sync(0xffffffffu, t[g], o);
    if ((tid & 31) == 0)
        for (int g = 0; g < 3; ++g) red[g][tid >> 5] = t[g];
    __syncthreads();
    if (tid == 0) {
        const int cnt[3] = {nbs, nbu, nb - nbs - nbu};
        double mean[3], tr = 0.0, ref = 0.0;
        for (int g = 0; g < 3; ++g) {
            double a = 0.0;
            for (int w = 0; w < (blockDim.x + 31) / 32; ++w) a += red[g][w];
            mean[g] = cnt[g] > 0 ? a / cnt[g] : 0.0;
            if (ref == 0.0 && mean[g] > 0.0) ref = mean[g];
        }
        for (int g = 0; g < 3; ++g) {
            s_sc[g] = (mean[g] > 0.0 && ref > 0.0) ? sqrt(ref / mean[g]) : 1.0;
            tr += s_sc[g] * s_sc[g] * mean[g] * cnt[g];
        }
        s_rho = (tr > 0.0) ? (double)nb / tr : 1.0;
        rho2[c] = s_rho;
    }
    __syncthreads();
    const double rho = s_rho;
    for (int j = tid; j < nb; j += blockDim.x) {
        const double r = s_sc[group(j)];
        const double l = r * blo[j], h = r * bhi[j];
        rs[(long)c * nb + j] = r;
        lo[(long)c * nb + j] = l;
        hi[(long)c * nb + j] = h;
        if (isfinite(l)) bm = fmax(bm, fabs(l));
        if (isfinite(h)) bm = fmax(bm, fabs(h));
    }
    for (int o = 16; o > 0; o >>= 1) bm = fmax(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if ((tid & 31) == 0) red[3][tid >> 5] = bm;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) m = fmax(m, red[3][w]);
        bmax[c] = m;
    }
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        const int i = e / nb, j = e % nb;
        const double v = s_sc[group(i)] * s_sc[group(j)] * 0.5 * (Y[(long)bpos[i] * nb + j] + Y[(long)bpos[j] * nb + i]);
        Lam[e] = v;
        Mm[e] = rho * v + (i == j ? 1.0 : 0.0);
    }
}

// Gather the operators the solver reads.
__global__ void k_extract(FillArgs f, int nf, int nb, int Lm, const int *__restrict__ invperm,
                          const int *__restrict__ bpos, const double *__restrict__ X0f, long bsX0f,
                          const double *__restrict__ Cp, const double *__restrict__ Y, long bsY,
                          const double *__restrict__ rho2, const double *__restrict__ rs,
                          double *__restrict__ Ku, double *__restrict__ X0, double *__restrict__ Ks,
                          double *__restrict__ Psi, double *__restrict__ Yf) {
    const int c = blockIdx.y;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nm = f.n * f.m;
    X0f += (long)c * bsX0f;
    if (e < (long)f.nx * f.nth) {
        const int i = (int)(e / f.nth), t = (int)(e % f.nth);
        const int a = invperm[i];
        const double v = a < nf ? X0f[(long)a * f.nth + t] : Cp[(long)a * f.nth + t];
        X0[(long)c * f.nx * f.nth + e] = v;
        if (i >= nm && i < f.nu) Ku[(long)c * Lm * f.nth + (long)(i - nm) * f.nth + t] = v;
    }
    if (nb > 0) {
        const double rho = rho2[c];
        Y += (long)c * bsY;
        if (e < (long)nb * f.nth) {
            const int j = (int)(e / f.nth), t = (int)(e % f.nth);
            Ks[(long)c * nb * f.nth + e] = rs[(long)c * nb + j] * X0f[(long)bpos[j] * f.nth + t];
        }
        if (e < (long)f.nx * nb) {
            const int i = (int)(e / nb), j = (int)(e % nb);
            const int a = invperm[i];
            const double v = a < nf ? rho * rs[(long)c * nb + j] * Y[(long)a * nb + j] : 0.0;
            Yf[(long)c * f.nx * nb + e] = v;
            if (i >= nm && i < f.nu) Psi[(long)c * Lm * nb + (long)(i - nm) * nb + j] = v;
        }
    }
}

// Fnz[c] = 1 when the feasibility map F of controller c has an entry above 1e-13 (rank-deficient H: some windows are
// not trajectories of the data); with noisy data H has full row rank, F vanishes and the per-solve check is skipped.
__global__ void k_flag_nonzero(long n, const double *__restrict__ F, int *__restrict__ flag) {
    const int c = blockIdx.x;
    F += (long)c * n;
    int nz = 0;
    for (long e = threadIdx.x; e < n; e += blockDim.x) nz |= fabs(F[e]) > 1e-13;
    nz = __syncthreads_or(nz);
    if (threadIdx.x == 0) flag[c] = nz;
}

__global__ void k_sub_identity(int n, double *__restrict__ A, long bs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(long)blockIdx.y * bs + (long)i * n + i] -= 1.0;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// gains of controllers whose setup failed -> NaN
__global__ void k_poison(int count, long per, const int *__restrict__ cstat, double *__restrict__ M) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (long)count * per && cstat[e / per] != 0) M[e] = __longlong_as_double(0x7ff8000000000000LL);
}

int closed_loop_fast_prepare(ddmpc_set *set, cudaStream_t st);   // fast_loop.cu
int closed_loop_cvx_prepare(ddmpc_set *set, cudaStream_t st);    // cvx_loop.cu

static Dims make_dims(const ddmpc_params &q) {
    Dims d{};
    d.n = q.n; d.m = q.m; d.p = q.p; d.N = q.N; d.L = q.L; d.Lp = q.L + q.n;
    d.nu = d.Lp * q.m; d.ny = d.Lp * q.p;
    d.r = d.nu + d.ny; d.cols = q.N - q.L - q.n + 1;
    d.robust = q.controller_type == DDMPC_ROBUST;
    d.convex = d.robust && q.slack_type == DDMPC_SLACK_CONVEX;
    d.terminal = q.use_terminal ? 1 : 0;
    d.nx = d.r + (d.robust ? d.ny : 0);
    d.nfix = q.n * (q.m + q.p) * (d.terminal ? 2 : 1);
    d.nf = d.nx - d.nfix;
    d.nth = q.n * (q.m + q.p) + q.m + q.p;
    d.nbs = d.convex ? q.L * q.p : 0;
    d.nbu = (q.u_min || q.u_max) ? (d.terminal ? (q.L - q.n) * q.m : q.L * q.m) : 0;
    d.nby = (q.y_min || q.y_max) ? (d.terminal ? (q.L - q.n) * q.p : q.L * q.p) : 0;
    d.nb = d.nbs + d.nbu + d.nby;
    d.Lm = q.L * q.m;
    return d;
}

static int validate(const ddmpc_params &q) {
    // order of checks = order in the reference constructor (controller.py:165-240)
    if (q.n <= 0 || q.m <= 0 || q.p <= 0 || q.L <= 0 || q.N <= 0 || q.n_mpc_step <= 0)
        return fail(DDMPC_ERR_INVALID_ARG, "n, m, p, L, N, n_mpc_step must be positive");
    if (q.controller_type != DDMPC_NOMINAL && q.controller_type != DDMPC_ROBUST)
        return fail(DDMPC_ERR_CONTROLLER_TYPE, "Unsupported controller type.");
    if (q.slack_type < 0 || q.slack_type > 2)
        return fail(DDMPC_ERR_SLACK_TYPE, "Unsupported slack variable constraint type.");
    if (q.controller_type == DDMPC_ROBUST &&
        (std::isnan(q.eps_max) || std::isnan(q.lamb_alpha) || std::isnan(q.lamb_sigma) || std::isnan(q.c)))
        return fail(DDMPC_ERR_ROBUST_PARAMS,
                    "All robust MPC parameters (eps_max, lamb_alpha, lamb_sigma, c) must be provided for a "
                    "'ROBUST' controller.");
    if (q.u_min || q.u_max || q.y_min || q.y_max) {
        if (q.controller_type != DDMPC_ROBUST)
            return fail(DDMPC_ERR_NOT_IMPLEMENTED, "The input / output box constraints are only implemented for 'ROBUST' controllers.");
        for (int j = 0; j < q.m && (q.u_min || q.u_max); ++j) {
            const double l = q.u_min ? q.u_min[j] : -INFINITY, h = q.u_max ? q.u_max[j] : INFINITY;
            if (std::isnan(l) || std::isnan(h) || l > h)
                return fail(DDMPC_ERR_INVALID_ARG, "input box: u_min[%d] = %g must not exceed u_max[%d] = %g", j, l, j, h);
        }
        for (int j = 0; j < q.p && (q.y_min || q.y_max); ++j) {
            const double l = q.y_min ? q.y_min[j] : -INFINITY, h = q.y_max ? q.y_max[j] : INFINITY;
            if (std::isnan(l) || std::isnan(h) || l > h)
                return fail(DDMPC_ERR_INVALID_ARG, "output box: y_min[%d] = %g must not exceed y_max[%d] = %g", j, l, j, h);
        }
    }
    return DDMPC_OK;
}

int pe_rank_device(cudaStream_t st, int batch, const double *X, long bsX, int N, int nch, int order,
                   int *rank_dev /* device, batch ints */) {
    const int rows = order * nch, cols = N - order + 1;
    DevBuf Hpe, G;
    DDMPC_CUDA(Hpe.alloc(sizeof(double) * (size_t)batch * rows * cols));
    DDMPC_CUDA(G.alloc(sizeof(double) * (size_t)batch * rows * rows));
    DDMPC_TRY(launch_hankel(st, batch, X, bsX, N, nch, order, 0, nullptr, Hpe.d(), cols, (long)rows * cols));
    Mat Hm = mat(Hpe.d(), cols, 1, (long)rows * cols);
    DDMPC_TRY(gemm(st, batch, rows, rows, cols, 1.0, Hm, tr(Hm), 0.0, G.d(), rows, 1, (long)rows * rows, nullptr, 0, true));
    DDMPC_TRY(symmetrize(st, batch, rows, G.d(), rows, (long)rows * rows));
    // Stage 0: Cholesky of G - tau I, tau = 1e-9 max diag(G) (blocked, trailing updates on the tensor cores).  Success
    // proves lambda_min(G) > tau (the rounding error of the factorisation is ~ n^2 eps max diag = 2e-11 for n = 320), far
    // above anything stage 1 or np.linalg.matrix_rank would call rank deficient, so such an entry is full rank and the
    // n^3 one-CTA elimination below (11 ms for the 320 x 320 Gram matrix of config 4) is skipped for it.
    {
        DevBuf G0, info0;
        const long sG = (long)rows * rows;
        DDMPC_CUDA(G0.alloc(sizeof(double) * (size_t)batch * sG));
        DDMPC_CUDA(info0.alloc(sizeof(int) * (size_t)batch));
        DDMPC_TRY(copy_bcast(st, batch, sG, G.d(), sG, G0.d(), sG));
        k_shift_diag<<<batch, 256, 0, st>>>(rows, G0.d(), rows, sG, 1e-9);
        DDMPC_LAUNCH_CHECK();
        DDMPC_TRY(potrf(st, batch, rows, G0.d(), rows, sG, info0.i()));
        k_rank_from_info<<<ceil_div(batch, 128), 128, 0, st>>>(batch, rows, info0.i(), rank_dev);
        DDMPC_LAUNCH_CHECK();
    }
    // Stage 1: number of pivots of a diagonally pivoted elimination of the Gram matrix above 1e-12 x the largest.
    // The Gram matrix squares the singular values, so this proves full rank only for sigma_min/sigma_max > 1e-6 -
    // true for every well-excited data set, and cheap (n^3 per controller, batched).
    // Stage 2 (only for entries stage 1 could not certify): pivoted Gram-Schmidt on H itself with the cut of
    // np.linalg.matrix_rank (hankel_matrix.py:82), so ill-conditioned but full-rank data is accepted exactly as
    // the reference accepts it, and rank-deficient data reports the reference's rank.
    DDMPC_TRY(pivot_rank(st, batch, rows, G.d(), rows, (long)rows * rows, 1e-12, rank_dev));
    DDMPC_TRY(rowqr_rank(st, batch, rows, cols, Hpe.d(), (long)rows * cols, rows, rank_dev));
    DDMPC_CUDA(cudaStreamSynchronize(st));
    return DDMPC_OK;
}

// Short data (fewer Hankel columns than rows): rows of the equality constraint "t = T x stays in range(H)", one per
// eigenvector of W, zero for the eigenvectors that span the range (keep[k] = 1):  Aeq[k][a] = (1 - keep[k]) V[trow(perm[a])][k]
__global__ void k_fill_Aeq(FillArgs f, const int *__restrict__ perm, const double *__restrict__ V, long bsV,
                           const double *__restrict__ keep, long bsk, double *__restrict__ Aeq, long bsA) {
    const int r = f.nu + f.ny, c = blockIdx.y;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)r * f.nx) return;
    const int k = (int)(e / f.nx), a = (int)(e % f.nx);
    const int i = perm[a], ti = i < r ? i : i - f.ny;
    Aeq[(long)c * bsA + e] = (1.0 - keep[(long)c * bsk + k]) * V[(long)c * bsV + (long)ti * r + k];
}

// dst (rows x cols, per batch entry) = src^T, src (cols x rows) with row stride lds
__global__ void k_transpose(int rows, int cols, const double *__restrict__ src, long lds, long bs_src,
                            double *__restrict__ dst, long bs_dst) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)rows * cols) return;
    const int i = (int)(e / cols), k = (int)(e % cols);
    dst[(long)blockIdx.y * bs_dst + e] = src[(long)blockIdx.y * bs_src + (long)k * lds + i];
}

// S[k][k] += keep[k]: the rows of the Schur complement that belong to no constraint become identity rows
__global__ void k_add_diag(int n, const double *__restrict__ dvec, long bsd, double *__restrict__ S, long bsS) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) S[(long)blockIdx.y * bsS + (long)k * n + k] += dvec[(long)blockIdx.y * bsd + k];
}

static int build_robust(cudaStream_t st, Plan &pl, const FillArgs &fa, const int *perm_d, const int *invperm_d,
                        const int *bpos_d, const double *Rd, const double *Qd, int *info_d) {
    const Dims &d = pl.d;
    const int C = pl.count, r = d.r, nx = d.nx, nf = d.nf, nfix = d.nfix, nth = d.nth, nb = d.nb;
    const long sW = (long)r * r, sP = (long)nx * nx;
    DevBuf Lw, P, Cp, DTp, ZT, Nn, X0f, G1, Y, Mm;
    DDMPC_CUDA(Lw.alloc(sizeof(double) * pl.data_count * sW));
    DDMPC_CUDA(P.alloc(sizeof(double) * C * sP));
    DDMPC_CUDA(Cp.alloc(sizeof(double) * (size_t)nx * nth));
    DDMPC_CUDA(DTp.alloc(sizeof(double) * (size_t)nx * nth));
    DDMPC_CUDA(ZT.alloc(sizeof(double) * (size_t)nth * nth));
    DDMPC_CUDA(Nn.alloc(sizeof(double) * (size_t)C * nf * nth));
    DDMPC_CUDA(X0f.alloc(sizeof(double) * (size_t)C * nf * nth));
    DDMPC_CUDA(G1.alloc(sizeof(double) * (size_t)C * nfix * nth));

    // W = H H^T ; Om = W^-1.  With one data set shared by the whole set (lambda sweeps, BASELINE config 5) the Hankel
    // stack, its Gram matrix, the Cholesky factor and the r x r inverse exist once instead of `count` times.
    const int CW = pl.data_count;
    const long sOm = CW == 1 ? 0 : sW;
    Mat Hm = mat(pl.H.d(), d.cols, 1, (long)r * d.cols);
    DDMPC_TRY(gemm(st, CW, r, r, d.cols, 1.0, Hm, tr(Hm), 0.0, pl.W.d(), r, 1, sW, nullptr, 0, true));
    DDMPC_TRY(symmetrize(st, CW, r, pl.W.d(), r, sW));
    DDMPC_TRY(copy_bcast(st, CW, sW, pl.W.d(), sW, Lw.d(), sW));
    // Short data - fewer Hankel columns than rows, which the reference accepts down to N_min (controller.py:275-283):
    // W is singular, ||alpha||^2 = t^T W^+ t on range(H) and t = T x has to stay there.  Om becomes the pseudo-inverse
    // (eigen-decomposition instead of Cholesky) and the null-space rows of W an equality constraint Aeq x = 0 that is
    // resolved further down by a Schur complement on top of the same reduced Hessian.
    // The same path takes over when W has enough columns but is numerically singular all the same (noise-free data
    // under a ROBUST controller): its Cholesky factorisation fails for some data set of the set.
    bool deficient = d.cols < r;
    DevBuf Vw, lw, dinv, keep, Aeq, Yq, Sq, Rq, AcC, Rb;
    const long sAq = (long)r * nx, sAe = CW == 1 ? 0 : sAq;
    if (!deficient) {
        DDMPC_TRY(potrf(st, CW, r, Lw.d(), r, sW, info_d));
        std::vector<int> hinfo(CW);
        DDMPC_CUDA(cudaMemcpyAsync(hinfo.data(), info_d, sizeof(int) * CW, cudaMemcpyDeviceToHost, st));
        DDMPC_CUDA(cudaStreamSynchronize(st));
        for (int v : hinfo) deficient = deficient || v != 0;
        if (deficient) {
            DDMPC_CUDA(cudaMemsetAsync(info_d, 0, sizeof(int) * C, st));
            DDMPC_TRY(copy_bcast(st, CW, sW, pl.W.d(), sW, Lw.d(), sW));
        }
    }
    pl.range_constrained = deficient;
    if (!deficient) {
        DDMPC_TRY(set_identity(st, CW, r, pl.Om.d(), r, sW));
        DDMPC_TRY(potrs(st, CW, r, r, Lw.d(), r, sW, pl.Om.d(), r, sW));
    } else {
        DDMPC_CUDA(Vw.alloc(sizeof(double) * CW * sW));
        for (DevBuf *b : {&lw, &dinv, &keep}) DDMPC_CUDA(b->alloc(sizeof(double) * (size_t)CW * r));
        DDMPC_TRY(jacobi_eig(st, CW, r, Lw.d(), r, sW, Vw.d(), r, sW, lw.d(), r));
        k_spectrum<<<CW, 32, 0, st>>>(r, lw.d(), r, 1e-12, 1, dinv.d(), r, nullptr);
        DDMPC_LAUNCH_CHECK();
        k_spectrum<<<CW, 32, 0, st>>>(r, lw.d(), r, 1e-12, 0, keep.d(), r, nullptr);
        DDMPC_LAUNCH_CHECK();
        Mat Vm = mat(Vw.d(), r, 1, sW);
        DDMPC_TRY(gemm(st, CW, r, r, r, 1.0, Vm, tr(Vm), 0.0, pl.Om.d(), r, 1, sW, dinv.d(), r));
        DDMPC_CUDA(Aeq.alloc(sizeof(double) * CW * sAq));
        dim3 g(ceil_div(sAq, 256), CW);
        k_fill_Aeq<<<g, 256, 0, st>>>(fa, perm_d, Vw.d(), sW, keep.d(), r, Aeq.d(), sAq);
        DDMPC_LAUNCH_CHECK();
    }
    DDMPC_TRY(symmetrize(st, CW, r, pl.Om.d(), r, sW));

    // reduced Hessian (permuted [free; fixed]) and the theta maps
    {
        dim3 g(ceil_div(sP, 256), C);
        k_fill_P<<<g, 256, 0, st>>>(fa, perm_d, Rd, Qd, pl.lamA.d(), pl.lamS.d(), pl.Om.d(), sOm, P.d(), sP);
        DDMPC_LAUNCH_CHECK();
        k_fill_CDT<<<ceil_div((long)nx * nth, 256), 256, 0, st>>>(fa, perm_d, Rd, Qd, Cp.d(), DTp.d());
        DDMPC_LAUNCH_CHECK();
        k_fill_ZT<<<ceil_div(nth * nth, 256), 256, 0, st>>>(fa, invperm_d, DTp.d(), ZT.d());
        DDMPC_LAUNCH_CHECK();
    }
    const double *Cc = Cp.d() + (long)nf * nth;    // (nfix, nth)
    const double *DTc = DTp.d() + (long)nf * nth;  // (nfix, nth)
    Mat Pfc = mat(P.d() + nf, nx, 1, sP);
    Mat Pcc = mat(P.d() + (long)nf * nx + nf, nx, 1, sP);
    Mat Ccm = mat(Cc, nth, 1, 0), DTcm = mat(DTc, nth, 1, 0);
    // N' = DT_f - P_fc C_c
    DDMPC_TRY(copy_bcast(st, C, (long)nf * nth, DTp.d(), 0, Nn.d(), (long)nf * nth));
    DDMPC_TRY(gemm(st, C, nf, nth, nfix, -1.0, Pfc, Ccm, 1.0, Nn.d(), nth, 1, (long)nf * nth));
    // G1 = P_cc C_c
    DDMPC_TRY(gemm(st, C, nfix, nth, nfix, 1.0, Pcc, Ccm, 0.0, G1.d(), nth, 1, (long)nfix * nth));
    // A = L L^T (in place, top-left block of P), X0f = A^-1 N'
    DDMPC_TRY(potrf(st, C, nf, P.d(), nx, sP, info_d + C));
    DDMPC_TRY(copy_bcast(st, C, (long)nf * nth, Nn.d(), (long)nf * nth, X0f.d(), (long)nf * nth));
    DDMPC_TRY(potrs(st, C, nf, nth, P.d(), nx, sP, X0f.d(), nth, (long)nf * nth));
    Mat Af{};
    if (deficient) {
        Af = mat(Aeq.d(), nx, 1, sAe);                    // (r x nf): the free columns; the fixed ones follow at + nf
        // Yq = A^-1 Af^T, S = Af Yq (+ identity on the rows without a constraint), mu = S^-1 (Af X0f + Ac C_c), X0f -= Yq mu
        DDMPC_CUDA(Yq.alloc(sizeof(double) * (size_t)C * nf * r));
        DDMPC_CUDA(Sq.alloc(sizeof(double) * (size_t)C * r * r));
        DDMPC_CUDA(Rq.alloc(sizeof(double) * (size_t)C * r * nth));
        DDMPC_CUDA(AcC.alloc(sizeof(double) * (size_t)CW * r * nth));
        {
            dim3 g(ceil_div((long)nf * r, 256), C);
            k_transpose<<<g, 256, 0, st>>>(nf, r, Aeq.d(), nx, sAe, Yq.d(), (long)nf * r);
            DDMPC_LAUNCH_CHECK();
        }
        DDMPC_TRY(potrs(st, C, nf, r, P.d(), nx, sP, Yq.d(), r, (long)nf * r));
        const Mat Yqm = mat(Yq.d(), r, 1, (long)nf * r);
        DDMPC_TRY(gemm(st, C, r, r, nf, 1.0, Af, Yqm, 0.0, Sq.d(), r, 1, (long)r * r));
        {
            dim3 g(ceil_div(r, 128), C);
            k_add_diag<<<g, 128, 0, st>>>(r, keep.d(), CW == 1 ? 0 : r, Sq.d(), (long)r * r);
            DDMPC_LAUNCH_CHECK();
        }
        DDMPC_TRY(symmetrize(st, C, r, Sq.d(), r, (long)r * r));
        DDMPC_TRY(potrf(st, C, r, Sq.d(), r, (long)r * r, info_d));
        DDMPC_TRY(gemm(st, CW, r, nth, nfix, 1.0, mat(Aeq.d() + nf, nx, 1, sAq), Ccm, 0.0, AcC.d(), nth, 1, (long)r * nth));
        DDMPC_TRY(copy_bcast(st, C, (long)r * nth, AcC.d(), CW == 1 ? 0 : (long)r * nth, Rq.d(), (long)r * nth));
        DDMPC_TRY(gemm(st, C, r, nth, nf, 1.0, Af, mat(X0f.d(), nth, 1, (long)nf * nth), 1.0, Rq.d(), nth, 1, (long)r * nth));
        DDMPC_TRY(potrs(st, C, r, nth, Sq.d(), r, (long)r * r, Rq.d(), nth, (long)r * nth));
        DDMPC_TRY(gemm(st, C, nf, nth, r, -1.0, Yqm, mat(Rq.d(), nth, 1, (long)r * nth), 1.0, X0f.d(), nth, 1, (long)nf * nth));
    }
    // Z = ZT + C_c^T G1 - C_c^T DT_c - DT_c^T C_c - N'^T X0f   (+ (Ac C_c)^T mu with the range constraint)
    const long sZ = (long)nth * nth;
    DDMPC_TRY(copy_bcast(st, C, sZ, ZT.d(), 0, pl.Z.d(), sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, nfix, 1.0, tr(Ccm), mat(G1.d(), nth, 1, (long)nfix * nth), 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, nfix, -1.0, tr(Ccm), DTcm, 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, nfix, -1.0, tr(DTcm), Ccm, 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, nf, -1.0, tr(mat(Nn.d(), nth, 1, (long)nf * nth)),
                   mat(X0f.d(), nth, 1, (long)nf * nth), 1.0, pl.Z.d(), nth, 1, sZ));
    if (deficient)
        DDMPC_TRY(gemm(st, C, nth, nth, r, 1.0, tr(mat(AcC.d(), nth, 1, CW == 1 ? 0 : (long)r * nth)),
                       mat(Rq.d(), nth, 1, (long)r * nth), 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(symmetrize(st, C, nth, pl.Z.d(), nth, sZ));

    if (nb > 0) {
        DDMPC_CUDA(Y.alloc(sizeof(double) * (size_t)C * nf * nb));
        DDMPC_CUDA(Mm.alloc(sizeof(double) * (size_t)C * nb * nb));
        dim3 g(ceil_div((long)nf * nb, 256), C);
        k_fill_Bt<<<g, 256, 0, st>>>(nf, nb, bpos_d, Y.d(), (long)nf * nb);
        DDMPC_LAUNCH_CHECK();
        DDMPC_TRY(potrs(st, C, nf, nb, P.d(), nx, sP, Y.d(), nb, (long)nf * nb));
        if (deficient) {   // the response to a force on the box rows, with the range constraint: Y -= Yq S^-1 (Af Y)
            DDMPC_CUDA(Rb.alloc(sizeof(double) * (size_t)C * r * nb));
            DDMPC_TRY(gemm(st, C, r, nb, nf, 1.0, Af, mat(Y.d(), nb, 1, (long)nf * nb), 0.0, Rb.d(), nb, 1, (long)r * nb));
            DDMPC_TRY(potrs(st, C, r, nb, Sq.d(), r, (long)r * r, Rb.d(), nb, (long)r * nb));
            DDMPC_TRY(gemm(st, C, nf, nb, r, -1.0, mat(Yq.d(), r, 1, (long)nf * r), mat(Rb.d(), nb, 1, (long)r * nb), 1.0,
                           Y.d(), nb, 1, (long)nf * nb));
        }
        k_lam_rho<<<C, 256, 0, st>>>(nf, nb, d.nbs, d.nbu, bpos_d, Y.d(), (long)nf * nb, pl.blo.d(), pl.bhi.d(), pl.Lam.d(),
                                     Mm.d(), pl.rho2.d(), pl.rs.d(), pl.lo.d(), pl.hi.d(), pl.bmax.d());
        DDMPC_LAUNCH_CHECK();
        DDMPC_TRY(potrf(st, C, nb, Mm.d(), nb, (long)nb * nb, info_d + 2 * C));
        DDMPC_TRY(set_identity(st, C, nb, pl.Phi.d(), nb, (long)nb * nb));
        DDMPC_TRY(potrs(st, C, nb, nb, Mm.d(), nb, (long)nb * nb, pl.Phi.d(), nb, (long)nb * nb));
        DDMPC_TRY(symmetrize(st, C, nb, pl.Phi.d(), nb, (long)nb * nb));
    }
    {
        long tot = (long)nx * nth;
        if (nb > 0) tot = std::max(tot, (long)nx * nb);
        dim3 g(ceil_div(tot, 256), C);
        k_extract<<<g, 256, 0, st>>>(fa, nf, nb, d.Lm, invperm_d, bpos_d, X0f.d(), (long)nf * nth, Cp.d(),
                                     nb > 0 ? Y.d() : nullptr, (long)nf * nb, pl.rho2.d(), pl.rs.d(), pl.Ku.d(), pl.X0.d(),
                                     pl.Ks.d(), pl.Psi.d(), pl.Yf.d());
        DDMPC_LAUNCH_CHECK();
    }
    DDMPC_CUDA(cudaStreamSynchronize(st));  // scratch buffers die with this scope
    return DDMPC_OK;
}

static int build_nominal(cudaStream_t st, Plan &pl, const FillArgs &fa, const int *perm_d, const int *invperm_d,
                         const double *ud, long bs_ud, const double *yd, long bs_yd, const double *Rd,
                         const double *Qd) {
    // t = [ubar; ybar] constrained to range(H) (rank-revealing, DESIGN.md "NOMINAL").
    const Dims &d = pl.d;
    const int C = pl.count, r = d.r, nf = d.nf, nfix = d.nfix, nth = d.nth;
    const long sW = (long)r * r, sF = (long)nfix * nfix, sT = (long)r * nth;
    DevBuf Hp, Wp, V1, l1, d1, Pi, Mc, Mc0, Vc, lc, d2, Mcp, G2, tp, G3, Pi0, Dp, G4, G0, V0, l0, d3, Cp, DTp, ZT, G5, G6,
        G7, G8, X0p, G9, G10;
    DDMPC_CUDA(Hp.alloc(sizeof(double) * (size_t)C * r * d.cols));
    for (DevBuf *b : {&Wp, &V1, &Pi, &Pi0, &Dp, &G4, &G0, &V0}) DDMPC_CUDA(b->alloc(sizeof(double) * C * sW));
    for (DevBuf *b : {&l1, &d1, &l0, &d3}) DDMPC_CUDA(b->alloc(sizeof(double) * (size_t)C * r));
    for (DevBuf *b : {&Mc, &Mc0, &Vc, &Mcp, &G10}) DDMPC_CUDA(b->alloc(sizeof(double) * C * sF));
    for (DevBuf *b : {&lc, &d2}) DDMPC_CUDA(b->alloc(sizeof(double) * (size_t)C * nfix));
    DDMPC_CUDA(G2.alloc(sizeof(double) * (size_t)C * nfix * nth));
    DDMPC_CUDA(G3.alloc(sizeof(double) * (size_t)C * nfix * r));
    for (DevBuf *b : {&tp, &G5, &G6, &G7, &G8, &X0p, &G9}) DDMPC_CUDA(b->alloc(sizeof(double) * C * sT));
    DDMPC_CUDA(Cp.alloc(sizeof(double) * (size_t)r * nth));
    DDMPC_CUDA(DTp.alloc(sizeof(double) * (size_t)r * nth));
    DDMPC_CUDA(ZT.alloc(sizeof(double) * (size_t)nth * nth));

    // permuted Hankel stack and its Gram matrix
    const long sH = (long)r * d.cols;
    DDMPC_TRY(launch_hankel(st, C, ud, bs_ud, d.N, d.m, d.Lp, 0, invperm_d, Hp.d(), d.cols, sH));
    DDMPC_TRY(launch_hankel(st, C, yd, bs_yd, d.N, d.p, d.Lp, d.nu, invperm_d, Hp.d(), d.cols, sH));
    Mat Hm = mat(Hp.d(), d.cols, 1, sH);
    DDMPC_TRY(gemm(st, C, r, r, d.cols, 1.0, Hm, tr(Hm), 0.0, Wp.d(), r, 1, sW, nullptr, 0, true));
    DDMPC_TRY(symmetrize(st, C, r, Wp.d(), r, sW));
    DDMPC_TRY(jacobi_eig(st, C, r, Wp.d(), r, sW, V1.d(), r, sW, l1.d(), r));
    k_spectrum<<<C, 32, 0, st>>>(r, l1.d(), r, 1e-11, 0, d1.d(), r, nullptr);
    DDMPC_LAUNCH_CHECK();
    Mat V1m = mat(V1.d(), r, 1, sW);
    DDMPC_TRY(gemm(st, C, r, r, r, 1.0, V1m, tr(V1m), 0.0, Pi.d(), r, 1, sW, d1.d(), r));   // projector on range(H)
    DDMPC_TRY(symmetrize(st, C, r, Pi.d(), r, sW));
    // Mc = Pi[fix, fix] and its pseudo-inverse
    DDMPC_TRY(copy2d(st, C, nfix, nfix, Pi.d() + (long)nf * r + nf, r, sW, Mc.d(), nfix, sF));
    DDMPC_TRY(copy_bcast(st, C, sF, Mc.d(), sF, Mc0.d(), sF));
    DDMPC_TRY(jacobi_eig(st, C, nfix, Mc.d(), nfix, sF, Vc.d(), nfix, sF, lc.d(), nfix));
    k_spectrum<<<C, 32, 0, st>>>(nfix, lc.d(), nfix, 1e-9, 1, d2.d(), nfix, nullptr);
    DDMPC_LAUNCH_CHECK();
    Mat Vcm = mat(Vc.d(), nfix, 1, sF);
    DDMPC_TRY(gemm(st, C, nfix, nfix, nfix, 1.0, Vcm, tr(Vcm), 0.0, Mcp.d(), nfix, 1, sF, d2.d(), nfix));
    DDMPC_TRY(symmetrize(st, C, nfix, Mcp.d(), nfix, sF));
    // theta maps
    k_fill_CDT<<<ceil_div((long)r * nth, 256), 256, 0, st>>>(fa, perm_d, Rd, Qd, Cp.d(), DTp.d());
    DDMPC_LAUNCH_CHECK();
    k_fill_ZT<<<ceil_div(nth * nth, 256), 256, 0, st>>>(fa, invperm_d, DTp.d(), ZT.d());
    DDMPC_LAUNCH_CHECK();
    {
        dim3 g(ceil_div(sW, 256), C);
        k_fill_P<<<g, 256, 0, st>>>(fa, perm_d, Rd, Qd, nullptr, nullptr, nullptr, 0, Dp.d(), sW);
        DDMPC_LAUNCH_CHECK();
    }
    Mat Ccm = mat(Cp.d() + (long)nf * nth, nth, 1, 0);
    Mat Mcpm = mat(Mcp.d(), nfix, 1, sF);
    Mat Pifix_cols = mat(Pi.d() + nf, r, 1, sW);             // Pi[:, fix]  (r, nfix)
    Mat Pifix_rows = mat(Pi.d() + (long)nf * r, r, 1, sW);   // Pi[fix, :]  (nfix, r)
    // tp = Pi[:,fix] Mc^+ C_c
    DDMPC_TRY(gemm(st, C, nfix, nth, nfix, 1.0, Mcpm, Ccm, 0.0, G2.d(), nth, 1, (long)nfix * nth));
    DDMPC_TRY(gemm(st, C, r, nth, nfix, 1.0, Pifix_cols, mat(G2.d(), nth, 1, (long)nfix * nth), 0.0, tp.d(), nth, 1, sT));
    // Pi0 = Pi - Pi[:,fix] Mc^+ Pi[fix,:]
    DDMPC_TRY(gemm(st, C, nfix, r, nfix, 1.0, Mcpm, Pifix_rows, 0.0, G3.d(), r, 1, (long)nfix * r));
    DDMPC_TRY(copy_bcast(st, C, sW, Pi.d(), sW, Pi0.d(), sW));
    DDMPC_TRY(gemm(st, C, r, r, nfix, -1.0, Pifix_cols, mat(G3.d(), r, 1, (long)nfix * r), 1.0, Pi0.d(), r, 1, sW));
    DDMPC_TRY(symmetrize(st, C, r, Pi0.d(), r, sW));
    // G0 = Pi0 D Pi0 and its pseudo-inverse spectrum
    Mat Pi0m = mat(Pi0.d(), r, 1, sW), Dpm = mat(Dp.d(), r, 1, sW);
    DDMPC_TRY(gemm(st, C, r, r, r, 1.0, Dpm, Pi0m, 0.0, G4.d(), r, 1, sW));
    DDMPC_TRY(gemm(st, C, r, r, r, 1.0, Pi0m, mat(G4.d(), r, 1, sW), 0.0, G0.d(), r, 1, sW));
    DDMPC_TRY(symmetrize(st, C, r, G0.d(), r, sW));
    DDMPC_TRY(jacobi_eig(st, C, r, G0.d(), r, sW, V0.d(), r, sW, l0.d(), r));
    k_spectrum<<<C, 32, 0, st>>>(r, l0.d(), r, 1e-11, 1, d3.d(), r, nullptr);
    DDMPC_LAUNCH_CHECK();
    // X0p = tp + Pi0 V0 diag(d3) V0^T Pi0 (DT - D tp)
    DDMPC_TRY(copy_bcast(st, C, sT, DTp.d(), 0, G5.d(), sT));
    DDMPC_TRY(gemm(st, C, r, nth, r, -1.0, Dpm, mat(tp.d(), nth, 1, sT), 1.0, G5.d(), nth, 1, sT));
    DDMPC_TRY(gemm(st, C, r, nth, r, 1.0, Pi0m, mat(G5.d(), nth, 1, sT), 0.0, G6.d(), nth, 1, sT));
    Mat V0m = mat(V0.d(), r, 1, sW);
    DDMPC_TRY(gemm(st, C, r, nth, r, 1.0, tr(V0m), mat(G6.d(), nth, 1, sT), 0.0, G7.d(), nth, 1, sT));
    DDMPC_TRY(gemm(st, C, r, nth, r, 1.0, V0m, mat(G7.d(), nth, 1, sT), 0.0, G8.d(), nth, 1, sT, d3.d(), r));
    DDMPC_TRY(copy_bcast(st, C, sT, tp.d(), sT, X0p.d(), sT));
    DDMPC_TRY(gemm(st, C, r, nth, r, 1.0, Pi0m, mat(G8.d(), nth, 1, sT), 1.0, X0p.d(), nth, 1, sT));
    // Z = ZT + X0p^T D X0p - X0p^T DT - DT^T X0p
    const long sZ = (long)nth * nth;
    Mat X0m = mat(X0p.d(), nth, 1, sT), DTm = mat(DTp.d(), nth, 1, 0);
    DDMPC_TRY(gemm(st, C, r, nth, r, 1.0, Dpm, X0m, 0.0, G9.d(), nth, 1, sT));
    DDMPC_TRY(copy_bcast(st, C, sZ, ZT.d(), 0, pl.Z.d(), sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, r, 1.0, tr(X0m), mat(G9.d(), nth, 1, sT), 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, r, -1.0, tr(X0m), DTm, 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(gemm(st, C, nth, nth, r, -1.0, tr(DTm), X0m, 1.0, pl.Z.d(), nth, 1, sZ));
    DDMPC_TRY(symmetrize(st, C, nth, pl.Z.d(), nth, sZ));
    // F = (Mc Mc^+ - I) C_c
    DDMPC_TRY(gemm(st, C, nfix, nfix, nfix, 1.0, mat(Mc0.d(), nfix, 1, sF), Mcpm, 0.0, G10.d(), nfix, 1, sF));
    {
        dim3 g(ceil_div(nfix, 128), C);
        k_sub_identity<<<g, 128, 0, st>>>(nfix, G10.d(), sF);
        DDMPC_LAUNCH_CHECK();
    }
    DDMPC_TRY(gemm(st, C, nfix, nth, nfix, 1.0, mat(G10.d(), nfix, 1, sF), Ccm, 0.0, pl.F.d(), nth, 1, (long)nfix * nth));
    k_flag_nonzero<<<C, 256, 0, st>>>((long)nfix * nth, pl.F.d(), pl.Fnz.i());
    DDMPC_LAUNCH_CHECK();
    // extract: X0p holds every (permuted) row, so "nf" = nx here
    {
        dim3 g(ceil_div((long)r * nth, 256), C);
        k_extract<<<g, 256, 0, st>>>(fa, r, 0, d.Lm, invperm_d, nullptr, X0p.d(), sT, Cp.d(), nullptr, 0, nullptr, nullptr,
                                     pl.Ku.d(), pl.X0.d(), nullptr, nullptr, nullptr);
        DDMPC_LAUNCH_CHECK();
    }
    DDMPC_CUDA(cudaStreamSynchronize(st));
    return DDMPC_OK;
}

int set_create_device(const ddmpc_params *prm, int count, const double *u_d, size_t ud_stride, const double *y_d,
                      size_t yd_stride, const double *Q, const double *R, const double *lamb_alpha,
                      const double *lamb_sigma, cudaStream_t st, ddmpc_set **out) {
    if (!prm || !out || !u_d || !y_d || !Q || !R || count <= 0) return fail(DDMPC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    const ddmpc_params &q = *prm;
    DDMPC_TRY(validate(q));
    ScratchStreamScope scratch_scope(st);   // every DevBuf below is allocated and freed in the order of `st`
    // Remark 1 of the paper / controller.py:275-283
    const long N_min = (long)q.m * (q.L + 2 * q.n) + q.L + 2 * q.n - 1;
    if (q.check_pe && q.N < N_min)
        return fail(DDMPC_ERR_N_TOO_SMALL,
                    "Initial input trajectory data is not persistently exciting of order (L + 2 * n). It does not "
                    "satisfy the inequality: N - L - 2 * n + 1 >= m * (L + 2 * n). The required minimum N is %ld, "
                    "but got %d.", N_min, q.N);
    if (q.N < q.L + q.n) return fail(DDMPC_ERR_HANKEL_WINDOW, "N must be greater than or equal to L.");

    std::unique_ptr<ddmpc_set> set(new ddmpc_set());
    set->prm = q;
    set->prm.u_min = set->prm.u_max = set->prm.y_min = set->prm.y_max = nullptr;   // caller memory: the plan keeps its own copy
    Plan &pl = set->plan;
    pl.d = make_dims(q);
    pl.count = count;
    pl.data_count = (ud_stride == 0 && yd_stride == 0) ? 1 : count;
    pl.pe_rank.assign(count, -1);
    pl.status.assign(count, DDMPC_OK);
    const Dims &d = pl.d;
    pl.bound = d.convex ? q.c * q.eps_max : 0.0;

    // persistency of excitation of order L + 2n (controller.py:285-296)
    if (q.check_pe) {
        const int nb_pe = ud_stride == 0 ? 1 : count;
        DevBuf rk;
        DDMPC_CUDA(rk.alloc(sizeof(int) * nb_pe));
        DDMPC_TRY(pe_rank_device(st, nb_pe, u_d, (long)ud_stride, q.N, q.m, q.L + 2 * q.n, rk.i()));
        std::vector<int> hr(nb_pe);
        DDMPC_CUDA(cudaMemcpy(hr.data(), rk.p, sizeof(int) * nb_pe, cudaMemcpyDeviceToHost));
        for (int c = 0; c < count; ++c) {
            pl.pe_rank[c] = hr[ud_stride == 0 ? 0 : c];
            if (pl.pe_rank[c] != q.m * (q.L + 2 * q.n)) pl.status[c] = DDMPC_ERR_NOT_PE;
        }
    }
    if (count == 1 && pl.status[0] == DDMPC_ERR_NOT_PE)
        return fail(DDMPC_ERR_NOT_PE,
                    "Initial input trajectory data is not persistently exciting of order (L + 2 * n). The rank of "
                    "its induced Hankel matrix (%d) does not match the expected rank (%d).",
                    pl.pe_rank[0], q.m * (q.L + 2 * q.n));
    // controller.py:298-325
    if (q.controller_type == DDMPC_NOMINAL && q.L < q.n)
        return fail(DDMPC_ERR_HORIZON,
                    "The prediction horizon (`L`) must be greater than or equal to the estimated system order `n`.");
    if (q.controller_type == DDMPC_ROBUST && q.L < 2 * q.n)
        return fail(DDMPC_ERR_HORIZON,
                    "The prediction horizon (`L`) must be greater than or equal to two times the estimated system "
                    "order `n`.");
    // controller.py:664-670
    if (d.robust && q.slack_type == DDMPC_SLACK_NON_CONVEX)
        return fail(DDMPC_ERR_NOT_IMPLEMENTED,
                    "Robust Data-Driven MPC with a Non-Convex slack variable constraint is not currently implemented, "
                    "since it cannot be efficiently solved.");

    // index maps: permutation [free; fixed]
    std::vector<int> perm(d.nx), invperm(d.nx), bpos(std::max(d.nb, 1));
    {
        std::vector<char> fixed(d.nx, 0);
        for (int i = 0; i < q.n * q.m; ++i) fixed[i] = 1;
        for (int i = 0; i < q.n * q.p; ++i) fixed[d.nu + i] = 1;
        if (d.terminal) {
            for (int i = q.L * q.m; i < d.nu; ++i) fixed[i] = 1;
            for (int i = d.nu + q.L * q.p; i < d.nu + d.ny; ++i) fixed[i] = 1;
        }
        int a = 0;
        for (int i = 0; i < d.nx; ++i) if (!fixed[i]) perm[a++] = i;
        if (a != d.nf) return fail(DDMPC_ERR_INVALID_ARG, "internal: free count mismatch");
        for (int i = 0; i < d.nx; ++i) if (fixed[i]) perm[a++] = i;
        for (int k = 0; k < d.nx; ++k) invperm[perm[k]] = k;
        for (int j = 0; j < d.nbs; ++j) bpos[j] = invperm[d.nu + d.ny + q.n * q.p + j];   // sigma_pred
        for (int j = 0; j < d.nbu; ++j) bpos[d.nbs + j] = invperm[q.n * q.m + j];          // free predicted inputs
        for (int j = 0; j < d.nby; ++j) bpos[d.nbs + d.nbu + j] = invperm[d.nu + q.n * q.p + j];   // free predicted outputs
    }
    DevBuf perm_d, invperm_d, bpos_d, info_d;
    DDMPC_CUDA(perm_d.alloc(sizeof(int) * d.nx));
    DDMPC_CUDA(invperm_d.alloc(sizeof(int) * d.nx));
    DDMPC_CUDA(bpos_d.alloc(sizeof(int) * bpos.size()));
    DDMPC_CUDA(info_d.alloc(sizeof(int) * 3 * count));
    DDMPC_CUDA(cudaMemcpyAsync(perm_d.p, perm.data(), sizeof(int) * d.nx, cudaMemcpyHostToDevice, st));
    DDMPC_CUDA(cudaMemcpyAsync(invperm_d.p, invperm.data(), sizeof(int) * d.nx, cudaMemcpyHostToDevice, st));
    DDMPC_CUDA(cudaMemcpyAsync(bpos_d.p, bpos.data(), sizeof(int) * bpos.size(), cudaMemcpyHostToDevice, st));
    DDMPC_CUDA(cudaMemsetAsync(info_d.p, 0, sizeof(int) * 3 * count, st));

    // per-controller weights
    {
        std::vector<double> la(count), ls(count);
        for (int c = 0; c < count; ++c) {
            const double a = lamb_alpha ? lamb_alpha[c] : q.lamb_alpha;
            la[c] = d.robust ? a * q.eps_max : 0.0;   // controller.py:714: lamb_alpha * eps_max
            ls[c] = d.robust ? (lamb_sigma ? lamb_sigma[c] : q.lamb_sigma) : 0.0;
        }
        DDMPC_CUDA(pl.lamA.alloc(sizeof(double) * count));
        DDMPC_CUDA(pl.lamS.alloc(sizeof(double) * count));
        DDMPC_CUDA(cudaMemcpyAsync(pl.lamA.p, la.data(), sizeof(double) * count, cudaMemcpyHostToDevice, st));
        DDMPC_CUDA(cudaMemcpyAsync(pl.lamS.p, ls.data(), sizeof(double) * count, cudaMemcpyHostToDevice, st));
        DDMPC_CUDA(cudaStreamSynchronize(st));
    }

    // operator storage
    const size_t C = count;
    const size_t CD = pl.data_count;
    DDMPC_CUDA(pl.H.alloc(sizeof(double) * CD * d.r * d.cols));
    DDMPC_CUDA(pl.Ku.alloc(sizeof(double) * C * d.Lm * d.nth));
    DDMPC_CUDA(pl.Z.alloc(sizeof(double) * C * d.nth * d.nth));
    DDMPC_CUDA(pl.X0.alloc(sizeof(double) * C * d.nx * d.nth));
    DDMPC_CUDA(pl.rho2.alloc(sizeof(double) * C));
    DDMPC_CUDA(cudaMemsetAsync(pl.rho2.p, 0, sizeof(double) * C, st));
    if (d.robust) {
        DDMPC_CUDA(pl.W.alloc(sizeof(double) * CD * d.r * d.r));
        DDMPC_CUDA(pl.Om.alloc(sizeof(double) * CD * d.r * d.r));
    } else {
        DDMPC_CUDA(pl.F.alloc(sizeof(double) * C * d.nfix * d.nth));
        DDMPC_CUDA(pl.Fnz.alloc(sizeof(int) * C));
    }
    if (d.nb > 0) {
        // box rows: sigma_pred in [-c eps_max, c eps_max] (controller.py:659-675), predicted inputs in [u_min, u_max]
        std::vector<double> blo(d.nb), bhi(d.nb);
        for (int j = 0; j < d.nbs; ++j) { blo[j] = -pl.bound; bhi[j] = pl.bound; }
        for (int j = 0; j < d.nbu; ++j) {
            blo[d.nbs + j] = q.u_min ? q.u_min[j % q.m] : -INFINITY;
            bhi[d.nbs + j] = q.u_max ? q.u_max[j % q.m] : INFINITY;
        }
        for (int j = 0; j < d.nby; ++j) {
            blo[d.nbs + d.nbu + j] = q.y_min ? q.y_min[j % q.p] : -INFINITY;
            bhi[d.nbs + d.nbu + j] = q.y_max ? q.y_max[j % q.p] : INFINITY;
        }
        if (d.nby > 0) {
            std::vector<double> yl(q.p, -INFINITY), yh(q.p, INFINITY);
            for (int j = 0; j < q.p; ++j) {
                if (q.y_min) yl[j] = q.y_min[j];
                if (q.y_max) yh[j] = q.y_max[j];
            }
            DDMPC_CUDA(pl.ymin.alloc(sizeof(double) * q.p));
            DDMPC_CUDA(pl.ymax.alloc(sizeof(double) * q.p));
            DDMPC_CUDA(cudaMemcpy(pl.ymin.p, yl.data(), sizeof(double) * q.p, cudaMemcpyHostToDevice));
            DDMPC_CUDA(cudaMemcpy(pl.ymax.p, yh.data(), sizeof(double) * q.p, cudaMemcpyHostToDevice));
        }
        DDMPC_CUDA(pl.blo.alloc(sizeof(double) * d.nb));
        DDMPC_CUDA(pl.bhi.alloc(sizeof(double) * d.nb));
        DDMPC_CUDA(cudaMemcpyAsync(pl.blo.p, blo.data(), sizeof(double) * d.nb, cudaMemcpyHostToDevice, st));
        DDMPC_CUDA(cudaMemcpyAsync(pl.bhi.p, bhi.data(), sizeof(double) * d.nb, cudaMemcpyHostToDevice, st));
        if (d.nbu > 0) {
            pl.u_min.assign(q.m, -INFINITY);
            pl.u_max.assign(q.m, INFINITY);
            for (int j = 0; j < q.m; ++j) {
                if (q.u_min) pl.u_min[j] = q.u_min[j];
                if (q.u_max) pl.u_max[j] = q.u_max[j];
            }
            DDMPC_CUDA(pl.umin.alloc(sizeof(double) * q.m));
            DDMPC_CUDA(pl.umax.alloc(sizeof(double) * q.m));
            DDMPC_CUDA(cudaMemcpyAsync(pl.umin.p, pl.u_min.data(), sizeof(double) * q.m, cudaMemcpyHostToDevice, st));
            DDMPC_CUDA(cudaMemcpyAsync(pl.umax.p, pl.u_max.data(), sizeof(double) * q.m, cudaMemcpyHostToDevice, st));
        }
        DDMPC_CUDA(cudaStreamSynchronize(st));   // blo / bhi are stack vectors
        DDMPC_CUDA(pl.rs.alloc(sizeof(double) * C * d.nb));
        DDMPC_CUDA(pl.lo.alloc(sizeof(double) * C * d.nb));
        DDMPC_CUDA(pl.hi.alloc(sizeof(double) * C * d.nb));
        DDMPC_CUDA(pl.bmax.alloc(sizeof(double) * C));
        DDMPC_CUDA(pl.Ks.alloc(sizeof(double) * C * d.nb * d.nth));
        DDMPC_CUDA(pl.Phi.alloc(sizeof(double) * C * d.nb * d.nb));
        DDMPC_CUDA(pl.Psi.alloc(sizeof(double) * C * d.Lm * d.nb));
        DDMPC_CUDA(pl.Lam.alloc(sizeof(double) * C * d.nb * d.nb));
        DDMPC_CUDA(pl.Yf.alloc(sizeof(double) * C * d.nx * d.nb));
    }

    // K1: Hankel matrices H_{L+n}(u^d), H_{L+n}(y^d)  (controller.py:376-377)
    const long sH = (long)d.r * d.cols;
    DDMPC_TRY(launch_hankel(st, pl.data_count, u_d, (long)ud_stride, q.N, q.m, d.Lp, 0, nullptr, pl.H.d(), d.cols, sH));
    DDMPC_TRY(launch_hankel(st, pl.data_count, y_d, (long)yd_stride, q.N, q.p, d.Lp, d.nu, nullptr, pl.H.d(), d.cols, sH));

    FillArgs fa{q.n, q.m, q.p, q.L, d.nu, d.ny, d.nx, d.nth, d.terminal, d.robust};
    if (d.robust) {
        DDMPC_TRY(build_robust(st, pl, fa, perm_d.i(), invperm_d.i(), bpos_d.i(), R, Q, info_d.i()));
        std::vector<int> info(3 * count);
        DDMPC_CUDA(cudaMemcpy(info.data(), info_d.p, sizeof(int) * 3 * count, cudaMemcpyDeviceToHost));
        // slot 0: Cholesky of W, one per data set - or, with short data (build_robust), of the Schur complement, one per controller
        const bool per_ctrl = pl.data_count != 1 || pl.range_constrained;
        for (int c = 0; c < count; ++c)
            if (pl.status[c] == DDMPC_OK && (info[per_ctrl ? c : 0] || info[count + c] || info[2 * count + c]))
                pl.status[c] = DDMPC_ERR_FACTORIZATION;
    } else {
        DDMPC_TRY(build_nominal(st, pl, fa, perm_d.i(), invperm_d.i(), u_d, (long)ud_stride, y_d, (long)yd_stride, R, Q));
    }
    // Per-controller verdicts.  A single controller that cannot be set up is an error, as in the reference
    // (controller.py:285-296 raises at construction).  In a larger set the failed controllers are poisoned: their gains
    // become NaN, so every solve / closed loop that uses one returns NaN and DDMPC_SOLVE_NONFINITE instead of finite
    // garbage under status "optimal"; ddmpc_set_failed_count / ddmpc_set_info report which ones.
    for (int c = 0; c < count; ++c) set->n_failed += pl.status[c] != DDMPC_OK;
    if (count == 1 && pl.status[0] == DDMPC_ERR_FACTORIZATION)
        return fail(DDMPC_ERR_FACTORIZATION,
                    "The Gram matrix of the data or the reduced Hessian is not positive definite: the data are too "
                    "ill-conditioned for the robust setup.");
    if (set->n_failed > 0) {
        DDMPC_CUDA(set->ctrl_status.alloc(sizeof(int) * count));
        DDMPC_CUDA(cudaMemcpyAsync(set->ctrl_status.p, pl.status.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
        const int T = 256;
        k_poison<<<ceil_div((long)count * d.Lm * d.nth, T), T, 0, st>>>(count, (long)d.Lm * d.nth, set->ctrl_status.i(), pl.Ku.d());
        DDMPC_LAUNCH_CHECK();
        if (d.nb > 0) {
            k_poison<<<ceil_div((long)count * d.nb * d.nth, T), T, 0, st>>>(count, (long)d.nb * d.nth, set->ctrl_status.i(), pl.Ks.d());
            DDMPC_LAUNCH_CHECK();
        }
        DDMPC_CUDA(cudaStreamSynchronize(st));
    }
    DDMPC_TRY(closed_loop_fast_prepare(set.get(), st));
    DDMPC_TRY(closed_loop_cvx_prepare(set.get(), st));
    DDMPC_CUDA(cudaStreamSynchronize(st));
    set->detach_streams();                  // the set may outlive `st`: it is freed on the legacy stream after a device sync
    *out = set.release();
    return DDMPC_OK;
}

}  // namespace ddmpc

using namespace ddmpc;

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *ddmpc_version(void) { return "ddmpc-b200 0.1 (sm_100a)"; }
const char *ddmpc_last_error(void) { return g_last_error; }
uint64_t ddmpc_kernel_launches(void) { return g_launches.load(); }

int ddmpc_trim_memory(void) {
    int dev = 0;
    DDMPC_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    DDMPC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    DDMPC_CUDA(cudaDeviceSynchronize());
    DDMPC_CUDA(cudaMemPoolTrimTo(pool, 0));
    return DDMPC_OK;
}

const char *ddmpc_strerror(int s) {
    switch (s) {
        case DDMPC_OK: return "ok";
        case DDMPC_ERR_INVALID_ARG: return "invalid argument";
        case DDMPC_ERR_CUDA: return "CUDA error";
        case DDMPC_ERR_CONTROLLER_TYPE: return "Unsupported controller type.";
        case DDMPC_ERR_SLACK_TYPE: return "Unsupported slack variable constraint type.";
        case DDMPC_ERR_ROBUST_PARAMS: return "robust parameters missing";
        case DDMPC_ERR_N_TOO_SMALL: return "input data too short for persistency of excitation";
        case DDMPC_ERR_NOT_PE: return "input data not persistently exciting";
        case DDMPC_ERR_HORIZON: return "prediction horizon too short";
        case DDMPC_ERR_NOT_IMPLEMENTED: return "non-convex slack constraint not implemented";
        case DDMPC_ERR_FACTORIZATION: return "factorisation failed (matrix not positive definite)";
        case DDMPC_ERR_HANKEL_WINDOW: return "N must be greater than or equal to L.";
        default: return "unknown status";
    }
}

int ddmpc_hankel(const double *X, int N, int n_ch, int L, double *H, void *stream) {
    if (!X || !H || n_ch <= 0 || L <= 0) return fail(DDMPC_ERR_INVALID_ARG, "hankel: bad argument");
    if (N < L) return fail(DDMPC_ERR_HANKEL_WINDOW, "N must be greater than or equal to L.");
    return launch_hankel((cudaStream_t)stream, 1, X, 0, N, n_ch, L, 0, nullptr, H, N - L + 1, 0);
}

int ddmpc_hankel_host(const double *X, int N, int n_ch, int L, double *H) {
    if (!X || !H || n_ch <= 0 || L <= 0) return fail(DDMPC_ERR_INVALID_ARG, "hankel: bad argument");
    if (N < L) return fail(DDMPC_ERR_HANKEL_WINDOW, "N must be greater than or equal to L.");
    DevBuf dx, dh;
    const size_t nx = (size_t)N * n_ch, nh = (size_t)L * n_ch * (N - L + 1);
    DDMPC_CUDA(dx.alloc(sizeof(double) * nx));
    DDMPC_CUDA(dh.alloc(sizeof(double) * nh));
    DDMPC_CUDA(cudaMemcpy(dx.p, X, sizeof(double) * nx, cudaMemcpyHostToDevice));
    DDMPC_TRY(ddmpc_hankel(dx.d(), N, n_ch, L, dh.d(), nullptr));
    DDMPC_CUDA(cudaMemcpy(H, dh.p, sizeof(double) * nh, cudaMemcpyDeviceToHost));
    return DDMPC_OK;
}

int ddmpc_pe_rank_host(const double *X, int N, int n_ch, int order, int *rank) {
    if (!X || !rank || n_ch <= 0 || order <= 0) return fail(DDMPC_ERR_INVALID_ARG, "pe_rank: bad argument");
    if (N < order) return fail(DDMPC_ERR_HANKEL_WINDOW, "N must be greater than or equal to L.");
    DevBuf dx, rk;
    DDMPC_CUDA(dx.alloc(sizeof(double) * (size_t)N * n_ch));
    DDMPC_CUDA(rk.alloc(sizeof(int)));
    DDMPC_CUDA(cudaMemcpy(dx.p, X, sizeof(double) * (size_t)N * n_ch, cudaMemcpyHostToDevice));
    DDMPC_TRY(pe_rank_device(nullptr, 1, dx.d(), 0, N, n_ch, order, rk.i()));
    DDMPC_CUDA(cudaMemcpy(rank, rk.p, sizeof(int), cudaMemcpyDeviceToHost));
    return DDMPC_OK;
}

int ddmpc_set_create(const ddmpc_params *params, int count, const double *u_d, size_t ud_stride, const double *y_d,
                     size_t yd_stride, const double *Q, const double *R, const double *lamb_alpha,
                     const double *lamb_sigma, void *stream, ddmpc_set **out) {
    return set_create_device(params, count, u_d, ud_stride, y_d, yd_stride, Q, R, lamb_alpha, lamb_sigma,
                             (cudaStream_t)stream, out);
}

int ddmpc_set_create_host(const ddmpc_params *params, int count, const double *u_d, size_t ud_stride,
                          const double *y_d, size_t yd_stride, const double *Q, const double *R,
                          const double *lamb_alpha, const double *lamb_sigma, ddmpc_set **out) {
    if (!params || !u_d || !y_d || !Q || !R || count <= 0) return fail(DDMPC_ERR_INVALID_ARG, "null argument");
    const ddmpc_params &q = *params;
    if (q.N <= 0 || q.m <= 0 || q.p <= 0 || q.L <= 0) return fail(DDMPC_ERR_INVALID_ARG, "bad sizes");
    const size_t nud = (size_t)q.N * q.m, nyd = (size_t)q.N * q.p;
    const size_t tot_u = ud_stride ? ud_stride * (count - 1) + nud : nud;
    const size_t tot_y = yd_stride ? yd_stride * (count - 1) + nyd : nyd;
    const size_t nQ = (size_t)q.p * q.L * q.p * q.L, nR = (size_t)q.m * q.L * q.m * q.L;
    DevBuf du, dy, dQ, dR;
    DDMPC_CUDA(du.alloc(sizeof(double) * tot_u));
    DDMPC_CUDA(dy.alloc(sizeof(double) * tot_y));
    DDMPC_CUDA(dQ.alloc(sizeof(double) * nQ));
    DDMPC_CUDA(dR.alloc(sizeof(double) * nR));
    DDMPC_CUDA(cudaMemcpy(du.p, u_d, sizeof(double) * tot_u, cudaMemcpyHostToDevice));
    DDMPC_CUDA(cudaMemcpy(dy.p, y_d, sizeof(double) * tot_y, cudaMemcpyHostToDevice));
    DDMPC_CUDA(cudaMemcpy(dQ.p, Q, sizeof(double) * nQ, cudaMemcpyHostToDevice));
    DDMPC_CUDA(cudaMemcpy(dR.p, R, sizeof(double) * nR, cudaMemcpyHostToDevice));
    return set_create_device(params, count, du.d(), ud_stride, dy.d(), yd_stride, dQ.d(), dR.d(), lamb_alpha,
                             lamb_sigma, nullptr, out);
}

void ddmpc_set_destroy(ddmpc_set *set) {
    if (!set) return;
    cudaDeviceSynchronize();   // the plan's buffers go back to the stream-ordered pool: nothing may still be using them
    delete set;
}

int ddmpc_set_count(const ddmpc_set *set) { return set ? set->plan.count : 0; }
int ddmpc_set_failed_count(const ddmpc_set *set) { return set ? set->n_failed : 0; }

int ddmpc_set_option(ddmpc_set *set, const char *name, int value) {
    if (!set || !name) return fail(DDMPC_ERR_INVALID_ARG, "set_option: null argument");
    const std::string nm(name);
    if (nm == "closed_loop_path") {
        if (value < DDMPC_PATH_AUTO || value > DDMPC_PATH_TC) return fail(DDMPC_ERR_INVALID_ARG, "set_option: unknown path %d", value);
        set->opt_path = value;
    } else if (nm == "dmma_warps") {
        if (value != 1 && value != 2 && value != 4) return fail(DDMPC_ERR_INVALID_ARG, "set_option: dmma_warps must be 1, 2 or 4");
        set->opt_dmma_warps = value;
    } else if (nm == "loops_per_thread") {
        if (value < 0 || value > 2) return fail(DDMPC_ERR_INVALID_ARG, "set_option: loops_per_thread must be 0, 1 or 2");
        set->opt_lpt = value;
    } else if (nm == "cvx_ctas_per_sm") {
        if (value != 2 && value != 3) return fail(DDMPC_ERR_INVALID_ARG, "set_option: cvx_ctas_per_sm must be 2 or 3");
        set->opt_cvx_ctas = value;
    } else if (nm == "trajectory_layout") {
        if (value != 0 && value != 1) return fail(DDMPC_ERR_INVALID_ARG, "set_option: trajectory_layout must be 0 (loop-major) or 1 (step-major)");
        set->opt_layout = value;
    } else if (nm == "tc_passes") {
        if (value < 1 || value > 3) return fail(DDMPC_ERR_INVALID_ARG, "set_option: tc_passes must be 1, 2 or 3");
        set->opt_tc_passes = value;
    } else if (nm == "solve_path") {
        if (value < 0 || value > 2) return fail(DDMPC_ERR_INVALID_ARG, "set_option: solve_path must be 0, 1 or 2");
        set->opt_solve = value;
    } else {
        return fail(DDMPC_ERR_INVALID_ARG, "set_option: unknown option '%s'", name);
    }
    return DDMPC_OK;
}

int ddmpc_set_info(const ddmpc_set *set, int index, int *pe_rank, int *status) {
    if (!set || index < 0 || index >= set->plan.count) return fail(DDMPC_ERR_INVALID_ARG, "set_info: bad index");
    if (pe_rank) *pe_rank = set->plan.pe_rank[index];
    if (status) *status = set->plan.status[index];
    return DDMPC_OK;
}

int ddmpc_set_get(const ddmpc_set *set, const char *name, int index, double *out, size_t capacity, size_t *n_elem) {
    if (!set || !name || index < 0 || index >= set->plan.count) return fail(DDMPC_ERR_INVALID_ARG, "set_get: bad argument");
    const Plan &pl = set->plan;
    const Dims &d = pl.d;
    const std::string nm(name);
    const double *src = nullptr;
    size_t n = 0;
    const size_t sH = (size_t)d.r * d.cols;
    const size_t di = pl.data_count == 1 ? 0 : (size_t)index;   // shared data: one H / W / Om for the whole set
    if (nm == "H") { src = pl.H.d() + di * sH; n = sH; }
    else if (nm == "HLn_ud") { src = pl.H.d() + di * sH; n = (size_t)d.nu * d.cols; }
    else if (nm == "HLn_yd") { src = pl.H.d() + di * sH + (size_t)d.nu * d.cols; n = (size_t)d.ny * d.cols; }
    else if (nm == "W" && pl.W.p) { n = (size_t)d.r * d.r; src = pl.W.d() + di * n; }
    else if (nm == "Om" && pl.Om.p) { n = (size_t)d.r * d.r; src = pl.Om.d() + di * n; }
    else if (nm == "Ku") { n = (size_t)d.Lm * d.nth; src = pl.Ku.d() + index * n; }
    else if (nm == "Z") { n = (size_t)d.nth * d.nth; src = pl.Z.d() + index * n; }
    else if (nm == "X0") { n = (size_t)d.nx * d.nth; src = pl.X0.d() + index * n; }
    else if (nm == "Ks" && pl.Ks.p) { n = (size_t)d.nb * d.nth; src = pl.Ks.d() + index * n; }
    else if (nm == "Phi" && pl.Phi.p) { n = (size_t)d.nb * d.nb; src = pl.Phi.d() + index * n; }
    else if (nm == "Psi" && pl.Psi.p) { n = (size_t)d.Lm * d.nb; src = pl.Psi.d() + index * n; }
    else if (nm == "Lam" && pl.Lam.p) { n = (size_t)d.nb * d.nb; src = pl.Lam.d() + index * n; }
    else if (nm == "Yf" && pl.Yf.p) { n = (size_t)d.nx * d.nb; src = pl.Yf.d() + index * n; }
    else if (nm == "F" && pl.F.p) { n = (size_t)d.nfix * d.nth; src = pl.F.d() + index * n; }
    else if (nm == "rho2") { n = 1; src = pl.rho2.d() + index; }
    else if (nm == "row_scale" && pl.rs.p) { n = (size_t)d.nb; src = pl.rs.d() + index * n; }
    else if (nm == "box_lo" && pl.lo.p) { n = (size_t)d.nb; src = pl.lo.d() + index * n; }
    else if (nm == "box_hi" && pl.hi.p) { n = (size_t)d.nb; src = pl.hi.d() + index * n; }
    else return fail(DDMPC_ERR_INVALID_ARG, "set_get: unknown or unavailable matrix '%s'", name);
    if (n_elem) *n_elem = n;
    if (!out) return DDMPC_OK;
    if (capacity < n) return fail(DDMPC_ERR_INVALID_ARG, "set_get: capacity %zu < %zu", capacity, n);
    DDMPC_CUDA(cudaMemcpy(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return DDMPC_OK;
}

}  // extern "C"
