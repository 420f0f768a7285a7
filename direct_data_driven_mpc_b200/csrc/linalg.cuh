// Batched FP64 dense linear algebra used by the per-controller setup:
// strided GEMM (optional diagonal scaling), Cholesky, triangular solves and a
// parallel-ordered cyclic Jacobi eigensolver.  One batch entry per controller.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace ddmpc {

// ---------------------------------------------------------------------------
// C(MxN) = alpha * A(MxK) * diag(d) * B(KxN) + beta * C     (d optional)
// element (i,j) of X lives at X[i*rsX + j*csX]; batch b adds b*bsX.
//
// FP64 tensor-core GEMM: 64x64 CTA tile, 4 warps (2x2), each warp a 32x32 block of 4x4
// mma.sync.m8n8k4.f64 tiles (DMMA), K staged through shared memory 16 at a time.
// Fragment layout of m8n8k4 (g = lane/4, q = lane%4):  A[g][q], B[q][g], C[g][2q], C[g][2q+1].
// Shared row strides are 4 (mod 16) doubles so the 4x4 (row, k) patch a half-warp reads hits
// 16 distinct 8-byte bank pairs.
// ---------------------------------------------------------------------------
constexpr int GK = 16;
constexpr int G_LDA = GK + 4;   // As[m][k]

// GT = CTA tile edge: 64 (warp tile 32x32) or 32 (warp tile 16x16, for products too small to fill
// 148 SMs with 64x64 tiles)
template <int GT>
static __global__ void __launch_bounds__(128)
k_gemm(int M, int N, int K, double alpha,
       const double *__restrict__ A, long rsA, long csA, long bsA,
       const double *__restrict__ Bm, long rsB, long csB, long bsB,
       const double *__restrict__ dvec, long bsd,
       double beta, double *__restrict__ C, long rsC, long csC, long bsC, int sym) {
    // sym: the product is known to be symmetric (Gram matrices A A^T): tiles above the diagonal are skipped and the
    // tiles below it are written to both triangles
    if (sym && blockIdx.x > blockIdx.y) return;
    constexpr int G_LDB = GT + 4;   // Bs[k][n]
    constexpr int WT = GT / 2;      // warp tile edge
    constexpr int NF = WT / 8;      // m8n8 fragments per warp tile edge
    __shared__ double As[GT][G_LDA];
    __shared__ double Bs[GK][G_LDB];
    const int b = blockIdx.z;
    A += (long)b * bsA;
    Bm += (long)b * bsB;
    C += (long)b * bsC;
    if (dvec) dvec += (long)b * bsd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * WT, wn = (warp & 1) * WT;
    const int g = lane >> 2, q = lane & 3;
    const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
    double acc[NF][NF][2];
#pragma unroll
    for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int c = 0; c < NF; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

    for (int k0 = 0; k0 < K; k0 += GK) {
        // stage A tile (GT x GK) and B tile (GK x GT); consecutive threads follow the unit stride
        for (int e = tid; e < GT * GK; e += 128) {
            int i, k;
            if (csA == 1) { i = e / GK; k = e % GK; } else { k = e / GT; i = e % GT; }
            double v = 0.0;
            if (i0 + i < M && k0 + k < K) {
                v = A[(long)(i0 + i) * rsA + (long)(k0 + k) * csA];
                if (dvec) v *= dvec[k0 + k];
            }
            As[i][k] = v;
        }
        for (int e = tid; e < GT * GK; e += 128) {
            int j, k;
            if (csB == 1) { k = e / GT; j = e % GT; } else { j = e / GK; k = e % GK; }
            double v = 0.0;
            if (j0 + j < N && k0 + k < K) v = Bm[(long)(k0 + k) * rsB + (long)(j0 + j) * csB];
            Bs[k][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK / 4; ++ks) {
            double af[NF], bf[NF];
#pragma unroll
            for (int a = 0; a < NF; ++a) af[a] = As[wm + 8 * a + g][4 * ks + q];
#pragma unroll
            for (int c = 0; c < NF; ++c) bf[c] = Bs[4 * ks + q][wn + 8 * c + g];
#pragma unroll
            for (int a = 0; a < NF; ++a)
#pragma unroll
                for (int c = 0; c < NF; ++c)
                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(acc[a][c][0]), "+d"(acc[a][c][1])
                        : "d"(af[a]), "d"(bf[c]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < NF; ++a) {
        const int i = i0 + wm + 8 * a + g;
        if (i >= M) continue;
#pragma unroll
        for (int c = 0; c < NF; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + wn + 8 * c + 2 * q + h;
                if (j >= N) continue;
                double *cp = C + (long)i * rsC + (long)j * csC;
                double v = alpha * acc[a][c][h];
                if (beta != 0.0) v += beta * (*cp);
                *cp = v;
                if (sym && blockIdx.x != blockIdx.y) C[(long)j * rsC + (long)i * csC] = v;
            }
        }
    }
}

struct Mat {  // strided batched matrix view
    const double *p;
    long rs, cs, bs;
};
inline Mat mat(const double *p, long rs, long cs, long bs) { return Mat{p, rs, cs, bs}; }
inline Mat tr(Mat a) { return Mat{a.p, a.cs, a.rs, a.bs}; }

inline int gemm(cudaStream_t st, int batch, int M, int N, int K, double alpha, Mat A, Mat B,
                double beta, double *C, long rsC, long csC, long bsC,
                const double *dvec = nullptr, long bsd = 0, bool symmetric = false) {
    if (M <= 0 || N <= 0 || batch <= 0) return DDMPC_OK;
    const int sym = (symmetric && M == N && beta == 0.0) ? 1 : 0;
    const long tiles64 = (long)ceil_div(N, 64) * ceil_div(M, 64) * batch;
    // executed area with 64- and 32-wide tiles: small matrices (136 rows = 3 x 64 or 5 x 32) waste less with the latter
    const long area64 = (long)ceil_div(N, 64) * ceil_div(M, 64) * 4096, area32 = (long)ceil_div(N, 32) * ceil_div(M, 32) * 1024;
    if (tiles64 >= 148 && 4 * area32 > 3 * area64) {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 64), batch);
        k_gemm<64><<<grid, 128, 0, st>>>(M, N, K, alpha, A.p, A.rs, A.cs, A.bs, B.p, B.rs, B.cs, B.bs, dvec, bsd, beta,
                                         C, rsC, csC, bsC, sym);
    } else {   // not enough 64x64 tiles for one wave, or too much padding in them: quarter-size tiles
        dim3 grid(ceil_div(N, 32), ceil_div(M, 32), batch);
        k_gemm<32><<<grid, 128, 0, st>>>(M, N, K, alpha, A.p, A.rs, A.cs, A.bs, B.p, B.rs, B.cs, B.bs, dvec, bsd, beta,
                                         C, rsC, csC, bsC, sym);
    }
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// ---------------------------------------------------------------------------
// In-place lower Cholesky of a row-major n x n matrix (ld), one CTA per batch
// entry.  info[b] = 0 or (1 + index of the first non-positive pivot).
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(1024)
k_potrf(int n, double *__restrict__ A, long ld, long bs, int *__restrict__ info) {
    A += (long)blockIdx.x * bs;
    const int tid = threadIdx.x, T = blockDim.x;
    __shared__ double piv;
    __shared__ int bad;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            double d = A[(long)j * ld + j];
            if (!(d > 0.0) || !isfinite(d)) {
                if (!bad) bad = j + 1;
                d = 1.0;  // keep going with finite numbers; the verdict is in info
            }
            d = sqrt(d);
            A[(long)j * ld + j] = d;
            piv = d;
        }
        __syncthreads();
        const double inv = 1.0 / piv;
        for (int i = j + 1 + tid; i < n; i += T) A[(long)i * ld + j] *= inv;
        __syncthreads();
        const int rem = n - j - 1;
        for (long idx = tid; idx < (long)rem * rem; idx += T) {
            const int i = j + 1 + (int)(idx / rem);
            const int k = j + 1 + (int)(idx % rem);
            if (k <= i) A[(long)i * ld + k] = fma(-A[(long)i * ld + j], A[(long)k * ld + j], A[(long)i * ld + k]);
        }
        __syncthreads();
    }
    if (tid == 0 && info) info[blockIdx.x] = bad;
}

// Same factorisation with the lower triangle held in shared memory (packed, row-major: (i, j) at i(i+1)/2 + j) for
// matrices that fit (n <= 208): the unblocked right-looking update moves n^3/6 elements, which is L2 traffic in
// k_potrf and shared-memory traffic here.  One CTA of 256 threads per matrix; a warp per row of the trailing update.
// off > 0: the matrix is a diagonal block (first row `off`) of a blocked factorisation: info keeps the first failure.
static __global__ void __launch_bounds__(256)
k_potrf_smem(int n, double *__restrict__ A, long ld, long bs, int *__restrict__ info, int off = 0) {
    extern __shared__ double sh[];          // Lp[n(n+1)/2], col[n], dg[n]
    A += (long)blockIdx.x * bs;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
    double *Lp = sh, *col = sh + (size_t)n * (n + 1) / 2, *dg = col + n;
    __shared__ int bad;
    if (tid == 0) bad = 0;
    for (int i = warp; i < n; i += NW) {
        const double *row = A + (long)i * ld;
        double *dst = Lp + (size_t)i * (i + 1) / 2;
        for (int j = lane; j <= i; j += 32) dst[j] = row[j];
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        double d = Lp[(size_t)j * (j + 1) / 2 + j];      // not written during this column (the square root goes to dg)
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0 && !bad) bad = j + 1;
            d = 1.0;                                     // keep going with finite numbers; the verdict is in info
        }
        d = sqrt(d);
        const double inv = 1.0 / d;
        if (tid == 0) dg[j] = d;
        for (int i = j + 1 + tid; i < n; i += T) {
            const size_t e = (size_t)i * (i + 1) / 2 + j;
            const double v = Lp[e] * inv;
            Lp[e] = v;
            col[i] = v;
        }
        __syncthreads();
        for (int i = j + 1 + warp; i < n; i += NW) {
            const double ci = col[i];
            double *row = Lp + (size_t)i * (i + 1) / 2;
            for (int k = j + 1 + lane; k <= i; k += 32) row[k] = fma(-ci, col[k], row[k]);
        }
        __syncthreads();
    }
    for (int i = warp; i < n; i += NW) {
        double *row = A + (long)i * ld;
        const double *src = Lp + (size_t)i * (i + 1) / 2;
        for (int j = lane; j < i; j += 32) row[j] = src[j];
        if (lane == 0) row[i] = dg[i];
    }
    if (tid == 0 && info) {
        if (off == 0) info[blockIdx.x] = bad;
        else if (bad && !info[blockIdx.x]) info[blockIdx.x] = bad + off;
    }
}

// Row-wise triangular solve of a panel: each row a of A21 (m x jb, ld) <- a L11^-T, i.e. x L11^T = a (forward
// substitution along the row).  One thread per row; L11 (jb x jb, ld) is read by every thread at the same address.
static __global__ void __launch_bounds__(64)
k_trsm_rows(int m, int jb, const double *__restrict__ L11, double *__restrict__ A21, long ld, long bs) {
    extern __shared__ double xs[];          // [jb][blockDim]: the row being solved stays on chip (a thread re-reads every
                                            // entry it has produced: through global memory that was an L2 round trip each)
    const int i = blockIdx.x * blockDim.x + threadIdx.x, T = blockDim.x, tid = threadIdx.x;
    if (i >= m) return;
    L11 += (long)blockIdx.y * bs;
    double *row = A21 + (long)blockIdx.y * bs + (long)i * ld;
    for (int k = 0; k < jb; ++k) xs[k * T + tid] = row[k];
    for (int k = 0; k < jb; ++k) {
        const double *Lk = L11 + (long)k * ld;
        double a0 = xs[k * T + tid], a1 = 0.0;
        int l = 0;
        for (; l + 1 < k; l += 2) {
            a0 = fma(-__ldg(Lk + l), xs[l * T + tid], a0);
            a1 = fma(-__ldg(Lk + l + 1), xs[(l + 1) * T + tid], a1);
        }
        if (l < k) a0 = fma(-__ldg(Lk + l), xs[l * T + tid], a0);
        xs[k * T + tid] = (a0 + a1) / __ldg(Lk + k);
    }
    for (int k = 0; k < jb; ++k) row[k] = xs[k * T + tid];
}

inline int potrf(cudaStream_t st, int batch, int n, double *A, long ld, long bs, int *info) {
    if (n <= 0 || batch <= 0) return DDMPC_OK;
    const size_t sh = sizeof(double) * ((size_t)n * (n + 1) / 2 + 2 * (size_t)n);
    if (sh <= 200 * 1024) {
        static std::atomic<unsigned long long> attr_done{0};
        if (first_time_on_device(attr_done))
            DDMPC_CUDA(cudaFuncSetAttribute(k_potrf_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k_potrf_smem<<<batch, 256, sh, st>>>(n, A, ld, bs, info, 0);
        DDMPC_LAUNCH_CHECK();
        return DDMPC_OK;
    }
    // Larger matrices (config 4: r = 480, n_f = 560): blocked right-looking factorisation.  The diagonal block is
    // factorised in shared memory, the panel below it solved one row per thread, and the trailing update
    // A22 -= L21 L21^T - where the flops are - runs on the FP64 tensor cores (k_gemm).  The unblocked one-CTA kernel
    // took 14.9 ms for n = 480; this takes well under a millisecond.
    constexpr int NBK = 64;
    const size_t shb = sizeof(double) * ((size_t)NBK * (NBK + 1) / 2 + 2 * (size_t)NBK);
    for (int j0 = 0; j0 < n; j0 += NBK) {
        const int jb = std::min(NBK, n - j0), m2 = n - j0 - jb;
        double *A11 = A + (long)j0 * ld + j0;
        k_potrf_smem<<<batch, 256, shb, st>>>(jb, A11, ld, bs, info, j0);
        DDMPC_LAUNCH_CHECK();
        if (m2 <= 0) break;
        double *A21 = A + (long)(j0 + jb) * ld + j0, *A22 = A + (long)(j0 + jb) * ld + (j0 + jb);
        dim3 gr(ceil_div(m2, 64), batch);
        k_trsm_rows<<<gr, 64, sizeof(double) * jb * 64, st>>>(m2, jb, A11, A21, ld, bs);
        DDMPC_LAUNCH_CHECK();
        const Mat P21 = mat(A21, ld, 1, bs);
        DDMPC_TRY(gemm(st, batch, m2, m2, jb, -1.0, P21, tr(P21), 1.0, A22, ld, 1, bs));
    }
    return DDMPC_OK;
}

// ---------------------------------------------------------------------------
// Triangular solves with the lower Cholesky factor L (row-major, ldl):
//   trans == 0:  L   X = B        trans == 1:  L^T X = B
// B (n x nrhs, row-major, ldb) is overwritten by X.  One thread per RHS column.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(128)
k_trsm(int n, int nrhs, int trans, const double *__restrict__ Lm, long ldl, long bsl,
       double *__restrict__ Bm, long ldb, long bsb) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nrhs) return;
    Lm += (long)blockIdx.y * bsl;
    Bm += (long)blockIdx.y * bsb;
    if (!trans) {
        for (int i = 0; i < n; ++i) {
            double a0 = Bm[(long)i * ldb + j], a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double *Li = Lm + (long)i * ldl;
            int k = 0;
            for (; k + 3 < i; k += 4) {
                a0 = fma(-Li[k], Bm[(long)k * ldb + j], a0);
                a1 = fma(-Li[k + 1], Bm[(long)(k + 1) * ldb + j], a1);
                a2 = fma(-Li[k + 2], Bm[(long)(k + 2) * ldb + j], a2);
                a3 = fma(-Li[k + 3], Bm[(long)(k + 3) * ldb + j], a3);
            }
            for (; k < i; ++k) a0 = fma(-Li[k], Bm[(long)k * ldb + j], a0);
            Bm[(long)i * ldb + j] = ((a0 + a1) + (a2 + a3)) / Li[i];
        }
    } else {
        for (int i = n - 1; i >= 0; --i) {
            double a0 = Bm[(long)i * ldb + j], a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = i + 1;
            for (; k + 3 < n; k += 4) {
                a0 = fma(-Lm[(long)k * ldl + i], Bm[(long)k * ldb + j], a0);
                a1 = fma(-Lm[(long)(k + 1) * ldl + i], Bm[(long)(k + 1) * ldb + j], a1);
                a2 = fma(-Lm[(long)(k + 2) * ldl + i], Bm[(long)(k + 2) * ldb + j], a2);
                a3 = fma(-Lm[(long)(k + 3) * ldl + i], Bm[(long)(k + 3) * ldb + j], a3);
            }
            for (; k < n; ++k) a0 = fma(-Lm[(long)k * ldl + i], Bm[(long)k * ldb + j], a0);
            Bm[(long)i * ldb + j] = ((a0 + a1) + (a2 + a3)) / Lm[(long)i * ldl + i];
        }
    }
}

// Same solves for a small diagonal block (n <= 64) with the right-hand side held in shared memory [n][blockDim].
static __global__ void __launch_bounds__(64)
k_trsm_blk(int n, int nrhs, int trans, const double *__restrict__ Lm, long ldl, long bsl,
           double *__restrict__ Bm, long ldb, long bsb) {
    extern __shared__ double xs[];
    const int j = blockIdx.x * blockDim.x + threadIdx.x, T = blockDim.x, tid = threadIdx.x;
    if (j >= nrhs) return;
    Lm += (long)blockIdx.y * bsl;
    Bm += (long)blockIdx.y * bsb;
    for (int i = 0; i < n; ++i) xs[i * T + tid] = Bm[(long)i * ldb + j];
    if (!trans) {
        for (int i = 0; i < n; ++i) {
            const double *Li = Lm + (long)i * ldl;
            double a0 = xs[i * T + tid], a1 = 0.0;
            int k = 0;
            for (; k + 1 < i; k += 2) {
                a0 = fma(-__ldg(Li + k), xs[k * T + tid], a0);
                a1 = fma(-__ldg(Li + k + 1), xs[(k + 1) * T + tid], a1);
            }
            if (k < i) a0 = fma(-__ldg(Li + k), xs[k * T + tid], a0);
            xs[i * T + tid] = (a0 + a1) / __ldg(Li + i);
        }
    } else {
        for (int i = n - 1; i >= 0; --i) {
            double a0 = xs[i * T + tid], a1 = 0.0;
            int k = i + 1;
            for (; k + 1 < n; k += 2) {
                a0 = fma(-__ldg(Lm + (long)k * ldl + i), xs[k * T + tid], a0);
                a1 = fma(-__ldg(Lm + (long)(k + 1) * ldl + i), xs[(k + 1) * T + tid], a1);
            }
            if (k < n) a0 = fma(-__ldg(Lm + (long)k * ldl + i), xs[k * T + tid], a0);
            xs[i * T + tid] = (a0 + a1) / __ldg(Lm + (long)i * ldl + i);
        }
    }
    for (int i = 0; i < n; ++i) Bm[(long)i * ldb + j] = xs[i * T + tid];
}

// X = (L L^T)^-1 B in place
inline int potrs(cudaStream_t st, int batch, int n, int nrhs, const double *L, long ldl, long bsl,
                 double *B, long ldb, long bsb) {
    if (n <= 0 || nrhs <= 0 || batch <= 0) return DDMPC_OK;
    const int threads = nrhs >= 128 ? 128 : (nrhs >= 64 ? 64 : 32);
    if (n >= 96 && (long)batch * nrhs < 8192) {
        // Few right-hand sides in total (a single controller, or a small set): one thread per right-hand side leaves
        // the GPU to a handful of threads walking n^2 / 2 dependent steps each (0.6-1.1 ms for n = 136, 12 ms for
        // n = 480).  Blocked substitution: the diagonal block is solved per thread as before (NB steps), the rest of
        // the right-hand side is updated with one tensor-core GEMM per block.
        const int NB = n >= 256 ? 64 : 32;
        const Mat Lall = mat(L, ldl, 1, bsl);
        for (int j0 = 0; j0 < n; j0 += NB) {                       // L X = B, top to bottom
            const int jb = std::min(NB, n - j0), m2 = n - j0 - jb;
            dim3 grid(ceil_div(nrhs, 64), batch);
            k_trsm_blk<<<grid, 64, sizeof(double) * jb * 64, st>>>(jb, nrhs, 0, L + (long)j0 * ldl + j0, ldl, bsl, B + (long)j0 * ldb, ldb, bsb);
            DDMPC_LAUNCH_CHECK();
            if (m2 > 0)
                DDMPC_TRY(gemm(st, batch, m2, nrhs, jb, -1.0, mat(L + (long)(j0 + jb) * ldl + j0, ldl, 1, bsl),
                               mat(B + (long)j0 * ldb, ldb, 1, bsb), 1.0, B + (long)(j0 + jb) * ldb, ldb, 1, bsb));
        }
        const int last = ((n - 1) / NB) * NB;
        for (int j0 = last; j0 >= 0; j0 -= NB) {                    // L^T X = B, bottom to top
            const int jb = std::min(NB, n - j0);
            dim3 grid(ceil_div(nrhs, 64), batch);
            k_trsm_blk<<<grid, 64, sizeof(double) * jb * 64, st>>>(jb, nrhs, 1, L + (long)j0 * ldl + j0, ldl, bsl, B + (long)j0 * ldb, ldb, bsb);
            DDMPC_LAUNCH_CHECK();
            if (j0 > 0)     // B[0:j0, :] -= L[j0:j0+jb, 0:j0]^T X[j0:j0+jb, :]
                DDMPC_TRY(gemm(st, batch, j0, nrhs, jb, -1.0, tr(mat(L + (long)j0 * ldl, ldl, 1, bsl)),
                               mat(B + (long)j0 * ldb, ldb, 1, bsb), 1.0, B, ldb, 1, bsb));
        }
        (void)Lall;
        return DDMPC_OK;
    }
    dim3 grid(ceil_div(nrhs, threads), batch);
    k_trsm<<<grid, threads, 0, st>>>(n, nrhs, 0, L, ldl, bsl, B, ldb, bsb);
    DDMPC_LAUNCH_CHECK();
    k_trsm<<<grid, threads, 0, st>>>(n, nrhs, 1, L, ldl, bsl, B, ldb, bsb);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// ---------------------------------------------------------------------------
// Numerical rank of a symmetric PSD matrix by elimination with diagonal pivoting (the pivots
// of a rank-revealing Cholesky / LDL^T).  One CTA per batch entry; A (n x n, ld) is destroyed.
// rank[b] = number of pivots > rel * (first pivot).  Rows/columns are never swapped: an `alive`
// flag marks what is left, and each step is one rank-1 update of the alive block.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(1024)
k_pivot_rank(int n, double *__restrict__ A, long ld, long bs, double rel, int *__restrict__ rank) {
    extern __shared__ double sh[];          // colp[n], alive[n] (as double), red[64]
    A += (long)blockIdx.x * bs;
    const int tid = threadIdx.x, T = blockDim.x;
    double *colp = sh, *alive = sh + n, *red = sh + 2 * n;
    int *redi = reinterpret_cast<int *>(red + 32);
    __shared__ double s_piv, s_first;
    __shared__ int s_idx, s_stop;
    if (rank[blockIdx.x] == n) return;      // already certified full rank by the shifted Cholesky (pe_rank_device stage 0)
    for (int i = tid; i < n; i += T) alive[i] = 1.0;
    if (tid == 0) { s_stop = 0; s_first = 0.0; }
    __syncthreads();
    int r = 0;
    for (; r < n; ++r) {
        // pivot = largest alive diagonal entry
        double best = -1.0;
        int bi = -1;
        for (int i = tid; i < n; i += T)
            if (alive[i] != 0.0) {
                const double d = A[(long)i * ld + i];
                if (d > best) { best = d; bi = i; }
            }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
        }
        if ((tid & 31) == 0) { red[tid >> 5] = best; redi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = -1.0;
            int ix = -1;
            for (int w = 0; w < (T + 31) / 32; ++w)
                if (red[w] > b || (red[w] == b && redi[w] >= 0 && (ix < 0 || redi[w] < ix))) { b = red[w]; ix = redi[w]; }
            if (r == 0) s_first = b;
            s_piv = b;
            s_idx = ix;
            s_stop = (ix < 0 || !(b > rel * s_first) || !(b > 0.0)) ? 1 : 0;
        }
        __syncthreads();
        if (s_stop) break;
        const int pv = s_idx;
        const double inv = 1.0 / s_piv;
        for (int i = tid; i < n; i += T) colp[i] = (alive[i] != 0.0 && i != pv) ? A[(long)i * ld + pv] : 0.0;
        __syncthreads();
        if (tid == 0) alive[pv] = 0.0;
        // rank-1 update of the alive block: A[i][j] -= A[i][pv] A[pv][j] / A[pv][pv]
        for (long e = tid; e < (long)n * n; e += T) {
            const int i = (int)(e / n), j = (int)(e % n);
            const double ci = colp[i], cj = colp[j];
            if (ci != 0.0 && cj != 0.0) A[(long)i * ld + j] = fma(-ci * inv, cj, A[(long)i * ld + j]);
        }
        __syncthreads();
    }
    if (tid == 0) rank[blockIdx.x] = r;
}

inline int pivot_rank(cudaStream_t st, int batch, int n, double *A, long ld, long bs, double rel, int *rank) {
    if (n <= 0 || batch <= 0) return DDMPC_OK;
    const size_t sh = (size_t)(2 * n + 64) * sizeof(double);
    const int threads = n >= 128 ? 1024 : (n >= 48 ? 512 : 256);
    k_pivot_rank<<<batch, threads, sh, st>>>(n, A, ld, bs, rel, rank);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// Stage 0 of the rank test: A <- A - tau I with tau = rel * max diag(A) (one CTA per batch entry).  If the Cholesky
// factorisation of the shifted matrix succeeds, lambda_min(A) > tau (inertia), i.e. A is full rank with a wide margin.
static __global__ void __launch_bounds__(256)
k_shift_diag(int n, double *__restrict__ A, long ld, long bs, double rel) {
    A += (long)blockIdx.x * bs;
    __shared__ double red[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, A[(long)i * ld + i]);
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    const double tau = rel * red[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) A[(long)i * ld + i] -= tau;
}
// rank[b] = n when the factorisation of entry b succeeded (info[b] == 0), else -1 (undecided)
static __global__ void k_rank_from_info(int batch, int n, const int *__restrict__ info, int *__restrict__ rank) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) rank[b] = info[b] == 0 ? n : -1;
}

// ---------------------------------------------------------------------------
// Second-stage rank test on the matrix itself (not its Gram matrix): row-pivoted modified Gram-Schmidt (an LQ
// factorisation with pivoting) of M (R x C, row-major, destroyed), one CTA per batch entry, one warp per row.
// Entries whose first-stage verdict rank_io[b] already equals `full` are skipped.  A pivot counts while the
// largest remaining row norm (recomputed exactly every step) exceeds max(R, C) * eps * ||M||_F, the cut of
// np.linalg.matrix_rank (S > S.max() * max(M, N) * eps, hankel_matrix.py:82) with ||M||_F standing in for S.max().
// Resolves singular-value ratios down to ~1e-13, where the Gram test stops at ~1e-6.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(512)
k_rowqr_rank(int R, int C, double *__restrict__ M, long bs, int full, int *__restrict__ rank_io) {
    extern __shared__ double sh[];          // nrm2[R], alive[R]
    const int b = blockIdx.x;
    if (rank_io[b] == full) return;
    M += (long)b * bs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *nrm2 = sh, *alive = sh + R;
    __shared__ double s_thr, s_piv2;
    __shared__ int s_pv;
    for (int i = tid; i < R; i += blockDim.x) alive[i] = 1.0;
    __syncthreads();
    int r = 0;
    for (; r < R; ++r) {
        for (int i = warp; i < R; i += NW) {
            if (alive[i] == 0.0) continue;
            const double *row = M + (long)i * C;
            double acc = 0.0;
            for (int j = lane; j < C; j += 32) acc = fma(row[j], row[j], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) nrm2[i] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            double best = -1.0, fro2 = 0.0;
            int bi = -1;
            for (int i = 0; i < R; ++i)
                if (alive[i] != 0.0) {
                    fro2 += nrm2[i];
                    if (nrm2[i] > best) { best = nrm2[i]; bi = i; }
                }
            if (r == 0) s_thr = (double)(R > C ? R : C) * 2.220446049250313e-16 * sqrt(fro2);
            s_pv = (bi >= 0 && sqrt(best) > s_thr) ? bi : -1;
            s_piv2 = best;
            if (s_pv >= 0) alive[s_pv] = 0.0;
        }
        __syncthreads();
        const int pv = s_pv;
        if (pv < 0) break;
        const double inv = 1.0 / s_piv2;
        const double *prow = M + (long)pv * C;
        for (int i = warp; i < R; i += NW) {
            if (alive[i] == 0.0) continue;
            double *row = M + (long)i * C;
            double acc = 0.0;
            for (int j = lane; j < C; j += 32) acc = fma(row[j], prow[j], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const double d = acc * inv;
            for (int j = lane; j < C; j += 32) row[j] = fma(-d, prow[j], row[j]);
        }
        __syncthreads();
    }
    if (tid == 0) rank_io[b] = r;
}

inline int rowqr_rank(cudaStream_t st, int batch, int R, int C, double *M, long bs, int full, int *rank_io) {
    if (R <= 0 || batch <= 0) return DDMPC_OK;
    k_rowqr_rank<<<batch, 512, sizeof(double) * 2 * R, st>>>(R, C, M, bs, full, rank_io);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

// ---------------------------------------------------------------------------
// Symmetric eigen-decomposition by parallel-ordered cyclic Jacobi.  One CTA per
// batch entry; A (n x n, ld) is destroyed (its diagonal ends as the spectrum),
// V (n x n, ldv; may be NULL) receives the eigenvectors as columns, lam (n) the
// eigenvalues (unsorted).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_pair(int n2, int r, int k, int &p, int &q) {
    int a, b;
    if (k == 0) { a = n2 - 1; b = r; }
    else { a = (r + k) % (n2 - 1); b = (r - k + (n2 - 1)) % (n2 - 1); }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

static __global__ void __launch_bounds__(1024)
k_jacobi(int n, double *__restrict__ A, long ld, long bsA, double *__restrict__ V, long ldv, long bsV,
         double *__restrict__ lam, long bsl, int max_sweeps) {
    extern __shared__ double sh[];  // c[n2/2], s[n2/2], red[32]
    A += (long)blockIdx.x * bsA;
    if (V) V += (long)blockIdx.x * bsV;
    lam += (long)blockIdx.x * bsl;
    const int tid = threadIdx.x, T = blockDim.x;
    const int n2 = n + (n & 1), np2 = n2 / 2;
    double *cs_c = sh, *cs_s = sh + np2, *red = sh + 2 * np2;
    __shared__ int done;
    __shared__ double prev_off;
    if (threadIdx.x == 0) prev_off = 1e300;
    if (V)
        for (long e = tid; e < (long)n * n; e += T) V[(e / n) * ldv + (e % n)] = (e / n == e % n) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        // off-diagonal and total Frobenius mass
        double off = 0.0, tot = 0.0;
        for (long e = tid; e < (long)n * n; e += T) {
            const int i = (int)(e / n), j = (int)(e % n);
            const double v = A[(long)i * ld + j];
            tot += v * v;
            if (i != j) off += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if ((tid & 31) == 0) { red[tid >> 5] = off; red[32 + (tid >> 5)] = tot; }
        __syncthreads();
        if (tid == 0) {
            double so = 0.0, stt = 0.0;
            for (int w = 0; w < (T + 31) / 32; ++w) { so += red[w]; stt += red[32 + w]; }
            // converged, or at the rounding floor: the off-diagonal mass falls quadratically until rotations can no longer
            // reduce it (~ n eps^2 of the total, which for n > 100 lies above 1e-29) and then stalls
            done = (so <= 1e-29 * stt || (so <= 1e-22 * stt && so > 0.25 * prev_off)) ? 1 : 0;
            prev_off = so;
        }
        __syncthreads();
        if (done) break;
        for (int r = 0; r < n2 - 1; ++r) {
            for (int k = tid; k < np2; k += T) {
                int p, q;
                jacobi_pair(n2, r, k, p, q);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = A[(long)p * ld + q];
                    if (fabs(apq) > 1e-300) {
                        const double tau = (A[(long)q * ld + q] - A[(long)p * ld + p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                cs_c[k] = c;
                cs_s[k] = s;
            }
            __syncthreads();
            // column rotations  A <- A J,  V <- V J
            for (long e = tid; e < (long)np2 * n; e += T) {
                // a warp works on ONE row and 32 different pairs: its loads stay inside a few cache lines
                // (one row per lane would touch 32 different lines per load)
                const int i = (int)(e / np2), k = (int)(e % np2);
                int p, q;
                jacobi_pair(n2, r, k, p, q);
                if (q >= n) continue;
                const double c = cs_c[k], s = cs_s[k];
                if (s == 0.0) continue;
                const double ap = A[(long)i * ld + p], aq = A[(long)i * ld + q];
                A[(long)i * ld + p] = c * ap - s * aq;
                A[(long)i * ld + q] = s * ap + c * aq;
                if (V) {
                    const double vp = V[(long)i * ldv + p], vq = V[(long)i * ldv + q];
                    V[(long)i * ldv + p] = c * vp - s * vq;
                    V[(long)i * ldv + q] = s * vp + c * vq;
                }
            }
            __syncthreads();
            // row rotations  A <- J^T A
            for (long e = tid; e < (long)np2 * n; e += T) {
                const int k = (int)(e / n), j = (int)(e % n);
                int p, q;
                jacobi_pair(n2, r, k, p, q);
                if (q >= n) continue;
                const double c = cs_c[k], s = cs_s[k];
                if (s == 0.0) continue;
                const double ap = A[(long)p * ld + j], aq = A[(long)q * ld + j];
                A[(long)p * ld + j] = c * ap - s * aq;
                A[(long)q * ld + j] = s * ap + c * aq;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += T) lam[i] = A[(long)i * ld + i];
}

// Same algorithm with A resident in shared memory (n <= 158: 200 KB), one warp per row (column rotations) or per pair
// (row rotations), the pairs of a round tabulated once: no global traffic for A, no integer division per element.  The
// four-tank Gram matrices (136 x 136) take ~1/5 of the time of the global-memory version.
static __global__ void __launch_bounds__(1024)
k_jacobi_smem(int n, double *__restrict__ A, long ld, long bsA, double *__restrict__ V, long ldv, long bsV,
              double *__restrict__ lam, long bsl, int max_sweeps) {
    extern __shared__ __align__(16) double sh[];
    A += (long)blockIdx.x * bsA;
    if (V) V += (long)blockIdx.x * bsV;
    lam += (long)blockIdx.x * bsl;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int n2 = n + (n & 1), np2 = n2 / 2;
    const int lds = n | 1;                            // odd row stride: a column read by 32 lanes (rows) hits 32 different banks
    double *As = sh;                                  // n x lds
    double *cs_c = As + (size_t)n * lds, *cs_s = cs_c + np2, *red = cs_s + np2;   // red[64]
    int *pq = reinterpret_cast<int *>(red + 64);      // (p, q) of the pairs of the current round
    __shared__ int done;
    __shared__ double prev_off;
    if (threadIdx.x == 0) prev_off = 1e300;
    for (int i = warp; i < n; i += W)
        for (int j = lane; j < n; j += 32) {
            As[(size_t)i * lds + j] = A[(long)i * ld + j];
            if (V) V[(long)i * ldv + j] = (i == j) ? 1.0 : 0.0;
        }
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, tot = 0.0;
        for (int i = warp; i < n; i += W)
            for (int j = lane; j < n; j += 32) {
                const double v = As[(size_t)i * lds + j];
                tot += v * v;
                if (i != j) off += v * v;
            }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if (lane == 0) { red[warp] = off; red[32 + warp] = tot; }
        __syncthreads();
        if (tid == 0) {
            double so = 0.0, stt = 0.0;
            for (int w = 0; w < W; ++w) { so += red[w]; stt += red[32 + w]; }
            // converged, or at the rounding floor: the off-diagonal mass falls quadratically until rotations can no longer
            // reduce it (~ n eps^2 of the total, which for n > 100 lies above 1e-29) and then stalls
            done = (so <= 1e-29 * stt || (so <= 1e-22 * stt && so > 0.25 * prev_off)) ? 1 : 0;
            prev_off = so;
        }
        __syncthreads();
        if (done) break;
        for (int r = 0; r < n2 - 1; ++r) {
            for (int k = tid; k < np2; k += T) {
                int p, q;
                jacobi_pair(n2, r, k, p, q);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = As[(size_t)p * lds + q];
                    if (fabs(apq) > 1e-300) {
                        const double tau = (As[(size_t)q * lds + q] - As[(size_t)p * lds + p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                cs_c[k] = c;
                cs_s[k] = (q < n) ? s : 0.0;
                pq[2 * k] = p;
                pq[2 * k + 1] = q;
            }
            __syncthreads();
            // column rotations  A <- A J : a warp per pair, lanes down the two columns (conflict-free with the odd stride;
            // a lane per pair along a row costs ~4 shared-memory wavefronts per access and was 3/4 of the kernel's time)
            for (int k = warp; k < np2; k += W) {
                const double s = cs_s[k];
                if (s == 0.0) continue;
                const double c = cs_c[k];
                double *cp = As + pq[2 * k], *cq = As + pq[2 * k + 1];
                for (int i = lane; i < n; i += 32) {
                    const double ap = cp[(size_t)i * lds], aq = cq[(size_t)i * lds];
                    cp[(size_t)i * lds] = c * ap - s * aq;
                    cq[(size_t)i * lds] = s * ap + c * aq;
                }
            }
            __syncthreads();
            // row rotations  A <- J^T A  and  V^T <- J^T V^T : a warp per pair, lanes along the row.  The eigenvectors are
            // accumulated TRANSPOSED (rows p, q of V^T are two contiguous runs in global memory: fully coalesced, where
            // columns p, q of V cost four times the L2 traffic in partial sectors) and transposed in place at the end.
            for (int k = warp; k < np2; k += W) {
                const double s = cs_s[k];
                if (s == 0.0) continue;
                const double c = cs_c[k];
                const int p = pq[2 * k], q = pq[2 * k + 1];
                double *rp = As + (size_t)p * lds, *rq = As + (size_t)q * lds;
                for (int j = lane; j < n; j += 32) {
                    const double ap = rp[j], aq = rq[j];
                    rp[j] = c * ap - s * aq;
                    rq[j] = s * ap + c * aq;
                }
                if (V) {
                    double *vp = V + (long)p * ldv, *vq = V + (long)q * ldv;
                    for (int j = lane; j < n; j += 32) {
                        const double a = vp[j], b = vq[j];
                        vp[j] = c * a - s * b;
                        vq[j] = s * a + c * b;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += T) lam[i] = As[(size_t)i * lds + i];
    // (A is documented as destroyed: its copy in global memory is left as it was)
    if (V) {                                          // V^T -> V, in place
        __syncthreads();
        for (int i = warp; i < n; i += W)
            for (int j = lane; j < i; j += 32) {
                const double a = V[(long)i * ldv + j], b = V[(long)j * ldv + i];
                V[(long)i * ldv + j] = b;
                V[(long)j * ldv + i] = a;
            }
    }
}

inline int jacobi_eig(cudaStream_t st, int batch, int n, double *A, long ld, long bsA, double *V,
                      long ldv, long bsV, double *lam, long bsl) {
    if (n <= 0 || batch <= 0) return DDMPC_OK;
    {
        const int n2s = n + (n & 1);
        const size_t shs = sizeof(double) * ((size_t)n * (n | 1) + n2s + 64) + sizeof(int) * (size_t)n2s;
        if (shs <= 200 * 1024) {
            // (set on every call: the kernel is a per-translation-unit static, a shared "done" flag would not do)
            if (shs > 48 * 1024)
                DDMPC_CUDA(cudaFuncSetAttribute(k_jacobi_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            const int threads = n >= 96 ? 1024 : (n >= 48 ? 512 : 256);
            k_jacobi_smem<<<batch, threads, shs, st>>>(n, A, ld, bsA, V, ldv, bsV, lam, bsl, 40);
            DDMPC_LAUNCH_CHECK();
            return DDMPC_OK;
        }
    }
    const int n2 = n + (n & 1);
    const size_t sh = (size_t)(n2 + 64) * sizeof(double);
    const int threads = n >= 128 ? 1024 : (n >= 48 ? 512 : 256);
    k_jacobi<<<batch, threads, sh, st>>>(n, A, ld, bsA, V, ldv, bsV, lam, bsl, 40);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

}  // namespace ddmpc
