/*
 * ddmpc.h - C ABI of the B200-native DD-MPC hot path (libddmpc.so).
 *
 * The reference (pavelacamposp/direct_data_driven_mpc) is pure Python and has
 * no native layer, so there is no pre-existing FFI to mirror.  Each entry point
 * below names the reference interface it replaces (file:line, relative to the
 * reference checkout).  The reference-side binding a maintainer would add is a
 * ctypes stub; it is shown in INTEGRATION.md and shipped as
 * direct_data_driven_mpc_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.
 *   - all matrices are C-order (row-major) FP64, exactly the NumPy layout the
 *     reference uses.
 *   - functions without the _host suffix take DEVICE pointers (any allocator,
 *     e.g. torch) and enqueue on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  They do not synchronise unless stated.
 *   - *_host variants take HOST pointers, copy in/out and synchronise.
 *   - every function returns a ddmpc_status code, never throws; the message
 *     of the last failure on the calling thread is ddmpc_last_error().
 *   - error codes map one-to-one onto the reference's exceptions (see
 *     INTEGRATION.md): the Python facade raises the same ValueError /
 *     NotImplementedError the reference raises.
 */
#ifndef DDMPC_H
#define DDMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DDMPC_OK = 0,
    DDMPC_ERR_INVALID_ARG = 1,      /* shapes / sizes / NULLs                       */
    DDMPC_ERR_CUDA = 2,             /* CUDA runtime failure (message has details)   */
    DDMPC_ERR_CONTROLLER_TYPE = 3,  /* controller.py:165-168                        */
    DDMPC_ERR_SLACK_TYPE = 4,       /* controller.py:211-215                        */
    DDMPC_ERR_ROBUST_PARAMS = 5,    /* controller.py:217-222                        */
    DDMPC_ERR_N_TOO_SMALL = 6,      /* controller.py:275-283  (N < N_min)           */
    DDMPC_ERR_NOT_PE = 7,           /* controller.py:285-296  (Hankel rank)         */
    DDMPC_ERR_HORIZON = 8,          /* controller.py:316-325                        */
    DDMPC_ERR_NOT_IMPLEMENTED = 9,  /* controller.py:664-670  (NON_CONVEX)          */
    DDMPC_ERR_FACTORIZATION = 10,   /* reduced Hessian / Gram not positive definite */
    DDMPC_ERR_HANKEL_WINDOW = 11    /* hankel_matrix.py:43-44 (N < L)               */
} ddmpc_status;

/* integer codes follow the reference's YAML maps (controller_creation.py:12-23) */
enum { DDMPC_NOMINAL = 0, DDMPC_ROBUST = 1 };
enum { DDMPC_SLACK_NONE = 0, DDMPC_SLACK_CONVEX = 1, DDMPC_SLACK_NON_CONVEX = 2 };

/* per-scenario solve status (cvxpy status strings, controller.py:743-767) */
enum {
    DDMPC_SOLVE_OPTIMAL = 0,            /* "optimal"            */
    DDMPC_SOLVE_OPTIMAL_INACCURATE = 1, /* "optimal_inaccurate" (max_iter reached)  */
    DDMPC_SOLVE_INFEASIBLE = 2,         /* "infeasible"         */
    DDMPC_SOLVE_NONFINITE = 3           /* non-finite data / diverged loop          */
};

/* Constructor arguments of DirectDataDrivenMPCController (controller.py:95-116)
 * that do not carry array data. */
typedef struct {
    int32_t n, m, p;          /* estimated order, inputs, outputs                */
    int32_t N, L;             /* data length, prediction horizon                 */
    int32_t controller_type;  /* DDMPC_NOMINAL / DDMPC_ROBUST                    */
    int32_t slack_type;       /* DDMPC_SLACK_*                                   */
    int32_t use_terminal;     /* use_terminal_constraint                         */
    int32_t n_mpc_step;       /* n-step scheme: inputs applied per solve         */
    int32_t check_pe;         /* 1: run the persistency-of-excitation rank test  */
    double eps_max, lamb_alpha, lamb_sigma, c;
    /* Optional input box u_min <= ubar[k] <= u_max on every predicted input (paper Eq. 6, u in U; the
     * reference has no such constraint: controller.py:447-504).  HOST pointers to m values each, copied at
     * creation; NULL / NULL (the default) leaves the problem exactly as the reference states it.  Entries may
     * be -inf / +inf.  ROBUST controllers only (DDMPC_ERR_NOT_IMPLEMENTED for NOMINAL). */
    const double *u_min, *u_max;
    /* Optional output box y_min <= ybar[k] <= y_max on every predicted output (paper Eq. 6, y in Y), same
     * conventions: HOST pointers to p values each, NULL / NULL = as the reference, ROBUST only. */
    const double *y_min, *y_max;
} ddmpc_params;

/* LTI plant of utilities/model_simulation.py:31-98, row-major host arrays. */
typedef struct {
    int32_t n_x, m, p;
    const double *A, *B, *C, *D;  /* (n_x,n_x) (n_x,m) (p,n_x) (p,m), HOST memory */
} ddmpc_plant;

/* A set of `count` controllers with identical structure (params) that may
 * differ in data (u_d, y_d) and in the regularisation weights. */
typedef struct ddmpc_set ddmpc_set;

const char *ddmpc_version(void);
const char *ddmpc_strerror(int status);
const char *ddmpc_last_error(void);

/* ---- hankel_matrix(X, L)  (direct_data_driven_mpc/utilities/hankel_matrix.py:5-53)
 * X (N, n_ch) -> H (L*n_ch, N-L+1); bit-exact copy kernel. */
int ddmpc_hankel(const double *X, int N, int n_ch, int L, double *H, void *stream);
int ddmpc_hankel_host(const double *X, int N, int n_ch, int L, double *H);

/* ---- evaluate_persistent_excitation(X, order)  (hankel_matrix.py:55-87)
 * numerical rank of H_order(X) (pivots of a rank-revealing elimination of its
 * Gram matrix, see DESIGN.md "Numerics"); *rank is a HOST int; synchronises. */
int ddmpc_pe_rank_host(const double *X, int N, int n_ch, int order, int *rank);

/* ---- DirectDataDrivenMPCController.__init__ + initialize_data_driven_mpc
 * (controller.py:95-240, 345-387): validation, PE test, Hankel matrices, Gram
 * matrix, factorisations and the condensed solve operators, all on the GPU.
 * u_d (count or 1, N, m), y_d (count or 1, N, p): stride 0 shares one data set.
 * Q (p*L, p*L), R (m*L, m*L) shared by the set.  lamb_alpha / lamb_sigma may
 * be NULL (use params) or arrays of `count` values (HOST memory in both
 * variants).  Synchronises (the PE verdict is needed on the host). */
int ddmpc_set_create(const ddmpc_params *params, int count,
                     const double *u_d, size_t ud_stride,
                     const double *y_d, size_t yd_stride,
                     const double *Q, const double *R,
                     const double *lamb_alpha, const double *lamb_sigma,
                     void *stream, ddmpc_set **out);
int ddmpc_set_create_host(const ddmpc_params *params, int count,
                          const double *u_d, size_t ud_stride,
                          const double *y_d, size_t yd_stride,
                          const double *Q, const double *R,
                          const double *lamb_alpha, const double *lamb_sigma,
                          ddmpc_set **out);
void ddmpc_set_destroy(ddmpc_set *set);
int ddmpc_set_count(const ddmpc_set *set);
/* per-controller verdicts: PE rank (-1 if not tested) and DDMPC_OK / error */
int ddmpc_set_info(const ddmpc_set *set, int index, int *pe_rank, int *status);
/* Copy a named per-controller device matrix to the host (HLn_ud, HLn_yd, W,
 * Ku, Z, Ks, Phi, Psi, Lam, X0 ...).  Returns its element count in *n_elem. */
int ddmpc_set_get(const ddmpc_set *set, const char *name, int index,
                  double *out, size_t capacity, size_t *n_elem);

/* ---- update_and_solve_data_driven_mpc for a batch  (controller.py:389-407,
 * 739-808): B independent QP solves.  ctrl_idx (B) selects the controller of
 * each solve (NULL = controller 0).  u_past (B, n*m), y_past (B, n*p),
 * u_s (B, m), y_s (B, p) -> optimal_u (B, L*m), cost (B), status (B), iters (B).
 * cost / status / iters may be NULL. */
int ddmpc_solve_batch(const ddmpc_set *set, int B, const int32_t *ctrl_idx,
                      const double *u_past, const double *y_past,
                      const double *u_s, const double *y_s,
                      double tol, int max_iter,
                      double *optimal_u, double *cost, int32_t *status, int32_t *iters,
                      void *stream);
int ddmpc_solve_batch_host(const ddmpc_set *set, int B, const int32_t *ctrl_idx,
                           const double *u_past, const double *y_past,
                           const double *u_s, const double *y_s,
                           double tol, int max_iter,
                           double *optimal_u, double *cost, int32_t *status, int32_t *iters);

/* Full primal solution of each solve: ubar (B, (L+n)*m), ybar (B, (L+n)*p),
 * sigma (B, (L+n)*p) [robust], alpha (B, N-L-n+1) [robust: min-norm alpha].
 * Any output may be NULL.  Requires the set to have been created with the
 * environment default (operators for the full primal are always kept). */
int ddmpc_solve_full_batch(const ddmpc_set *set, int B, const int32_t *ctrl_idx,
                           const double *u_past, const double *y_past,
                           const double *u_s, const double *y_s,
                           double tol, int max_iter,
                           double *ubar, double *ybar, double *sigma, double *alpha,
                           void *stream);

/* ---- simulate_data_driven_mpc_control_loop for a batch
 * (utilities/controller/controller_operation.py:201-331; plant step
 * utilities/model_simulation.py:93-98; window update controller.py:893-895).
 * B closed loops advance in lockstep on the device for n_steps steps.
 *   x0 (B, n_x)  plant state at loop start
 *   u_past0 (B, n*m), y_past0 (B, n*p)  initial measurement window
 *   u_s (B, m), y_s (B, p)  set-points
 *   noise: w != NULL  -> w (B, n_steps, p) measurement noise, already scaled
 *                        (parity mode: NumPy PCG64 draws uploaded by the caller)
 *          w == NULL  -> device Philox4x32-10, key = noise_seed, stream id =
 *                        scenario_id0 + b, scaled by noise_eps (throughput mode)
 * Outputs u_sys (B, n_steps, m), y_sys (B, n_steps, p), status (B) = worst
 * solve status of the loop, iters (B) = total solver iterations (may be NULL).
 */
int ddmpc_closed_loop_batch(const ddmpc_set *set, const ddmpc_plant *plant, int B,
                            const int32_t *ctrl_idx,
                            const double *x0, const double *u_past0, const double *y_past0,
                            const double *u_s, const double *y_s,
                            const double *w, uint64_t noise_seed, uint64_t scenario_id0,
                            double noise_eps,
                            int n_steps, double tol, int max_iter,
                            double *u_sys, double *y_sys, int32_t *status, int32_t *iters,
                            double *x_final, void *stream);
int ddmpc_closed_loop_batch_host(const ddmpc_set *set, const ddmpc_plant *plant, int B,
                                 const int32_t *ctrl_idx,
                                 const double *x0, const double *u_past0, const double *y_past0,
                                 const double *u_s, const double *y_s,
                                 const double *w, uint64_t noise_seed, uint64_t scenario_id0,
                                 double noise_eps,
                                 int n_steps, double tol, int max_iter,
                                 double *u_sys, double *y_sys, int32_t *status, int32_t *iters,
                                 double *x_final);

/* ---- On-device scenario generation with NumPy-compatible streams (next-tier row):
 * stages 1-3 of examples/direct_data_driven_mpc_example.py:263-300 for S seeds at once =
 * randomize_initial_system_state (utilities/controller/controller_operation.py:59-75) then
 * generate_initial_input_output_data (:126-133).  Stream s is np.random.default_rng(seeds[s])
 * (SeedSequence -> PCG64 -> Generator.uniform, restated bit-for-bit), draws in the reference's
 * order.  plant matrices, pinv_Ot (n_x, p*n_x), Tt (p*n_x, m*n_x) and seeds are HOST arrays;
 * outputs are DEVICE arrays x0 (S, n_x), u_d (S, N, m), y_d (S, N, p), x_end (S, n_x) (plant
 * state after the data run) and rng_state (S, 4) uint64 (generator state to continue from).
 * Synchronises. */
int ddmpc_generate_example_data(const ddmpc_plant *plant, const double *pinv_Ot, const double *Tt, int S,
                                const uint64_t *seeds, int N, double u_lo, double u_hi, double eps,
                                double *x0, double *u_d, double *y_d, double *x_end, uint64_t *rng_state,
                                void *stream);
/* Continue the S streams: out (S, count) = scale * Generator.uniform(lo, hi, count), e.g. the loop noise
 * w_sys = eps_max * U(-1, 1, (n_steps, p)) of controller_operation.py:263.  Device arrays. */
int ddmpc_pcg64_uniform(uint64_t *rng_state, int S, int count, double lo, double hi, double scale, double *out,
                        void *stream);

/* Kernel selection of ddmpc_closed_loop_batch.  DDMPC_PATH_AUTO (the default) picks by controller structure, shape and
 * batch size; the others force one implementation where it applies (else the generic kernel runs) -
 * the parity tests use this to compare every kernel with every other on the same inputs. */
enum {
    DDMPC_PATH_AUTO = 0,
    DDMPC_PATH_GENERIC = 1,   /* k_closed_loop: thread per loop, any sizes / controller kinds                    */
    DDMPC_PATH_FAST = 2,      /* k_closed_loop_fast: register-resident hybrid DFMA + DMMA, shared controller      */
    DDMPC_PATH_WS = 3,        /* k_closed_loop_ws: warp-specialised all-tensor-core, four-tank n-step shape       */
    DDMPC_PATH_PERLOOP = 4,   /* k_closed_loop_perloop: 8 lanes per loop, per-loop controllers / small batches    */
    DDMPC_PATH_DMMA = 5,      /* k_closed_loop_dmma: config-4 shapes, warp per 8 loops on the FP64 tensor cores   */
    DDMPC_PATH_GEMM = 6,      /* closed_loop_gemm: batch as the N dimension of FP64 tensor-core GEMMs             */
    DDMPC_PATH_CVX = 7,       /* k_closed_loop_cvx: shared CONVEX controller, slack check + ADMM on the tensor cores */
    DDMPC_PATH_TC = 8         /* k_closed_loop_tc: config-4 shape on tcgen05 / TMEM, TF32x3 arithmetic (u within ~5e-7 of
                                 the FP64 kernels: inside the 1e-5 tolerance, but never selected automatically)           */
};
/* Per-set options: "closed_loop_path" (DDMPC_PATH_*), "dmma_warps" (1, 2 or 4: CTA size of the config-4 kernel),
 * "loops_per_thread" (0 = automatic, 1, 2: hybrid kernel), "solve_path" (0 = automatic, 1 = thread / CTA per solve kernels,
 * 2 = tensor-core ADMM pipeline where it applies), "cvx_ctas_per_sm" (2 or 3: register budget of k_closed_loop_cvx), "tc_passes" (1..3: TF32 passes of the tcgen05 path),
 * "trajectory_layout" (0 = u_sys [B, n_steps, m] / y_sys [B, n_steps, p] as documented at ddmpc_closed_loop_batch - the
 * reference's per-loop arrays stacked; 1 = step-major [n_steps, B, m] / [n_steps, B, p] for consumers that stay on the
 * device: written by k_closed_loop_ws only, DDMPC_ERR_NOT_IMPLEMENTED elsewhere).  Not thread-safe against running calls
 * on the same set. */
int ddmpc_set_option(ddmpc_set *set, const char *name, int value);

/* Number of controllers of the set whose setup failed (not persistently exciting / factorisation); their indices
 * are reported by ddmpc_set_info.  Solves and closed loops that use such a controller return status
 * DDMPC_SOLVE_NONFINITE and NaN outputs. */
int ddmpc_set_failed_count(const ddmpc_set *set);

/* ---- Roofline denominators, measured on the current device (csrc/probes.cu; bench.py calls them in the run whose
 * fractions it reports).  Both synchronise.
 *   ddmpc_probe_fp64_tflops: sustained FP64 TFLOP/s of the chip issued as DFMA (use_dmma = 0) or as DMMA m8n8k4.
 *   ddmpc_probe_store_ms:    best-of-5 time to write u (B, n_steps, 2) and y (B, n_steps, 2) (device buffers, 32-byte
 *                            aligned) with no compute, in the closed-loop kernels' own pattern (coalesced = 0: a thread
 *                            owns a loop and writes one 32-byte sector at a time) or fully coalesced (= 1). */
int ddmpc_probe_fp64_tflops(int use_dmma, double *tflops, void *stream);
int ddmpc_probe_store_ms(int B, int n_steps, int coalesced, double *u, double *y, double *ms, void *stream);

/* Number of kernels this library has launched since load (bench bookkeeping). */
uint64_t ddmpc_kernel_launches(void);

/* Plan storage and setup scratch come from the device's stream-ordered memory pool, which keeps freed memory for the
 * next setup (re-mapping gigabytes per call is what made batched setups slow).  This returns the unused part to the
 * driver (e.g. before handing the GPU to another allocator).  Synchronises the device. */
int ddmpc_trim_memory(void);

#ifdef __cplusplus
}
#endif
#endif /* DDMPC_H */
