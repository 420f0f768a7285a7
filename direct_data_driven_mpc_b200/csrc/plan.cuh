// Per-controller "plan": dimensions, index maps and the device-resident
// condensed solve operators produced by the setup pipeline (setup.cu) and
// consumed by the batched solver / fused closed loop (solve.cu).
#pragma once
#include <initializer_list>
#include <vector>

#include "common.cuh"

namespace ddmpc {

struct Dims {
    int n, m, p, N, L, Lp;
    int nu, ny;        // (L+n)*m, (L+n)*p
    int r, cols;       // Hankel stack: r = nu+ny rows, cols = N-L-n+1
    int nx;            // primal x = [ubar; ybar; sigma(robust)]
    int nfix, nf;      // fixed (initial/terminal) and free coordinates
    int nth;           // theta = [u_past; y_past; u_s; y_s]
    int nb;            // box rows = nbs + nbu + nby
    int nbs, nbu, nby; // CONVEX: sigma_pred rows (L*p) / input box: free predicted-input rows ((L-n)*m or L*m) /
                       // output box: free predicted-output rows ((L-n)*p or L*p)
    int Lm;            // L*m = len(optimal_u)
    int robust, convex, terminal;
};

// Device operators, each stored as [count][rows][cols] row-major FP64.
struct Plan {
    Dims d{};
    int count = 0;
    int data_count = 0;   // 1 when every controller shares one (u_d, y_d): H, W, Om exist once (lambda sweeps), else count
    DevBuf H;      // (r, cols)      stacked Hankel [H_u; H_y]   (HLn_ud / HLn_yd)
    DevBuf Om;     // (r, r)         W^-1 (W^+ with short / rank-deficient data), W = H H^T   (robust; alpha recovery)
    DevBuf W;      // (r, r)         Gram matrix (kept for inspection)
    DevBuf Ku;     // (Lm, nth)      optimal_u = Ku theta
    DevBuf Z;      // (nth, nth)     unconstrained optimal cost = theta^T Z theta
    DevBuf X0;     // (nx, nth)      full primal x0 = X0 theta
    DevBuf Ks;     // (nb, nth)      s_unc = Ks theta             (convex)
    DevBuf Phi;    // (nb, nb)       (I + rho/2 Lam)^-1           (convex)
    DevBuf Psi;    // (Lm, nb)       u = u0 + Psi (v - s)         (convex)
    DevBuf Lam;    // (nb, nb)       B A^-1 B^T                   (convex)
    DevBuf Yf;     // (nx, nb)       rho/2 * A^-1 B^T scattered   (convex; full primal)
    DevBuf rho2;   // (1)            rho/2
    DevBuf rs;     // (nb)           row scale of the box rows (sigma rows 1, input rows sqrt(mean Lam_ss / mean Lam_uu))
    DevBuf lo, hi; // (nb)           scaled bounds of the box rows
    DevBuf bmax;   // (1)            largest finite |scaled bound| (residual tolerance scale)
    DevBuf blo, bhi;   // (nb)       unscaled bounds, shared by the set
    DevBuf umin, umax; // (m)        input box (device copy), empty when absent
    DevBuf ymin, ymax; // (p)        output box (device copy), empty when absent
    DevBuf F;      // (nfix, nth)    feasibility residual map     (nominal)
    DevBuf Fnz;    // (1) int        1 when F has a non-zero entry (rank-deficient data); 0 = every window is feasible
    DevBuf lamA, lamS;  // per-controller lamb_alpha*eps_max, lamb_sigma
    std::vector<int> pe_rank, status;
    bool range_constrained = false;   // robust setup with a singular W: t kept in range(H) by a Schur complement (setup.cu)
    void detach_streams() {
        for (DevBuf *b : {&H, &Om, &W, &Ku, &Z, &X0, &Ks, &Phi, &Psi, &Lam, &Yf, &rho2, &rs, &lo, &hi, &bmax, &blo, &bhi,
                          &umin, &umax, &ymin, &ymax, &F, &Fnz, &lamA, &lamS})
            b->detach_stream();
    }
    double bound = 0.0;   // c * eps_max
    std::vector<double> u_min, u_max;   // host copy of the input box (empty = none)
};

}  // namespace ddmpc

struct ddmpc_set {
    ddmpc_params prm;
    ddmpc::Plan plan;
    // ddmpc_set_option(): kernel selection of ddmpc_closed_loop_batch (DDMPC_PATH_*), CTA size of the config-4 kernel,
    // loops per thread of the hybrid kernel (0 = automatic)
    int opt_path = 0, opt_dmma_warps = 1, opt_lpt = 0, opt_solve = 0, opt_cvx_ctas = 3, opt_tc_passes = 3, opt_layout = 0;
    // per-controller setup verdicts on the device (count ints, DDMPC_OK or the error): kernels report failed
    // controllers as DDMPC_SOLVE_NONFINITE with NaN outputs instead of finite garbage
    ddmpc::DevBuf ctrl_status;
    int n_failed = 0;
    // last plant uploaded by ddmpc_closed_loop_batch (re-used while unchanged, so the
    // launch path stays asynchronous)
    mutable std::vector<double> plant_host;
    mutable ddmpc::DevBuf plant_dev;
    // fast_loop.cu: host copy of the applied gain rows + device copy of their set-point block
    mutable std::vector<double> fast_host;
    mutable ddmpc::DevBuf fast_ksp;
    // gemm_loop.cu: block maps of the plant and their host key (the per-call loop state is allocated per call,
    // stream-ordered: two streams may run closed loops of one set at the same time)
    mutable ddmpc::DevBuf gemm_ws;
    mutable std::vector<double> gemm_host;
    // tc_loop.cu: TF32 hi / lo images of the gain block and the plant block map (+ the plant in FP64) and their host key
    mutable ddmpc::DevBuf tc_ws;
    mutable std::vector<double> tc_key;
    // cvx_loop.cu: Phi zero-padded to 64 x 64 (A operand of the ADMM iteration when it is streamed from L1), built with the set
    ddmpc::DevBuf cvx_phi64;
    // dmma_loop.cu: packed A fragments (gain rows + block maps of the plant) and their host key
    mutable ddmpc::DevBuf dmma_ws;
    mutable std::vector<double> dmma_key;
    // solve.cu: staging of the B = 1 host path (pinned host, device, private stream)
    mutable void *stage_host = nullptr;
    mutable ddmpc::DevBuf stage_dev;
    mutable cudaStream_t stage_stream = nullptr;
    void detach_streams() {
        plan.detach_streams();
        for (ddmpc::DevBuf *b : {&ctrl_status, &plant_dev, &fast_ksp, &gemm_ws, &dmma_ws, &tc_ws, &cvx_phi64, &stage_dev}) b->detach_stream();
    }
    ~ddmpc_set() {
        if (stage_host) cudaFreeHost(stage_host);
        if (stage_stream) cudaStreamDestroy(stage_stream);
    }
};
