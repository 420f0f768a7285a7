"""FP64 NumPy/SciPy oracle of the DD-MPC hot path.  TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``: pinned against the unmodified reference class run with a cvxpy stand-in;
parity UNPINNED only at the level of cvxpy's own solver tolerance).

Every function cites the reference lines it restates (paths relative to the
reference checkout).  The QP is solved literally as stated by the reference:
variables (alpha, ubar, ybar[, sigma]), the Hankel dynamics equality, the
initial/terminal equalities and, for CONVEX, the inf-norm slack bound - through
one dense symmetric-indefinite KKT solve (plus an active-set loop on the slack
bound).  Nothing here shares code or a formulation with the CUDA product, which
works on a condensed row-space problem: agreement between the two is a genuine
cross-check.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.linalg as sla

# Integer codes follow the reference's YAML maps
# (utilities/controller/controller_creation.py:12-23).
NOMINAL, ROBUST = 0, 1
SLACK_NONE, SLACK_CONVEX, SLACK_NON_CONVEX = 0, 1, 2


# --------------------------------------------------------------------------
# Hankel matrix / persistency of excitation
# --------------------------------------------------------------------------
def hankel_matrix(X: np.ndarray, L: int) -> np.ndarray:
    """direct_data_driven_mpc/utilities/hankel_matrix.py:39-53.

    Column i is the flattened window X[i:i+L, :]; output (L*n_ch, N-L+1) FP64.
    """
    N, n_ch = X.shape
    if N < L:
        raise ValueError("N must be greater than or equal to L.")
    cols = N - L + 1
    flat = np.ascontiguousarray(X, dtype=np.float64).reshape(-1)
    # H[r, c] = flat[r + c * n_ch]  (an overlapping strided view, copied out)
    idx = np.arange(L * n_ch)[:, None] + n_ch * np.arange(cols)[None, :]
    return flat[idx]


def evaluate_persistent_excitation(X: np.ndarray, order: int) -> Tuple[int, bool]:
    """hankel_matrix.py:77-87: rank(H_order(X)) == n_ch * order (SVD rank)."""
    n_ch = X.shape[1]
    rank = int(np.linalg.matrix_rank(hankel_matrix(X, order)))
    return rank, rank == n_ch * order


# --------------------------------------------------------------------------
# LTI plant + observer helpers
# --------------------------------------------------------------------------
def observability_matrix(A, C):
    """utilities/initial_state_estimation.py:3-24."""
    n = A.shape[0]
    return np.vstack([C @ np.linalg.matrix_power(A, i) for i in range(n)])


def toeplitz_input_output_matrix(A, B, C, D, t):
    """utilities/initial_state_estimation.py:26-93."""
    if t <= 0:
        raise ValueError("The number of time steps t must be positive.")
    m, p = B.shape[1], C.shape[0]
    Tt = np.zeros((p * t, m * t))
    for i in range(t):
        Tt[i * p:(i + 1) * p, i * m:(i + 1) * m] = D
        for j in range(i):
            Tt[i * p:(i + 1) * p, j * m:(j + 1) * m] = (
                C @ np.linalg.matrix_power(A, i - j - 1) @ B)
    return Tt


class Plant:
    """utilities/model_simulation.py:31-98 (LTIModel): y uses the pre-update x."""

    def __init__(self, A, B, C, D, eps_max=0.0):
        self.A, self.B, self.C, self.D = (np.asarray(M, dtype=float) for M in (A, B, C, D))
        self.eps_max = eps_max
        self.n, self.m, self.p = self.A.shape[0], self.B.shape[1], self.C.shape[0]
        self.x = np.zeros(self.n)
        self.Ot = observability_matrix(self.A, self.C)
        self.Tt = toeplitz_input_output_matrix(self.A, self.B, self.C, self.D, self.n)

    def simulate_step(self, u, w):
        y = self.C @ self.x + self.D @ u + w            # model_simulation.py:94
        self.x = self.A @ self.x + self.B @ u           # model_simulation.py:96
        return y

    def simulate(self, U, W, steps):
        Y = np.zeros((steps, self.p))
        for k in range(steps):                          # model_simulation.py:125-131
            Y[k, :] = self.simulate_step(U[k, :], W[k, :])
        return Y

    def initial_state_from_trajectory(self, U, Y):
        """initial_state_estimation.py:131: pinv(Ot) (Y - Tt U)."""
        return np.linalg.pinv(self.Ot) @ (Y - self.Tt @ U)

    def gain(self):
        return self.C @ np.linalg.inv(np.eye(self.n) - self.A) @ self.B + self.D

    def equilibrium_output_from_input(self, u_eq):
        """initial_state_estimation.py:162-169."""
        return self.gain() @ u_eq

    def equilibrium_input_from_output(self, y_eq):
        """initial_state_estimation.py:198-205."""
        return np.linalg.pinv(self.gain()) @ y_eq

    # reference-compatible accessors so the reference's own loop driver can use it
    def get_system_order(self): return self.n
    def get_number_inputs(self): return self.m
    def get_number_outputs(self): return self.p
    def get_eps_max(self): return self.eps_max
    def get_state(self): return self.x
    def set_state(self, state): self.x = state


FOUR_TANK = dict(
    A=[[0.921, 0, 0.041, 0], [0, 0.918, 0, 0.033], [0, 0, 0.924, 0], [0, 0, 0, 0.937]],
    B=[[0.017, 0.001], [0.001, 0.023], [0, 0.061], [0.072, 0]],
    C=[[1, 0, 0, 0], [0, 1, 0, 0]],
    D=[[0, 0], [0, 0]],
    eps_max=0.002,
)  # examples/config/models/four_tank_system_params.yaml:9-26


def four_tank_plant() -> Plant:
    return Plant(**FOUR_TANK)


def four_tank_params(m=2, p=2) -> Dict:
    """Parameter derivation of controller_creation.py:110-168 applied to
    examples/config/controllers/data_driven_mpc_example_params.yaml:9-22."""
    L, eps = 30, 0.002
    return dict(
        u_range=[-1, 1], N=400, n=4, eps_max=eps, L=L,
        Q=3 * np.eye(p * L), R=0.0001 * np.eye(m * L),
        lamb_alpha=0.1 / eps, lamb_sigma=1000, c=1.0,
        slack_type=SLACK_NONE, ctrl_type=ROBUST, n_mpc_step=4,
        u_s=np.array([[1.0], [1.0]]), y_s=np.array([[0.65], [0.77]]),
    )


# --------------------------------------------------------------------------
# Scenario generators (RNG draw order = SURVEY Appendix B)
# --------------------------------------------------------------------------
def randomize_initial_system_state(plant: Plant, u_range, rng) -> np.ndarray:
    """utilities/controller/controller_operation.py:59-75."""
    ns, m, p = plant.n, plant.m, plant.p
    plant.x = rng.uniform(-1.0, 1.0, size=ns)
    u_i = rng.uniform(*u_range, (ns, m))
    w_i = plant.eps_max * rng.uniform(-1.0, 1.0, (ns, p))
    y_i = plant.simulate(u_i, w_i, ns)
    return plant.initial_state_from_trajectory(u_i.flatten(), y_i.flatten())


def generate_initial_input_output_data(plant: Plant, N, u_range, rng):
    """controller_operation.py:126-133."""
    u_d = rng.uniform(*u_range, (N, plant.m))
    w_d = plant.eps_max * rng.uniform(-1.0, 1.0, (N, plant.p))
    y_d = plant.simulate(u_d, w_d, N)
    return u_d, y_d


def simulate_n_input_output_measurements(plant: Plant, n, u_s, rng):
    """controller_operation.py:190-197."""
    U_n = np.tile(u_s, (n, 1)).reshape(n, plant.m)
    W_n = plant.eps_max * rng.uniform(-1.0, 1.0, (n, plant.p))
    Y_n = plant.simulate(U_n, W_n, n)
    return U_n, Y_n


def equilibrium_state_from_output(plant: Plant, y_eq):
    """utilities/reproduction/paper_reproduction.py:104-114."""
    u_eq = plant.equilibrium_input_from_output(y_eq)
    return plant.initial_state_from_trajectory(np.tile(u_eq, plant.n), np.tile(y_eq, plant.n))


# --------------------------------------------------------------------------
# The QP - literal KKT
# --------------------------------------------------------------------------
@dataclass
class QPSolution:
    status: str
    optimal_u: Optional[np.ndarray] = None     # ubar[n*m:] flattened, (L*m,)
    cost: float = float("nan")
    alpha: Optional[np.ndarray] = None
    ubar: Optional[np.ndarray] = None
    ybar: Optional[np.ndarray] = None
    sigma: Optional[np.ndarray] = None
    n_active: int = 0
    kkt_residual: float = float("nan")


class OracleQP:
    """Literal restatement of the reference QP.

    Variables  z = [alpha (c); ubar (Lp*m); ybar (Lp*p); sigma (Lp*p, robust)]
    (direct_data_driven_mpc_controller.py:433-445), cost :703-722, constraints
    :533-545 (dynamics), :577-581 (initial n blocks), :612-627 (terminal n
    blocks), :659-675 (CONVEX slack bound on sigma[n*p:]).
    """

    def __init__(self, n, m, p, u_d, y_d, L, Q, R, eps_max=None, lamb_alpha=None,
                 lamb_sigma=None, c=None, slack_type=SLACK_CONVEX, ctrl_type=NOMINAL,
                 use_terminal=True, cache_factor=True, input_bounds=None, output_bounds=None):
        """input_bounds = (u_min, u_max) / output_bounds = (y_min, y_max): optional boxes on every predicted input
        ubar[n*m:] / output ybar[n*p:] (paper Eq. 6, u in U, y in Y).  NOT in the reference (its constraint list,
        controller.py:447-504, has no such entry): extensions that only the oracle pins, off by default."""
        if slack_type == SLACK_NON_CONVEX and ctrl_type == ROBUST:
            raise NotImplementedError("NON_CONVEX slack constraint (controller.py:664-670)")
        self.n, self.m, self.p, self.L = n, m, p, L
        self.N = u_d.shape[0]
        self.Lp = Lp = L + n
        self.robust = ctrl_type == ROBUST
        self.convex = self.robust and slack_type == SLACK_CONVEX
        self.use_terminal = use_terminal
        self.Q = 0.5 * (np.asarray(Q, float) + np.asarray(Q, float).T)
        self.R = 0.5 * (np.asarray(R, float) + np.asarray(R, float).T)
        self.eps_max, self.lamb_alpha, self.lamb_sigma, self.c = eps_max, lamb_alpha, lamb_sigma, c
        self.Hu = hankel_matrix(u_d, Lp)                 # controller.py:376
        self.Hy = hankel_matrix(y_d, Lp)                 # controller.py:377
        self.H = np.vstack([self.Hu, self.Hy])
        self.r, self.cols = self.H.shape
        c_ = self.cols
        nu, ny = Lp * m, Lp * p
        self.o_a, self.o_u, self.o_y, self.o_s = 0, c_, c_ + nu, c_ + nu + ny
        self.nz = c_ + nu + ny + (ny if self.robust else 0)
        self.bound = (c * eps_max) if self.convex else None
        # inequality rows lo <= z[idx] <= hi: sigma[n*p:] (:659-675) and, if given, the free predicted inputs
        idx, lo, hi = [], [], []
        if self.convex:
            for j in range(L * p):
                idx.append(self.o_s + n * p + j); lo.append(-self.bound); hi.append(self.bound)
        self.u_box = None
        if input_bounds is not None:
            if not self.robust:
                raise NotImplementedError("input box: ROBUST controllers only")
            bl, bh = input_bounds
            bl = np.full(m, -np.inf) if bl is None else np.broadcast_to(np.asarray(bl, float).reshape(-1), (m,))
            bh = np.full(m, np.inf) if bh is None else np.broadcast_to(np.asarray(bh, float).reshape(-1), (m,))
            if np.any(bl > bh):
                raise ValueError("input box: u_min must not exceed u_max")
            self.u_box = (bl, bh)
            for i in range((L - n) * m if use_terminal else L * m):     # terminal blocks are equalities
                idx.append(self.o_u + n * m + i); lo.append(bl[i % m]); hi.append(bh[i % m])
        self.y_box = None
        if output_bounds is not None:
            if not self.robust:
                raise NotImplementedError("output box: ROBUST controllers only")
            bl, bh = output_bounds
            bl = np.full(p, -np.inf) if bl is None else np.broadcast_to(np.asarray(bl, float).reshape(-1), (p,))
            bh = np.full(p, np.inf) if bh is None else np.broadcast_to(np.asarray(bh, float).reshape(-1), (p,))
            if np.any(bl > bh):
                raise ValueError("output box: y_min must not exceed y_max")
            self.y_box = (bl, bh)
            for i in range((L - n) * p if use_terminal else L * p):     # terminal blocks are equalities
                idx.append(self.o_y + n * p + i); lo.append(bl[i % p]); hi.append(bh[i % p])
        self.ineq_idx, self.ineq_lo, self.ineq_hi = np.array(idx, int), np.array(lo, float), np.array(hi, float)
        self._cache_factor = cache_factor
        self._lu = None
        if self.robust:
            self._assemble_kkt()
        else:
            self._prepare_nominal()

    # ---- robust: full KKT ------------------------------------------------
    def _cost_matrices(self):
        n, m, p, L, Lp = self.n, self.m, self.p, self.L, self.Lp
        P = np.zeros((self.nz, self.nz))
        ou, oy = self.o_u + n * m, self.o_y + n * p
        P[ou:ou + L * m, ou:ou + L * m] = self.R          # :708-709
        P[oy:oy + L * p, oy:oy + L * p] = self.Q          # :710
        if self.robust:
            w_alpha = self.lamb_alpha * self.eps_max      # :714
            P[self.o_a:self.o_a + self.cols, self.o_a:self.o_a + self.cols] = w_alpha * np.eye(self.cols)
            P[self.o_s:, self.o_s:] = self.lamb_sigma * np.eye(Lp * p)   # :715 (all Lp blocks)
        return P

    def _constraint_matrix(self):
        n, m, p, L, Lp = self.n, self.m, self.p, self.L, self.Lp
        nu, ny = Lp * m, Lp * p
        rows = [np.zeros((self.r, self.nz))]
        dyn = rows[0]
        dyn[:, self.o_a:self.o_a + self.cols] = -self.H   # [ubar; ybar+sigma] = H alpha
        dyn[:nu, self.o_u:self.o_u + nu] = np.eye(nu)
        dyn[nu:, self.o_y:self.o_y + ny] = np.eye(ny)
        if self.robust:
            dyn[nu:, self.o_s:self.o_s + ny] = np.eye(ny)
        init = np.zeros((n * (m + p), self.nz))
        init[:n * m, self.o_u:self.o_u + n * m] = np.eye(n * m)
        init[n * m:, self.o_y:self.o_y + n * p] = np.eye(n * p)
        rows.append(init)
        if self.use_terminal:
            term = np.zeros((n * (m + p), self.nz))
            term[:n * m, self.o_u + L * m:self.o_u + Lp * m] = np.eye(n * m)
            term[n * m:, self.o_y + L * p:self.o_y + Lp * p] = np.eye(n * p)
            rows.append(term)
        return np.vstack(rows)

    def _assemble_kkt(self):
        self.P = self._cost_matrices()
        self.Aeq = self._constraint_matrix()
        ne = self.Aeq.shape[0]
        K = np.zeros((self.nz + ne, self.nz + ne))
        K[:self.nz, :self.nz] = 2.0 * self.P
        K[:self.nz, self.nz:] = self.Aeq.T
        K[self.nz:, :self.nz] = self.Aeq
        self.K = K
        self._pinv = None
        if self.robust and self.lamb_alpha * self.eps_max == 0.0:
            # eps_max = 0 under a ROBUST controller (controller_creation.py:129-136 anticipates it): alpha carries no
            # weight, the KKT matrix is singular in the alpha directions of null(H) while (ubar, ybar, sigma) stay
            # unique - any KKT point will do, the minimum-norm one is taken
            self._pinv = np.linalg.pinv(K, rcond=1e-13)
        elif self._cache_factor:
            self._lu = sla.lu_factor(K)

    def _rhs(self, u_past, y_past, u_s, y_s):
        n, m, p, L = self.n, self.m, self.p, self.L
        us_L = np.tile(np.reshape(u_s, (-1,)), L)
        ys_L = np.tile(np.reshape(y_s, (-1,)), L)
        q = np.zeros(self.nz)
        ou, oy = self.o_u + n * m, self.o_y + n * p
        q[ou:ou + L * m] = -2.0 * self.R @ us_L
        q[oy:oy + L * p] = -2.0 * self.Q @ ys_L
        const = us_L @ self.R @ us_L + ys_L @ self.Q @ ys_L
        b = [np.zeros(self.r), np.reshape(u_past, (-1,)), np.reshape(y_past, (-1,))]
        if self.use_terminal:
            b += [np.tile(np.reshape(u_s, (-1,)), n), np.tile(np.reshape(y_s, (-1,)), n)]
        return q, np.concatenate(b), const

    def _kkt_solve(self, rhs):
        if self._pinv is not None:
            return self._pinv @ rhs
        if self._lu is not None:
            return sla.lu_solve(self._lu, rhs)
        return np.linalg.solve(self.K, rhs)

    def _solve_robust(self, u_past, y_past, u_s, y_s) -> QPSolution:
        q, b, const = self._rhs(u_past, y_past, u_s, y_s)
        rhs = np.concatenate([-q, b])
        zt0 = self._kkt_solve(rhs)
        n_act = 0
        zt = zt0
        if self.u_box is not None and self.use_terminal:
            us = np.reshape(u_s, (-1,))
            if np.any(us < self.u_box[0] - 1e-9 * (1 + np.abs(us))) or np.any(us > self.u_box[1] + 1e-9 * (1 + np.abs(us))):
                return QPSolution(status="infeasible")             # terminal equality ubar = u_s violates the box
        if self.y_box is not None and self.use_terminal:
            ysv = np.reshape(y_s, (-1,))
            if np.any(ysv < self.y_box[0] - 1e-9 * (1 + np.abs(ysv))) or np.any(ysv > self.y_box[1] + 1e-9 * (1 + np.abs(ysv))):
                return QPSolution(status="infeasible")             # terminal equality ybar = y_s violates the box
        if self.ineq_idx.size:
            zi, lo, hi = self.ineq_idx, self.ineq_lo, self.ineq_hi
            nb = zi.size
            fin = np.concatenate([np.abs(lo[np.isfinite(lo)]), np.abs(hi[np.isfinite(hi)]), [0.0]])
            vtol = 1e-13 * np.maximum(1.0, np.maximum(np.where(np.isfinite(lo), np.abs(lo), 0), np.where(np.isfinite(hi), np.abs(hi), 0)))
            active: Dict[int, int] = {}                            # row -> +1 (upper bound) / -1 (lower bound)
            vcache: Dict[int, np.ndarray] = {}
            for _ in range(20 * nb + 20):
                if active:
                    idx = sorted(active)
                    for j in idx:
                        if j not in vcache:
                            e = np.zeros(self.K.shape[0]); e[zi[j]] = 1.0
                            vcache[j] = self._kkt_solve(e)
                    V = np.stack([vcache[j] for j in idx], axis=1)
                    d = np.array([hi[j] if active[j] > 0 else lo[j] for j in idx])
                    S = V[zi[idx], :]
                    mu = np.linalg.solve(S, zt0[zi[idx]] - d)
                    zt = zt0 - V @ mu
                else:
                    idx, mu, zt = [], np.zeros(0), zt0
                vals = zt[zi]
                viol = np.maximum(vals - hi, lo - vals) - vtol
                for j in idx:
                    viol[j] = -np.inf
                jmax = int(np.argmax(viol))
                if viol[jmax] > 0.0:
                    active[jmax] = 1 if vals[jmax] > hi[jmax] else -1
                    continue
                bad = [(active[j] * mu[k], j) for k, j in enumerate(idx) if active[j] * mu[k] < -1e-12]
                if bad:
                    del active[min(bad)[1]]
                    continue
                break
            else:
                return QPSolution(status="solver_error")
            n_act = len(active)
        z = zt[:self.nz]
        cost = float(z @ self.P @ z + q @ z + const)
        sol = self._unpack(z)
        sol.cost, sol.n_active, sol.status = cost, n_act, "optimal"
        res_eq = np.linalg.norm(self.Aeq @ z - b, np.inf)
        sol.kkt_residual = float(res_eq)
        return sol

    def _unpack(self, z) -> QPSolution:
        n, m, p, L, Lp = self.n, self.m, self.p, self.L, self.Lp
        ubar = z[self.o_u:self.o_u + Lp * m]
        ybar = z[self.o_y:self.o_y + Lp * p]
        sigma = z[self.o_s:self.o_s + Lp * p] if self.robust else None
        return QPSolution(status="optimal", optimal_u=ubar[n * m:].copy(),   # :799-806
                          alpha=z[:self.cols].copy(), ubar=ubar.copy(), ybar=ybar.copy(),
                          sigma=None if sigma is None else sigma.copy())

    # ---- nominal: rank-revealing elimination of alpha ---------------------
    def _prepare_nominal(self):
        U, s, _ = np.linalg.svd(self.H, full_matrices=False)
        tol = s.max() * max(self.H.shape) * np.finfo(float).eps     # numpy matrix_rank default
        self.rank_H = int(np.sum(s > tol))
        self.Uk = U[:, :self.rank_H]
        n, m, p, L, Lp = self.n, self.m, self.p, self.L, self.Lp
        nu = Lp * m
        fixed = list(range(0, n * m)) + list(range(nu, nu + n * p))
        if self.use_terminal:
            fixed += list(range(L * m, Lp * m)) + list(range(nu + L * p, nu + Lp * p))
        self.fixed_rows = np.array(fixed)
        self.Cfix = self.Uk[self.fixed_rows, :]
        Uc, sc, Vct = np.linalg.svd(self.Cfix, full_matrices=True)
        tolc = sc.max() * 1e-9
        rc = int(np.sum(sc > tolc))
        self.C_pinv = (Vct[:rc].T / sc[:rc]) @ Uc[:, :rc].T
        self.Znull = Vct[rc:].T
        # weighted cost rows:  D^(1/2) S t
        D = np.zeros((self.r, self.r))
        D[n * m:nu, n * m:nu] = self.R
        D[nu + n * p:, nu + n * p:] = self.Q
        self.Dfull = D

    def _solve_nominal(self, u_past, y_past, u_s, y_s) -> QPSolution:
        n, m, p, L, Lp = self.n, self.m, self.p, self.L, self.Lp
        nu = Lp * m
        b = [np.reshape(u_past, (-1,)), np.reshape(y_past, (-1,))]
        if self.use_terminal:
            b += [np.tile(np.reshape(u_s, (-1,)), n), np.tile(np.reshape(y_s, (-1,)), n)]
        b = np.concatenate(b)
        g_p = self.C_pinv @ b
        feas = np.linalg.norm(self.Cfix @ g_p - b, np.inf)
        if feas > 1e-6 * (1.0 + np.linalg.norm(b, np.inf)):
            return QPSolution(status="infeasible", kkt_residual=float(feas))
        d = np.zeros(self.r)
        d[n * m:nu] = np.tile(np.reshape(u_s, (-1,)), L)
        d[nu + n * p:] = np.tile(np.reshape(y_s, (-1,)), L)
        # minimise (t-d)^T D (t-d), t = Uk (g_p + Z zeta)
        UZ = self.Uk @ self.Znull
        Hs = UZ.T @ self.Dfull @ UZ
        gs = UZ.T @ self.Dfull @ (d - self.Uk @ g_p)
        zeta = np.linalg.lstsq(Hs, gs, rcond=1e-11)[0] if Hs.size else np.zeros(0)
        t = self.Uk @ (g_p + self.Znull @ zeta)
        ubar, ybar = t[:nu], t[nu:]
        cost = float((t - d) @ self.Dfull @ (t - d))
        alpha = np.linalg.pinv(self.H) @ t
        return QPSolution(status="optimal", optimal_u=ubar[n * m:].copy(), cost=cost, alpha=alpha,
                          ubar=ubar.copy(), ybar=ybar.copy(), kkt_residual=float(feas))

    def solve(self, u_past, y_past, u_s, y_s) -> QPSolution:
        if self.robust:
            return self._solve_robust(u_past, y_past, u_s, y_s)
        return self._solve_nominal(u_past, y_past, u_s, y_s)


# --------------------------------------------------------------------------
# Controller façade with the reference's method names, so the reference's own
# (unmodified) loop driver can drive the oracle when fixtures are generated.
# --------------------------------------------------------------------------
class OracleController:
    def __init__(self, n, m, p, u_d, y_d, L, Q, R, u_s, y_s, eps_max=None, lamb_alpha=None,
                 lamb_sigma=None, c=None, slack_type=SLACK_CONVEX, ctrl_type=NOMINAL,
                 n_mpc_step=1, use_terminal=True, cache_factor=True, check_pe=True, input_bounds=None,
                 output_bounds=None):
        self.n, self.m, self.p, self.L = n, m, p, L
        self.u_s, self.y_s, self.n_mpc_step = u_s, y_s, n_mpc_step
        N = u_d.shape[0]
        if check_pe:                                      # controller.py:242-296
            N_min = m * (L + 2 * n) + L + 2 * n - 1
            if N < N_min:
                raise ValueError("N < N_min")
            rank, ok = evaluate_persistent_excitation(u_d, L + 2 * n)
            if not ok:
                raise ValueError("not persistently exciting")
        self.qp = OracleQP(n, m, p, u_d, y_d, L, Q, R, eps_max, lamb_alpha, lamb_sigma, c,
                           slack_type, ctrl_type, use_terminal, cache_factor, input_bounds, output_bounds)
        self.u_past = u_d[-n:, :].reshape(-1, 1)          # controller.py:184
        self.y_past = y_d[-n:, :].reshape(-1, 1)          # controller.py:185
        self.solution: Optional[QPSolution] = None
        self.optimal_u = None
        self.history = []
        self.update_and_solve_data_driven_mpc()           # controller.py:385-387 (solve #0)
        self.history.clear()

    def update_and_solve_data_driven_mpc(self):           # controller.py:389-407
        sol = self.qp.solve(self.u_past, self.y_past, self.u_s, self.y_s)
        if sol.status not in ("optimal", "optimal_inaccurate"):
            raise ValueError("MPC problem was not solved optimally.")
        self.solution, self.optimal_u = sol, sol.optimal_u
        self.history.append((self.u_past.copy(), self.y_past.copy(), sol.optimal_u.copy(), sol.cost))

    def get_optimal_control_input_at_step(self, n_step=0):   # controller.py:810-842
        if not 0 <= n_step < self.L:
            raise ValueError("n_step out of range")
        return self.optimal_u[n_step * self.m:(n_step + 1) * self.m]

    def store_input_output_measurement(self, u_current, y_current):   # controller.py:844-895
        self.u_past = np.vstack([self.u_past[self.m:], u_current])
        self.y_past = np.vstack([self.y_past[self.p:], y_current])

    def set_past_input_output_data(self, u_past, y_past):   # controller.py:897-943
        self.u_past, self.y_past = u_past, y_past

    def get_optimal_cost_value(self):
        return self.solution.cost


def closed_loop(plant: Plant, ctrl, n_steps: int, w_sys: np.ndarray):
    """controller_operation.py:269-305 with the noise (already scaled by
    eps_max, :263) supplied by the caller."""
    m, p = plant.m, plant.p
    u_sys = np.zeros((n_steps, m))
    y_sys = np.zeros((n_steps, p))
    for t in range(0, n_steps, ctrl.n_mpc_step):
        ctrl.update_and_solve_data_driven_mpc()
        for k in range(t, min(t + ctrl.n_mpc_step, n_steps)):
            u_sys[k, :] = ctrl.get_optimal_control_input_at_step(n_step=k - t)
            y_sys[k, :] = plant.simulate_step(u_sys[k, :], w_sys[k, :])
            ctrl.store_input_output_measurement(u_sys[k, :].reshape(-1, 1), y_sys[k, :].reshape(-1, 1))
    return u_sys, y_sys


def make_controller(params: Dict, u_d, y_d, **over) -> OracleController:
    """controller_creation.py:255-273 with the params dict of four_tank_params()."""
    kw = dict(params)
    kw.update(over)
    m, p = u_d.shape[1], y_d.shape[1]
    return OracleController(
        n=kw["n"], m=m, p=p, u_d=u_d, y_d=y_d, L=kw["L"], Q=kw["Q"], R=kw["R"], u_s=kw["u_s"],
        y_s=kw["y_s"], eps_max=kw["eps_max"], lamb_alpha=kw["lamb_alpha"], lamb_sigma=kw["lamb_sigma"],
        c=kw["c"], slack_type=kw["slack_type"], ctrl_type=kw["ctrl_type"], n_mpc_step=kw["n_mpc_step"],
        use_terminal=kw.get("use_terminal", True), cache_factor=kw.get("cache_factor", True),
        input_bounds=kw.get("input_bounds"), output_bounds=kw.get("output_bounds"))


def example_scenario(seed: int, params: Optional[Dict] = None, plant: Optional[Plant] = None):
    """Stages 1-3 of examples/direct_data_driven_mpc_example.py:263-300 (RNG draws 1-5)."""
    params = params or four_tank_params()
    plant = plant or four_tank_plant()
    rng = np.random.default_rng(seed)
    x0 = randomize_initial_system_state(plant, params["u_range"], rng)
    plant.set_state(x0)
    u_d, y_d = generate_initial_input_output_data(plant, params["N"], params["u_range"], rng)
    return plant, params, rng, x0, u_d, y_d


def run_example(seed=0, t_sim=400, **over):
    """The whole BASELINE config 1 run: returns (u_sys, y_sys, controller, extras)."""
    plant, params, rng, x0, u_d, y_d = example_scenario(seed)
    ctrl = make_controller(params, u_d, y_d, **over)
    n_steps = t_sim + 1
    x_loop0 = plant.x.copy()
    w_sys = plant.eps_max * rng.uniform(-1.0, 1.0, (n_steps, plant.p))   # controller_operation.py:263
    u_sys, y_sys = closed_loop(plant, ctrl, n_steps, w_sys)
    return u_sys, y_sys, ctrl, dict(x0=x0, u_d=u_d, y_d=y_d, w_sys=w_sys, x_loop0=x_loop0, params=params)


def run_reproduction(seed=4, t_sim=600):
    """examples/robust_data_driven_mpc_reproduction.py:126-295 (TEC, TEC-n-step, UCON)."""
    plant, params, rng, x0, u_d, y_d = example_scenario(seed)
    n = params["n"]
    schemes = [("TEC", 1, True), ("TEC_N_STEP", n, True), ("UCON", 1, False)]
    ctrls = [make_controller(params, u_d, y_d, n_mpc_step=s[1], use_terminal=s[2]) for s in schemes]
    plant.set_state(equilibrium_state_from_output(plant, np.array([0.4, 0.4])))
    U_n, Y_n = simulate_n_input_output_measurements(plant, n, params["u_s"], rng)
    for c_ in ctrls:
        c_.set_past_input_output_data(U_n.reshape(-1, 1), Y_n.reshape(-1, 1))
    x_start = plant.x.copy()
    n_steps = t_sim + 1 - n
    out = {}
    for (name, _, _), c_ in zip(schemes, ctrls):
        plant.set_state(x_start.copy())
        w_sys = plant.eps_max * rng.uniform(-1.0, 1.0, (n_steps, plant.p))
        u_sys, y_sys = closed_loop(plant, c_, n_steps, w_sys)
        out[name] = dict(u_sys=u_sys, y_sys=y_sys, w_sys=w_sys)
    return out, dict(u_d=u_d, y_d=y_d, U_n=U_n, Y_n=Y_n, x_start=x_start, params=params)


# --------------------------------------------------------------------------
# Philox4x32-10 counter RNG (throughput-mode measurement noise).  This is the
# product's own noise-stream definition (the reference only has NumPy PCG64
# draws, which the parity mode uploads); the oracle restates it so that
# Philox-mode closed loops can be replayed on the CPU.
# --------------------------------------------------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (...,4) uint32, key (...,2) uint32 -> (...,4) uint32, 10 rounds."""
    c = [ctr[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32)
    k1 = key[..., 1].astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c[0].astype(np.uint64)
            p1 = _PH_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PH_W0).astype(np.uint32)
            k1 = (k1 + _PH_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_noise(seed: int, scenario_ids: np.ndarray, n_steps: int, p: int, eps_max: float) -> np.ndarray:
    """w[b, k, j] = eps_max * (2*v - 3) where, with q = k*p + j, v in [1, 2) is the double whose
    top 32 mantissa bits are word (q & 3) of Philox4x32-10(counter = (q >> 2, 0, id_lo, id_hi),
    key = (seed_lo, seed_hi)) and whose low 20 mantissa bits are zero."""
    ids = np.asarray(scenario_ids, dtype=np.uint64)
    B = ids.shape[0]
    nq = n_steps * p
    ncall = (nq + 3) // 4
    ctr = np.zeros((B, ncall, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(ncall, dtype=np.uint32)[None, :]
    ctr[..., 2] = (ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[..., 3] = (ids >> np.uint64(32)).astype(np.uint32)[:, None]
    key = np.zeros((B, ncall, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32(ctr, key).astype(np.uint64).reshape(B, ncall * 4)[:, :nq]
    v = ((x << np.uint64(20)) | np.uint64(0x3FF0000000000000)).view(np.float64)
    return (eps_max * (2.0 * v - 3.0)).reshape(B, n_steps, p)
