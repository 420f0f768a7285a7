"""Drop-in ``DirectDataDrivenMPCController`` backed by libddmpc (B = 1 case of
the batched CUDA solver).

Mirrors the public surface of the reference class
(``direct_data_driven_mpc/direct_data_driven_mpc_controller.py:22-981``): same
constructor signature, method names, attributes, exception types and messages,
so the reference's own ``utilities/controller/*`` and example scripts can drive
it unchanged.  What differs is what sits behind ``solve_mpc_problem``: no cvxpy
objects are built; the constructor factorises the problem once on the GPU and
every solve is one kernel launch.  The ``cp.Variable`` attributes of the
reference (``alpha, ubar, ybar, sigma``) are exposed as small value holders
whose ``.value`` is filled lazily from the device (``None`` before a solve).
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import Optional

import numpy as np

from . import _lib
from .hankel import evaluate_persistent_excitation


class DataDrivenMPCType(Enum):
    # The reference declares these with trailing commas (controller.py:11-13), which
    # makes NOMINAL tuple-valued.  Kept so `.value` comparisons behave identically.
    NOMINAL = 0,
    ROBUST = 1


class SlackVarConstraintTypes(Enum):
    # controller.py:15-20 (same trailing-comma quirk)
    NON_CONVEX = 0,
    CONVEX = 1,
    NONE = 2


_CTRL_CODE = {DataDrivenMPCType.NOMINAL: _lib.NOMINAL, DataDrivenMPCType.ROBUST: _lib.ROBUST}
_SLACK_CODE = {SlackVarConstraintTypes.NONE: _lib.SLACK_NONE, SlackVarConstraintTypes.CONVEX: _lib.SLACK_CONVEX,
               SlackVarConstraintTypes.NON_CONVEX: _lib.SLACK_NON_CONVEX}


class _Var:
    """Stand-in for a solved ``cp.Variable``: ``.value`` is a column vector."""

    def __init__(self, owner: "DirectDataDrivenMPCController", name: str, rows: int):
        self._owner, self._name, self.shape = owner, name, (rows, 1)

    @property
    def value(self) -> Optional[np.ndarray]:
        return self._owner._primal(self._name)


class _Problem:
    """Stand-in for ``cp.Problem``: carries ``status`` and ``value``."""

    def __init__(self):
        self.status: Optional[str] = None
        self.value: Optional[float] = None


def _nan_if_none(x) -> float:
    return float("nan") if x is None else float(x)


class DirectDataDrivenMPCController:
    """Nominal / robust direct data-driven MPC controller (Berberich et al. 2021).

    Same arguments as the reference constructor (controller.py:95-116).
    """

    def __init__(self, n: int, m: int, p: int, u_d: np.ndarray, y_d: np.ndarray, L: int, Q: np.ndarray,
                 R: np.ndarray, u_s: np.ndarray, y_s: np.ndarray, eps_max: Optional[float] = None,
                 lamb_alpha: Optional[float] = None, lamb_sigma: Optional[float] = None,
                 c: Optional[float] = None,
                 slack_var_constraint_type: SlackVarConstraintTypes = SlackVarConstraintTypes.CONVEX,
                 controller_type: DataDrivenMPCType = DataDrivenMPCType.NOMINAL, n_mpc_step: int = 1,
                 use_terminal_constraint: bool = True, input_bounds=None, output_bounds=None):
        # input_bounds: optional (u_min, u_max) box on every predicted input (paper Eq. 6).  An extension: the
        # reference class has no such argument (controller.py:95-116) and never constrains ubar (:447-504).
        self._set = None
        self.input_bounds = None
        if input_bounds is not None:
            lo, hi = input_bounds
            lo = np.full(m, -np.inf) if lo is None else np.broadcast_to(np.asarray(lo, dtype=np.float64).reshape(-1), (m,)).copy()
            hi = np.full(m, np.inf) if hi is None else np.broadcast_to(np.asarray(hi, dtype=np.float64).reshape(-1), (m,)).copy()
            self.input_bounds = (lo, hi)
        self.output_bounds = None
        if output_bounds is not None:                                 # same extension for the predicted outputs ybar
            lo, hi = output_bounds
            lo = np.full(p, -np.inf) if lo is None else np.broadcast_to(np.asarray(lo, dtype=np.float64).reshape(-1), (p,)).copy()
            hi = np.full(p, np.inf) if hi is None else np.broadcast_to(np.asarray(hi, dtype=np.float64).reshape(-1), (p,)).copy()
            self.output_bounds = (lo, hi)
        self._solve_tol, self._solve_max_iter = 1e-8, 2000
        self.controller_type = controller_type
        if controller_type not in _CTRL_CODE:                         # controller.py:165-168
            raise ValueError("Unsupported controller type.")
        self.n, self.m, self.p = n, m, p
        self.u_d, self.y_d = u_d, y_d
        self.N = u_d.shape[0]                                         # controller.py:179
        self.u_past = u_d[-n:, :].reshape(-1, 1)                      # controller.py:184
        self.y_past = y_d[-n:, :].reshape(-1, 1)                      # controller.py:185
        self.L, self.Q, self.R = L, Q, R
        self.u_s, self.y_s = u_s, y_s
        self.eps_max, self.lamb_alpha, self.lamb_sigma, self.c = eps_max, lamb_alpha, lamb_sigma, c
        self.slack_var_constraint_type = slack_var_constraint_type
        if slack_var_constraint_type not in _SLACK_CODE:              # controller.py:211-215
            raise ValueError("Unsupported slack variable constraint type.")
        if self.controller_type == DataDrivenMPCType.ROBUST:          # controller.py:217-222
            if None in (eps_max, lamb_alpha, lamb_sigma, c):
                raise ValueError("All robust MPC parameters (eps_max, lamb_alpha, lamb_sigma, c) must be "
                                 "provided for a 'ROBUST' controller.")
        self.n_mpc_step = n_mpc_step
        self.use_terminal_constraint = use_terminal_constraint
        self.optimal_u = None
        self.problem = _Problem()
        self._pe_checked = False

        self.evaluate_input_persistent_excitation()
        self.check_prediction_horizon_length()
        self.check_weighting_matrices_dimensions()
        self.initialize_data_driven_mpc()

    # ---- validation (controller.py:242-343) ---------------------------------
    def evaluate_input_persistent_excitation(self) -> None:
        u_d_n = self.u_d.shape[1]
        if u_d_n != self.m:
            raise ValueError("The length of the elements of the data "
                             f"sequence ({u_d_n}) should match the number of "
                             f"inputs of the system ({self.m}).")
        N_min = self.m * (self.L + 2 * self.n) + self.L + 2 * self.n - 1
        if self.N < N_min:
            raise ValueError(
                "Initial input trajectory data is not persistently exciting "
                "of order (L + 2 * n). It does not satisfy the inequality: "
                "N - L - 2 * n + 1 ≥ m * (L + 2 * n). The required minimum N "
                f"is {N_min}, but got {self.N}.")
        expected_order = self.L + 2 * self.n
        in_hankel_rank, in_pers_exc = evaluate_persistent_excitation(X=self.u_d, order=expected_order)
        if not in_pers_exc:
            raise ValueError(
                "Initial input trajectory data is not persistently exciting "
                "of order (L + 2 * n). The rank of its induced Hankel matrix "
                f"({in_hankel_rank}) does not match the expected rank ("
                f"{u_d_n * expected_order}).")
        self._pe_checked = True

    def check_prediction_horizon_length(self) -> None:
        if self.controller_type == DataDrivenMPCType.NOMINAL:
            if self.L < self.n:
                raise ValueError("The prediction horizon (`L`) must be greater than or equal to the estimated "
                                 "system order `n`.")
        elif self.controller_type == DataDrivenMPCType.ROBUST:
            if self.L < 2 * self.n:
                raise ValueError("The prediction horizon (`L`) must be greater than or equal to two times the "
                                 "estimated system order `n`.")

    def check_weighting_matrices_dimensions(self) -> None:
        if self.Q.shape != (self.p * self.L, self.p * self.L):
            raise ValueError("Output weighting square matrix Q should beof order (p * L)")   # sic: controller.py:338-339
        if self.R.shape != (self.m * self.L, self.m * self.L):
            raise ValueError("Input weighting square matrix R should beof order (m * L)")   # sic: controller.py:342-343

    # ---- construction of the device plan (controller.py:345-387) ------------
    def initialize_data_driven_mpc(self) -> None:
        self._destroy()
        robust = self.controller_type == DataDrivenMPCType.ROBUST
        prm = _lib.Params(
            n=self.n, m=self.m, p=self.p, N=self.N, L=self.L,
            controller_type=_CTRL_CODE[self.controller_type],
            slack_type=_SLACK_CODE[self.slack_var_constraint_type],
            use_terminal=1 if self.use_terminal_constraint else 0,
            n_mpc_step=int(self.n_mpc_step), check_pe=0 if self._pe_checked else 1,
            eps_max=_nan_if_none(self.eps_max) if robust else 0.0,
            lamb_alpha=_nan_if_none(self.lamb_alpha) if robust else 0.0,
            lamb_sigma=_nan_if_none(self.lamb_sigma) if robust else 0.0,
            c=_nan_if_none(self.c) if robust else 0.0)
        if self.input_bounds is not None:
            prm.u_min, prm.u_max = self.input_bounds[0].ctypes.data, self.input_bounds[1].ctypes.data
        if self.output_bounds is not None:
            prm.y_min, prm.y_max = self.output_bounds[0].ctypes.data, self.output_bounds[1].ctypes.data
        ud = np.ascontiguousarray(self.u_d, dtype=np.float64)
        yd = np.ascontiguousarray(self.y_d, dtype=np.float64)
        Q = np.ascontiguousarray(self.Q, dtype=np.float64)
        R = np.ascontiguousarray(self.R, dtype=np.float64)
        handle = C.c_void_p()
        _lib.check(_lib.lib.ddmpc_set_create_host(C.byref(prm), 1, ud.ctypes.data, 0, yd.ctypes.data, 0,
                                                  Q.ctypes.data, R.ctypes.data, None, None, C.byref(handle)))
        self._set = handle
        rank, status = C.c_int(), C.c_int()
        _lib.check(_lib.lib.ddmpc_set_info(self._set, 0, C.byref(rank), C.byref(status)))
        _lib.check(status.value)
        Lp = self.L + self.n
        cols = self.N - Lp + 1
        self.HLn_ud = self._get("HLn_ud").reshape(Lp * self.m, cols)   # controller.py:376
        self.HLn_yd = self._get("HLn_yd").reshape(Lp * self.p, cols)   # controller.py:377
        self.define_optimization_variables()
        # "solve once to ensure the formulation is valid" (controller.py:385-387)
        self.solve_mpc_problem()
        self.get_optimal_control_input()

    def _get(self, name: str) -> np.ndarray:
        n_elem = C.c_size_t()
        _lib.check(_lib.lib.ddmpc_set_get(self._set, name.encode(), 0, None, 0, C.byref(n_elem)))
        out = np.empty(n_elem.value, dtype=np.float64)
        _lib.check(_lib.lib.ddmpc_set_get(self._set, name.encode(), 0, out.ctypes.data, out.size, C.byref(n_elem)))
        return out

    def define_optimization_variables(self) -> None:
        Lp = self.L + self.n
        self.alpha = _Var(self, "alpha", self.N - Lp + 1)
        self.ubar = _Var(self, "ubar", Lp * self.m)
        self.ybar = _Var(self, "ybar", Lp * self.p)
        if self.controller_type == DataDrivenMPCType.ROBUST:
            self.sigma = _Var(self, "sigma", Lp * self.p)

    # The reference rebuilds cvxpy constraint / problem objects here every step
    # (controller.py:447-504, 724-737); the device plan depends on none of the
    # quantities that change between steps, so these are no-ops kept for API parity.
    def define_mpc_constraints(self) -> None:
        return None

    def define_cost_function(self) -> None:
        return None

    def define_mpc_problem(self) -> None:
        return None

    # ---- per-step path (controller.py:389-407, 739-842) ---------------------
    def update_and_solve_data_driven_mpc(self) -> None:
        self.define_mpc_constraints()
        self.define_mpc_problem()
        self.solve_mpc_problem()
        self.get_optimal_control_input()

    def _theta(self):
        f = lambda a, k: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1)[:k])
        return (f(self.u_past, self.n * self.m), f(self.y_past, self.n * self.p), f(self.u_s, self.m),
                f(self.y_s, self.p))

    def _solve_buffers(self):
        # per-step path: persistent host buffers and their raw addresses (no allocation, no ctypes conversion per call)
        b = getattr(self, "_sbuf", None)
        if b is None:
            arr = {"up": np.empty(self.n * self.m), "yp": np.empty(self.n * self.p), "us": np.empty(self.m),
                   "ys": np.empty(self.p), "out": np.empty(self.L * self.m), "cost": np.empty(1),
                   "status": np.zeros(1, dtype=np.int32), "iters": np.zeros(1, dtype=np.int32)}
            b = self._sbuf = (arr, {k: v.ctypes.data for k, v in arr.items()})
        return b

    def solve_mpc_problem(self) -> str:
        arr, ptr = self._solve_buffers()
        arr["up"][:] = np.asarray(self.u_past, dtype=np.float64).reshape(-1)[:self.n * self.m]
        arr["yp"][:] = np.asarray(self.y_past, dtype=np.float64).reshape(-1)[:self.n * self.p]
        arr["us"][:] = np.asarray(self.u_s, dtype=np.float64).reshape(-1)[:self.m]
        arr["ys"][:] = np.asarray(self.y_s, dtype=np.float64).reshape(-1)[:self.p]
        _lib.check(_lib.lib.ddmpc_solve_batch_host(
            self._set, 1, None, ptr["up"], ptr["yp"], ptr["us"], ptr["ys"], self._solve_tol, self._solve_max_iter,
            ptr["out"], ptr["cost"], ptr["status"], ptr["iters"]))
        self._last_u = arr["out"].copy()
        self._last_theta = (arr["up"].copy(), arr["yp"].copy(), arr["us"].copy(), arr["ys"].copy())
        self._primal_cache = None
        self.problem.status = _lib.STATUS_STRINGS.get(int(arr["status"][0]), "solver_error")
        self.problem.value = float(arr["cost"][0])
        self.solver_iterations = int(arr["iters"][0])
        return self.problem.status

    def get_problem_solve_status(self) -> str:
        return self.problem.status

    def get_optimal_cost_value(self) -> float:
        return self.problem.value

    def get_optimal_control_input(self) -> np.ndarray:
        if self.problem.status in ["optimal", "optimal_inaccurate"]:
            self.optimal_u = self._last_u
            return self.optimal_u
        raise ValueError("MPC problem was not solved optimally.")      # controller.py:808

    def get_optimal_control_input_at_step(self, n_step: int = 0) -> np.ndarray:
        if not 0 <= n_step < self.L:
            raise ValueError(f"The specified prediction time step ({n_step}) is out of range. It should be "
                             f"within [0, {self.L - 1}].")
        return self.optimal_u[n_step * self.m:(n_step + 1) * self.m]

    def store_input_output_measurement(self, u_current: np.ndarray, y_current: np.ndarray) -> None:
        expected_u0_dim, expected_y0_dim = (self.m, 1), (self.p, 1)
        if u_current.shape != expected_u0_dim or y_current.shape != expected_y0_dim:
            raise ValueError(f"Incorrect dimensions. Expected dimensions are {expected_u0_dim} for u_current and "
                             f"{expected_y0_dim} for y_current, but got {u_current.shape} and {y_current.shape} "
                             "instead.")
        # controller.py:893-895 (np.vstack there; both operands are 2-D columns here, so concatenate is the same array
        # without vstack's atleast_2d pass: 1 us less per call on the per-step path)
        self.u_past = np.concatenate((self.u_past[self.m:], u_current))
        self.y_past = np.concatenate((self.y_past[self.p:], y_current))

    def set_past_input_output_data(self, u_past: np.ndarray, y_past: np.ndarray) -> None:
        expected_u_dim, expected_y_dim = (self.n * self.m, 1), (self.n * self.p, 1)
        if u_past.shape != expected_u_dim:
            raise ValueError(f"Incorrect dimensions. u_past must be shaped as {expected_u_dim}. Got "
                             f"{u_past.shape}. instead")
        if y_past.shape != expected_y_dim:
            raise ValueError(f"Incorrect dimensions. y_past must be shaped as {expected_y_dim}. Got "
                             f"{y_past.shape} instead.")
        self.u_past, self.y_past = u_past, y_past

    def set_input_output_setpoints(self, u_s: np.ndarray, y_s: np.ndarray) -> None:
        if u_s.shape != self.u_s.shape:
            raise ValueError(f"Incorrect dimensions. u_s must have shape {self.u_s.shape}, got {u_s.shape}")
        if y_s.shape != self.y_s.shape:
            raise ValueError(f"Incorrect dimensions. y_s must have shape {self.y_s.shape}, got {y_s.shape}")
        self.u_s, self.y_s = u_s, y_s
        # The reference re-runs initialize_data_driven_mpc() here (controller.py:982).  Set-points
        # only enter the right-hand side of the condensed problem, so the factorisation is kept
        # and only the validation solve is repeated.
        self.solve_mpc_problem()
        self.get_optimal_control_input()

    # ---- full primal solution on demand -------------------------------------
    def _primal(self, name: str) -> Optional[np.ndarray]:
        if self._set is None or getattr(self, "_last_theta", None) is None:
            return None
        if self._primal_cache is None:
            import torch  # device scratch for the full-primal kernels
            dev = torch.device("cuda")
            up, yp, us, ys = (torch.from_numpy(a).to(dev).reshape(1, -1) for a in self._last_theta)
            Lp = self.L + self.n
            robust = self.controller_type == DataDrivenMPCType.ROBUST
            ub = torch.empty(1, Lp * self.m, dtype=torch.float64, device=dev)
            yb = torch.empty(1, Lp * self.p, dtype=torch.float64, device=dev)
            sg = torch.empty(1, Lp * self.p, dtype=torch.float64, device=dev) if robust else None
            al = torch.empty(1, self.N - Lp + 1, dtype=torch.float64, device=dev) if robust else None
            _lib.check(_lib.lib.ddmpc_solve_full_batch(
                self._set, 1, None, up.data_ptr(), yp.data_ptr(), us.data_ptr(), ys.data_ptr(), self._solve_tol,
                self._solve_max_iter, ub.data_ptr(), yb.data_ptr(), sg.data_ptr() if robust else None,
                al.data_ptr() if robust else None, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            self._primal_cache = {"ubar": ub.cpu().numpy().reshape(-1, 1), "ybar": yb.cpu().numpy().reshape(-1, 1),
                                  "sigma": sg.cpu().numpy().reshape(-1, 1) if robust else None,
                                  "alpha": al.cpu().numpy().reshape(-1, 1) if robust else None}
        return self._primal_cache.get(name)

    # ---- lifetime ------------------------------------------------------------
    def _destroy(self) -> None:
        if getattr(self, "_set", None):
            _lib.lib.ddmpc_set_destroy(self._set)
            self._set = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass
