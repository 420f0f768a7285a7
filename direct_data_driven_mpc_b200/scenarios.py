"""Workload generation for the BASELINE configs (host side, NumPy; not the hot path).

Mirrors the data-generation stages of the reference's example script
(``examples/direct_data_driven_mpc_example.py:263-300`` ->
``utilities/controller/controller_operation.py:13-135``): same NumPy PCG64 draw
order, so ``example_data(seed)`` reproduces the ``(u_d, y_d)``
the reference would hand to the controller for ``--seed <seed>``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .batch import LTIPlant

# examples/config/models/four_tank_system_params.yaml:9-26
FOUR_TANK_A = [[0.921, 0, 0.041, 0], [0, 0.918, 0, 0.033], [0, 0, 0.924, 0], [0, 0, 0, 0.937]]
FOUR_TANK_B = [[0.017, 0.001], [0.001, 0.023], [0, 0.061], [0.072, 0]]
FOUR_TANK_C = [[1, 0, 0, 0], [0, 1, 0, 0]]
FOUR_TANK_D = [[0, 0], [0, 0]]
FOUR_TANK_EPS = 0.002


def four_tank_plant() -> LTIPlant:
    return LTIPlant(FOUR_TANK_A, FOUR_TANK_B, FOUR_TANK_C, FOUR_TANK_D, FOUR_TANK_EPS)


def four_tank_controller_params(m: int = 2, p: int = 2) -> Dict:
    """examples/config/controllers/data_driven_mpc_example_params.yaml:9-22 through the
    derivation of utilities/controller/controller_creation.py:110-168."""
    L, eps = 30, 0.002
    return dict(u_range=(-1.0, 1.0), N=400, n=4, eps_max=eps, L=L, Q=3.0 * np.eye(p * L), R=1e-4 * np.eye(m * L),
                lamb_alpha=0.1 / eps, lamb_sigma=1000.0, c=1.0, slack_type=0, controller_type=1, n_mpc_step=4,
                u_s=np.array([[1.0], [1.0]]), y_s=np.array([[0.65], [0.77]]))


def _observer_matrices(pl: LTIPlant) -> Tuple[np.ndarray, np.ndarray]:
    """utilities/initial_state_estimation.py:3-93 (observability / Toeplitz matrices, t = n_x)."""
    n, m, p = pl.n_x, pl.m, pl.p
    pw = [np.linalg.matrix_power(pl.A, i) for i in range(n)]
    Ot = np.vstack([pl.C @ pw[i] for i in range(n)])
    Tt = np.zeros((p * n, m * n))
    for i in range(n):
        Tt[i * p:(i + 1) * p, i * m:(i + 1) * m] = pl.D
        for j in range(i):
            Tt[i * p:(i + 1) * p, j * m:(j + 1) * m] = pl.C @ pw[i - j - 1] @ pl.B
    return Ot, Tt


def simulate(pl: LTIPlant, x: np.ndarray, U: np.ndarray, W: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """utilities/model_simulation.py:93-131; returns (Y, final state)."""
    Y = np.zeros((U.shape[0], pl.p))
    for k in range(U.shape[0]):
        Y[k] = pl.C @ x + pl.D @ U[k] + W[k]
        x = pl.A @ x + pl.B @ U[k]
    return Y, x


def example_data(seed: int, pl: LTIPlant = None, N: int = 400, u_range=(-1.0, 1.0)):
    """RNG draws 1-5 of the example script: returns (rng, x0, u_d, y_d, x_after_data)."""
    pl = pl or four_tank_plant()
    rng = np.random.default_rng(seed)
    ns = pl.n_x
    x_i0 = rng.uniform(-1.0, 1.0, size=ns)
    u_i = rng.uniform(*u_range, (ns, pl.m))
    w_i = pl.eps_max * rng.uniform(-1.0, 1.0, (ns, pl.p))
    y_i, _ = simulate(pl, x_i0, u_i, w_i)
    Ot, Tt = _observer_matrices(pl)
    x0 = np.linalg.pinv(Ot) @ (y_i.flatten() - Tt @ u_i.flatten())
    u_d = rng.uniform(*u_range, (N, pl.m))
    w_d = pl.eps_max * rng.uniform(-1.0, 1.0, (N, pl.p))
    y_d, x_end = simulate(pl, x0, u_d, w_d)
    return rng, x0, u_d, y_d, x_end


def setpoint_grid(pl: LTIPlant, side: int = 16, lo: float = 0.5, hi: float = 1.5, first=None):
    """u_s on a side x side grid over [lo, hi]^m (m = 2), y_s the true
    equilibrium output C (I-A)^-1 B u_s + D u_s; element 0 is replaced by `first` = (u_s, y_s)."""
    g = np.linspace(lo, hi, side)
    us = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
    ys = us @ pl.equilibrium_gain().T
    if first is not None:
        us[0], ys[0] = np.reshape(first[0], -1), np.reshape(first[1], -1)
    return us, ys


def config3_batch(B: int = 65536, seed: int = 0, n_setpoints: int = 256):
    """four-tank robust n-step DD-MPC, shared data (seed), B closed loops = set-point grid x
    noise realisations; scenario b uses set-point b % n_setpoints and Philox stream id0 + b."""
    pl = four_tank_plant()
    prm = four_tank_controller_params()
    rng, x0, u_d, y_d, x_end = example_data(seed, pl, prm["N"], prm["u_range"])
    side = int(round(np.sqrt(n_setpoints)))
    us_g, ys_g = setpoint_grid(pl, side, first=(prm["u_s"], prm["y_s"]))
    idx = np.arange(B) % us_g.shape[0]
    n = prm["n"]
    return dict(plant=pl, params=prm, u_d=u_d, y_d=y_d,
                x0=np.tile(x_end, (B, 1)), u_past0=np.tile(u_d[-n:].reshape(1, -1), (B, 1)),
                y_past0=np.tile(y_d[-n:].reshape(1, -1), (B, 1)), u_s=us_g[idx], y_s=ys_g[idx])


def synthetic_plant(seed: int = 0, n: int = 20, m: int = 4, p: int = 4, eps_max: float = 0.002) -> LTIPlant:
    """SURVEY 8d config 4 recipe: A = 0.9 G / rho(G), B, C ~ N(0, 1/n), D = 0."""
    rng = np.random.default_rng(seed)
    G = rng.normal(size=(n, n))
    A = 0.9 * G / np.abs(np.linalg.eigvals(G)).max()
    Bm = rng.normal(scale=np.sqrt(1.0 / n), size=(n, m))
    Cm = rng.normal(scale=np.sqrt(1.0 / n), size=(p, n))
    return LTIPlant(A, Bm, Cm, np.zeros((p, m)), eps_max)


def config4_batch(B: int = 16384, seed: int = 0, N: int = 2000, L: int = 40, n_mpc_step: int = 20):
    pl = synthetic_plant(seed)
    rng = np.random.default_rng(seed + 1)
    n, m, p = pl.n_x, pl.m, pl.p
    u_d = rng.uniform(-1.0, 1.0, (N, m))
    w_d = pl.eps_max * rng.uniform(-1.0, 1.0, (N, p))
    y_d, x_end = simulate(pl, np.zeros(n), u_d, w_d)
    u_s = np.ones(m)
    y_s = pl.equilibrium_gain() @ u_s
    prm = dict(N=N, n=n, eps_max=pl.eps_max, L=L, Q=3.0 * np.eye(p * L), R=1e-4 * np.eye(m * L),
               lamb_alpha=0.1 / pl.eps_max, lamb_sigma=1000.0, c=1.0, slack_type=0, controller_type=1,
               n_mpc_step=n_mpc_step, u_s=u_s.reshape(-1, 1), y_s=y_s.reshape(-1, 1))
    return dict(plant=pl, params=prm, u_d=u_d, y_d=y_d, x0=np.tile(x_end, (B, 1)),
                u_past0=np.tile(u_d[-n:].reshape(1, -1), (B, 1)), y_past0=np.tile(y_d[-n:].reshape(1, -1), (B, 1)),
                u_s=np.tile(u_s, (B, 1)), y_s=np.tile(y_s, (B, 1)))


class DeviceScenarios:
    """S example-script scenarios generated on the GPU with the reference's NumPy streams
    (``ddmpc_generate_example_data``): tensors ``x0 (S, n_x)``, ``u_d (S, N, m)``, ``y_d (S, N, p)``,
    ``x_end (S, n_x)`` and the generator states, from which further draws continue in the
    reference's order (``uniform``)."""

    def __init__(self, seeds, plant: LTIPlant = None, N: int = 400, u_range=(-1.0, 1.0), device=None):
        import ctypes as C
        import torch
        from . import _lib
        self.plant = plant or four_tank_plant()
        pl = self.plant
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(-1))
        S = self.S = self.seeds.size
        self.N = N
        Ot, Tt = _observer_matrices(pl)
        pinv = np.ascontiguousarray(np.linalg.pinv(Ot))
        Tt = np.ascontiguousarray(Tt)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.x0 = torch.empty(S, pl.n_x, **f64)
        self.u_d = torch.empty(S, N, pl.m, **f64)
        self.y_d = torch.empty(S, N, pl.p, **f64)
        self.x_end = torch.empty(S, pl.n_x, **f64)
        self.rng_state = torch.empty(S, 4, dtype=torch.int64, device=self.device)   # raw uint64 words
        ps = pl.c_struct()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.ddmpc_generate_example_data(
                C.byref(ps), pinv.ctypes.data, Tt.ctypes.data, S, self.seeds.ctypes.data, N, float(u_range[0]),
                float(u_range[1]), float(pl.eps_max), self.x0.data_ptr(), self.u_d.data_ptr(), self.y_d.data_ptr(),
                self.x_end.data_ptr(), self.rng_state.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def uniform(self, count: int, lo: float = -1.0, hi: float = 1.0, scale: float = 1.0):
        """Next ``count`` draws of every stream: (S, count) = scale * Generator.uniform(lo, hi, count)."""
        import torch
        from . import _lib
        out = torch.empty(self.S, count, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.ddmpc_pcg64_uniform(self.rng_state.data_ptr(), self.S, count, float(lo), float(hi),
                                                    float(scale), out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
        return out
