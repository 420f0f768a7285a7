"""NumPy recipes of the BASELINE workloads for the CPU side.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`bench.py --impl reference` / `cpu_baseline` and the golden-fixture generators build their inputs from this module
and `oracle.ddmpc_oracle` alone, so that neither ever imports the product package (which loads libddmpc.so).  The
recipes restate SURVEY 8d; `direct_data_driven_mpc_b200/scenarios.py` holds the product-side copies and
`tests/test_oracle_golden.py::test_workload_recipes_agree` checks the two produce identical arrays.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import ddmpc_oracle as O


def setpoint_grid(plant: O.Plant, side: int = 16, lo: float = 0.5, hi: float = 1.5, first=None) -> Tuple[np.ndarray, np.ndarray]:
    """Config 3: u_s on a side x side grid over [lo, hi]^2, y_s the true equilibrium output
    (C (I - A)^-1 B + D) u_s  (utilities/initial_state_estimation.py:162-169); element 0 is replaced by `first`."""
    g = np.linspace(lo, hi, side)
    us = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
    gain = plant.C @ np.linalg.solve(np.eye(plant.A.shape[0]) - plant.A, plant.B) + plant.D
    ys = us @ gain.T
    if first is not None:
        us[0], ys[0] = np.reshape(first[0], -1), np.reshape(first[1], -1)
    return us, ys


def config3(seed: int = 0) -> Dict:
    """Four-tank robust n-step DD-MPC on shared data (example script, --seed `seed`): data, plant state at loop start,
    the 256 set-point pairs.  Scenario b uses set-point b % 256 and Philox stream b."""
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(seed)
    us, ys = setpoint_grid(plant, 16, first=(prm["u_s"], prm["y_s"]))
    return dict(plant=plant, params=prm, u_d=u_d, y_d=y_d, x_start=plant.x.copy(), u_s=us, y_s=ys)


def synthetic_plant(seed: int = 0, n: int = 20, m: int = 4, p: int = 4, eps_max: float = 0.002) -> O.Plant:
    """Config 4 recipe: A = 0.9 G / rho(G), B, C ~ N(0, 1/n), D = 0."""
    rng = np.random.default_rng(seed)
    G = rng.normal(size=(n, n))
    A = 0.9 * G / np.abs(np.linalg.eigvals(G)).max()
    Bm = rng.normal(scale=np.sqrt(1.0 / n), size=(n, m))
    Cm = rng.normal(scale=np.sqrt(1.0 / n), size=(p, n))
    return O.Plant(A, Bm, Cm, np.zeros((p, m)), eps_max)


def config4(seed: int = 0, N: int = 2000, L: int = 40, n_mpc_step: int = 20) -> Dict:
    """Config 4: synthetic stable LTI n = 20, m = p = 4; data from zero state, set-point u_s = 1 and its equilibrium."""
    pl = synthetic_plant(seed)
    rng = np.random.default_rng(seed + 1)
    n, m, p = pl.A.shape[0], pl.B.shape[1], pl.C.shape[0]
    u_d = rng.uniform(-1.0, 1.0, (N, m))
    w_d = pl.eps_max * rng.uniform(-1.0, 1.0, (N, p))
    x = np.zeros(n)
    y_d = np.zeros((N, p))
    for k in range(N):                                             # utilities/model_simulation.py:93-98
        y_d[k] = pl.C @ x + pl.D @ u_d[k] + w_d[k]
        x = pl.A @ x + pl.B @ u_d[k]
    u_s = np.ones(m)
    y_s = (pl.C @ np.linalg.solve(np.eye(n) - pl.A, pl.B) + pl.D) @ u_s
    prm = dict(N=N, n=n, eps_max=pl.eps_max, L=L, Q=3.0 * np.eye(p * L), R=1e-4 * np.eye(m * L),
               lamb_alpha=0.1 / pl.eps_max, lamb_sigma=1000.0, c=1.0, slack_type=0, controller_type=1,
               n_mpc_step=n_mpc_step, u_s=u_s.reshape(-1, 1), y_s=y_s.reshape(-1, 1))
    return dict(plant=pl, params=prm, u_d=u_d, y_d=y_d, x_end=x, u_s=u_s, y_s=y_s)
