"""Batched paper reproduction: the three robust schemes of the reference's
``examples/robust_data_driven_mpc_reproduction.py`` (TEC, TEC-n-step, UCON) for many seeds at once.

Per seed the semantics are the script's: stages 1-3 with ``default_rng(seed)``, three
controllers on the same data (``utilities/reproduction/paper_reproduction.py:43-59, 167-199``), plant at
the equilibrium of y_0 = [0.4, 0.4] (``paper_reproduction.py:104-114``), n warm-up steps at u_s with fresh
noise (``controller_operation.py:190-197``), then one closed loop per scheme of ``t_sim + 1 - n`` steps,
each drawing its own noise from the SAME generator in the order TEC, TEC-n-step, UCON
(``paper_reproduction.py:243-270``).  Data generation, controller set-up, noise draws and the loops all
run on the GPU; only the 4-step warm-up is a (vectorised) host computation.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from . import _lib
from .batch import ControllerSet
from . import scenarios as S

SCHEMES = (("TEC", 1, True), ("TEC_N_STEP", None, True), ("UCON", 1, False))   # (name, n_mpc_step, terminal)


def equilibrium_state_from_output(plant, y_eq: np.ndarray) -> np.ndarray:
    """paper_reproduction.py:104-114: u_eq = pinv(G) y_eq, x_eq = pinv(Ot) (Y_eq - Tt U_eq)."""
    Ot, Tt = S._observer_matrices(plant)
    u_eq = np.linalg.pinv(plant.equilibrium_gain()) @ y_eq
    return np.linalg.pinv(Ot) @ (np.tile(y_eq, plant.n_x) - Tt @ np.tile(u_eq, plant.n_x))


def run_reproduction_batch(seeds: Sequence[int], t_sim: int = 600, y_0=(0.4, 0.4), device=None) -> Dict:
    plant = S.four_tank_plant()
    prm = S.four_tank_controller_params()
    n, m, p = prm["n"], plant.m, plant.p
    ds = S.DeviceScenarios(seeds, plant, prm["N"], prm["u_range"], device=device)
    n_seeds, dev = ds.S, ds.device
    sets = {}
    for name, nmpc, term in SCHEMES:
        sets[name] = ControllerSet(n, m, p, ds.u_d, ds.y_d, prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                                   prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], _lib.SLACK_NONE, _lib.ROBUST,
                                   n if nmpc is None else nmpc, term, device=dev)
    # warm-up: n steps at constant u_s from the equilibrium of y_0, noise = next (n, p) draws of each stream
    x = np.tile(equilibrium_state_from_output(plant, np.asarray(y_0, dtype=float)), (n_seeds, 1))
    W_n = ds.uniform(n * p, -1.0, 1.0, plant.eps_max).cpu().numpy().reshape(n_seeds, n, p)
    u_s = prm["u_s"].reshape(-1)
    U_n = np.tile(u_s, (n_seeds, n, 1))
    Y_n = np.zeros((n_seeds, n, p))
    for k in range(n):                                        # model_simulation.py:93-98, all seeds at once
        Y_n[:, k] = x @ plant.C.T + U_n[:, k] @ plant.D.T + W_n[:, k]
        x = x @ plant.A.T + U_n[:, k] @ plant.B.T
    n_steps = t_sim + 1 - n
    us = np.tile(u_s, (n_seeds, 1))
    ys = np.tile(prm["y_s"].reshape(-1), (n_seeds, 1))
    out = {"U_n": U_n, "Y_n": Y_n, "x_start": x, "u_d": ds.u_d, "y_d": ds.y_d, "schemes": {}}
    idx = np.arange(n_seeds)
    for name, _, _ in SCHEMES:
        w = ds.uniform(n_steps * p, -1.0, 1.0, plant.eps_max).reshape(n_seeds, n_steps, p)   # this scheme's draw
        u, y, status, iters = sets[name].closed_loop(plant, x, U_n.reshape(n_seeds, -1), Y_n.reshape(n_seeds, -1), us, ys,
                                                     n_steps, w=w, ctrl_idx=idx)
        ys_t = torch.from_numpy(ys).to(y.device)
        out["schemes"][name] = {
            "u_sys": u, "y_sys": y, "status": status, "iters": iters,
            # device-side metrics: final tracking error, input peak, divergence flag
            "final_error": (y[:, -1] - ys_t).abs().amax(dim=1), "u_peak": u.abs().amax(dim=(1, 2)),
            "diverged": (y.abs().amax(dim=(1, 2)) > 10.0) | (status >= _lib.SOLVE_INFEASIBLE),
        }
    return out
