"""Generates tests/golden/*.npz by running the UNMODIFIED reference code that
is importable in the build container (it is absent on the GPU box).

    python tests/golden/make_golden.py [/root/reference]

What comes from the live reference:
  * hankel_matrix (direct_data_driven_mpc/utilities/hankel_matrix.py)
  * LTIModel / LTISystemModel, observer + equilibrium helpers
    (utilities/model_simulation.py, utilities/initial_state_estimation.py)
  * randomize_initial_system_state, generate_initial_input_output_data,
    simulate_n_input_output_measurements and the closed-loop driver
    simulate_data_driven_mpc_control_loop (utilities/controller/controller_operation.py),
    imported with a stub `cvxpy` module in sys.modules (none of these functions
    touches cvxpy; only the reference controller class does).
  * the YAML parameter derivation (utilities/controller/controller_creation.py).
The QP solves inside the closed loops are done by the oracle's literal-KKT
controller (cvxpy is not installable here), driven by the reference's own loop
function - so the loop / window / plant / RNG-order semantics in the fixtures
are the reference's, and the solve is the oracle's ("parity unpinned", see
oracle/__init__.py).
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
_cp = types.ModuleType("cvxpy")
_cp.__getattr__ = lambda name: type(name, (), {})   # annotations such as cp.Constraint resolve to dummies
sys.modules.setdefault("cvxpy", _cp)

from direct_data_driven_mpc.utilities.hankel_matrix import hankel_matrix as ref_hankel  # noqa: E402
assert "reference" in sys.modules["direct_data_driven_mpc.utilities.hankel_matrix"].__file__, "shadow package picked up"
from utilities.model_simulation import LTISystemModel  # noqa: E402
from utilities.initial_state_estimation import toeplitz_input_output_matrix  # noqa: E402
from utilities.controller.controller_creation import get_data_driven_mpc_controller_params  # noqa: E402
from utilities.controller import controller_operation as ref_op  # noqa: E402

from oracle import ddmpc_oracle as O  # noqa: E402

MODEL_YAML = os.path.join(REF, "examples/config/models/four_tank_system_params.yaml")
CTRL_YAML = os.path.join(REF, "examples/config/controllers/data_driven_mpc_example_params.yaml")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def golden_hankel():
    out = {}
    # docstring known answer (hankel_matrix.py:26-37)
    X = np.random.default_rng(0).uniform(-1, 1, (4, 2))
    out["doc_X"], out["doc_H"] = X, ref_hankel(X, 2)
    cases = [(1, 7, 1, 3), (2, 9, 3, 4), (3, 12, 2, 12), (4, 5, 5, 1), (5, 400, 2, 34), (6, 400, 2, 38),
             (7, 2000, 4, 60), (8, 2000, 4, 80), (9, 64, 3, 17)]
    out["cases"] = np.array(cases)
    for (seed, N, nch, L) in cases:
        X = np.random.default_rng(seed).normal(size=(N, nch))
        H = ref_hankel(X, L)
        if H.size <= 400:
            out[f"H_{seed}"] = H
        out[f"sha_{seed}"] = np.array(sha(H))
    np.savez_compressed(os.path.join(HERE, "hankel.npz"), **out)


def golden_toeplitz():
    # docstring example of toeplitz_input_output_matrix (initial_state_estimation.py:57-70)
    A = np.array([[0.9, 0.1, 0.0], [0.0, 0.8, 0.2], [0.0, 0.0, 0.7]])
    B = np.array([[1.0], [0.5], [0.2]])
    Cm = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    D = np.array([[0.0], [0.1]])
    np.savez_compressed(os.path.join(HERE, "toeplitz.npz"), A=A, B=B, C=Cm, D=D,
                        T3=toeplitz_input_output_matrix(A, B, Cm, D, 3))


def make_oracle_ctrl(cfg, u_d, y_d, n_mpc_step, use_terminal, slack_type=O.SLACK_NONE):
    return O.OracleController(
        n=cfg["n"], m=u_d.shape[1], p=y_d.shape[1], u_d=u_d, y_d=y_d, L=cfg["L"], Q=cfg["Q"], R=cfg["R"],
        u_s=cfg["u_s"], y_s=cfg["y_s"], eps_max=cfg["eps_max"], lamb_alpha=cfg["lamb_alpha"],
        lamb_sigma=cfg["lamb_sigma"], c=cfg["c"], slack_type=slack_type, ctrl_type=O.ROBUST,
        n_mpc_step=n_mpc_step, use_terminal=use_terminal)


def golden_example(seed=0, t_sim=400):
    """examples/direct_data_driven_mpc_example.py --seed 0 --t_sim 400, stages 1-5."""
    model = LTISystemModel(config_file=MODEL_YAML, model_key_value="FourTankSystem")
    cfg = get_data_driven_mpc_controller_params(CTRL_YAML, "data_driven_mpc_params", m=2, p=2)
    rng = np.random.default_rng(seed)
    x0 = ref_op.randomize_initial_system_state(model, cfg, rng)
    model.set_state(state=x0)
    u_d, y_d = ref_op.generate_initial_input_output_data(model, cfg, rng)
    x_loop0 = model.get_state().copy()
    ctrl = make_oracle_ctrl(cfg, u_d, y_d, cfg["n_mpc_step"], True)
    # peek at the noise the loop will draw (same generator state)
    state = rng.bit_generator.state
    w_sys = model.get_eps_max() * rng.uniform(-1.0, 1.0, (t_sim + 1, 2))
    rng.bit_generator.state = state
    u_sys, y_sys = ref_op.simulate_data_driven_mpc_control_loop(model, ctrl, t_sim + 1, rng, 0)
    opt_u = np.stack([h[2] for h in ctrl.history])
    costs = np.array([h[3] for h in ctrl.history])
    np.savez_compressed(os.path.join(HERE, f"example_seed{seed}.npz"), x0=x0, u_d=u_d, y_d=y_d, x_loop0=x_loop0,
                        w_sys=w_sys, u_sys=u_sys, y_sys=y_sys, optimal_u=opt_u, cost=costs,
                        lamb_alpha=cfg["lamb_alpha"], Q0=cfg["Q"][0, 0], R0=cfg["R"][0, 0],
                        A=model.A, B=model.B, C=model.C, D=model.D, eps_max=model.get_eps_max(),
                        Ot=model.Ot, Tt=model.Tt)


def golden_reproduction(seed=4, t_sim=600):
    """examples/robust_data_driven_mpc_reproduction.py defaults (TEC, TEC-n-step, UCON)."""
    model = LTISystemModel(config_file=MODEL_YAML, model_key_value="FourTankSystem")
    cfg = get_data_driven_mpc_controller_params(CTRL_YAML, "data_driven_mpc_params", m=2, p=2)
    rng = np.random.default_rng(seed)
    x0 = ref_op.randomize_initial_system_state(model, cfg, rng)
    model.set_state(state=x0)
    u_d, y_d = ref_op.generate_initial_input_output_data(model, cfg, rng)
    n = cfg["n"]
    schemes = [("TEC", 1, True), ("TEC_N_STEP", n, True), ("UCON", 1, False)]
    ctrls = [make_oracle_ctrl(cfg, u_d, y_d, s[1], s[2]) for s in schemes]
    # paper_reproduction.py:104-114 (matplotlib import keeps that module unimportable here;
    # the three lines are restated with the live LTIModel methods)
    y_0 = np.array([0.4, 0.4])
    u_eq = model.get_equilibrium_input_from_output(y_eq=y_0)
    x_eq = model.get_initial_state_from_trajectory(U=np.tile(u_eq, n), Y=np.tile(y_0, n))
    model.set_state(x_eq)
    U_n, Y_n = ref_op.simulate_n_input_output_measurements(model, cfg, rng)
    for c_ in ctrls:
        c_.set_past_input_output_data(U_n.reshape(-1, 1), Y_n.reshape(-1, 1))
    x_start = model.get_state().copy()
    n_steps = t_sim + 1 - n
    out = dict(u_d=u_d, y_d=y_d, U_n=U_n, Y_n=Y_n, x_start=x_start, x_eq=x_eq, u_eq=u_eq)
    for (name, _, _), c_ in zip(schemes, ctrls):
        model.set_state(state=x_start)
        state = rng.bit_generator.state
        out[f"w_{name}"] = model.get_eps_max() * rng.uniform(-1.0, 1.0, (n_steps, 2))
        rng.bit_generator.state = state
        u_sys, y_sys = ref_op.simulate_data_driven_mpc_control_loop(model, c_, n_steps, rng, 0)
        out[f"u_{name}"], out[f"y_{name}"] = u_sys, y_sys
    np.savez_compressed(os.path.join(HERE, f"reproduction_seed{seed}.npz"), **out)


if __name__ == "__main__":
    golden_hankel()
    golden_toeplitz()
    golden_example()
    golden_reproduction()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
