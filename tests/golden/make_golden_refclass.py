"""Fixtures from the UNMODIFIED reference controller CLASS (tests/golden/refclass_*.npz).

    python tests/golden/make_golden_refclass.py [/root/reference]

The reference class (direct_data_driven_mpc/direct_data_driven_mpc_controller.py) needs cvxpy, which cannot be
installed here.  ``mini_cvxpy.py`` (same directory) stands in for it: an affine-expression tracker with a generic
dense QP solve that knows nothing about MPC.  With it in ``sys.modules["cvxpy"]`` the reference's own code runs
end to end - ``create_data_driven_mpc_controller`` (controller_creation.py:192-273), the constructor with its
validation and Hankel matrices, ``define_mpc_constraints`` / ``define_cost_function`` / ``define_mpc_problem``
rebuilt at every step (controller.py:389-407), ``get_optimal_control_input``, ``store_input_output_measurement``
and the loop driver ``simulate_data_driven_mpc_control_loop`` (controller_operation.py:201-331).

So these fixtures pin what the oracle could only restate: the QP *as the reference's code builds it* (variable
layout, slices, weights, which blocks are fixed, the sigma rows the inf-norm bound acts on), the class's window and
set-point handling and its exception messages.  What stays unpinned is only the numerical tolerance of cvxpy's
backend (the fixtures hold the exact optimum of the reference's QP, to ~1e-10).

The script also checks the oracle against every fixture while it writes them and prints the deviations.
"""
import os
import sys

import numpy as np

REF = next((a for a in sys.argv[1:] if not a.startswith("--")), "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import mini_cvxpy  # noqa: E402

sys.modules["cvxpy"] = mini_cvxpy

from direct_data_driven_mpc.direct_data_driven_mpc_controller import (  # noqa: E402
    DataDrivenMPCType, DirectDataDrivenMPCController, SlackVarConstraintTypes)
assert "reference" in sys.modules["direct_data_driven_mpc.direct_data_driven_mpc_controller"].__file__, "shadow package picked up"
from utilities.model_simulation import LTISystemModel  # noqa: E402
from utilities.controller.controller_creation import (  # noqa: E402
    create_data_driven_mpc_controller, get_data_driven_mpc_controller_params)
from utilities.controller import controller_operation as ref_op  # noqa: E402

from oracle import ddmpc_oracle as O  # noqa: E402

MODEL_YAML = os.path.join(REF, "examples/config/models/four_tank_system_params.yaml")
CTRL_YAML = os.path.join(REF, "examples/config/controllers/data_driven_mpc_example_params.yaml")


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max()))


class Recorder:
    """Wraps a reference controller: records (u_past, y_past, optimal_u, cost, status) of every solve."""

    def __init__(self, ctrl):
        self._c, self.log = ctrl, []

    def __getattr__(self, name):
        return getattr(self._c, name)

    def update_and_solve_data_driven_mpc(self):
        self._c.update_and_solve_data_driven_mpc()
        self.log.append((self._c.u_past.copy().ravel(), self._c.y_past.copy().ravel(), self._c.optimal_u.copy(),
                         self._c.get_optimal_cost_value(), self._c.get_problem_solve_status()))


def setup(seed):
    model = LTISystemModel(config_file=MODEL_YAML, model_key_value="FourTankSystem")
    cfg = get_data_driven_mpc_controller_params(CTRL_YAML, "data_driven_mpc_params", m=2, p=2)
    rng = np.random.default_rng(seed)
    x0 = ref_op.randomize_initial_system_state(model, cfg, rng)
    model.set_state(state=x0)
    u_d, y_d = ref_op.generate_initial_input_output_data(model, cfg, rng)
    return model, cfg, rng, u_d, y_d


def oracle_ctrl(cfg, u_d, y_d, **over):
    kw = dict(n=cfg["n"], m=2, p=2, u_d=u_d, y_d=y_d, L=cfg["L"], Q=cfg["Q"], R=cfg["R"], u_s=cfg["u_s"], y_s=cfg["y_s"],
              eps_max=cfg["eps_max"], lamb_alpha=cfg["lamb_alpha"], lamb_sigma=cfg["lamb_sigma"], c=cfg["c"],
              slack_type=O.SLACK_NONE, ctrl_type=O.ROBUST, n_mpc_step=cfg["n_mpc_step"], use_terminal=True)
    kw.update(over)
    return O.OracleController(**kw)


def run_loop(model, ctrl, n_steps, rng):
    state = rng.bit_generator.state
    w = model.get_eps_max() * rng.uniform(-1.0, 1.0, (n_steps, 2))
    rng.bit_generator.state = state
    rec = Recorder(ctrl)
    u, y = ref_op.simulate_data_driven_mpc_control_loop(model, rec, n_steps, rng, 0)
    return u, y, w, rec.log


def pack(log):
    return dict(up=np.stack([l[0] for l in log]), yp=np.stack([l[1] for l in log]), opt_u=np.stack([l[2] for l in log]),
                cost=np.array([l[3] for l in log]))


def fixture_example(seed=0, t_sim=400):
    """config 1: examples/direct_data_driven_mpc_example.py --seed 0 --t_sim 400, reference class end to end."""
    model, cfg, rng, u_d, y_d = setup(seed)
    ctrl = create_data_driven_mpc_controller(controller_config=cfg, u_d=u_d, y_d=y_d)
    x_loop0 = model.get_state().copy()
    u, y, w, log = run_loop(model, ctrl, t_sim + 1, rng)
    assert all(l[4] == "optimal" for l in log)
    # oracle on the same inputs
    po = O.four_tank_plant(); po.x = x_loop0.copy()
    uo, yo = O.closed_loop(po, oracle_ctrl(cfg, u_d, y_d), t_sim + 1, w)
    print(f"example seed {seed}: {len(log)} reference-class solves; oracle vs reference class: u {rel(uo, u):.2e} y {rel(yo, y):.2e}")
    np.savez_compressed(os.path.join(HERE, f"refclass_example_seed{seed}.npz"), u_d=u_d, y_d=y_d, x_loop0=x_loop0, w_sys=w,
                        u_sys=u, y_sys=y, HLn_ud_sha=np.array(O.hankel_matrix(u_d, 34).tobytes().hex()[:64]), **pack(log))


def fixture_reproduction(seed=4, t_sim=600):
    """examples/robust_data_driven_mpc_reproduction.py: TEC, TEC-n-step, UCON with the reference class."""
    model, cfg, rng, u_d, y_d = setup(seed)
    n = cfg["n"]
    schemes = [("TEC", 1, True), ("TEC_N_STEP", n, True), ("UCON", 1, False)]
    ctrls = []
    for _, nmpc, term in schemes:
        c2 = dict(cfg); c2["n_mpc_step"] = nmpc
        ctrls.append(create_data_driven_mpc_controller(controller_config=c2, u_d=u_d, y_d=y_d, use_terminal_constraint=term))
    y_0 = np.array([0.4, 0.4])
    u_eq = model.get_equilibrium_input_from_output(y_eq=y_0)
    x_eq = model.get_initial_state_from_trajectory(U=np.tile(u_eq, n), Y=np.tile(y_0, n))
    model.set_state(x_eq)
    U_n, Y_n = ref_op.simulate_n_input_output_measurements(model, cfg, rng)
    for c_ in ctrls:
        c_.set_past_input_output_data(u_past=U_n.reshape(-1, 1), y_past=Y_n.reshape(-1, 1))
    x_start = model.get_state().copy()
    n_steps = t_sim + 1 - n
    out = dict(u_d=u_d, y_d=y_d, U_n=U_n, Y_n=Y_n, x_start=x_start)
    for (name, nmpc, term), c_ in zip(schemes, ctrls):
        model.set_state(state=x_start)
        steps = n_steps if name != "UCON" else 150          # UCON diverges by design; its first 150 steps are compared
        u, y, w, log = run_loop(model, c_, steps, rng)
        oc = oracle_ctrl(cfg, u_d, y_d, n_mpc_step=nmpc, use_terminal=term)
        oc.set_past_input_output_data(U_n.reshape(-1, 1), Y_n.reshape(-1, 1))
        po = O.four_tank_plant(); po.x = x_start.copy()
        uo, yo = O.closed_loop(po, oc, steps, w)
        print(f"reproduction {name}: {len(log)} solves; oracle vs reference class: u {rel(uo, u):.2e} y {rel(yo, y):.2e}")
        out[f"u_{name}"], out[f"y_{name}"], out[f"w_{name}"] = u, y, w
        out[f"cost_{name}"] = np.array([l[3] for l in log])
    np.savez_compressed(os.path.join(HERE, f"refclass_reproduction_seed{seed}.npz"), **out)


def fixture_variants(seed=1):
    """Single solves of every controller variant on random windows / set-points: CONVEX (class default slack type)
    with active and inactive bound, UCON, NOMINAL on noise-free and noisy data, general (non-diagonal) Q and R."""
    model, cfg, rng, u_d, y_d = setup(seed)
    r = np.random.default_rng(100 + seed)
    out = dict(u_d=u_d, y_d=y_d)
    # noise-free data for the NOMINAL scheme
    m0 = LTISystemModel(config_file=MODEL_YAML, model_key_value="FourTankSystem")
    m0.eps_max = 0.0
    r0 = np.random.default_rng(5)
    u_nf = r0.uniform(-1, 1, (400, 2))
    m0.set_state(r0.uniform(-0.1, 0.1, 4))
    y_nf = m0.simulate(U=u_nf, W=np.zeros((400, 2)), steps=400)
    out["u_nf"], out["y_nf"] = u_nf, y_nf
    G = r.normal(size=(60, 60)); Qg = G @ G.T / 60 + np.eye(60)
    G = r.normal(size=(60, 60)); Rg = 1e-3 * (G @ G.T / 60 + np.eye(60))
    out["Qg"], out["Rg"] = Qg, Rg
    cases = [  # name, ctrl type, slack, c, terminal, data, Q, R
        ("convex_c1", DataDrivenMPCType.ROBUST, SlackVarConstraintTypes.CONVEX, 1.0, True, "noisy", None, None),
        ("convex_c03", DataDrivenMPCType.ROBUST, SlackVarConstraintTypes.CONVEX, 0.3, True, "noisy", None, None),
        ("convex_ucon", DataDrivenMPCType.ROBUST, SlackVarConstraintTypes.CONVEX, 0.3, False, "noisy", None, None),
        ("none_ucon", DataDrivenMPCType.ROBUST, SlackVarConstraintTypes.NONE, 1.0, False, "noisy", None, None),
        ("general_QR", DataDrivenMPCType.ROBUST, SlackVarConstraintTypes.NONE, 1.0, True, "noisy", Qg, Rg),
        ("nominal_nf", DataDrivenMPCType.NOMINAL, SlackVarConstraintTypes.NONE, 1.0, True, "nf", None, None),
        ("nominal_noisy", DataDrivenMPCType.NOMINAL, SlackVarConstraintTypes.NONE, 1.0, True, "noisy", None, None),
    ]
    gain = model.C @ np.linalg.inv(np.eye(4) - model.A) @ model.B
    for name, ctype, slack, c, term, data, Qx, Rx in cases:
        ud, yd = (u_nf, y_nf) if data == "nf" else (u_d, y_d)
        robust = ctype == DataDrivenMPCType.ROBUST
        ys0 = gain @ cfg["u_s"] if data == "nf" else cfg["y_s"]            # noise-free data: the set-point must be an exact equilibrium
        ctrl = DirectDataDrivenMPCController(
            n=4, m=2, p=2, u_d=ud, y_d=yd, L=30, Q=cfg["Q"] if Qx is None else Qx, R=cfg["R"] if Rx is None else Rx,
            u_s=cfg["u_s"], y_s=ys0, eps_max=cfg["eps_max"] if robust else None,
            lamb_alpha=cfg["lamb_alpha"] if robust else None, lamb_sigma=cfg["lamb_sigma"] if robust else None,
            c=c if robust else None, slack_var_constraint_type=slack, controller_type=ctype, n_mpc_step=1,
            use_terminal_constraint=term)
        oc = O.OracleQP(4, 2, 2, ud, yd, 30, cfg["Q"] if Qx is None else Qx, cfg["R"] if Rx is None else Rx,
                        cfg["eps_max"], cfg["lamb_alpha"], cfg["lamb_sigma"], c,
                        {SlackVarConstraintTypes.NONE: O.SLACK_NONE, SlackVarConstraintTypes.CONVEX: O.SLACK_CONVEX}[slack],
                        O.ROBUST if robust else O.NOMINAL, term)
        ups, yps, uss, yss, us_out, costs, nact = [], [], [], [], [], [], []
        worst = 0.0
        for _ in range(4):
            k = int(r.integers(0, 396))
            if data == "nf":   # a window that IS a trajectory of the noise-free system, set-point = an equilibrium
                us = cfg["u_s"] * r.uniform(0.8, 1.2)
                ys = gain @ us
            else:
                us, ys = cfg["u_s"] * r.uniform(0.8, 1.2), cfg["y_s"] * r.uniform(0.8, 1.2)
            ctrl.set_past_input_output_data(u_past=ud[k:k + 4].reshape(-1, 1), y_past=yd[k:k + 4].reshape(-1, 1))
            # (assigning ctrl.u_s directly would change the terminal constraint but not the cost, which the
            #  reference builds once per initialisation: controller.py:382, 709-710)
            ctrl.set_input_output_setpoints(u_s=us, y_s=ys)                 # re-initialises and solves (controller.py:945-982)
            ctrl.update_and_solve_data_driven_mpc()
            assert ctrl.get_problem_solve_status() == "optimal", name
            so = oc.solve(ud[k:k + 4].reshape(-1, 1), yd[k:k + 4].reshape(-1, 1), us, ys)
            worst = max(worst, rel(so.optimal_u, ctrl.optimal_u))
            ups.append(ud[k:k + 4].ravel()); yps.append(yd[k:k + 4].ravel()); uss.append(us.ravel()); yss.append(ys.ravel())
            us_out.append(ctrl.optimal_u.copy()); costs.append(ctrl.get_optimal_cost_value())
            nact.append(getattr(ctrl.problem, "n_active", 0))
        print(f"variant {name}: oracle vs reference class: u {worst:.2e}  active bounds {nact}")
        out.update({f"{name}_up": np.stack(ups), f"{name}_yp": np.stack(yps), f"{name}_us": np.stack(uss),
                    f"{name}_ys": np.stack(yss), f"{name}_opt_u": np.stack(us_out), f"{name}_cost": np.array(costs),
                    f"{name}_nact": np.array(nact)})
    np.savez_compressed(os.path.join(HERE, "refclass_variants.npz"), **out)


def fixture_errors():
    """Exception types and messages of the reference class for invalid use (controller.py:165-343, 804-937)."""
    model, cfg, rng, u_d, y_d = setup(0)
    base = dict(n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=cfg["Q"], R=cfg["R"], u_s=cfg["u_s"], y_s=cfg["y_s"],
                eps_max=cfg["eps_max"], lamb_alpha=cfg["lamb_alpha"], lamb_sigma=cfg["lamb_sigma"], c=cfg["c"],
                slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
                n_mpc_step=4, use_terminal_constraint=True)
    t = np.arange(400)[:, None]
    cases = {
        "missing_robust_params": dict(eps_max=None),
        "channel_mismatch": dict(u_d=np.hstack([u_d, u_d[:, :1]])),
        "short_data": dict(u_d=u_d[:100], y_d=y_d[:100]),
        "not_pe": dict(u_d=np.hstack([np.sin(0.3 * t), np.cos(0.2 * t)])),
        "horizon_too_short": dict(L=6, Q=np.eye(12), R=np.eye(12)),
        "bad_Q": dict(Q=np.eye(10)),
        "bad_R": dict(R=np.eye(10)),
        "non_convex": dict(slack_var_constraint_type=SlackVarConstraintTypes.NON_CONVEX),
    }
    out = {}
    for name, over in cases.items():
        kw = dict(base); kw.update(over)
        try:
            DirectDataDrivenMPCController(**kw)
            out[name] = "no error"
        except Exception as exc:  # noqa: BLE001
            out[name] = f"{type(exc).__name__}|{exc}"
    ctrl = DirectDataDrivenMPCController(**base)
    for name, fn in {
        "step_out_of_range": lambda: ctrl.get_optimal_control_input_at_step(n_step=30),
        "bad_measurement": lambda: ctrl.store_input_output_measurement(u_current=np.zeros((2,)), y_current=np.zeros((2, 1))),
        "bad_past_u": lambda: ctrl.set_past_input_output_data(u_past=np.zeros((7, 1)), y_past=np.zeros((8, 1))),
        "bad_past_y": lambda: ctrl.set_past_input_output_data(u_past=np.zeros((8, 1)), y_past=np.zeros((7, 1))),
        "bad_setpoint_u": lambda: ctrl.set_input_output_setpoints(u_s=np.zeros((3, 1)), y_s=np.zeros((2, 1))),
        "bad_setpoint_y": lambda: ctrl.set_input_output_setpoints(u_s=np.zeros((2, 1)), y_s=np.zeros((3, 1))),
    }.items():
        try:
            fn()
            out[name] = "no error"
        except Exception as exc:  # noqa: BLE001
            out[name] = f"{type(exc).__name__}|{exc}"
    # set_input_output_setpoints re-initialises and solves (controller.py:945-982)
    us2, ys2 = np.array([[0.8], [1.1]]), np.array([[0.6], [0.7]])
    ctrl.set_input_output_setpoints(u_s=us2, y_s=ys2)
    for k, v in out.items():
        print(f"error case {k}: {v[:110]}")
    np.savez_compressed(os.path.join(HERE, "refclass_errors.npz"), u_d=u_d, y_d=y_d, setpoint_us=us2, setpoint_ys=ys2,
                        setpoint_opt_u=ctrl.optimal_u.copy(), setpoint_u_past=ctrl.u_past.copy(),
                        **{f"err_{k}": np.array(v) for k, v in out.items()})


def fixture_config4(n_steps=3):
    """BASELINE config 4 (synthetic n = 20, m = p = 4, N = 2000, L = 40): the reference class on the large problem
    (2661 variables, 800 equalities), three closed-loop steps of the 1-step scheme and one 20-step block."""
    from oracle import workloads as W     # NumPy recipe of the config-4 data (the product package is not imported here)
    out = {}
    for nmpc, steps in ((1, n_steps), (20, 20)):
        sc = W.config4(n_mpc_step=nmpc)
        prm, pl = sc["params"], sc["plant"]
        ctrl = DirectDataDrivenMPCController(
            n=20, m=4, p=4, u_d=sc["u_d"], y_d=sc["y_d"], L=40, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
            eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
            slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
            n_mpc_step=nmpc, use_terminal_constraint=True)
        r = np.random.default_rng(3)
        x0 = sc["x_end"] + 0.1 * r.normal(size=20)
        w = pl.eps_max * r.uniform(-1, 1, (steps, 4))
        plant = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max); plant.x = x0.copy()
        rec = Recorder(ctrl)
        u, y = O.closed_loop(plant, rec, steps, w)      # same loop as controller_operation.py:269-305, noise supplied
        oc = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["u_s"], prm["y_s"], prm["eps_max"],
                                prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST, nmpc, True, check_pe=False)
        po = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max); po.x = x0.copy()
        uo, yo = O.closed_loop(po, oc, steps, w)
        print(f"config 4 n_mpc={nmpc}: {len(rec.log)} reference-class solves; oracle vs reference class: u {rel(uo, u):.2e} y {rel(yo, y):.2e}")
        out.update({f"x0_{nmpc}": x0, f"w_{nmpc}": w, f"u_{nmpc}": u, f"y_{nmpc}": y,
                    f"opt_u_{nmpc}": np.stack([l[2] for l in rec.log]), f"cost_{nmpc}": np.array([l[3] for l in rec.log])})
    np.savez_compressed(os.path.join(HERE, "refclass_config4.npz"), **out)


def fixture_short_data(seed=2):
    """Data the reference class accepts although the Gram matrix of the stacked Hankel matrix is singular: N down to
    N_min = 113 (fewer Hankel columns than rows), noise-free data under the ROBUST controller, and the eps_max = 0
    configuration controller_creation.py:129-136 anticipates (lamb_alpha = 1000: alpha carries no weight).  A short closed
    loop of the reference class per case, every solve recorded."""
    out = {}
    cases = [  # name, N, plant noise, eps_max, lamb_alpha, slack, c
        ("nmin", 113, 0.002, 0.002, 50.0, SlackVarConstraintTypes.NONE, 1.0),
        ("n150", 150, 0.002, 0.002, 50.0, SlackVarConstraintTypes.NONE, 1.0),
        ("n150_convex", 150, 0.002, 0.002, 50.0, SlackVarConstraintTypes.CONVEX, 0.3),
        ("exact", 400, 0.0, 0.002, 50.0, SlackVarConstraintTypes.NONE, 1.0),
        ("eps0", 400, 0.0, 0.0, 1000.0, SlackVarConstraintTypes.NONE, 1.0),
    ]
    for name, N, noise, eps, lam_a, slack, c in cases:
        model = LTISystemModel(config_file=MODEL_YAML, model_key_value="FourTankSystem")
        model.eps_max = noise
        cfg = get_data_driven_mpc_controller_params(CTRL_YAML, "data_driven_mpc_params", m=2, p=2)
        rng = np.random.default_rng(seed)
        model.set_state(rng.uniform(-1.0, 1.0, 4))
        u_d = rng.uniform(-1.0, 1.0, (N, 2))
        y_d = model.simulate(U=u_d, W=noise * rng.uniform(-1.0, 1.0, (N, 2)), steps=N)
        x_loop0 = model.get_state().copy()
        ctrl = DirectDataDrivenMPCController(
            n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=cfg["Q"], R=cfg["R"], u_s=cfg["u_s"], y_s=cfg["y_s"], eps_max=eps,
            lamb_alpha=lam_a, lamb_sigma=cfg["lamb_sigma"], c=c, slack_var_constraint_type=slack,
            controller_type=DataDrivenMPCType.ROBUST, n_mpc_step=4, use_terminal_constraint=True)
        n_steps = 21
        u, y, w, log = run_loop(model, ctrl, n_steps, rng)
        assert all(l[4] == "optimal" for l in log), name
        oc = O.OracleController(4, 2, 2, u_d, y_d, 30, cfg["Q"], cfg["R"], cfg["u_s"], cfg["y_s"], eps, lam_a, cfg["lamb_sigma"], c,
                                {SlackVarConstraintTypes.NONE: O.SLACK_NONE, SlackVarConstraintTypes.CONVEX: O.SLACK_CONVEX}[slack],
                                O.ROBUST, 4, True)
        po = O.four_tank_plant()
        po.eps_max = noise
        po.x = x_loop0.copy()
        uo, yo = O.closed_loop(po, oc, n_steps, w)
        print(f"short data {name}: N {N}, Hankel columns {N - 33} of 136 rows: oracle vs reference class: u {rel(uo, u):.2e}  y {rel(yo, y):.2e}")
        out.update({f"{name}_u_d": u_d, f"{name}_y_d": y_d, f"{name}_x0": x_loop0, f"{name}_w": w, f"{name}_u": u, f"{name}_y": y,
                    f"{name}_opt_u": np.stack([l[2] for l in log]), f"{name}_cost": np.array([l[3] for l in log]),
                    f"{name}_params": np.array([N, noise, eps, lam_a, 1.0 if slack == SlackVarConstraintTypes.CONVEX else 0.0, c])})
    np.savez_compressed(os.path.join(HERE, "refclass_short_data.npz"), **out)


if __name__ == "__main__":
    if "--config4-only" in sys.argv:
        fixture_config4()
        sys.exit(0)
    if "--short-data-only" in sys.argv:
        fixture_short_data()
        sys.exit(0)
    fixture_errors()
    fixture_variants()
    fixture_example()
    fixture_reproduction()
    fixture_config4()
    fixture_short_data()
    for f in sorted(os.listdir(HERE)):
        if f.startswith("refclass_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
