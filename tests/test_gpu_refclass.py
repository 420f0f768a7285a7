"""GPU: the CUDA path (through the C ABI) against fixtures produced by the UNMODIFIED reference controller class
(tests/golden/make_golden_refclass.py; see tests/test_refclass_golden.py).  Tolerance: 1e-5 relative on u
(north_star) for the iterative CONVEX solves at tol 1e-8, 1e-8 for the direct (equality-only) variants."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O
from test_refclass_golden import SHORT_DATA, VARIANTS

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


def _plant():
    from direct_data_driven_mpc_b200 import LTIPlant
    return LTIPlant(**{k: O.FOUR_TANK[k] for k in "ABCD"}, eps_max=0.002)


def _set(u_d, y_d, slack=0, term=True, n_mpc=4, c=1.0, ctype=1, Q=None, R=None):
    from direct_data_driven_mpc_b200 import ControllerSet
    prm = O.four_tank_params()
    return ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"] if Q is None else Q, prm["R"] if R is None else R, prm["eps_max"],
                         prm["lamb_alpha"], prm["lamb_sigma"], c, slack, ctype, n_mpc, term), prm


def test_config1_closed_loop_and_every_solve_vs_reference_class(refclass):
    g = refclass["example_seed0"]
    cs, prm = _set(g["u_d"], g["y_d"])
    u, y, st, it = cs.closed_loop(_plant(), g["x_loop0"][None], g["u_d"][-4:].reshape(1, -1), g["y_d"][-4:].reshape(1, -1),
                                  prm["u_s"].reshape(1, -1), prm["y_s"].reshape(1, -1), 401, w=g["w_sys"][None])
    assert int(st[0]) == 0 and int(it[0]) == 101
    assert _rel(u.cpu().numpy()[0], g["u_sys"]) < 1e-8 and _rel(y.cpu().numpy()[0], g["y_sys"]) < 1e-8
    # every solve of the run as one batch: the whole L*m prediction and problem.value
    B = g["up"].shape[0]
    uo, cost, st, _ = cs.solve_batch(g["up"], g["yp"], np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1)))
    assert int(st.max()) == 0
    assert _rel(uo.cpu().numpy(), g["opt_u"]) < 1e-8
    assert np.abs(cost.cpu().numpy() - g["cost"]).max() <= 1e-7 * max(1.0, np.abs(g["cost"]).max())
    # and through the large-batch kernels (fused / warp-specialised): 16,384 copies of the run
    Bb = 16384 + 5
    ub, yb, sb, _ = cs.closed_loop(_plant(), np.tile(g["x_loop0"], (Bb, 1)), np.tile(g["u_d"][-4:].reshape(1, -1), (Bb, 1)),
                                   np.tile(g["y_d"][-4:].reshape(1, -1), (Bb, 1)), np.tile(prm["u_s"].T, (Bb, 1)),
                                   np.tile(prm["y_s"].T, (Bb, 1)), 401, w=np.tile(g["w_sys"][None], (Bb, 1, 1)))
    assert int(sb.max()) == 0
    for b in (0, 8191, Bb - 1):
        assert _rel(ub[b].cpu().numpy(), g["u_sys"]) < 1e-8 and _rel(yb[b].cpu().numpy(), g["y_sys"]) < 1e-8


@pytest.mark.parametrize("name,n_mpc,term,steps", [("TEC", 1, True, 597), ("TEC_N_STEP", 4, True, 597), ("UCON", 1, False, 150)])
def test_reproduction_schemes_vs_reference_class(refclass, name, n_mpc, term, steps):
    g = refclass["reproduction_seed4"]
    cs, prm = _set(g["u_d"], g["y_d"], 0, term, n_mpc)
    u, y, st, _ = cs.closed_loop(_plant(), g["x_start"][None], g["U_n"].reshape(1, -1), g["Y_n"].reshape(1, -1),
                                 prm["u_s"].reshape(1, -1), prm["y_s"].reshape(1, -1), steps, w=g[f"w_{name}"][None])
    assert int(st[0]) == 0
    assert _rel(u.cpu().numpy()[0], g[f"u_{name}"]) < 1e-8 and _rel(y.cpu().numpy()[0], g[f"y_{name}"]) < 1e-8


@pytest.mark.parametrize("name,ctype,slack,c,term,tol", VARIANTS)
def test_controller_variants_vs_reference_class(refclass, name, ctype, slack, c, term, tol):
    g = refclass["variants"]
    ud, yd = (g["u_nf"], g["y_nf"]) if name == "nominal_nf" else (g["u_d"], g["y_d"])
    Q, R = (g["Qg"], g["Rg"]) if name == "general_QR" else (None, None)
    cs, _ = _set(ud, yd, slack, term, 1, c, ctype, Q, R)
    u, cost, st, it = cs.solve_batch(g[f"{name}_up"], g[f"{name}_yp"], g[f"{name}_us"], g[f"{name}_ys"], tol=1e-8)
    assert int(st.max()) == 0
    gpu_tol = 1e-5 if slack == O.SLACK_CONVEX else (1e-6 if ctype == O.NOMINAL else 1e-8)
    assert _rel(u.cpu().numpy(), g[f"{name}_opt_u"]) < gpu_tol, name
    assert np.abs(cost.cpu().numpy() - g[f"{name}_cost"]).max() <= 1e-5 * max(1.0, np.abs(g[f"{name}_cost"]).max())
    if slack == O.SLACK_CONVEX:
        assert int(it.min()) > 1                                               # the bound binds in every fixture solve


def test_class_facade_errors_and_setpoints_vs_reference_class(refclass):
    """Exception types/texts of every invalid use and set_input_output_setpoints, reference class vs our class."""
    from direct_data_driven_mpc_b200 import DataDrivenMPCType, DirectDataDrivenMPCController, SlackVarConstraintTypes
    g, prm = refclass["errors"], O.four_tank_params()
    u_d, y_d = g["u_d"], g["y_d"]
    base = dict(n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
                eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
                slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
                n_mpc_step=4, use_terminal_constraint=True)
    t = np.arange(400)[:, None]
    ctor = {"missing_robust_params": dict(eps_max=None), "channel_mismatch": dict(u_d=np.hstack([u_d, u_d[:, :1]])),
            "short_data": dict(u_d=u_d[:100], y_d=y_d[:100]), "not_pe": dict(u_d=np.hstack([np.sin(0.3 * t), np.cos(0.2 * t)])),
            "horizon_too_short": dict(L=6, Q=np.eye(12), R=np.eye(12)), "bad_Q": dict(Q=np.eye(10)), "bad_R": dict(R=np.eye(10)),
            "non_convex": dict(slack_var_constraint_type=SlackVarConstraintTypes.NON_CONVEX)}
    ctrl = DirectDataDrivenMPCController(**base)
    calls = {"step_out_of_range": lambda: ctrl.get_optimal_control_input_at_step(n_step=30),
             "bad_measurement": lambda: ctrl.store_input_output_measurement(u_current=np.zeros((2,)), y_current=np.zeros((2, 1))),
             "bad_past_u": lambda: ctrl.set_past_input_output_data(u_past=np.zeros((7, 1)), y_past=np.zeros((8, 1))),
             "bad_past_y": lambda: ctrl.set_past_input_output_data(u_past=np.zeros((8, 1)), y_past=np.zeros((7, 1))),
             "bad_setpoint_u": lambda: ctrl.set_input_output_setpoints(u_s=np.zeros((3, 1)), y_s=np.zeros((2, 1))),
             "bad_setpoint_y": lambda: ctrl.set_input_output_setpoints(u_s=np.zeros((2, 1)), y_s=np.zeros((3, 1)))}
    for name in list(ctor) + list(calls):
        kind, text = str(g[f"err_{name}"]).split("|", 1)
        exc = {"ValueError": ValueError, "NotImplementedError": NotImplementedError}[kind]
        with pytest.raises(exc) as ei:
            if name in ctor:
                DirectDataDrivenMPCController(**{**base, **ctor[name]})
            else:
                calls[name]()
        assert str(ei.value) == text, (name, str(ei.value))
    ctrl.set_input_output_setpoints(u_s=g["setpoint_us"], y_s=g["setpoint_ys"])
    assert _rel(ctrl.optimal_u, g["setpoint_opt_u"]) < 1e-8
    assert np.array_equal(ctrl.u_past, g["setpoint_u_past"])


@pytest.mark.parametrize("n_mpc", [1, 20])
def test_config4_large_problem_vs_reference_class(refclass, n_mpc):
    """BASELINE config 4 (2661 variables / 800 equalities in the reference's formulation): the fused FP64 tensor-core
    kernel (batch of 512 copies) and the batched solve against closed-loop steps of the reference class."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    g = refclass["config4"]
    sc = S.config4_batch(1, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    w, steps = g[f"w_{n_mpc}"], g[f"w_{n_mpc}"].shape[0]
    for B in (1, 512):                                   # generic kernel / fused DMMA kernel
        t = lambda a: np.tile(np.asarray(a).reshape(1, -1), (B, 1))
        u, y, st, _ = cs.closed_loop(pl, t(g[f"x0_{n_mpc}"]), t(sc["u_past0"]), t(sc["y_past0"]), t(prm["u_s"]), t(prm["y_s"]),
                                     steps, w=np.tile(w[None], (B, 1, 1)))
        assert int(st.max()) == 0
        for b in {0, B - 1}:
            assert _rel(u[b].cpu().numpy(), g[f"u_{n_mpc}"]) < 1e-8 and _rel(y[b].cpu().numpy(), g[f"y_{n_mpc}"]) < 1e-8
    uo, cost, st, _ = cs.solve_batch(sc["u_past0"], sc["y_past0"], prm["u_s"].T, prm["y_s"].T)
    assert _rel(uo.cpu().numpy()[0], g[f"opt_u_{n_mpc}"][0]) < 1e-8              # first solve: the whole L*m prediction
    assert abs(float(cost[0]) - g[f"cost_{n_mpc}"][0]) <= 1e-7 * max(1.0, abs(g[f"cost_{n_mpc}"][0]))


@pytest.mark.parametrize("name", SHORT_DATA)
def test_singular_gram_matrix_cases_vs_reference_class(refclass, name):
    """Data the reference class accepts although W = H H^T is singular - N = N_min = 113 and N = 150 (fewer Hankel columns
    than rows), noise-free data under the ROBUST controller, the eps_max = 0 configuration of the YAML loader: the drop-in
    class driven step by step exactly as the reference's loop drives it (controller_operation.py:269-305) against the
    recorded run of the unmodified reference class, every solve (whole prediction and problem.value) included."""
    from direct_data_driven_mpc_b200 import DataDrivenMPCType, DirectDataDrivenMPCController, SlackVarConstraintTypes
    g = refclass["short_data"]
    N, noise, eps, lam_a, convex, c = g[f"{name}_params"]
    prm = O.four_tank_params()
    ctrl = DirectDataDrivenMPCController(
        n=4, m=2, p=2, u_d=g[f"{name}_u_d"], y_d=g[f"{name}_y_d"], L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
        eps_max=float(eps), lamb_alpha=float(lam_a), lamb_sigma=prm["lamb_sigma"], c=float(c),
        slack_var_constraint_type=SlackVarConstraintTypes.CONVEX if convex else SlackVarConstraintTypes.NONE,
        controller_type=DataDrivenMPCType.ROBUST, n_mpc_step=4, use_terminal_constraint=True)
    plant = O.four_tank_plant()
    plant.x = g[f"{name}_x0"].copy()
    w = g[f"{name}_w"]
    n_steps = w.shape[0]
    u, y, opt, cost = np.zeros((n_steps, 2)), np.zeros((n_steps, 2)), [], []
    for t in range(0, n_steps, 4):
        ctrl.update_and_solve_data_driven_mpc()
        assert ctrl.get_problem_solve_status() in ("optimal", "optimal_inaccurate")
        opt.append(np.asarray(ctrl.optimal_u).reshape(-1).copy())
        cost.append(ctrl.get_optimal_cost_value())
        for k in range(t, min(t + 4, n_steps)):
            u[k] = ctrl.get_optimal_control_input_at_step(n_step=k - t).reshape(-1)
            y[k] = plant.simulate_step(u[k], w[k])
            ctrl.store_input_output_measurement(u[k].reshape(-1, 1), y[k].reshape(-1, 1))
    tol = 1e-5 if convex else 1e-7
    assert _rel(u, g[f"{name}_u"]) < tol and _rel(y, g[f"{name}_y"]) < tol, (name, _rel(u, g[f"{name}_u"]))
    assert _rel(np.stack(opt), g[f"{name}_opt_u"]) < tol
    assert np.abs(np.array(cost) - g[f"{name}_cost"]).max() <= 1e-5 * max(1.0, np.abs(g[f"{name}_cost"]).max())
