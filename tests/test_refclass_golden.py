"""CPU: the oracle against fixtures produced by the UNMODIFIED reference controller class.

`tests/golden/make_golden_refclass.py` runs direct_data_driven_mpc_controller.DirectDataDrivenMPCController itself
(constructor, constraint/cost/problem builders, loop driver) with a minimal cvxpy-compatible expression layer
(`tests/golden/mini_cvxpy.py`: generic dense QP solve, no MPC knowledge).  These tests pin the oracle's restatement
of the QP - variable layout, fixed blocks, weights, slack rows, window semantics - to what the reference's own
code builds.  Tolerance 1e-9 relative (both sides solve the same strictly convex QP exactly in FP64; NOMINAL on
noisy data: 1e-7, its optimum is only unique up to the rank decision)."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


def test_config1_closed_loop_and_every_solve(refclass):
    g = refclass["example_seed0"]
    prm = O.four_tank_params()
    ctrl = O.make_controller(prm, g["u_d"], g["y_d"])
    po = O.four_tank_plant()
    po.x = g["x_loop0"].copy()
    u, y = O.closed_loop(po, ctrl, 401, g["w_sys"])
    assert _rel(u, g["u_sys"]) < 1e-9 and _rel(y, g["y_sys"]) < 1e-9
    assert len(ctrl.history) == 101 == g["opt_u"].shape[0]
    for k, (up, yp, opt_u, cost) in enumerate(ctrl.history):
        assert _rel(up.ravel(), g["up"][k]) < 1e-9 and _rel(yp.ravel(), g["yp"][k]) < 1e-9       # window semantics
        assert _rel(opt_u, g["opt_u"][k]) < 1e-9, k                                                # whole L*m prediction
        assert abs(cost - g["cost"][k]) <= 1e-8 * max(1.0, abs(g["cost"][k])), k                   # problem.value


@pytest.mark.parametrize("name,n_mpc,term,steps", [("TEC", 1, True, 597), ("TEC_N_STEP", 4, True, 597), ("UCON", 1, False, 150)])
def test_reproduction_schemes(refclass, name, n_mpc, term, steps):
    g = refclass["reproduction_seed4"]
    prm = O.four_tank_params()
    ctrl = O.make_controller(prm, g["u_d"], g["y_d"], n_mpc_step=n_mpc, use_terminal=term)
    ctrl.set_past_input_output_data(g["U_n"].reshape(-1, 1), g["Y_n"].reshape(-1, 1))
    po = O.four_tank_plant()
    po.x = g["x_start"].copy()
    u, y = O.closed_loop(po, ctrl, steps, g[f"w_{name}"])
    assert _rel(u, g[f"u_{name}"]) < 1e-9 and _rel(y, g[f"y_{name}"]) < 1e-9
    costs = np.array([h[3] for h in ctrl.history])
    assert np.abs(costs - g[f"cost_{name}"]).max() <= 1e-8 * max(1.0, np.abs(g[f"cost_{name}"]).max())


VARIANTS = [("convex_c1", O.ROBUST, O.SLACK_CONVEX, 1.0, True, 1e-9), ("convex_c03", O.ROBUST, O.SLACK_CONVEX, 0.3, True, 1e-9),
            ("convex_ucon", O.ROBUST, O.SLACK_CONVEX, 0.3, False, 1e-9), ("none_ucon", O.ROBUST, O.SLACK_NONE, 1.0, False, 1e-9),
            ("general_QR", O.ROBUST, O.SLACK_NONE, 1.0, True, 1e-9), ("nominal_nf", O.NOMINAL, O.SLACK_NONE, 1.0, True, 1e-9),
            ("nominal_noisy", O.NOMINAL, O.SLACK_NONE, 1.0, True, 1e-7)]


def variant_qp(g, name, ctype, slack, c, term):
    prm = O.four_tank_params()
    ud, yd = (g["u_nf"], g["y_nf"]) if name == "nominal_nf" else (g["u_d"], g["y_d"])
    Q, R = (g["Qg"], g["Rg"]) if name == "general_QR" else (prm["Q"], prm["R"])
    return O.OracleQP(4, 2, 2, ud, yd, 30, Q, R, prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], c, slack, ctype, term)


@pytest.mark.parametrize("name,ctype,slack,c,term,tol", VARIANTS)
def test_controller_variants(refclass, name, ctype, slack, c, term, tol):
    g = refclass["variants"]
    qp = variant_qp(g, name, ctype, slack, c, term)
    for k in range(g[f"{name}_up"].shape[0]):
        so = qp.solve(g[f"{name}_up"][k], g[f"{name}_yp"][k], g[f"{name}_us"][k], g[f"{name}_ys"][k])
        assert so.status == "optimal"
        assert _rel(so.optimal_u, g[f"{name}_opt_u"][k]) < tol, (name, k)
        assert abs(so.cost - g[f"{name}_cost"][k]) <= 1e-6 * max(1.0, abs(g[f"{name}_cost"][k]))
        if slack == O.SLACK_CONVEX:
            assert so.n_active == int(g[f"{name}_nact"][k]) > 0                 # same active slack rows, bound really binds


def test_reference_error_messages_on_host(refclass):
    """Exception types and texts of the reference class for the checks our facade performs before any device work."""
    from direct_data_driven_mpc_b200 import DataDrivenMPCType, DirectDataDrivenMPCController, SlackVarConstraintTypes
    g, prm = refclass["errors"], O.four_tank_params()
    base = dict(n=4, m=2, p=2, u_d=g["u_d"], y_d=g["y_d"], L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
                eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
                slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
                n_mpc_step=4, use_terminal_constraint=True)
    cases = {"missing_robust_params": dict(eps_max=None),
             "channel_mismatch": dict(u_d=np.hstack([g["u_d"], g["u_d"][:, :1]])),
             "short_data": dict(u_d=g["u_d"][:100], y_d=g["y_d"][:100])}
    for name, over in cases.items():
        kind, text = str(g[f"err_{name}"]).split("|", 1)
        with pytest.raises(ValueError if kind == "ValueError" else NotImplementedError) as ei:
            DirectDataDrivenMPCController(**{**base, **over})
        assert str(ei.value) == text, name


@pytest.mark.parametrize("n_mpc", [1, 20])
def test_config4_large_problem(refclass, n_mpc):
    """BASELINE config 4 (n = 20, m = p = 4, N = 2000, L = 40): oracle vs closed-loop steps of the reference class."""
    from direct_data_driven_mpc_b200 import scenarios as S
    g = refclass["config4"]
    sc = S.config4_batch(1, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    oc = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["u_s"], prm["y_s"], prm["eps_max"],
                            prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST, n_mpc, True, check_pe=False)
    po = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max)
    po.x = g[f"x0_{n_mpc}"].copy()
    u, y = O.closed_loop(po, oc, g[f"w_{n_mpc}"].shape[0], g[f"w_{n_mpc}"])
    assert _rel(u, g[f"u_{n_mpc}"]) < 1e-9 and _rel(y, g[f"y_{n_mpc}"]) < 1e-9
    assert _rel(oc.history[0][2], g[f"opt_u_{n_mpc}"][0]) < 1e-9


SHORT_DATA = ["nmin", "n150", "n150_convex", "exact", "eps0"]


@pytest.mark.parametrize("name", SHORT_DATA)
def test_singular_gram_matrix_cases_vs_reference_class(refclass, name):
    """Data the reference class accepts although W = H H^T is singular (N down to N_min, noise-free data, eps_max = 0):
    the oracle's closed loop against the recorded run of the unmodified class, every solve of it included."""
    g = refclass["short_data"]
    N, noise, eps, lam_a, convex, c = g[f"{name}_params"]
    prm = O.four_tank_params()
    ctrl = O.OracleController(4, 2, 2, g[f"{name}_u_d"], g[f"{name}_y_d"], 30, prm["Q"], prm["R"], prm["u_s"], prm["y_s"], eps, lam_a,
                              prm["lamb_sigma"], c, O.SLACK_CONVEX if convex else O.SLACK_NONE, O.ROBUST, 4, True)
    plant = O.four_tank_plant()
    plant.x = g[f"{name}_x0"].copy()
    u, y = O.closed_loop(plant, ctrl, g[f"{name}_u"].shape[0], g[f"{name}_w"])
    assert _rel(u, g[f"{name}_u"]) < 1e-9 and _rel(y, g[f"{name}_y"]) < 1e-9
    opt = np.stack([h[2] for h in ctrl.history])
    cost = np.array([h[3] for h in ctrl.history])
    assert _rel(opt, g[f"{name}_opt_u"]) < 1e-8
    assert np.abs(cost - g[f"{name}_cost"]).max() <= 1e-6 * max(1.0, np.abs(g[f"{name}_cost"]).max())
