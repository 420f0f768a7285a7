"""CPU: the oracle against every pinned vector the reference offers
(docstring known answers + fixtures generated from the live reference code)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ddmpc_oracle as O
import condensed_numpy as CN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def test_hankel_docstring_known_answer(golden_hankel):
    # hankel_matrix.py:26-37
    X = np.random.default_rng(0).uniform(-1, 1, (4, 2))
    expect = np.array([[0.27392337, -0.91805295, 0.62654048], [-0.46042657, -0.96694473, 0.82551115],
                       [-0.91805295, 0.62654048, 0.21327155], [-0.96694473, 0.82551115, 0.45899312]])
    H = O.hankel_matrix(X, 2)
    assert np.allclose(H, expect, atol=5e-9)
    assert np.array_equal(H, golden_hankel["doc_H"])


def test_hankel_bit_exact_vs_reference_fixtures(golden_hankel):
    for seed, N, nch, L in golden_hankel["cases"]:
        X = np.random.default_rng(int(seed)).normal(size=(int(N), int(nch)))
        H = O.hankel_matrix(X, int(L))
        assert H.shape == (L * nch, N - L + 1)
        assert sha(H) == str(golden_hankel[f"sha_{seed}"])
        if f"H_{seed}" in golden_hankel:
            assert np.array_equal(H, golden_hankel[f"H_{seed}"])


def test_hankel_window_error():
    with pytest.raises(ValueError):
        O.hankel_matrix(np.zeros((3, 2)), 4)


def test_toeplitz_docstring_known_answer():
    g = np.load(os.path.join(GOLDEN, "toeplitz.npz"))
    T3 = O.toeplitz_input_output_matrix(g["A"], g["B"], g["C"], g["D"], 3)
    assert np.array_equal(T3, g["T3"])
    # initial_state_estimation.py:57-70 printed values
    expect = np.array([[0., 0., 0.], [0.1, 0., 0.], [1., 0., 0.], [0.5, 0.1, 0.], [0.95, 1., 0.], [0.44, 0.5, 0.1]])
    assert np.allclose(T3, expect, atol=1e-12)


def test_scenario_generators_match_reference(golden_example):
    g = golden_example
    plant, params, rng, x0, u_d, y_d = O.example_scenario(0)
    assert np.array_equal(plant.Ot, g["Ot"]) and np.array_equal(plant.Tt, g["Tt"])
    assert np.allclose(x0, g["x0"], rtol=0, atol=1e-13)
    assert np.array_equal(u_d, g["u_d"])
    assert np.allclose(y_d, g["y_d"], rtol=0, atol=1e-13)
    assert np.allclose(plant.x, g["x_loop0"], rtol=0, atol=1e-13)
    w = plant.eps_max * rng.uniform(-1.0, 1.0, (401, 2))
    assert np.array_equal(w, g["w_sys"])
    assert params["lamb_alpha"] == float(g["lamb_alpha"])
    assert params["lamb_alpha"] * params["eps_max"] == 0.1


def test_oracle_loop_matches_reference_loop_driver(golden_example):
    """The oracle's restated loop == the reference's own loop function driving the same solver."""
    g = golden_example
    u_sys, y_sys, ctrl, ex = O.run_example(0, 400)
    assert np.allclose(u_sys, g["u_sys"], rtol=0, atol=1e-9)
    assert np.allclose(y_sys, g["y_sys"], rtol=0, atol=1e-11)
    assert len(ctrl.history) == 101
    # figure-level values (README robust_dd_mpc_example.png)
    assert np.allclose(u_sys[0], [21.22, 20.30], atol=0.01)
    assert np.allclose(y_sys[0], [-0.0349, 0.0219], atol=1e-4)
    assert np.allclose(y_sys[400], [0.652, 0.770], atol=1e-3)


def test_reproduction_matches_reference_loop_driver(golden_repro):
    g = golden_repro
    out, ex = O.run_reproduction(4, 600)
    assert np.array_equal(ex["u_d"], g["u_d"])
    assert np.allclose(ex["x_start"], g["x_start"], rtol=0, atol=1e-12)
    for name in ("TEC", "TEC_N_STEP"):
        assert np.array_equal(out[name]["w_sys"], g[f"w_{name}"])
        assert np.allclose(out[name]["u_sys"], g[f"u_{name}"], rtol=0, atol=1e-8)
        assert np.allclose(out[name]["y_sys"], g[f"y_{name}"], rtol=0, atol=1e-10)
        assert abs(np.abs(out[name]["u_sys"]).max() - 8.66) < 0.01           # figure-level
    # UCON diverges by design (reproduction.py:21-28): compare while it is still small
    k = 200
    assert np.allclose(out["UCON"]["y_sys"][:k], g["y_UCON"][:k], rtol=1e-6, atol=1e-8)
    assert np.abs(out["UCON"]["y_sys"]).max() > 10.0


def _qp(slack, ctrl, term, u_d, y_d, c=1.0):
    prm = O.four_tank_params()
    return O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                      prm["lamb_sigma"], c, slack, ctrl, term)


def test_kkt_optimality_robust_none():
    plant, params, rng, x0, u_d, y_d = O.example_scenario(0)
    qp = _qp(O.SLACK_NONE, O.ROBUST, True, u_d, y_d)
    up, yp = u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1)
    sol = qp.solve(up, yp, params["u_s"], params["y_s"])
    z = np.concatenate([sol.alpha, sol.ubar, sol.ybar, sol.sigma])
    q, b, const = qp._rhs(up, yp, params["u_s"], params["y_s"])
    assert np.abs(qp.Aeq @ z - b).max() < 1e-9
    # stationarity on the null space of the equalities: perturbations do not decrease the cost
    rng2 = np.random.default_rng(0)
    for _ in range(5):
        dz = rng2.normal(size=z.size)
        dz -= np.linalg.lstsq(qp.Aeq, qp.Aeq @ dz, rcond=None)[0]
        g = (2 * qp.P @ z + q) @ dz
        assert abs(g) < 1e-6 * np.linalg.norm(dz)
    assert abs(sol.cost - (z @ qp.P @ z + q @ z + const)) < 1e-9


def test_convex_active_set_properties():
    plant, params, rng, x0, u_d, y_d = O.example_scenario(0)
    qp = _qp(O.SLACK_CONVEX, O.ROBUST, True, u_d, y_d, c=0.5)
    qp0 = _qp(O.SLACK_NONE, O.ROBUST, True, u_d, y_d)
    up, yp = u_d[10:14].reshape(-1, 1), y_d[10:14].reshape(-1, 1)
    sol, sol0 = qp.solve(up, yp, params["u_s"], params["y_s"]), qp0.solve(up, yp, params["u_s"], params["y_s"])
    sp = sol.sigma[8:]
    assert sol.n_active > 0
    assert np.abs(sp).max() <= qp.bound * (1 + 1e-9)
    assert sol.cost >= sol0.cost - 1e-9           # a constrained optimum cannot be cheaper


@pytest.mark.parametrize("slack,term,c", [(O.SLACK_NONE, True, 1.0), (O.SLACK_NONE, False, 1.0),
                                          (O.SLACK_CONVEX, True, 1.0), (O.SLACK_CONVEX, True, 0.2),
                                          (O.SLACK_CONVEX, False, 0.3)])
def test_condensed_formulation_equals_literal_kkt(slack, term, c):
    """The row-space condensation the CUDA path implements == the reference formulation."""
    plant, params, rng, x0, u_d, y_d = O.example_scenario(1)
    qp = _qp(slack, O.ROBUST, term, u_d, y_d, c)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, params["Q"], params["R"], params["eps_max"], params["lamb_alpha"],
                       params["lamb_sigma"], c, slack, CN.ROBUST, term)
    r = np.random.default_rng(7)
    for _ in range(3):
        k = int(r.integers(0, 396))
        up, yp = u_d[k:k + 4].reshape(-1, 1), y_d[k:k + 4].reshape(-1, 1)
        us, ys = params["u_s"] * r.uniform(0.5, 1.5), params["y_s"] * r.uniform(0.5, 1.5)
        so = qp.solve(up, yp, us, ys)
        u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, up, yp, us, ys), tol=1e-10, max_iter=5000)
        assert st == "optimal"
        assert np.abs(u - so.optimal_u).max() <= 1e-9 * max(1.0, np.abs(so.optimal_u).max())
        assert abs(cost - so.cost) <= 1e-8 * max(1.0, abs(so.cost))


def test_condensed_formulation_with_zero_alpha_weight_equals_literal_kkt():
    """eps_max = 0 under a ROBUST controller, noise-free data (controller_creation.py:129-136 sets lamb_alpha = 1000 for
    it): alpha carries no weight, W is singular.  (ubar, ybar, sigma) stay unique; the condensation with W^+ and the range
    constraint finds them, the oracle takes the minimum-norm KKT point."""
    plant = O.Plant(**{**O.FOUR_TANK, "eps_max": 0.0})
    params = O.four_tank_params()
    rng = np.random.default_rng(1)
    plant.x = rng.uniform(-1, 1, 4)
    u_d, y_d = O.generate_initial_input_output_data(plant, 400, [-1, 1], rng)
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, params["Q"], params["R"], 0.0, 1000.0, 1000.0, 1.0, O.SLACK_NONE, O.ROBUST, True)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, params["Q"], params["R"], 0.0, 1000.0, 1000.0, 1.0, O.SLACK_NONE, CN.ROBUST, True)
    for k in (10, 200, 390):
        up, yp = u_d[k:k + 4].reshape(-1, 1), y_d[k:k + 4].reshape(-1, 1)
        so = qp.solve(up, yp, params["u_s"], params["y_s"])
        u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, up, yp, params["u_s"], params["y_s"]))
        assert so.status == "optimal" and st == "optimal"
        assert np.abs(u - so.optimal_u).max() <= 1e-8 * max(1.0, np.abs(so.optimal_u).max())
        assert abs(cost - so.cost) <= 1e-8 * max(1.0, abs(so.cost))


@pytest.mark.parametrize("N,slack,term,c", [(113, O.SLACK_NONE, True, 1.0), (150, O.SLACK_NONE, False, 1.0),
                                            (150, O.SLACK_CONVEX, True, 0.3)])
def test_condensed_formulation_with_short_data_equals_literal_kkt(N, slack, term, c):
    """Fewer Hankel columns than rows (the reference accepts N down to N_min = 113 here, controller.py:275-283): W = H H^T
    is singular; the condensation then uses W^+ and keeps t in range(H) by a Schur complement (what setup.cu build_robust
    does on the device).  Same optimum as the literal KKT system, which never forms W."""
    plant = O.four_tank_plant()
    params = O.four_tank_params()
    rng = np.random.default_rng(1)
    plant.x = rng.uniform(-1, 1, 4)
    u_d, y_d = O.generate_initial_input_output_data(plant, N, [-1, 1], rng)
    assert N - 34 + 1 < 136
    qp = _qp(slack, O.ROBUST, term, u_d, y_d, c)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, params["Q"], params["R"], params["eps_max"], params["lamb_alpha"],
                       params["lamb_sigma"], c, slack, CN.ROBUST, term)
    r = np.random.default_rng(N)
    active = 0
    for _ in range(3):
        k = int(r.integers(0, N - 4))
        up, yp = u_d[k:k + 4].reshape(-1, 1), y_d[k:k + 4].reshape(-1, 1)
        us, ys = params["u_s"] * r.uniform(0.5, 1.5), params["y_s"] * r.uniform(0.5, 1.5)
        so = qp.solve(up, yp, us, ys)
        u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, up, yp, us, ys), tol=1e-10, max_iter=5000)
        active += it > 1
        assert st == "optimal"
        assert np.abs(u - so.optimal_u).max() <= 1e-8 * max(1.0, np.abs(so.optimal_u).max())
        assert abs(cost - so.cost) <= 1e-8 * max(1.0, abs(so.cost))
    assert slack == O.SLACK_NONE or active > 0


@pytest.mark.parametrize("slack,term,box", [(O.SLACK_NONE, True, (-3.0, [5.0, 4.0])), (O.SLACK_CONVEX, True, (-3.0, 5.0)),
                                            (O.SLACK_NONE, False, (None, 6.0)), (O.SLACK_CONVEX, False, ([-1.0, -2.0], None))])
def test_input_box_condensed_admm_equals_oracle_active_set(slack, term, box):
    """Input box (paper Eq. 6; an extension, absent from the reference): the condensed box-row ADMM with group row
    scaling that the CUDA path implements == the oracle's active set on the literal KKT system."""
    plant, params, rng, x0, u_d, y_d = O.example_scenario(1)
    prm = O.four_tank_params()
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                    1.0, slack, O.ROBUST, term, input_bounds=box)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, params["Q"], params["R"], params["eps_max"], params["lamb_alpha"],
                       params["lamb_sigma"], 1.0, slack, CN.ROBUST, term, input_bounds=box)
    r = np.random.default_rng(3)
    n_bind = 0
    for _ in range(3):
        k = int(r.integers(0, 396))
        up, yp = u_d[k:k + 4].reshape(-1, 1), y_d[k:k + 4].reshape(-1, 1)
        us, ys = params["u_s"] * r.uniform(0.8, 1.2), params["y_s"] * r.uniform(0.8, 1.2)
        so = qp.solve(up, yp, us, ys)
        u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, up, yp, us, ys), tol=1e-10, max_iter=20000)
        assert st == "optimal" and so.status == "optimal"
        n_bind += so.n_active
        lo = -np.inf if box[0] is None else np.tile(np.broadcast_to(np.asarray(box[0], float).reshape(-1), (2,)), 30)
        hi = np.inf if box[1] is None else np.tile(np.broadcast_to(np.asarray(box[1], float).reshape(-1), (2,)), 30)
        nfree = (26 if term else 30) * 2
        assert np.all(so.optimal_u[:nfree] >= (lo if np.isscalar(lo) else lo[:nfree]) - 1e-9)
        assert np.all(so.optimal_u[:nfree] <= (hi if np.isscalar(hi) else hi[:nfree]) + 1e-9)
        assert np.abs(u - so.optimal_u).max() <= 1e-7 * max(1.0, np.abs(so.optimal_u).max())
        assert abs(cost - so.cost) <= 1e-7 * max(1.0, abs(so.cost))
    assert n_bind > 0                                   # the box really binds in these cases


@pytest.mark.parametrize("slack,ub,yb", [(0, None, (None, [0.66, 0.775])), (1, (-3.0, 5.0), (0.0, [0.68, 0.79])),
                                         (0, (-3.0, [5.0, 4.0]), ([0.1, 0.1], None))])
def test_output_box_condensed_admm_equals_oracle_active_set(slack, ub, yb):
    """Output box on ybar (paper Eq. 6, y in Y; an extension), alone and together with the input box and the CONVEX
    slack bound: three row groups, equilibrated as in k_lam_rho; condensed ADMM == oracle active set."""
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0,
                    slack, O.ROBUST, True, input_bounds=ub, output_bounds=yb)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0,
                       slack, CN.ROBUST, True, input_bounds=ub, output_bounds=yb)
    so = qp.solve(u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1), prm["u_s"], prm["y_s"])
    u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, u_d[-4:], y_d[-4:], prm["u_s"], prm["y_s"]), tol=1e-10, max_iter=50000)
    assert st == "optimal" and so.status == "optimal" and so.n_active > 0
    yfree = so.ybar[8:8 + 52].reshape(-1, 2)
    if yb[1] is not None:
        assert np.all(yfree <= np.asarray(yb[1]) + 1e-9)
    if yb[0] is not None:
        assert np.all(yfree >= np.asarray(yb[0]) - 1e-9)
    assert np.abs(u - so.optimal_u).max() <= 1e-7 * max(1.0, np.abs(so.optimal_u).max())
    assert abs(cost - so.cost) <= 1e-7 * max(1.0, abs(so.cost))
    # terminal equality ybar = y_s outside the output box: infeasible
    qp2 = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0,
                     slack, O.ROBUST, True, output_bounds=(None, 0.5))
    assert qp2.solve(u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1), prm["u_s"], prm["y_s"]).status == "infeasible"


def test_input_box_infeasible_setpoint_and_nominal():
    plant, params, rng, x0, u_d, y_d = O.example_scenario(0)
    prm = O.four_tank_params()
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                    1.0, O.SLACK_NONE, O.ROBUST, True, input_bounds=(-0.5, 0.5))
    sol = qp.solve(u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1), prm["u_s"], prm["y_s"])    # u_s = 1 > 0.5
    assert sol.status == "infeasible"
    with pytest.raises(NotImplementedError):
        O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], ctrl_type=O.NOMINAL, input_bounds=(-1, 1))


def test_nominal_noise_free_and_noisy():
    prm = O.four_tank_params()
    plant = O.four_tank_plant()
    plant.eps_max = 0.0
    r = np.random.default_rng(5)
    plant.x = r.uniform(-1, 1, 4)
    u_d = r.uniform(-1, 1, (400, 2))
    y_d = plant.simulate(u_d, np.zeros((400, 2)), 400)
    u_eq = np.array([[1.0], [1.0]])
    y_eq = plant.equilibrium_output_from_input(u_eq.ravel()).reshape(-1, 1)
    for term in (True, False):
        qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], ctrl_type=O.NOMINAL, use_terminal=term)
        assert qp.rank_H == 2 * 34 + 4                        # m (L+n) + n_sys  
        pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], None, None, None, None, 0, CN.NOMINAL, term)
        for k in (0, 100, 396):
            up, yp = u_d[k:k + 4].reshape(-1, 1), y_d[k:k + 4].reshape(-1, 1)
            so = qp.solve(up, yp, u_eq, y_eq)
            assert so.status == "optimal"
            u, cost, st, it = CN.solve(pl, CN.make_theta(4, 2, 2, up, yp, u_eq, y_eq))
            assert st == "optimal"
            assert np.abs(u - so.optimal_u).max() <= 1e-7 * max(1.0, np.abs(so.optimal_u).max())
    # an inconsistent initial window is reported infeasible by both
    bad = y_d[0:4].reshape(-1, 1) + 0.05
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], ctrl_type=O.NOMINAL, use_terminal=True)
    assert qp.solve(u_d[0:4].reshape(-1, 1), bad, u_eq, y_eq).status == "infeasible"


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = O.philox4x32(np.array([[0, 0, 0, 0]], dtype=np.uint32), np.array([[0, 0]], dtype=np.uint32))
    assert [hex(int(x)) for x in out[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = O.philox4x32(np.array([[0xffffffff] * 4], dtype=np.uint32), np.array([[0xffffffff] * 2], dtype=np.uint32))
    assert [hex(int(x)) for x in out[0]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    w = O.philox_noise(0, np.arange(4), 5, 3, 0.002)
    assert w.shape == (4, 5, 3) and np.abs(w).max() <= 0.002
