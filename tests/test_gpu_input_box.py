"""GPU: optional input box u_min <= ubar <= u_max (paper Eq. 6).

The reference has no such constraint (controller.py:447-504), so these tests are pinned by the oracle alone
(active set on the literal KKT system).  Tolerance: 1e-5 relative on u at solver tolerance 1e-8 (north_star)."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O

pytestmark = pytest.mark.gpu


def _plant():
    from direct_data_driven_mpc_b200 import LTIPlant
    return LTIPlant(**{k: O.FOUR_TANK[k] for k in "ABCD"}, eps_max=0.002)


def _pair(slack, term, box, n_mpc=1, seed=0, c=1.0, ybox=None):
    from direct_data_driven_mpc_b200 import ControllerSet
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(seed)
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, 1, n_mpc, term, input_bounds=box, output_bounds=ybox)
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                    c, slack, O.ROBUST, term, input_bounds=box, output_bounds=ybox)
    return cs, qp, prm, u_d, y_d, plant


@pytest.mark.parametrize("slack,term,box", [(0, True, (-3.0, [5.0, 4.0])), (1, True, (-3.0, 5.0)), (0, False, (None, 6.0)),
                                            (1, False, ([-1.0, -2.0], None))])
@pytest.mark.parametrize("B", [6, 80])     # one-CTA-per-solve kernel and thread-per-solve kernel
def test_solve_batch_with_input_box_vs_oracle(slack, term, box, B):
    cs, qp, prm, u_d, y_d, _ = _pair(slack, term, box)
    r = np.random.default_rng(4)
    ks = r.integers(0, 396, B)
    up = np.stack([u_d[k:k + 4].reshape(-1) for k in ks])
    yp = np.stack([y_d[k:k + 4].reshape(-1) for k in ks])
    us = prm["u_s"].reshape(1, -1) * r.uniform(0.8, 1.2, (B, 1))
    ys = prm["y_s"].reshape(1, -1) * r.uniform(0.8, 1.2, (B, 1))
    u, cost, status, iters = cs.solve_batch(up, yp, us, ys, tol=1e-8, max_iter=20000)
    u, cost, status, iters = u.cpu().numpy(), cost.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()
    assert (status == 0).all()
    n_active = 0
    for b in range(0, B, max(1, B // 12)):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        n_active += so.n_active
        rel = np.abs(u[b] - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max())
        assert rel < 1e-5, (b, rel, so.n_active, iters[b])
        assert abs(cost[b] - so.cost) <= 1e-5 * max(1.0, abs(so.cost))
    assert n_active > 0 and iters.max() > 1
    # the box holds on every free predicted input (scaled-row residual tolerance -> 1e-6 absolute is ample)
    nfree = (26 if term else 30) * 2
    lo = np.tile(np.broadcast_to(np.asarray(-np.inf if box[0] is None else box[0], float).reshape(-1), (2,)), 30)[:nfree]
    hi = np.tile(np.broadcast_to(np.asarray(np.inf if box[1] is None else box[1], float).reshape(-1), (2,)), 30)[:nfree]
    assert (u[:, :nfree] >= lo - 1e-6).all() and (u[:, :nfree] <= hi + 1e-6).all()


def test_box_that_never_binds_changes_nothing():
    cs0, _, prm, u_d, y_d, _ = _pair(0, True, None)
    cs1, _, _, _, _, _ = _pair(0, True, (-1e3, 1e3))
    up, yp = u_d[-4:].reshape(1, -1), y_d[-4:].reshape(1, -1)
    u0, c0, s0, i0 = cs0.solve_batch(up, yp, prm["u_s"].T, prm["y_s"].T)
    u1, c1, s1, i1 = cs1.solve_batch(up, yp, prm["u_s"].T, prm["y_s"].T)
    assert int(i1[0]) == 1 and int(s1[0]) == 0
    assert np.array_equal(u0.cpu().numpy(), u1.cpu().numpy())


@pytest.mark.parametrize("slack,n_mpc", [(0, 4), (1, 1)])
def test_closed_loop_with_input_box_vs_oracle(slack, n_mpc):
    """Whole closed loops (config-1 start: the unconstrained first move is u = 21, the box caps it at 5 / 4)."""
    box = (-3.0, [5.0, 4.0])
    cs, _, prm, u_d, y_d, plant_o = _pair(slack, True, box, n_mpc=n_mpc)
    B, n_steps = 5, 61
    r = np.random.default_rng(2)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    w = 0.002 * r.uniform(-1, 1, (B, n_steps, 2))
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    us, ys = np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1))
    u, y, st, it = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w, max_iter=20000)
    u, y = u.cpu().numpy(), y.cpu().numpy()
    assert int(st.max()) == 0
    assert u.max() <= 5.0 + 1e-6 and u[:, :, 1].max() <= 4.0 + 1e-6 and u.min() >= -3.0 - 1e-6
    assert u.max() > 5.0 - 1e-6                                    # the cap is really active
    for b in (0, 3):
        po = O.four_tank_plant()
        po.x = xs[b].copy()
        ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=n_mpc, slack_type=slack, input_bounds=box)
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[b])
        assert np.abs(u[b] - u_ref).max() / np.abs(u_ref).max() < 1e-5, b
        assert np.abs(y[b] - y_ref).max() / max(1.0, np.abs(y_ref).max()) < 1e-5, b


def test_infeasible_setpoint_nominal_and_bad_box():
    from direct_data_driven_mpc_b200 import ControllerSet
    cs, qp, prm, u_d, y_d, _ = _pair(0, True, (-0.5, 0.5))        # terminal equality ubar = u_s = 1 violates the box
    u, cost, status, iters = cs.solve_batch(u_d[-4:].reshape(1, -1), y_d[-4:].reshape(1, -1), prm["u_s"].T, prm["y_s"].T)
    assert int(status[0]) == 2                                      # DDMPC_SOLVE_INFEASIBLE
    assert qp.solve(u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1), prm["u_s"], prm["y_s"]).status == "infeasible"
    us_in = 0.3 * prm["u_s"].T                                     # a set-point pair inside the box
    u2, _, st2, _ = cs.solve_batch(u_d[-4:].reshape(1, -1), y_d[-4:].reshape(1, -1), us_in,
                                   us_in @ _plant().equilibrium_gain().T, max_iter=50000)
    assert int(st2[0]) == 0 and float(u2.abs().max()) <= 0.5 + 1e-6
    with pytest.raises(NotImplementedError):
        ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], controller_type=0, input_bounds=(-1.0, 1.0))
    with pytest.raises(ValueError):
        ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                      1.0, 0, 1, 1, True, input_bounds=(2.0, 1.0))


def test_controller_class_with_input_box():
    """B = 1 facade: same loop driver as the reference (controller_operation.py:269-305) with the box on."""
    from direct_data_driven_mpc_b200 import DirectDataDrivenMPCController, DataDrivenMPCType, SlackVarConstraintTypes
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    box = (-3.0, [5.0, 4.0])
    ctrl = DirectDataDrivenMPCController(
        n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
        eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
        slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
        n_mpc_step=4, use_terminal_constraint=True, input_bounds=box)
    ctrl._solve_max_iter = 20000
    ref = O.make_controller(prm, u_d, y_d, input_bounds=box)
    w = 0.002 * np.random.default_rng(1).uniform(-1, 1, (21, 2))
    x_start = plant_o.x.copy()
    po = O.four_tank_plant(); po.x = x_start.copy()
    u_ref, y_ref = O.closed_loop(po, ref, 21, w)
    pg = O.four_tank_plant(); pg.x = x_start.copy()
    u_g, y_g = O.closed_loop(pg, ctrl, 21, w)
    assert ctrl.get_problem_solve_status() == "optimal" and ctrl.solver_iterations >= 1
    assert np.abs(u_g - u_ref).max() / np.abs(u_ref).max() < 1e-5
    assert u_g.max() <= 5.0 + 1e-6


@pytest.mark.parametrize("slack,ub,yb", [(0, None, (None, [0.66, 0.775])), (1, (-3.0, 5.0), (0.0, [0.68, 0.79])),
                                         (0, (-3.0, [5.0, 4.0]), ([0.1, 0.1], None))])
def test_output_box_solves_and_closed_loop_vs_oracle(slack, ub, yb):
    """Output box on ybar (alone / with the input box / with the CONVEX slack bound): batched solve and a closed loop."""
    cs, qp, prm, u_d, y_d, plant_o = _pair(slack, True, ub, n_mpc=4, ybox=yb)
    B = 6
    r = np.random.default_rng(9)
    up, yp = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    us = prm["u_s"].reshape(1, -1) * r.uniform(0.97, 1.0, (B, 1))
    ys = prm["y_s"].reshape(1, -1) * r.uniform(0.97, 1.0, (B, 1))
    u, cost, st, it = cs.solve_batch(up, yp, us, ys, tol=1e-8, max_iter=50000)
    assert int(st.max()) == 0 and int(it.max()) > 1
    n_act = 0
    for b in range(B):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        n_act += so.n_active
        assert np.abs(u[b].cpu().numpy() - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max()) < 1e-5, b
        assert abs(float(cost[b]) - so.cost) <= 1e-5 * max(1.0, abs(so.cost))
    assert n_act > 0
    n_steps = 41
    xs = np.tile(plant_o.x, (2, 1))
    w = 0.002 * r.uniform(-1, 1, (2, n_steps, 2))
    uu, yy, st, _ = cs.closed_loop(_plant(), xs, up[:2], yp[:2], np.tile(prm["u_s"].T, (2, 1)), np.tile(prm["y_s"].T, (2, 1)),
                                   n_steps, w=w, max_iter=50000)
    assert int(st.max()) == 0
    po = O.four_tank_plant()
    po.x = xs[1].copy()
    ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=4, slack_type=slack, input_bounds=ub, output_bounds=yb)
    u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[1])
    assert np.abs(uu[1].cpu().numpy() - u_ref).max() / np.abs(u_ref).max() < 1e-5
    # set-point outside the output box with the terminal equality: infeasible
    cs2, qp2, *_ = _pair(0, True, None, ybox=(None, 0.5))
    _, _, st2, _ = cs2.solve_batch(up[:1], yp[:1], prm["u_s"].T, prm["y_s"].T)
    assert int(st2[0]) == 2
