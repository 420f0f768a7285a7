"""GPU: the 8-lanes-per-loop fused kernel (k_closed_loop_perloop, csrc/perloop_loop.cu) - the path of per-loop controllers
(BASELINE config 2 / 5), NOMINAL controllers and small batches of a shared controller - against the generic
thread-per-loop kernel on the same inputs (<= 1e-9: same maths, different FP64 summation order) and the oracle
(<= 1e-5 relative on u, north_star), through the C ABI."""
import warnings

import numpy as np
import pytest

from oracle import ddmpc_oracle as O

pytestmark = pytest.mark.gpu


def _plant():
    from direct_data_driven_mpc_b200 import LTIPlant
    return LTIPlant(**{k: O.FOUR_TANK[k] for k in "ABCD"}, eps_max=0.002)


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _launches():
    from direct_data_driven_mpc_b200 import _lib
    return _lib.kernel_launches()


def _per_seed(B):
    data = [O.example_scenario(s) for s in range(B)]
    return (np.stack([d[4] for d in data]), np.stack([d[5] for d in data]), np.stack([d[0].x for d in data]), data)


@pytest.mark.parametrize("ctype,term,n_mpc,n_steps", [(1, True, 4, 43), (1, True, 1, 23), (1, False, 1, 23), (1, True, 2, 21),
                                                      (0, True, 1, 23), (0, True, 4, 18), (1, True, 4, 4), (1, True, 4, 3)])
def test_per_seed_controllers_vs_generic_kernel(ctype, term, n_mpc, n_steps):
    """Config-2 style batch (own data, own controller per loop; ragged: 37 loops = 9 warps + 1 loop), Philox and uploaded
    noise, partial last block, x_final: one launch of the fused kernel equals the generic kernel."""
    from direct_data_driven_mpc_b200 import ControllerSet
    B = 37
    prm = O.four_tank_params()
    ud, yd, xs, _ = _per_seed(B)
    robust = ctype == 1
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"] if robust else None,
                       prm["lamb_alpha"] if robust else None, prm["lamb_sigma"] if robust else None,
                       1.0 if robust else None, 0, ctype, n_mpc, term)
    r = np.random.default_rng(3)
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.8, 1.2, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    idx = r.permutation(B)                                        # loop b uses controller idx[b]
    args = (_plant(), xs, ud[idx, -4:].reshape(B, -1), yd[idx, -4:].reshape(B, -1), us, ys, n_steps)
    for kw in (dict(noise_seed=7, scenario_id0=2 ** 33 + 5, noise_eps=0.002), dict(w=0.002 * r.uniform(-1, 1, (B, n_steps, 2)))):
        cs.set_option("closed_loop_path", "auto")
        l0 = _launches()
        u1, y1, s1, i1, x1 = cs.closed_loop(*args, ctrl_idx=idx, want_x_final=True, **kw)
        assert _launches() - l0 == 1
        cs.set_option("closed_loop_path", "perloop")
        u3, y3, s3, i3, x3 = cs.closed_loop(*args, ctrl_idx=idx, want_x_final=True, **kw)
        cs.set_option("closed_loop_path", "generic")
        u2, y2, s2, i2, x2 = cs.closed_loop(*args, ctrl_idx=idx, want_x_final=True, **kw)
        assert np.array_equal(u1.cpu().numpy(), u3.cpu().numpy())                 # auto IS the per-loop kernel here
        assert int(s1.max()) == 0 and int(s2.max()) == 0 and (i1 == i2).all()
        assert _rel(u1.cpu().numpy(), u2.cpu().numpy()) < 1e-9 and _rel(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-9
        assert _rel(x1.cpu().numpy(), x2.cpu().numpy()) < 1e-9


@pytest.mark.parametrize("n_mpc,c", [(4, 0.3), (1, 0.25), (4, 1.0)])
def test_per_seed_convex_controllers_vs_generic_kernel_and_oracle(n_mpc, c):
    """Per-loop CONVEX controllers (config 2, slack bound ||sigma_pred||_inf <= c eps_max): the 8 lanes of a loop run its
    box-row ADMM cooperatively with the loop's private Ks / Phi / Psi.  Same iterates as the generic kernel (equal
    iteration counts, 1e-8 on the trajectories); oracle active-set solution on a sample."""
    from direct_data_driven_mpc_b200 import ControllerSet
    B, n_steps = 21, 26
    prm = O.four_tank_params()
    ud, yd, xs, data = _per_seed(B)
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], c, 1, 1,
                       n_mpc, True)
    r = np.random.default_rng(4)
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.8, 1.2, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    w = 0.002 * r.uniform(-1, 1, (B, n_steps, 2))
    args = (_plant(), xs, ud[:, -4:].reshape(B, -1), yd[:, -4:].reshape(B, -1), us, ys, n_steps)
    l0 = _launches()
    u1, y1, s1, i1 = cs.closed_loop(*args, w=w, ctrl_idx=np.arange(B))
    assert _launches() - l0 == 1
    cs.set_option("closed_loop_path", "generic")
    u2, y2, s2, i2 = cs.closed_loop(*args, w=w, ctrl_idx=np.arange(B))
    cs.set_option("closed_loop_path", "auto")
    assert int(s1.max()) == 0 and int(s2.max()) == 0
    if c < 1.0:
        assert int(i1.max()) > -(-n_steps // n_mpc) + 5          # the box really binds somewhere
    assert (i1 == i2).all(), (i1 - i2).abs().max()
    assert _rel(u1.cpu().numpy(), u2.cpu().numpy()) < 1e-8 and _rel(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-8
    for b in (0, 7, B - 1):
        po = O.four_tank_plant()
        po.x = xs[b].copy()
        ctrl = O.make_controller(prm, ud[b], yd[b], n_mpc_step=n_mpc, slack_type=O.SLACK_CONVEX, c=c)
        ctrl.u_s, ctrl.y_s = us[b].reshape(-1, 1), ys[b].reshape(-1, 1)
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[b])
        assert _rel(u1[b].cpu().numpy(), u_ref) < 1e-5 and _rel(y1[b].cpu().numpy(), y_ref) < 1e-5, b


def test_nominal_noise_free_data_feasibility_check():
    """NOMINAL controller built from noise-free data (rank-deficient Hankel matrix: the feasibility map F is not zero):
    a consistent window is "optimal", an inconsistent one "infeasible" (status 2), as the generic kernel reports."""
    from direct_data_driven_mpc_b200 import ControllerSet
    prm = O.four_tank_params()
    pl = O.four_tank_plant()
    pl.eps_max = 0.0
    r = np.random.default_rng(5)
    pl.x = r.uniform(-1, 1, 4)
    u_d = r.uniform(-1, 1, (400, 2))
    y_d = pl.simulate(u_d, np.zeros((400, 2)), 400)
    x_end = pl.x.copy()
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], controller_type=0, n_mpc_step=1)
    B, n_steps = 5, 9
    u_eq = np.array([1.0, 1.0])
    y_eq = pl.equilibrium_output_from_input(u_eq)
    xs = np.tile(x_end, (B, 1))
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    yp0[3] += 0.05                                                 # loop 3 starts from a window that is no trajectory
    us, ys = np.tile(u_eq, (B, 1)), np.tile(y_eq, (B, 1))
    w = np.zeros((B, n_steps, 2))
    outs = {}
    for path in ("perloop", "generic"):
        cs.set_option("closed_loop_path", path)
        outs[path] = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)
    st1, st2 = outs["perloop"][2].cpu().numpy(), outs["generic"][2].cpu().numpy()
    assert np.array_equal(st1, st2) and st1[3] == 2 and (np.delete(st1, 3) == 0).all()
    good = [0, 1, 2, 4]
    assert _rel(outs["perloop"][0].cpu().numpy()[good], outs["generic"][0].cpu().numpy()[good]) < 1e-8


def test_siso_shape_vs_generic_and_oracle():
    """The second compiled shape family (n = 2, m = p = 1, 2 plant states; n_mpc_step 1 and 2) on a random stable plant."""
    from direct_data_driven_mpc_b200 import ControllerSet, LTIPlant
    r = np.random.default_rng(12)
    A = np.array([[0.8, 0.2], [-0.1, 0.7]])
    Bm, Cm, Dm = np.array([[0.5], [1.0]]), np.array([[1.0, 0.3]]), np.array([[0.0]])
    plant = LTIPlant(A, Bm, Cm, Dm, eps_max=0.002)
    po = O.Plant(A, Bm, Cm, Dm, 0.002)
    po.x = r.uniform(-1, 1, 2)
    N, L, n = 120, 10, 2
    u_d = r.uniform(-1, 1, (N, 1))
    y_d = po.simulate(u_d, 0.002 * r.uniform(-1, 1, (N, 1)), N)
    x_end = po.x.copy()
    Q, R = 3.0 * np.eye(L), 1e-2 * np.eye(L)
    u_s = np.array([[0.5]])
    y_s = (Cm @ np.linalg.solve(np.eye(2) - A, Bm) + Dm) @ u_s
    for n_mpc in (1, 2):
        cs = ControllerSet(n, 1, 1, u_d, y_d, L, Q, R, 0.002, 50.0, 1000.0, 1.0, 0, 1, n_mpc, True)
        B, n_steps = 11, 31
        xs = np.tile(x_end, (B, 1)) + 0.1 * r.normal(size=(B, 2))
        up0, yp0 = np.tile(u_d[-n:].reshape(1, -1), (B, 1)), np.tile(y_d[-n:].reshape(1, -1), (B, 1))
        us, ys = np.tile(u_s.T, (B, 1)), np.tile(y_s.T, (B, 1))
        w = 0.002 * r.uniform(-1, 1, (B, n_steps, 1))
        res = {}
        for path in ("perloop", "generic"):
            cs.set_option("closed_loop_path", path)
            l0 = _launches()
            res[path] = cs.closed_loop(plant, xs, up0, yp0, us, ys, n_steps, w=w, want_x_final=True)
            assert _launches() - l0 == 1
        for a, b_ in zip(res["perloop"], res["generic"]):
            assert _rel(a.cpu().numpy().astype(float), b_.cpu().numpy().astype(float)) < 1e-9
        qp = O.OracleController(n, 1, 1, u_d, y_d, L, Q, R, u_s, y_s, 0.002, 50.0, 1000.0, 1.0, O.SLACK_NONE, O.ROBUST,
                                n_mpc, True, check_pe=False)
        for b in (0, B - 1):
            p2 = O.Plant(A, Bm, Cm, Dm, 0.002)
            p2.x = xs[b].copy()
            qp.set_past_input_output_data(up0[b].reshape(-1, 1), yp0[b].reshape(-1, 1))
            u_ref, y_ref = O.closed_loop(p2, qp, n_steps, w[b])
            assert _rel(res["perloop"][0].cpu().numpy()[b], u_ref) < 1e-5 and _rel(res["perloop"][1].cpu().numpy()[b], y_ref) < 1e-5


def test_small_batches_of_a_shared_controller_take_the_per_loop_kernel():
    """Below 6,144 loops a shared ROBUST controller runs on the 8-lanes-per-loop kernel (the other kernels are bound by
    their serial chain there); it agrees with the hybrid kernel and is independent of the batch split."""
    from direct_data_driven_mpc_b200 import ControllerSet
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0,
                       0, 1, 4, True)
    B, n_steps = 4096 + 5, 101
    r = np.random.default_rng(2)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    kw = dict(noise_seed=4, scenario_id0=99, noise_eps=0.002)
    u1, y1, s1, i1 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, **kw)
    cs.set_option("closed_loop_path", "perloop")
    u2, y2, s2, i2 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, **kw)
    assert np.array_equal(u1.cpu().numpy(), u2.cpu().numpy())      # auto = per-loop kernel at this batch size
    lo = 2001                                                      # a shard with its id offset = the slice of the full run
    u3, y3, _, _ = cs.closed_loop(_plant(), xs[lo:], up0[lo:], yp0[lo:], us[lo:], ys[lo:], n_steps, noise_seed=4,
                                  scenario_id0=99 + lo, noise_eps=0.002)
    assert np.array_equal(u3.cpu().numpy(), u1.cpu().numpy()[lo:]) and np.array_equal(y3.cpu().numpy(), y1.cpu().numpy()[lo:])
    cs.set_option("closed_loop_path", "fast")
    u4, y4, s4, i4 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, **kw)
    assert int(s1.max()) == 0 and (i1 == i4).all()
    assert _rel(u1.cpu().numpy(), u4.cpu().numpy()) < 1e-9 and _rel(y1.cpu().numpy(), y4.cpu().numpy()) < 1e-9


def test_failed_controllers_in_a_set_return_nan_and_status_3():
    """A controller whose data are not persistently exciting raises at construction when it is alone (as the reference
    does, controller.py:285-296); inside a larger set it is reported (failed_mask, warning) and poisoned: every solve
    and closed loop that uses it returns NaN with status 3 instead of finite numbers under "optimal"."""
    from direct_data_driven_mpc_b200 import ControllerSet
    prm = O.four_tank_params()
    ud, yd, xs, _ = _per_seed(4)
    ud[2] = 0.3                                                    # constant input: rank 1 Hankel matrix
    with pytest.raises(ValueError):
        ControllerSet(4, 2, 2, ud[2], yd[2], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                      1.0, 0, 1, 4, True)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                           1.0, 0, 1, 4, True)
    assert cs.n_failed == 1 and cs.failed_mask.tolist() == [False, False, True, False]
    assert any("could not be set up" in str(w.message) for w in rec)
    B = 4
    args = (ud[:, -4:].reshape(B, -1), yd[:, -4:].reshape(B, -1), np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1)))
    uo, _, st, _ = cs.solve_batch(*args, ctrl_idx=np.arange(B))
    assert st.cpu().numpy().tolist() == [0, 0, 3, 0] and np.isnan(uo.cpu().numpy()[2]).all()
    assert np.isfinite(uo.cpu().numpy()[[0, 1, 3]]).all()
    for path in ("perloop", "generic"):
        cs.set_option("closed_loop_path", path)
        u, y, st, _ = cs.closed_loop(_plant(), xs, *args, 9, noise_seed=1, noise_eps=0.002, ctrl_idx=np.arange(B))
        assert st.cpu().numpy().tolist() == [0, 0, 3, 0] and np.isnan(u.cpu().numpy()[2]).all(), path
        assert np.isfinite(u.cpu().numpy()[[0, 1, 3]]).all() and np.isfinite(y.cpu().numpy()[[0, 1, 3]]).all()
    with pytest.raises(ValueError):
        cs.solve_batch(*args, ctrl_idx=np.array([0, 1, 2, 4]))    # index outside the set
    with pytest.raises(ValueError):
        ControllerSet(4, 2, 2, ud[0], yd, 30, prm["Q"][:50, :50], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                      1.0, 0, 1, 4, True)
    with pytest.raises(ValueError):
        ControllerSet(4, 2, 2, ud, yd[:3], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                      1.0, 0, 1, 4, True)                           # different numbers of data sets
