"""GPU: fused closed loops (C ABI) vs the golden fixtures and the oracle.

Tolerance: 1e-5 relative on u (north_star); trajectories of the stable schemes
are compared over the whole run ("no divergence over t_sim"), the unstable UCON
scheme over its first 150 steps and through per-step open-loop solves."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O

pytestmark = pytest.mark.gpu


def _plant():
    from direct_data_driven_mpc_b200 import LTIPlant
    return LTIPlant(**{k: O.FOUR_TANK[k] for k in "ABCD"}, eps_max=0.002)


def _set(u_d, y_d, slack=0, term=True, n_mpc=4, c=1.0):
    from direct_data_driven_mpc_b200 import ControllerSet
    prm = O.four_tank_params()
    return ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                         prm["lamb_sigma"], c, slack, 1, n_mpc, term), prm


def _launches():
    from direct_data_driven_mpc_b200 import _lib
    return _lib.kernel_launches()


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def test_config1_example_seed0_parity(golden_example):
    """BASELINE config 1: examples/direct_data_driven_mpc_example.py --seed 0 --t_sim 400."""
    g = golden_example
    cs, prm = _set(g["u_d"], g["y_d"])
    u, y, status, iters = cs.closed_loop(_plant(), g["x_loop0"][None], g["u_d"][-4:].reshape(1, -1),
                                         g["y_d"][-4:].reshape(1, -1), prm["u_s"].reshape(1, -1),
                                         prm["y_s"].reshape(1, -1), 401, w=g["w_sys"][None])
    u, y = u.cpu().numpy()[0], y.cpu().numpy()[0]
    assert int(status[0]) == 0 and int(iters[0]) == 101
    assert _rel(u, g["u_sys"]) < 1e-5 and _rel(y, g["y_sys"]) < 1e-5
    assert np.isfinite(y).all() and np.allclose(y[400], [0.652, 0.770], atol=1e-3)
    assert np.allclose(u[0], [21.22, 20.30], atol=0.01)


def test_reproduction_three_schemes(golden_repro):
    """examples/robust_data_driven_mpc_reproduction.py (seed 4, t_sim 600): TEC, TEC-n-step, UCON."""
    g = golden_repro
    prm = O.four_tank_params()
    for name, n_mpc, term in (("TEC", 1, True), ("TEC_N_STEP", 4, True), ("UCON", 1, False)):
        cs, _ = _set(g["u_d"], g["y_d"], 0, term, n_mpc)
        u, y, status, iters = cs.closed_loop(_plant(), g["x_start"][None], g["U_n"].reshape(1, -1),
                                             g["Y_n"].reshape(1, -1), prm["u_s"].reshape(1, -1),
                                             prm["y_s"].reshape(1, -1), 597, w=g[f"w_{name}"][None])
        u, y = u.cpu().numpy()[0], y.cpu().numpy()[0]
        if name == "UCON":
            assert _rel(u[:150], g["u_UCON"][:150]) < 1e-5
            assert np.abs(y).max() > 10.0                                  # diverges by design
        else:
            assert _rel(u, g[f"u_{name}"]) < 1e-5 and _rel(y, g[f"y_{name}"]) < 1e-5
            assert int(status[0]) == 0
            assert abs(np.abs(u).max() - 8.66) < 0.01


@pytest.mark.parametrize("n_mpc,n_steps", [(1, 23), (4, 41), (3, 20), (4, 4), (7, 9)])
def test_partial_blocks_and_step_counts(n_mpc, n_steps):
    """last n-step block may be partial (controller_operation.py:278)."""
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(2)
    w = plant_o.eps_max * rng.uniform(-1, 1, (n_steps, 2))
    xs = plant_o.x.copy()
    ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=n_mpc)
    u_ref, y_ref = O.closed_loop(plant_o, ctrl, n_steps, w)
    cs, _ = _set(u_d, y_d, 0, True, n_mpc)
    u, y, status, iters, xf = cs.closed_loop(_plant(), xs[None], u_d[-4:].reshape(1, -1), y_d[-4:].reshape(1, -1),
                                             prm["u_s"].reshape(1, -1), prm["y_s"].reshape(1, -1), n_steps, w=w[None],
                                             want_x_final=True)
    assert _rel(u.cpu().numpy()[0], u_ref) < 1e-7 and _rel(y.cpu().numpy()[0], y_ref) < 1e-7
    assert int(iters[0]) == -(-n_steps // n_mpc)
    assert np.allclose(xf.cpu().numpy()[0], plant_o.x, atol=1e-9)


def test_convex_closed_loop_vs_oracle():
    """CONVEX slack bound active in the loop (tight c so that it binds)."""
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    n_steps = 24
    w = plant_o.eps_max * rng.uniform(-1, 1, (n_steps, 2))
    xs = plant_o.x.copy()
    ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=2, slack_type=O.SLACK_CONVEX, c=0.3)
    u_ref, y_ref = O.closed_loop(plant_o, ctrl, n_steps, w)
    cs, _ = _set(u_d, y_d, 1, True, 2, c=0.3)
    u, y, status, iters = cs.closed_loop(_plant(), xs[None], u_d[-4:].reshape(1, -1), y_d[-4:].reshape(1, -1),
                                         prm["u_s"].reshape(1, -1), prm["y_s"].reshape(1, -1), n_steps, w=w[None])
    assert int(status[0]) == 0 and int(iters[0]) > 12
    assert _rel(u.cpu().numpy()[0], u_ref) < 1e-5 and _rel(y.cpu().numpy()[0], y_ref) < 1e-5


def test_batch_of_loops_per_seed_controllers_and_philox():
    """config-2 style: loop b has its own data/controller; Philox noise replayed by the oracle."""
    from direct_data_driven_mpc_b200 import ControllerSet
    B, n_steps = 6, 33
    prm = O.four_tank_params()
    data = [O.example_scenario(s) for s in range(B)]
    ud, yd = np.stack([d[4] for d in data]), np.stack([d[5] for d in data])
    xs = np.stack([d[0].x for d in data])
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                       1.0, 0, 1, 4, True)
    seed, id0 = 1234567890123, 77
    u, y, status, iters = cs.closed_loop(_plant(), xs, ud[:, -4:].reshape(B, -1), yd[:, -4:].reshape(B, -1),
                                         np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1)), n_steps,
                                         w=None, noise_seed=seed, scenario_id0=id0, noise_eps=0.002,
                                         ctrl_idx=np.arange(B))
    u, y = u.cpu().numpy(), y.cpu().numpy()
    w = O.philox_noise(seed, id0 + np.arange(B), n_steps, 2, 0.002)
    for b in range(B):
        plant_o = data[b][0]
        ctrl = O.make_controller(prm, ud[b], yd[b])
        u_ref, y_ref = O.closed_loop(plant_o, ctrl, n_steps, w[b])
        assert _rel(u[b], u_ref) < 1e-6 and _rel(y[b], y_ref) < 1e-6, b
    # sharding invariance: a shard launched with an id offset equals the slice of the full run
    u2, y2, _, _ = cs.closed_loop(_plant(), xs[3:], ud[3:, -4:].reshape(3, -1), yd[3:, -4:].reshape(3, -1),
                                  np.tile(prm["u_s"].T, (3, 1)), np.tile(prm["y_s"].T, (3, 1)), n_steps, w=None,
                                  noise_seed=seed, scenario_id0=id0 + 3, noise_eps=0.002, ctrl_idx=np.arange(3, 6))
    assert np.array_equal(u2.cpu().numpy(), u[3:]) and np.array_equal(y2.cpu().numpy(), y[3:])


def test_controller_class_drives_reference_style_loop(golden_example):
    """The drop-in class (B = 1) stepping through the reference's loop semantics."""
    from direct_data_driven_mpc_b200 import (DataDrivenMPCType, DirectDataDrivenMPCController,
                                             SlackVarConstraintTypes)
    g = golden_example
    prm = O.four_tank_params()
    ctrl = DirectDataDrivenMPCController(
        n=4, m=2, p=2, u_d=g["u_d"], y_d=g["y_d"], L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
        eps_max=0.002, lamb_alpha=prm["lamb_alpha"], lamb_sigma=1000, c=1.0,
        slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
        n_mpc_step=4, use_terminal_constraint=True)
    assert ctrl.HLn_ud.shape == (68, 367) and np.array_equal(ctrl.HLn_ud, O.hankel_matrix(g["u_d"], 34))
    assert ctrl.get_problem_solve_status() == "optimal"
    plant_o = O.four_tank_plant()
    plant_o.x = g["x_loop0"].copy()
    n_steps = 61
    u_sys, y_sys = O.closed_loop(plant_o, ctrl, n_steps, g["w_sys"][:n_steps])   # duck-typed reference loop
    assert _rel(u_sys, g["u_sys"][:n_steps]) < 1e-5 and _rel(y_sys, g["y_sys"][:n_steps]) < 1e-5
    assert abs(ctrl.get_optimal_cost_value() - g["cost"][15]) < 1e-6 * max(1.0, abs(g["cost"][15]))
    assert _rel(ctrl.optimal_u, g["optimal_u"][15]) < 1e-5
    # full primal on demand; dynamics constraint [ubar; ybar + sigma] = H alpha holds
    ub, yb, sg, al = ctrl.ubar.value, ctrl.ybar.value, ctrl.sigma.value, ctrl.alpha.value
    assert ub.shape == (68, 1) and al.shape == (367, 1)
    H = np.vstack([ctrl.HLn_ud, ctrl.HLn_yd])
    assert np.abs(H @ al - np.vstack([ub, yb + sg])).max() < 1e-7
    assert np.abs(ub[8:].ravel() - ctrl.optimal_u).max() < 1e-9
    with pytest.raises(ValueError):
        ctrl.get_optimal_control_input_at_step(30)
    with pytest.raises(ValueError):
        ctrl.store_input_output_measurement(np.zeros(2), np.zeros((2, 1)))
    with pytest.raises(ValueError):
        ctrl.set_past_input_output_data(np.zeros((8, 1)), np.zeros((7, 1)))
    ctrl.set_input_output_setpoints(np.array([[0.9], [1.1]]), np.array([[0.6], [0.8]]))
    assert ctrl.get_problem_solve_status() == "optimal"


def test_many_loops_linearity_property():
    """Size-independent property at a larger batch: equality-only closed loops are affine in
    (x0, set-point, noise): loop(a) + loop(b) - loop(c) == loop(a + b - c)."""
    import torch
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    B, n_steps = 4096, 101
    r = np.random.default_rng(11)
    x = r.uniform(-0.5, 0.5, (3, B, 4))
    us = r.uniform(0.5, 1.5, (3, B, 2))
    ys = r.uniform(0.4, 0.9, (3, B, 2))
    up = r.uniform(-1, 1, (3, B, 8))
    yp = r.uniform(-0.1, 0.1, (3, B, 8))
    w = 0.002 * r.uniform(-1, 1, (3, B, n_steps, 2))
    comb = lambda a: a[0] + a[1] - a[2]
    outs = [cs.closed_loop(_plant(), x[i], up[i], yp[i], us[i], ys[i], n_steps, w=w[i])[0] for i in range(3)]
    uc = cs.closed_loop(_plant(), comb(x), comb(up), comb(yp), comb(us), comb(ys), n_steps, w=comb(w))[0]
    lhs = outs[0] + outs[1] - outs[2]
    assert (torch.abs(lhs - uc).max() / torch.abs(uc).max()).item() < 1e-9


@pytest.mark.parametrize("n_mpc", [1, 20])
def test_config4_synthetic_system_vs_oracle(n_mpc):
    """BASELINE config 4 (n=20, m=p=4, N=2000, L=40) on a small batch: generic kernel vs the literal-KKT oracle."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B, n_steps = 3, 41 if n_mpc == 20 else 6
    sc = S.config4_batch(B, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    assert cs.info(0) == (320, 0)                                  # PE of order L + 2n = 80: rank 4 * 80
    r = np.random.default_rng(0)
    x0 = sc["x0"] + 0.1 * r.normal(size=sc["x0"].shape)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    w = pl.eps_max * r.uniform(-1, 1, (B, n_steps, 4))
    u, y, status, iters = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps, w=w)
    u, y = u.cpu().numpy(), y.cpu().numpy()
    assert (status.cpu().numpy() == 0).all()
    qp = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["u_s"], prm["y_s"],
                            prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST,
                            n_mpc, True, check_pe=False)
    for b in range(B):
        plant_o = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max)
        plant_o.x = x0[b].copy()
        qp.u_s, qp.y_s = u_s[b].reshape(-1, 1), y_s[b].reshape(-1, 1)
        qp.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
        u_ref, y_ref = O.closed_loop(plant_o, qp, n_steps, w[b])
        assert _rel(u[b], u_ref) < 1e-5 and _rel(y[b], y_ref) < 1e-5, (b, _rel(u[b], u_ref))


def test_config3_full_size_properties():
    """65,536 loops (BASELINE config 3 size): every loop converges to its set-point, the batch equals the
    concatenation of two half-batch shards (sharding invariance), and a sample matches the oracle."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B = 65536
    sc = S.config3_batch(B)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], prm["c"], 0, 1, 4, True)
    args = lambda lo, hi: (pl, sc["x0"][lo:hi], sc["u_past0"][lo:hi], sc["y_past0"][lo:hi], sc["u_s"][lo:hi],
                           sc["y_s"][lo:hi], 401)
    u, y, status, iters = cs.closed_loop(*args(0, B), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    assert int(status.max()) == 0 and int(iters.min()) == 101 and int(iters.max()) == 101
    ys = torch.from_numpy(sc["y_s"]).to(y.device)
    assert float((y[:, -1] - ys).abs().max()) < 0.02                 # settled near every set-point, no divergence
    half = B // 2
    ua, ya, _, _ = cs.closed_loop(*args(0, half), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    ub, yb, _, _ = cs.closed_loop(*args(half, B), noise_seed=0, scenario_id0=half, noise_eps=0.002)
    # shards may run a different kernel specialisation (one vs two loops per thread): same numbers to ~1e-12
    assert float((torch.cat([ua, ub]) - u).abs().max()) < 1e-9 and float((torch.cat([ya, yb]) - y).abs().max()) < 1e-9
    ids = [0, 1, 255, 256, 40000, 65535]
    w = O.philox_noise(0, np.array(ids), 401, 2, 0.002)
    ctrl = O.make_controller(O.four_tank_params(), sc["u_d"], sc["y_d"])
    for j, b in enumerate(ids):
        plant_o = O.four_tank_plant()
        plant_o.x = sc["x0"][b].copy()
        ctrl.u_s, ctrl.y_s = sc["u_s"][b].reshape(-1, 1), sc["y_s"][b].reshape(-1, 1)
        ctrl.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
        u_ref, y_ref = O.closed_loop(plant_o, ctrl, 401, w[j])
        assert _rel(u[b].cpu().numpy(), u_ref) < 1e-5 and _rel(y[b].cpu().numpy(), y_ref) < 1e-5, b


@pytest.mark.parametrize("c", [1.0, 0.3])
def test_config3_convex_full_size_properties(c):
    """BASELINE config 3 with the CONVEX slack bound, 65,536 loops through k_closed_loop_cvx: a loop's trajectory does not
    depend on which loops share its CTA (the ADMM runs 32 problems in lock-step, the TF32 screen votes over the CTA) -
    shards cut at a position that is NOT a multiple of the CTA size reproduce the full batch BIT for bit, iteration
    counts included; every loop settles; a sample with active bounds matches the oracle's active-set solution."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B = 65536
    sc = S.config3_batch(B)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, 1, 1, 4, True)
    args = lambda lo, hi: (pl, sc["x0"][lo:hi], sc["u_past0"][lo:hi], sc["y_past0"][lo:hi], sc["u_s"][lo:hi],
                           sc["y_s"][lo:hi], 401)
    l0 = _launches()
    u, y, status, iters = cs.closed_loop(*args(0, B), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    assert _launches() - l0 == 1
    assert int(status.max()) == 0 and int(iters.min()) >= 101 and int(iters.max()) > 101
    ys = torch.from_numpy(sc["y_s"]).to(y.device)
    assert float((y[:, -1] - ys).abs().max()) < 0.02
    cut = 30011                                                       # 30011 = 937 * 32 + 27: every later loop changes CTA and lane
    ua, ya, _, ia = cs.closed_loop(*args(0, cut), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    ub, yb, _, ib = cs.closed_loop(*args(cut, B), noise_seed=0, scenario_id0=cut, noise_eps=0.002)
    assert torch.equal(torch.cat([ua, ub]), u) and torch.equal(torch.cat([ya, yb]), y)
    assert torch.equal(torch.cat([ia, ib]), iters)
    active = torch.nonzero(iters > 101).flatten().cpu().numpy()
    ids = [int(active[0]), int(active[len(active) // 2]), int(active[-1])]
    w = O.philox_noise(0, np.array(ids), 401, 2, 0.002)
    for j, b in enumerate(ids):
        plant_o = O.four_tank_plant()
        plant_o.x = sc["x0"][b].copy()
        ctrl = O.make_controller(O.four_tank_params(), sc["u_d"], sc["y_d"], n_mpc_step=4, slack_type=O.SLACK_CONVEX, c=c)
        ctrl.u_s, ctrl.y_s = sc["u_s"][b].reshape(-1, 1), sc["y_s"][b].reshape(-1, 1)
        ctrl.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
        u_ref, y_ref = O.closed_loop(plant_o, ctrl, 401, w[j])
        assert _rel(u[b].cpu().numpy(), u_ref) < 1e-5 and _rel(y[b].cpu().numpy(), y_ref) < 1e-5, b


@pytest.mark.parametrize("n_mpc,n_steps", [(20, 47), (8, 20), (10, 10)])
def test_config4_gemm_path_vs_generic_and_oracle(n_mpc, n_steps):
    """Large-system path (batch as the N dimension of DMMA GEMMs, block-stepped plant) vs the generic
    thread-per-loop kernel on 512 loops, and vs the oracle on a sample; Philox and uploaded noise."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B = 512
    sc = S.config4_batch(B, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    r = np.random.default_rng(1)
    x0 = sc["x0"] + 0.1 * r.normal(size=sc["x0"].shape)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    w = pl.eps_max * r.uniform(-1, 1, (B, n_steps, 4))
    for noise in ("philox", "uploaded"):
        kw = dict(w=w) if noise == "uploaded" else dict(noise_seed=5, scenario_id0=1000, noise_eps=0.002)
        cs.set_option("closed_loop_path", "gemm")
        launches0 = _launches()
        u1, y1, s1, i1, xf1 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps,
                                             want_x_final=True, **kw)
        assert _launches() - launches0 > 3                         # the per-iteration GEMMs, not a fused kernel
        cs.set_option("closed_loop_path", "generic")
        u2, y2, s2, i2, xf2 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps,
                                             want_x_final=True, **kw)
        cs.set_option("closed_loop_path", "auto")
        assert int(s1.max()) == 0 and int(s2.max()) == 0
        assert (i1 == i2).all()
        assert _rel(u1.cpu().numpy(), u2.cpu().numpy()) < 1e-9 and _rel(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-9
        assert _rel(xf1.cpu().numpy(), xf2.cpu().numpy()) < 1e-9
    qp = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["u_s"], prm["y_s"],
                            prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST,
                            n_mpc, True, check_pe=False)
    u1, y1 = u1.cpu().numpy(), y1.cpu().numpy()
    for b in (0, 511):
        plant_o = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max)
        plant_o.x = x0[b].copy()
        qp.u_s, qp.y_s = u_s[b].reshape(-1, 1), y_s[b].reshape(-1, 1)
        qp.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
        u_ref, y_ref = O.closed_loop(plant_o, qp, n_steps, w[b])
        assert _rel(u1[b], u_ref) < 1e-5 and _rel(y1[b], y_ref) < 1e-5


@pytest.mark.parametrize("n_mpc,n_steps,B", [(1, 23, 515), (20, 47, 515), (20, 401, 264)])
def test_config4_fused_dmma_kernel_vs_generic_and_oracle(n_mpc, n_steps, B):
    """k_closed_loop_dmma (one launch, a warp per 8 loops, ring window in shared memory) vs the generic
    thread-per-loop kernel on a ragged batch, and vs the oracle on a sample; Philox and uploaded noise."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    sc = S.config4_batch(B, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    r = np.random.default_rng(1)
    x0 = sc["x0"] + 0.1 * r.normal(size=sc["x0"].shape)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    w = pl.eps_max * r.uniform(-1, 1, (B, n_steps, 4))
    for noise in ("philox", "uploaded"):
        kw = dict(w=w) if noise == "uploaded" else dict(noise_seed=5, scenario_id0=1000, noise_eps=0.002)
        launches0 = _launches()
        u1, y1, s1, i1, xf1 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps,
                                             want_x_final=True, **kw)
        assert _launches() - launches0 == 1                        # the fused kernel, not the per-iteration GEMMs
        cs.set_option("closed_loop_path", "generic")
        u2, y2, s2, i2, xf2 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps,
                                             want_x_final=True, **kw)
        cs.set_option("closed_loop_path", "auto")
        assert int(s1.max()) == 0 and int(s2.max()) == 0
        assert (i1 == i2).all()
        assert _rel(u1.cpu().numpy(), u2.cpu().numpy()) < 1e-9 and _rel(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-9
        assert _rel(xf1.cpu().numpy(), xf2.cpu().numpy()) < 1e-9
    qp = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["u_s"], prm["y_s"],
                            prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST,
                            n_mpc, True, check_pe=False)
    u1, y1 = u1.cpu().numpy(), y1.cpu().numpy()
    for b in (0, B - 1):
        plant_o = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max)
        plant_o.x = x0[b].copy()
        qp.u_s, qp.y_s = u_s[b].reshape(-1, 1), y_s[b].reshape(-1, 1)
        qp.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
        u_ref, y_ref = O.closed_loop(plant_o, qp, min(n_steps, 60), w[b][:60])
        assert _rel(u1[b][:60], u_ref) < 1e-5 and _rel(y1[b][:60], y_ref) < 1e-5


@pytest.mark.parametrize("n_steps,B", [(47, 300), (401, 140), (20, 128), (7, 5)])
def test_config4_tcgen05_path_within_tolerance_of_fp64_kernel(n_steps, B):
    """k_closed_loop_tc (opt-in path "tc": both products of an MPC iteration on tcgen05 / TMEM in TF32x3 arithmetic) against
    the FP64 tensor-core kernel: 1e-5 relative on u and y (north-star tolerance; scripts/tf32x3_emulation.py predicts
    ~5e-7), equal status and iteration counts; ragged batch, partial last block, Philox and uploaded noise, x_final."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    sc = S.config4_batch(B, n_mpc_step=20)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, 20, True)
    r = np.random.default_rng(1)
    x0 = sc["x0"] + 0.1 * r.normal(size=sc["x0"].shape)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    w = pl.eps_max * r.uniform(-1, 1, (B, n_steps, 4))
    for kw in (dict(noise_seed=5, scenario_id0=1000, noise_eps=0.002), dict(w=w)):
        cs.set_option("closed_loop_path", "auto")
        u1, y1, s1, i1, xf1 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps, want_x_final=True, **kw)
        cs.set_option("closed_loop_path", "tc")
        l0 = _launches()
        u2, y2, s2, i2, xf2 = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps, want_x_final=True, **kw)
        assert _launches() - l0 == 1
        assert int(s1.max()) == 0 and int(s2.max()) == 0 and (i1 == i2).all()
        eu, ey = _rel(u2.cpu().numpy(), u1.cpu().numpy()), _rel(y2.cpu().numpy(), y1.cpu().numpy())
        ex = _rel(xf2.cpu().numpy(), xf1.cpu().numpy())
        assert eu < 1e-5 and ey < 1e-5 and ex < 1e-5, (eu, ey, ex)
        assert eu > 1e-12                                          # (it really is the reduced-precision path)
    cs.set_option("closed_loop_path", "auto")


@pytest.mark.parametrize("c", [1.0, 0.3])
def test_convex_fused_path_vs_generic_and_oracle(c):
    """CONVEX slack bound inside the fused kernel (slack rows on the tensor cores, warp-cooperative ADMM for
    the violating loops) vs the generic thread-per-loop kernel (70 loops: several blocks, dead lanes) and vs
    the oracle's active-set solution on a sample."""
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    B, n_steps = 70, 37
    cs, _ = _set(u_d, y_d, 1, True, 4, c=c)
    r = np.random.default_rng(3)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    w = 0.002 * r.uniform(-1, 1, (B, n_steps, 2))
    l0 = _launches()
    u1, y1, s1, i1 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)        # auto = k_closed_loop_cvx
    assert _launches() - l0 == 1
    cs.set_option("closed_loop_path", "generic")
    u2, y2, s2, i2 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)
    cs.set_option("closed_loop_path", "fast")                  # hybrid kernel: slack rows on DMMA, one-loop-at-a-time warp ADMM
    u5, y5, s5, i5 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)
    cs.set_option("closed_loop_path", "cvx")
    u6, y6, s6, i6, x6 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w, want_x_final=True)
    cs.set_option("closed_loop_path", "auto")
    assert np.array_equal(u1.cpu().numpy(), u6.cpu().numpy())
    assert int(s1.max()) == 0 and int(s2.max()) == 0 and int(s5.max()) == 0
    assert int(i1.max()) > 10                                  # the box really binds somewhere
    assert (i1 == i2).all(), (i1 - i2).abs().max()
    assert (i5 == i2).all()
    assert _rel(u1.cpu().numpy(), u2.cpu().numpy()) < 1e-8 and _rel(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-8
    assert _rel(u5.cpu().numpy(), u2.cpu().numpy()) < 1e-8 and _rel(y5.cpu().numpy(), y2.cpu().numpy()) < 1e-8
    u1, y1 = u1.cpu().numpy(), y1.cpu().numpy()
    for b in (0, 33, 69):
        po = O.four_tank_plant()
        po.x = xs[b].copy()
        ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=4, slack_type=O.SLACK_CONVEX, c=c)
        ctrl.u_s, ctrl.y_s = us[b].reshape(-1, 1), ys[b].reshape(-1, 1)
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[b])
        assert _rel(u1[b], u_ref) < 1e-5 and _rel(y1[b], y_ref) < 1e-5, b
    # Philox noise through the fused path, large batch: equals the generic kernel on a slice
    Bb = 16384 + 40
    sc_x = np.tile(plant_o.x, (Bb, 1))
    args = (np.tile(u_d[-4:].reshape(1, -1), (Bb, 1)), np.tile(y_d[-4:].reshape(1, -1), (Bb, 1)),
            np.tile(prm["u_s"].T, (Bb, 1)), np.tile(prm["y_s"].T, (Bb, 1)))
    u3, y3, s3, i3 = cs.closed_loop(_plant(), sc_x, *args, 21, noise_seed=9, scenario_id0=5, noise_eps=0.002)
    cs.set_option("closed_loop_path", "generic")
    u4, y4, s4, i4 = cs.closed_loop(_plant(), sc_x[:64], *(a[:64] for a in args), 21, noise_seed=9, scenario_id0=5,
                                    noise_eps=0.002)
    cs.set_option("closed_loop_path", "auto")
    assert _rel(u3[:64].cpu().numpy(), u4.cpu().numpy()) < 1e-8 and (i3[:64] == i4).all()


def test_convex_screen_never_hides_a_violation():
    """The TF32 screen of k_closed_loop_cvx only ever SKIPS the exact FP64 slack check.  Bound placed 1e-4 (relative) below
    / 1e-9 above the largest slack of an unconstrained run - both far inside the screen's 4e-3 margin: in the first case the
    loops that reach it iterate (a violation of 1e-9 relative would converge in the first ADMM iteration and leave no
    trace in `iters`), in the second nobody does; the iteration counts equal those of the generic kernel, which has no screen."""
    import condensed_numpy as CN
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    B, n_steps, nblk = 70, 41, 11
    r = np.random.default_rng(12)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    w = 0.002 * r.uniform(-1, 1, (B, n_steps, 2))
    cs, _ = _set(u_d, y_d, 1, True, 4, c=1e6)
    u, y, st, it = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)
    assert int(st.max()) == 0 and int(it.max()) == nblk
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1e6,
                       CN.SLACK_CONVEX, CN.ROBUST, True)
    U = np.concatenate([up0.reshape(B, 4, 2), u.cpu().numpy()], axis=1)
    Y = np.concatenate([yp0.reshape(B, 4, 2), y.cpu().numpy()], axis=1)
    smax = np.zeros((B, nblk))
    for t in range(nblk):
        th = np.concatenate([U[:, 4 * t:4 * t + 4].reshape(B, -1), Y[:, 4 * t:4 * t + 4].reshape(B, -1), us, ys], axis=1)
        smax[:, t] = np.abs(th @ pl.Ks.T).max(axis=1)
    top = smax.max()
    for factor, binds in ((1.0 - 1e-4, True), (1.0 + 1e-9, False)):
        c = top * factor / prm["eps_max"]
        cs2, _ = _set(u_d, y_d, 1, True, 4, c=c)
        res = {}
        for path in ("cvx", "generic"):
            cs2.set_option("closed_loop_path", path)
            res[path] = cs2.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, w=w)
        assert (res["cvx"][3] == res["generic"][3]).all() and int(res["cvx"][2].max()) == 0
        extra = int(res["cvx"][3].sum()) - B * nblk
        assert (extra > 0) == binds, (factor, extra)
        assert _rel(res["cvx"][0].cpu().numpy(), res["generic"][0].cpu().numpy()) < 1e-8


def test_batched_reproduction_matches_reference_semantics(golden_repro):
    """reproduction.run_reproduction_batch (device data generation, shared generator order across the three
    schemes) vs the fixture produced by the reference's own functions for seed 4, plus figure-level facts."""
    from direct_data_driven_mpc_b200.reproduction import run_reproduction_batch
    g = golden_repro
    res = run_reproduction_batch([4, 0, 7, 11], t_sim=600)
    assert np.array_equal(res["u_d"][0].cpu().numpy(), g["u_d"])
    assert np.allclose(res["Y_n"][0], g["Y_n"], rtol=0, atol=1e-12) and np.allclose(res["x_start"][0], g["x_start"], atol=1e-12)
    for name in ("TEC", "TEC_N_STEP"):
        sc = res["schemes"][name]
        assert _rel(sc["u_sys"][0].cpu().numpy(), g[f"u_{name}"]) < 1e-5 and _rel(sc["y_sys"][0].cpu().numpy(), g[f"y_{name}"]) < 1e-5
        assert not bool(sc["diverged"].any()) and float(sc["final_error"].max()) < 0.03
        assert abs(float(sc["u_peak"][0]) - 8.66) < 0.01
    uc = res["schemes"]["UCON"]
    assert _rel(uc["u_sys"][0, :150].cpu().numpy(), g["u_UCON"][:150]) < 1e-5
    assert bool(uc["diverged"][0])                               # UCON diverges by design (reproduction.py:21-28)


def test_fused_four_tank_kernels_match_each_other():
    """The warp-specialised kernel (k_closed_loop_ws, the default for large batches: plant through the block map on the
    FP64 MMA pipe), the 8-lanes-per-loop kernel (k_closed_loop_perloop) and the generic thread-per-loop kernel vs the
    hybrid fused kernel (k_closed_loop_fast), including a partial last block, uploaded noise and a ragged batch.
    (The measured-and-dropped variants live in experiments/ and are compared by scripts/time_variants.py.)"""
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    B = 16384 + 70
    r = np.random.default_rng(5)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    for n_steps, kw in ((41, dict(noise_seed=3, scenario_id0=11, noise_eps=0.002)),
                        (401, dict(noise_seed=0, scenario_id0=0, noise_eps=0.002)),
                        (12, dict(w=0.002 * r.uniform(-1, 1, (B, 12, 2)))),
                        (3, dict(w=0.002 * r.uniform(-1, 1, (B, 3, 2))))):
        cs.set_option("closed_loop_path", "fast")
        u1, y1, s1, i1, x1 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, **kw)
        for path in ("ws", "auto", "perloop", "generic"):
            cs.set_option("closed_loop_path", path)
            launches0 = _launches()
            u2, y2, s2, i2, x2 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, **kw)
            assert _launches() - launches0 == 1, path
            assert int(s2.max()) == 0 and (i1 == i2).all(), path
            assert _rel(u2.cpu().numpy(), u1.cpu().numpy()) < 1e-9 and _rel(y2.cpu().numpy(), y1.cpu().numpy()) < 1e-9, path
            assert _rel(x2.cpu().numpy(), x1.cpu().numpy()) < 1e-9, path
        cs.set_option("closed_loop_path", "auto")


def test_step_major_trajectory_layout_is_the_same_numbers():
    """layout="step_major": k_closed_loop_ws stores (n_steps, B, m) - coalesced 1 KB runs per warp and step - and the API
    returns (B, n_steps, m) views of it.  Bit-identical to the loop-major result: even and odd batch sizes (an odd B
    breaks the 32-byte alignment of every other step), a ragged last CTA, a partial last block, uploaded noise; paths that
    cannot write the layout refuse."""
    import torch
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    r = np.random.default_rng(15)
    for B, n_steps, up in ((16384 + 70, 401, False), (8192 + 33, 42, False), (6401, 13, True)):
        xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
        us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
        ys = us @ _plant().equilibrium_gain().T
        up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
        kw = dict(w=0.002 * r.uniform(-1, 1, (B, n_steps, 2))) if up else dict(noise_seed=3, scenario_id0=11, noise_eps=0.002)
        cs.set_option("closed_loop_path", "ws")
        u1, y1, s1, i1, x1 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, **kw)
        l0 = _launches()
        u2, y2, s2, i2, x2 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, layout="step_major", **kw)
        assert _launches() - l0 == 1
        assert u2.shape == u1.shape and not u2.is_contiguous() and u2.permute(1, 0, 2).is_contiguous()
        assert torch.equal(u1, u2) and torch.equal(y1, y2) and torch.equal(x1, x2) and torch.equal(s1, s2) and torch.equal(i1, i2)
    cs.set_option("closed_loop_path", "perloop")
    with pytest.raises(NotImplementedError):
        cs.closed_loop(_plant(), xs, up0, yp0, us, ys, 13, layout="step_major", noise_seed=1, noise_eps=0.002)
    cs.set_option("closed_loop_path", "auto")
    u3, _, _, _ = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, 13, noise_seed=1, noise_eps=0.002)   # the option did not stick
    assert u3.is_contiguous()


@pytest.mark.parametrize("path", ["auto", "perloop", "fast"])
def test_warp_specialised_kernel_vs_oracle(path):
    """Large-batch paths (k_closed_loop_ws by default; the 8-lanes-per-loop and hybrid kernels when forced) against the
    literal-KKT oracle on whole 401-step loops."""
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    cs.set_option("closed_loop_path", path)
    B, n_steps = 16384 + 3, 401
    r = np.random.default_rng(8)
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
    ys = us @ _plant().equilibrium_gain().T
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
    u, y, st, it = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, noise_seed=21, scenario_id0=1000,
                                  noise_eps=0.002)
    assert int(st.max()) == 0 and int(it.min()) == 101 and int(it.max()) == 101
    u, y = u.cpu().numpy(), y.cpu().numpy()
    for b in (0, 1, 63, 64, 8191, B - 1):
        w = O.philox_noise(21, np.array([1000 + b]), n_steps, 2, 0.002)[0]
        po = O.four_tank_plant()
        po.x = xs[b].copy()
        ctrl = O.make_controller(prm, u_d, y_d, n_mpc_step=4)
        ctrl.u_s, ctrl.y_s = us[b].reshape(-1, 1), ys[b].reshape(-1, 1)
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w)
        assert _rel(u[b], u_ref) < 1e-9 and _rel(y[b], y_ref) < 1e-9, b


@pytest.mark.parametrize("chunks", [1, 3])
def test_closed_loop_host_matches_device_api(chunks):
    """Host-buffer entry point (the bench's `e2e` path): pinned or plain host arrays in, pinned trajectories out,
    chunked over two persistent streams; equals the device-resident call, also when called again with another size."""
    import torch
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    r = np.random.default_rng(11)
    for B, n_steps in ((1000, 41), (77, 12)):
        xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
        us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
        ys = us @ _plant().equilibrium_gain().T
        up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
        w = 0.002 * r.uniform(-1, 1, (B, n_steps, 2))
        for kw in (dict(w=w), dict(noise_seed=4, scenario_id0=123, noise_eps=0.002)):
            u1, y1, s1, _ = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, **kw)
            hu, hy, hs = cs.closed_loop_host(_plant(), torch.from_numpy(xs).pin_memory(), up0, yp0, us, ys, n_steps,
                                             chunks=chunks, **kw)
            assert hu.is_pinned() and tuple(hu.shape) == (B, n_steps, 2) and int(hs.max()) == 0
            assert _rel(hu.numpy(), u1.cpu().numpy()) < 1e-9 and _rel(hy.numpy(), y1.cpu().numpy()) < 1e-9


def test_closed_loop_host_chunks_on_the_gemm_path():
    """The chunked host API alternates its chunks between two streams.  On the GEMM path (large system, n_mpc_step = 8:
    not a compiled shape of the fused DMMA kernel) the per-call loop state used to live in a workspace cached in the
    set, which two concurrent chunks would have shared; it is now allocated per call in stream order."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B, n_steps, n_mpc = 1600, 20, 8
    sc = S.config4_batch(B, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    r = np.random.default_rng(2)
    x0 = sc["x0"] + 0.1 * r.normal(size=sc["x0"].shape)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    kw = dict(noise_seed=3, scenario_id0=50, noise_eps=0.002)
    l0 = _launches()
    u1, y1, s1, _ = cs.closed_loop(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps, **kw)
    assert _launches() - l0 > 3                                    # the per-iteration GEMMs, not a fused kernel
    for rep in range(3):
        hu, hy, hs = cs.closed_loop_host(pl, x0, sc["u_past0"], sc["y_past0"], u_s, y_s, n_steps, chunks=3, **kw)
        assert int(hs.max()) == 0
        assert _rel(hu.numpy(), u1.cpu().numpy()) < 1e-12 and _rel(hy.numpy(), y1.cpu().numpy()) < 1e-12, rep


def test_fused_kernels_are_run_to_run_deterministic():
    """Every fused kernel, run three times on the same inputs, must return bit-identical trajectories: the warp-
    specialised kernel rotates shared-memory buffers between warps and the config-4 kernel rotates a ring in place, so
    a missing barrier would show up as run-to-run differences (compute-sanitizer is not available on the GPU pool)."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    r = np.random.default_rng(2)
    B = 16384 + 9
    xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
    t = lambda a, b=B: np.tile(np.asarray(a).reshape(1, -1), (b, 1))
    runs = []
    cs, _ = _set(u_d, y_d)                                         # warp-specialised all-tensor-core kernel
    runs.append(lambda: cs.closed_loop(_plant(), xs, t(u_d[-4:]), t(y_d[-4:]), t(prm["u_s"]), t(prm["y_s"]), 45,
                                       noise_seed=7, noise_eps=0.002))
    cc, _ = _set(u_d, y_d, slack=1, c=0.3)                         # fused CONVEX kernel with active bounds
    runs.append(lambda: cc.closed_loop(_plant(), xs, t(u_d[-4:]), t(y_d[-4:]), t(prm["u_s"]), t(prm["y_s"]), 13,
                                       noise_seed=7, noise_eps=0.002))
    for n_mpc in (1, 20):                                          # config-4 fused DMMA kernel, both shapes
        sc = S.config4_batch(520, n_mpc_step=n_mpc)
        p4 = sc["params"]
        c4 = ControllerSet(20, 4, 4, sc["u_d"], sc["y_d"], 40, p4["Q"], p4["R"], p4["eps_max"], p4["lamb_alpha"],
                           p4["lamb_sigma"], p4["c"], 0, 1, n_mpc, True)
        runs.append(lambda c4=c4, sc=sc: c4.closed_loop(sc["plant"], sc["x0"], sc["u_past0"], sc["y_past0"], sc["u_s"],
                                                        sc["y_s"], 47, noise_seed=3, noise_eps=0.002))
    for k, fn in enumerate(runs):
        u0, y0, s0, i0 = fn()
        for _ in range(2):
            u1, y1, s1, i1 = fn()
            assert torch.equal(u0, u1) and torch.equal(y0, y1) and torch.equal(i0, i1), k


@pytest.mark.parametrize("name,ctype,slack,term,n_mpc,n_steps,tol", [
    ("ROBUST TEC n-step", 1, 0, True, 4, 401, 1e-6), ("ROBUST TEC 1-step", 1, 0, True, 1, 101, 1e-6),
    ("ROBUST UCON 1-step", 1, 0, False, 1, 101, 1e-6), ("ROBUST CONVEX n-step", 1, 1, True, 4, 201, 1e-5),
    ("NOMINAL 1-step", 0, 0, True, 1, 41, 1e-5)])
def test_config2_per_seed_variants_vs_oracle(name, ctype, slack, term, n_mpc, n_steps, tol):
    """BASELINE config 2 parity sample: loop b = the example script with `--seed b` - its own data, hence
    its own controller, and the measurement noise its generator draws next - for b < 32 and every controller
    variant, batched through ControllerSet(count=32) + ctrl_idx against the oracle."""
    from direct_data_driven_mpc_b200 import ControllerSet
    B = 32
    prm = O.four_tank_params()
    data = [O.example_scenario(s) for s in range(B)]              # (plant, params, rng, x0, u_d, y_d); rng continues
    ud, yd = np.stack([d[4] for d in data]), np.stack([d[5] for d in data])
    xs = np.stack([d[0].x for d in data])
    w = np.stack([0.002 * d[2].uniform(-1.0, 1.0, (n_steps, 2)) for d in data])      # controller_operation.py:263
    robust = ctype == 1
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"] if robust else None,
                       prm["lamb_alpha"] if robust else None, prm["lamb_sigma"] if robust else None, 1.0 if robust else None,
                       slack, ctype, n_mpc, term)
    assert (cs.statuses() == 0).all()
    u, y, st, it = cs.closed_loop(_plant(), xs, ud[:, -4:].reshape(B, -1), yd[:, -4:].reshape(B, -1),
                                  np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1)), n_steps, w=w,
                                  ctrl_idx=np.arange(B))
    u, y = u.cpu().numpy(), y.cpu().numpy()
    assert int(st.max()) == 0, name
    for b in range(B):
        ctrl = O.make_controller(prm, ud[b], yd[b], n_mpc_step=n_mpc, use_terminal=term, slack_type=slack, ctrl_type=ctype)
        u_ref, y_ref = O.closed_loop(data[b][0], ctrl, n_steps, w[b])
        assert _rel(u[b], u_ref) < tol and _rel(y[b], y_ref) < tol, (name, b, _rel(u[b], u_ref))


def test_config4_full_size_properties():
    """BASELINE config 4 at full size (16,384 loops x 401 steps, n_mpc_step = 20) through the fused DMMA kernel: every
    loop settles at its set-point, the batch equals the concatenation of two half-batch shards run with id offsets
    (sharding invariance), and a loop matches the oracle."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B, n_steps = 16384, 401
    sc = S.config4_batch(B, n_mpc_step=20)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], prm["c"], 0, 1, 20, True)
    r = np.random.default_rng(4)
    u_s = sc["u_s"] * r.uniform(0.8, 1.2, (B, 1))
    y_s = u_s @ pl.equilibrium_gain().T
    args = lambda lo, hi: (pl, sc["x0"][lo:hi], sc["u_past0"][lo:hi], sc["y_past0"][lo:hi], u_s[lo:hi], y_s[lo:hi], n_steps)
    u, y, st, it = cs.closed_loop(*args(0, B), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    assert int(st.max()) == 0 and int(it.min()) == 21 and int(it.max()) == 21
    assert float((y[:, -1] - torch.from_numpy(y_s).to(y.device)).abs().max()) < 0.02
    half = B // 2
    ua, ya, _, _ = cs.closed_loop(*args(0, half), noise_seed=0, scenario_id0=0, noise_eps=0.002)
    ub, yb, _, _ = cs.closed_loop(*args(half, B), noise_seed=0, scenario_id0=half, noise_eps=0.002)
    assert torch.equal(torch.cat([ua, ub]), u) and torch.equal(torch.cat([ya, yb]), y)
    b = B - 1
    w = O.philox_noise(0, np.array([b]), n_steps, 4, 0.002)[0]
    qp = O.OracleController(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], u_s[b].reshape(-1, 1), y_s[b].reshape(-1, 1),
                            prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST, 20, True,
                            check_pe=False)
    qp.set_past_input_output_data(sc["u_past0"][b].reshape(-1, 1), sc["y_past0"][b].reshape(-1, 1))
    po = O.Plant(pl.A, pl.B, pl.C, pl.D, pl.eps_max)
    po.x = sc["x0"][b].copy()
    u_ref, y_ref = O.closed_loop(po, qp, n_steps, w)
    assert _rel(u[b].cpu().numpy(), u_ref) < 1e-6 and _rel(y[b].cpu().numpy(), y_ref) < 1e-6


def test_set_lifecycle_returns_device_memory():
    """Create / use / destroy a controller set many times (batched solve, B = 1 staged host path with its mapped buffer
    and private stream, large fused closed loop): after ddmpc_trim_memory() the device's free memory is back where it
    started, i.e. nothing the library allocates outlives the set that owns it."""
    import torch
    from direct_data_driven_mpc_b200 import _lib
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    B = 20000
    xs = np.tile(plant_o.x, (B, 1))
    us, ys = np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1))
    up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))

    from direct_data_driven_mpc_b200 import (DataDrivenMPCType, DirectDataDrivenMPCController,
                                             SlackVarConstraintTypes)

    def cycle():
        ctrl = DirectDataDrivenMPCController(                            # B = 1: staged host path (solve #0 + one more)
            n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
            eps_max=0.002, lamb_alpha=prm["lamb_alpha"], lamb_sigma=1000, c=1.0,
            slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
            n_mpc_step=4, use_terminal_constraint=True)
        ctrl.update_and_solve_data_driven_mpc()
        del ctrl
        cs, _ = _set(u_d, y_d)
        cs.solve_batch(up0[:300], yp0[:300], us[:300], ys[:300])
        u, y, st, it = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, 41, noise_seed=1, noise_eps=0.002)
        assert int(st.max()) == 0
        del u, y, st, it
        cs.close()

    cycle()                                                              # first use: module load, pool creation
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    _lib.trim_memory()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        cycle()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    _lib.trim_memory()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 32 * 2**20, (free0, free1)


def test_nonfinite_inputs_flag_only_their_own_loops():
    """Status DDMPC_SOLVE_NONFINITE (3) for exactly the loops that were given a NaN / Inf initial state, set-point or
    window, on every closed-loop kernel: generic, 8-lanes-per-loop, hybrid and warp-specialised (four-tank), the fused
    FP64 tensor-core kernel and the generic one (config 4)."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    cs, _ = _set(u_d, y_d)
    bad = {5: "x0", 77: "u_s", 4100: "y_past", 16390: "x0"}
    for B, envs in ((300, ("generic", "auto", "fast", "perloop")),
                    (16384 + 9, ("auto", "fast", "perloop"))):
        xs = np.tile(plant_o.x, (B, 1))
        us, ys = np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1))
        up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
        expect = np.zeros(B, dtype=np.int32)
        for b, what in bad.items():
            if b >= B:
                continue
            expect[b] = 3
            if what == "x0":
                xs[b, 2] = np.nan
            elif what == "u_s":
                us[b, 0] = np.inf
            else:
                yp0[b, 3] = np.nan
        for env in envs:
            cs.set_option("closed_loop_path", env)
            u, y, st, it = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, 41, noise_seed=2, noise_eps=0.002)
            cs.set_option("closed_loop_path", "auto")
            assert np.array_equal(st.cpu().numpy(), expect), (B, env, np.nonzero(st.cpu().numpy() != expect)[0][:8])
            good = expect == 0
            assert np.isfinite(u.cpu().numpy()[good]).all() and np.isfinite(y.cpu().numpy()[good]).all()
    for n_mpc in (1, 20):
        B = 520
        sc = S.config4_batch(B, n_mpc_step=n_mpc)
        p4, pl = sc["params"], sc["plant"]
        c4 = ControllerSet(p4["n"], 4, 4, sc["u_d"], sc["y_d"], p4["L"], p4["Q"], p4["R"], p4["eps_max"],
                           p4["lamb_alpha"], p4["lamb_sigma"], p4["c"], 0, 1, n_mpc, True)
        xs, us, up0 = sc["x0"].copy(), sc["u_s"].copy(), sc["u_past0"].copy()
        expect = np.zeros(B, dtype=np.int32)
        xs[3, 7] = np.nan
        us[258, 1] = -np.inf
        up0[519, 11] = np.nan
        expect[[3, 258, 519]] = 3
        for env in ("auto", "generic") + (("tc",) if n_mpc == 20 else ()):
            c4.set_option("closed_loop_path", env)
            u, y, st, it = c4.closed_loop(pl, xs, up0, sc["y_past0"], us, sc["y_s"], 45, noise_seed=2, noise_eps=0.002)
            assert np.array_equal(st.cpu().numpy(), expect), (n_mpc, env, np.nonzero(st.cpu().numpy() != expect)[0][:8])


@pytest.mark.parametrize("n_mpc", [1, 20])
def test_dmma_kernel_cta_sizes_agree_bitwise(n_mpc):
    """k_closed_loop_dmma with CTAs of 1 (default), 2 and 4 warps (set option "dmma_warps"): the warps are independent, so the
    results must be identical bit for bit, ragged batch included."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    B, n_steps = 8 * 67 + 3, 45
    sc = S.config4_batch(B, n_mpc_step=n_mpc)
    prm, pl = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, n_mpc, True)
    outs = []
    for w in (1, 2, 4):
        cs.set_option("dmma_warps", w)
        u, y, st, it, xf = cs.closed_loop(pl, sc["x0"], sc["u_past0"], sc["y_past0"], sc["u_s"], sc["y_s"], n_steps,
                                          want_x_final=True, noise_seed=9, scenario_id0=77, noise_eps=0.002)
        assert int(st.max()) == 0
        outs.append((u.cpu().numpy(), y.cpu().numpy(), xf.cpu().numpy()))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o, outs[0]))


def test_randomised_batch_sizes_and_step_counts_agree_across_kernels():
    """Differential test: seeded random (batch size, step count, noise mode, slack) cases - batch sizes around the CTA and
    warp granularities of every kernel (1, 7, 31..33, 63..65, 6143..6145 loops), step counts from 1 - through every kernel
    that accepts the case, against the generic thread-per-loop kernel: same status and iteration counts, trajectories to 1e-9."""
    import torch
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    r = np.random.default_rng(2026)
    sizes = [1, 7, 31, 32, 33, 63, 64, 65, 127, 1000, 6143, 6144, 6145]
    sets = {0: _set(u_d, y_d)[0], 1: _set(u_d, y_d, 1, True, 4, c=0.5)[0]}
    for case in range(14):
        B = int(sizes[case % len(sizes)])
        n_steps = int(r.choice([1, 2, 3, 4, 5, 8, 9, 17, 40, 41]))
        slack = int(case % 3 == 2)
        cs = sets[slack]
        xs = np.tile(plant_o.x, (B, 1)) + 0.05 * r.normal(size=(B, 4))
        us = np.tile(prm["u_s"].T, (B, 1)) * r.uniform(0.7, 1.3, (B, 1))
        ys = us @ _plant().equilibrium_gain().T
        up0, yp0 = np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1))
        kw = dict(w=0.002 * r.uniform(-1, 1, (B, n_steps, 2))) if case % 2 else dict(noise_seed=case, scenario_id0=3 * case, noise_eps=0.002)
        cs.set_option("closed_loop_path", "generic")
        u0, y0, s0, i0, xf0 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, **kw)
        paths = ("auto", "cvx", "fast", "perloop") if slack else ("auto", "ws", "fast", "perloop")
        for path in paths:
            cs.set_option("closed_loop_path", path)
            u1, y1, s1, i1, xf1 = cs.closed_loop(_plant(), xs, up0, yp0, us, ys, n_steps, want_x_final=True, **kw)
            tag = (case, B, n_steps, slack, path)
            assert torch.equal(s0, s1) and torch.equal(i0, i1), tag
            assert _rel(u1.cpu().numpy(), u0.cpu().numpy()) < 1e-9 and _rel(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-9, tag
            assert _rel(xf1.cpu().numpy(), xf0.cpu().numpy()) < 1e-9, tag
        cs.set_option("closed_loop_path", "auto")
