"""GPU: Hankel kernel (bit-exact), PE rank test and every setup-stage operator
against the NumPy mirror of the pipeline, through the C ABI."""
import hashlib

import numpy as np
import pytest

from oracle import ddmpc_oracle as O
import condensed_numpy as CN

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def test_hankel_bit_exact_golden(golden_hankel):
    from direct_data_driven_mpc_b200 import hankel_matrix
    assert np.array_equal(hankel_matrix(golden_hankel["doc_X"], 2), golden_hankel["doc_H"])
    for seed, N, nch, L in golden_hankel["cases"]:
        X = np.random.default_rng(int(seed)).normal(size=(int(N), int(nch)))
        H = hankel_matrix(X, int(L))
        assert H.shape == (L * nch, N - L + 1)
        assert sha(H) == str(golden_hankel[f"sha_{seed}"]), (seed, N, nch, L)


@pytest.mark.parametrize("N,nch,L", [(400, 2, 34), (400, 2, 38), (2000, 4, 60), (2000, 4, 80), (37, 3, 5), (9, 1, 9),
                                     (50, 5, 50), (129, 7, 2)])
def test_hankel_bit_exact_oracle(N, nch, L):
    from direct_data_driven_mpc_b200 import hankel_matrix
    X = np.random.default_rng(N + nch + L).uniform(-1, 1, (N, nch))
    assert np.array_equal(hankel_matrix(X, L), O.hankel_matrix(X, L))
    # non-contiguous input view
    Xb = np.random.default_rng(1).normal(size=(N, 2 * nch))[:, ::2]
    assert np.array_equal(hankel_matrix(Xb, L), O.hankel_matrix(Xb, L))


def test_hankel_device_pointer_entry():
    import torch
    from direct_data_driven_mpc_b200 import _lib
    X = torch.rand(300, 3, dtype=torch.float64, device="cuda")
    H = torch.empty(3 * 20, 281, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib.ddmpc_hankel(X.data_ptr(), 300, 3, 20, H.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(H.cpu().numpy(), O.hankel_matrix(X.cpu().numpy(), 20))


def test_pe_rank():
    from direct_data_driven_mpc_b200 import evaluate_persistent_excitation
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (400, 2))
    assert evaluate_persistent_excitation(X, 38) == (76, True)
    assert evaluate_persistent_excitation(X, 38) == O.evaluate_persistent_excitation(X, 38)
    # constant input: rank n_ch? no - rank 1 Hankel
    Xc = np.ones((200, 2))
    r, ok = evaluate_persistent_excitation(Xc, 10)
    assert (r, ok) == O.evaluate_persistent_excitation(Xc, 10) and not ok
    # sum of two sinusoids per channel: low-rank
    t = np.arange(300)[:, None]
    Xs = np.hstack([np.sin(0.3 * t) + np.sin(0.7 * t), np.cos(0.2 * t)])
    r, ok = evaluate_persistent_excitation(Xs, 12)
    assert (r, ok) == O.evaluate_persistent_excitation(Xs, 12) and not ok
    # odd sizes
    X3 = rng.normal(size=(90, 3))
    assert evaluate_persistent_excitation(X3, 7) == O.evaluate_persistent_excitation(X3, 7)
    # ill-conditioned but full rank (sigma_min / sigma_max ~ 1e-9): the Gram test cannot certify it, the second
    # stage (pivoted Gram-Schmidt on H with matrix_rank's cut) accepts it exactly as the reference does
    for scale in (1e-7, 1e-9, 1e-11):
        Xi = Xs + scale * rng.normal(size=Xs.shape)
        ref = O.evaluate_persistent_excitation(Xi, 12)
        assert ref == (24, True)
        assert evaluate_persistent_excitation(Xi, 12) == ref, scale
    # the same verdicts through the controller constructor (controller.py:285-296)
    from direct_data_driven_mpc_b200 import ControllerSet
    ud = np.hstack([np.sin(0.3 * t) + np.sin(0.7 * t), np.cos(0.2 * t)])[:120] + 1e-9 * rng.normal(size=(120, 2))
    yd = rng.normal(size=(120, 2))
    try:                                                           # order L + 2n = 12: rank 2 * 12 -> PE accepted;
        cs = ControllerSet(2, 2, 2, ud, yd, 8, np.eye(16), np.eye(16), 0.01, 1.0, 10.0, 1.0, 0, 1, 1, True)
        assert cs.info(0) == (24, 0)
    except ValueError as exc:                                      # data this ill-conditioned may still fail the Gram
        assert "positive definite" in str(exc)                     # factorisation of the robust setup - but not the PE test
        assert "persistently exciting" not in str(exc)


def _four_tank_set(slack, term, c=1.0, n_mpc=4, seed=0, ctrl=1):
    from direct_data_driven_mpc_b200 import ControllerSet
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(seed)
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, ctrl, n_mpc, term)
    pl = CN.build_plan(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, ctrl, term)
    return cs, pl, prm, u_d, y_d


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("slack,term", [(0, True), (0, False), (1, True), (1, False)])
def test_setup_operators_robust(slack, term):
    cs, pl, prm, u_d, y_d = _four_tank_set(slack, term, c=0.5)
    assert cs.info(0) == (76, 0)
    H = cs.get("H").reshape(pl.H.shape)
    assert np.array_equal(H, pl.H)                                        # bit-exact
    assert np.array_equal(cs.get("HLn_ud").reshape(68, 367), O.hankel_matrix(u_d, 34))
    assert np.array_equal(cs.get("HLn_yd").reshape(68, 367), O.hankel_matrix(y_d, 34))
    assert _relerr(cs.get("W").reshape(pl.W.shape), pl.W) < 1e-13
    assert _relerr(cs.get("Om").reshape(pl.Om.shape), pl.Om) < 1e-8
    assert _relerr(cs.get("X0").reshape(pl.X0.shape), pl.X0) < 1e-8
    assert _relerr(cs.get("Ku").reshape(pl.Ku.shape), pl.Ku) < 1e-8
    assert _relerr(cs.get("Z").reshape(pl.Z.shape), pl.Z) < 1e-8
    if slack == 1:
        assert _relerr(cs.get("rho2"), np.array([pl.rho2])) < 1e-8
        assert _relerr(cs.get("Ks").reshape(pl.Ks.shape), pl.Ks) < 1e-8
        assert _relerr(cs.get("Lam").reshape(pl.Lam.shape), pl.Lam) < 1e-8
        assert _relerr(cs.get("Phi").reshape(pl.Phi.shape), pl.Phi) < 1e-8
        assert _relerr(cs.get("Psi").reshape(pl.Psi.shape), pl.Psi) < 1e-8


def test_setup_general_weights_and_odd_dims():
    """dense (non-diagonal) Q, R and m != p, exercised against the NumPy mirror."""
    from direct_data_driven_mpc_b200 import ControllerSet
    rng = np.random.default_rng(3)
    n, m, p, L, N = 3, 3, 2, 7, 160
    A = np.diag([0.9, 0.8, 0.7]) + 0.05 * rng.normal(size=(3, 3))
    plant = O.Plant(A, rng.normal(size=(3, m)), rng.normal(size=(p, 3)), np.zeros((p, m)), 0.01)
    plant.x = rng.normal(size=3)
    u_d = rng.uniform(-1, 1, (N, m))
    y_d = plant.simulate(u_d, 0.01 * rng.uniform(-1, 1, (N, p)), N)
    Mq, Mr = rng.normal(size=(p * L, p * L)), rng.normal(size=(m * L, m * L))
    Q, R = Mq @ Mq.T + np.eye(p * L), 0.1 * (Mr @ Mr.T) + 0.01 * np.eye(m * L)
    for slack, term in [(1, True), (0, False)]:
        cs = ControllerSet(n, m, p, u_d, y_d, L, Q, R, 0.01, 5.0, 200.0, 0.7, slack, 1, 2, term)
        pl = CN.build_plan(n, m, p, u_d, y_d, L, Q, R, 0.01, 5.0, 200.0, 0.7, slack, 1, term)
        assert cs.info(0)[1] == 0
        # this random plant gives a much worse conditioned Gram matrix than the four-tank data,
        # so the two FP64 pipelines agree to ~cond*eps rather than 1e-8
        assert _relerr(cs.get("Ku").reshape(pl.Ku.shape), pl.Ku) < 1e-5
        assert _relerr(cs.get("Z").reshape(pl.Z.shape), pl.Z) < 1e-5
        if slack == 1:
            assert _relerr(cs.get("Phi").reshape(pl.Phi.shape), pl.Phi) < 1e-5
            assert _relerr(cs.get("Psi").reshape(pl.Psi.shape), pl.Psi) < 1e-5
        # ... and the solves agree with the literal-KKT oracle
        qp = O.OracleQP(n, m, p, u_d, y_d, L, Q, R, 0.01, 5.0, 200.0, 0.7, slack, 1, term)
        for k in (0, 50, 150):
            up, yp = u_d[k:k + n].reshape(1, -1), y_d[k:k + n].reshape(1, -1)
            us, ys = np.array([[0.3, -0.2, 0.1]]), np.array([[0.5, -0.4]])
            u, cost, status, iters = cs.solve_batch(up, yp, us, ys)
            so = qp.solve(up[0], yp[0], us[0], ys[0])
            assert int(status[0]) == 0
            assert _relerr(u.cpu().numpy()[0], so.optimal_u) < 1e-5
            assert abs(float(cost[0]) - so.cost) < 1e-5 * max(1.0, abs(so.cost))


def test_setup_batch_of_controllers_matches_singles():
    """count > 1 (per-seed data, per-controller weights) == the same controllers built one by one."""
    from direct_data_driven_mpc_b200 import ControllerSet
    data = [O.example_scenario(s) for s in range(3)]
    prm = data[0][1]
    ud = np.stack([d[4] for d in data])
    yd = np.stack([d[5] for d in data])
    la = np.array([50.0, 20.0, 80.0])
    ls = np.array([1000.0, 500.0, 3000.0])
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], la, ls, 1.0, 1, 1, 4, True)
    assert cs.count == 3 and list(cs.statuses()) == [0, 0, 0]
    for i in range(3):
        one = ControllerSet(4, 2, 2, ud[i], yd[i], 30, prm["Q"], prm["R"], prm["eps_max"], la[i], ls[i], 1.0, 1, 1,
                            4, True)
        for name in ("Ku", "Z", "Phi", "Psi", "Ks"):
            assert np.array_equal(cs.get(name, i), one.get(name, 0)), name


def test_constructor_errors_on_device_path():
    from direct_data_driven_mpc_b200 import (DataDrivenMPCType, DirectDataDrivenMPCController,
                                             SlackVarConstraintTypes)
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    kw = dict(n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
              eps_max=0.002, lamb_alpha=50.0, lamb_sigma=1000, c=1.0,
              slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST)
    with pytest.raises(NotImplementedError):                             # controller.py:664-670
        DirectDataDrivenMPCController(**{**kw, "slack_var_constraint_type": SlackVarConstraintTypes.NON_CONVEX})
    with pytest.raises(ValueError, match="not persistently exciting"):   # controller.py:285-296
        DirectDataDrivenMPCController(**{**kw, "u_d": np.ones_like(u_d)})
    with pytest.raises(ValueError, match="two times the estimated system order"):
        DirectDataDrivenMPCController(**{**kw, "L": 6, "Q": 3 * np.eye(12), "R": 1e-4 * np.eye(12)})
    with pytest.raises(ValueError, match="Q should be"):
        DirectDataDrivenMPCController(**{**kw, "Q": np.eye(3)})


def test_on_device_scenario_generation_matches_numpy_streams(golden_example):
    """ddmpc_generate_example_data: thread s replays np.random.default_rng(seed_s) through the example
    script's stages 1-3.  Uniform draws are bit-exact; simulated outputs agree to rounding."""
    from direct_data_driven_mpc_b200 import scenarios as S
    seeds = [0, 1, 4, 12345, 2 ** 40 + 7, 2 ** 63 + 11]
    ds = S.DeviceScenarios(seeds)
    u_d, y_d, x0, x_end = (t.cpu().numpy() for t in (ds.u_d, ds.y_d, ds.x0, ds.x_end))
    w = ds.uniform(401 * 2, -1.0, 1.0, 0.002).cpu().numpy().reshape(len(seeds), 401, 2)
    for i, seed in enumerate(seeds):
        rng, x0_h, ud_h, yd_h, xe_h = S.example_data(seed)
        assert np.array_equal(u_d[i], ud_h), seed                       # PCG64 draws: bit-exact
        assert np.allclose(y_d[i], yd_h, rtol=0, atol=1e-13)
        assert np.allclose(x0[i], x0_h, rtol=0, atol=1e-12) and np.allclose(x_end[i], xe_h, rtol=0, atol=1e-12)
        assert np.array_equal(w[i], 0.002 * rng.uniform(-1.0, 1.0, (401, 2)))   # draw 7 continues the stream
    g = golden_example                                                   # ... and equals the live reference (seed 0)
    assert np.array_equal(u_d[0], g["u_d"]) and np.allclose(y_d[0], g["y_d"], rtol=0, atol=1e-13)
    assert np.array_equal(w[0], g["w_sys"])


def test_per_seed_controllers_from_device_data():
    """config-2 pipeline without host loops: generate S data sets on the device, build S controllers, run S loops."""
    from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
    n_seeds, n_steps = 48, 29
    ds = S.DeviceScenarios(np.arange(n_seeds))
    prm = S.four_tank_controller_params()
    cs = ControllerSet(4, 2, 2, ds.u_d, ds.y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], 1.0, 0, 1, 4, True)
    assert list(cs.statuses()) == [0] * n_seeds
    w = ds.uniform(n_steps * 2, -1.0, 1.0, 0.002).reshape(n_seeds, n_steps, 2)
    u, y, st, it = cs.closed_loop(ds.plant, ds.x_end, ds.u_d[:, -4:].reshape(n_seeds, -1), ds.y_d[:, -4:].reshape(n_seeds, -1),
                                  np.tile(prm["u_s"].T, (n_seeds, 1)), np.tile(prm["y_s"].T, (n_seeds, 1)), n_steps,
                                  w=w, ctrl_idx=np.arange(n_seeds))
    u, y = u.cpu().numpy(), y.cpu().numpy()
    for seed in (0, 17, 47):                                             # == `--seed <seed>` of the example script
        u_ref, y_ref, _, _ = O.run_example(seed, n_steps - 1)
        assert _relerr(u[seed], u_ref) < 1e-6 and _relerr(y[seed], y_ref) < 1e-6
