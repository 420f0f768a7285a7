"""GPU: batched QP solves (through the C ABI) against the literal-KKT oracle.

Tolerance (north_star): 1e-5 relative on u at solver tolerance 1e-8; the
equality-only variants are direct solves and are held to 1e-8."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O

pytestmark = pytest.mark.gpu


def _set(slack, term, c=1.0, n_mpc=1, seed=0, ctrl=1, data=None):
    from direct_data_driven_mpc_b200 import ControllerSet
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(seed)
    if data is not None:
        u_d, y_d = data
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, ctrl, n_mpc, term)
    qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                    c, slack, ctrl, term)
    return cs, qp, prm, u_d, y_d


def _thetas(u_d, y_d, prm, B, seed=0):
    r = np.random.default_rng(seed)
    ks = r.integers(0, u_d.shape[0] - 4, B)
    up = np.stack([u_d[k:k + 4].reshape(-1) for k in ks])
    yp = np.stack([y_d[k:k + 4].reshape(-1) for k in ks])
    us = prm["u_s"].reshape(1, -1) * r.uniform(0.5, 1.5, (B, 1))
    ys = prm["y_s"].reshape(1, -1) * r.uniform(0.5, 1.5, (B, 1))
    return up, yp, us, ys


@pytest.mark.parametrize("slack,term,c,tol", [(0, True, 1.0, 1e-8), (0, False, 1.0, 1e-8), (1, True, 1.0, 1e-5),
                                              (1, False, 0.3, 1e-5), (1, True, 0.2, 1e-5)])
def test_solve_batch_vs_oracle(slack, term, c, tol):
    cs, qp, prm, u_d, y_d = _set(slack, term, c)
    B = 24
    up, yp, us, ys = _thetas(u_d, y_d, prm, B)
    u, cost, status, iters = cs.solve_batch(up, yp, us, ys, tol=1e-8)
    u, cost, status, iters = u.cpu().numpy(), cost.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()
    assert (status == 0).all()
    n_active = 0
    for b in range(B):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        n_active += so.n_active
        rel = np.abs(u[b] - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max())
        assert rel < tol, (b, rel, so.n_active, iters[b])
        assert abs(cost[b] - so.cost) <= 1e-6 * max(1.0, abs(so.cost))
        if slack == 0:
            assert iters[b] == 1
    if slack == 1 and c < 1.0:
        assert n_active > 0 and iters.max() > 1          # the box path was really exercised


def test_solve_full_primal_vs_oracle():
    cs, qp, prm, u_d, y_d = _set(1, True, 0.3)
    up, yp, us, ys = _thetas(u_d, y_d, prm, 6, seed=3)
    ub, yb, sg, al = cs.solve_full_batch(up, yp, us, ys)
    for b in range(6):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        assert np.abs(ub[b].cpu().numpy() - so.ubar).max() < 1e-5 * max(1, np.abs(so.ubar).max())
        assert np.abs(yb[b].cpu().numpy() - so.ybar).max() < 1e-6
        assert np.abs(sg[b].cpu().numpy() - so.sigma).max() < 1e-7
        assert np.abs(al[b].cpu().numpy() - so.alpha).max() < 1e-6 * max(1, np.abs(so.alpha).max())
        assert np.abs(sg[b].cpu().numpy()[8:]).max() <= 0.3 * 0.002 * (1 + 1e-6)
    # a larger batch takes the GEMM route for alpha = (T x) W^-1 H (shared Hankel matrix x batch of iterates on the FP64
    # tensor cores): same numbers as the thread-per-element kernels of the small batch
    B = 100
    up, yp, us, ys = _thetas(u_d, y_d, prm, B, seed=4)
    ub, yb, sg, al = cs.solve_full_batch(up, yp, us, ys)
    ub2, yb2, sg2, al2 = cs.solve_full_batch(up[:40], yp[:40], us[:40], ys[:40])
    assert float((al[:40] - al2).abs().max()) <= 1e-10 * max(1.0, float(al2.abs().max()))
    H = np.vstack([cs.get("HLn_ud").reshape(68, -1), cs.get("HLn_yd").reshape(68, -1)])
    t = torch_cat(ub, yb + sg)
    assert np.abs(al.cpu().numpy() @ H.T - t).max() < 1e-7        # dynamics constraint [ubar; ybar + sigma] = H alpha


def torch_cat(a, b):
    import torch
    return torch.cat([a, b], dim=1).cpu().numpy()


def test_solve_batch_per_scenario_controller_index():
    from direct_data_driven_mpc_b200 import ControllerSet
    data = [O.example_scenario(s) for s in range(3)]
    prm = data[0][1]
    ud, yd = np.stack([d[4] for d in data]), np.stack([d[5] for d in data])
    cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                       1.0, 0, 1, 4, True)
    qps = [O.OracleQP(4, 2, 2, ud[i], yd[i], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                      prm["lamb_sigma"], 1.0, 0, 1, True) for i in range(3)]
    B = 9
    idx = np.arange(B) % 3
    up, yp, us, ys = _thetas(ud[0], yd[0], prm, B, seed=5)
    u, cost, status, iters = cs.solve_batch(up, yp, us, ys, ctrl_idx=idx)
    u = u.cpu().numpy()
    for b in range(B):
        so = qps[idx[b]].solve(up[b], yp[b], us[b], ys[b])
        assert np.abs(u[b] - so.optimal_u).max() < 1e-8 * max(1.0, np.abs(so.optimal_u).max())


def test_nominal_noisy_and_noise_free():
    from direct_data_driven_mpc_b200 import ControllerSet
    prm = O.four_tank_params()
    # noisy data: H has full row rank -> free coordinates sit on the set-point
    plant, _, rng, x0, u_d, y_d = O.example_scenario(0)
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], controller_type=0, n_mpc_step=1)
    up, yp, us, ys = _thetas(u_d, y_d, prm, 4)
    u, cost, status, iters = cs.solve_batch(up, yp, us, ys)
    assert (status.cpu().numpy() == 0).all()
    assert np.abs(u.cpu().numpy() - np.tile(us, (1, 30))).max() < 1e-7
    # noise-free data: rank(H) = m(L+n) + n_sys
    pl = O.four_tank_plant()
    pl.eps_max = 0.0
    r = np.random.default_rng(5)
    pl.x = r.uniform(-1, 1, 4)
    u_d = r.uniform(-1, 1, (400, 2))
    y_d = pl.simulate(u_d, np.zeros((400, 2)), 400)
    u_eq = np.array([1.0, 1.0])
    y_eq = pl.equilibrium_output_from_input(u_eq)
    for term in (True, False):
        cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], controller_type=0, use_terminal_constraint=term)
        qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], ctrl_type=O.NOMINAL, use_terminal=term)
        ks = [0, 100, 396]
        up = np.stack([u_d[k:k + 4].reshape(-1) for k in ks])
        yp = np.stack([y_d[k:k + 4].reshape(-1) for k in ks])
        us, ys = np.tile(u_eq, (3, 1)), np.tile(y_eq, (3, 1))
        u, cost, status, iters = cs.solve_batch(up, yp, us, ys)
        assert (status.cpu().numpy() == 0).all()
        for b in range(3):
            so = qp.solve(up[b], yp[b], us[b], ys[b])
            rel = np.abs(u[b].cpu().numpy() - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max())
            assert rel < 1e-5, (term, b, rel)
    # inconsistent window -> "infeasible" status, like the oracle
    yp_bad = yp + 0.05
    _, _, status, _ = cs.solve_batch(up, yp_bad, us, ys)
    cs_t = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], controller_type=0, use_terminal_constraint=True)
    _, _, status, _ = cs_t.solve_batch(up, yp_bad, us, ys)
    assert (status.cpu().numpy() == 2).all()


def test_empty_and_single():
    cs, qp, prm, u_d, y_d = _set(0, True)
    up, yp, us, ys = _thetas(u_d, y_d, prm, 1)
    u, cost, status, iters = cs.solve_batch(up, yp, us, ys)
    so = qp.solve(up[0], yp[0], us[0], ys[0])
    assert np.abs(u[0].cpu().numpy() - so.optimal_u).max() < 1e-8 * max(1.0, np.abs(so.optimal_u).max())
    u0, _, s0, _ = cs.solve_batch(up[:0], yp[:0], us[:0], ys[:0])
    assert u0.shape == (0, 60)


@pytest.mark.parametrize("slack,c", [(0, 1.0), (1, 0.3)])
def test_solve_batch_large_batch_kernel_matches_small_batch_kernel(slack, c):
    """B > 64 takes the thread-per-solve kernel, B <= 64 the CTA-per-solve kernel: same results; spot-check vs oracle."""
    cs, qp, prm, u_d, y_d = _set(slack, True, c)
    B = 200
    up, yp, us, ys = _thetas(u_d, y_d, prm, B, seed=11)
    u_big, cost_big, st_big, it_big = cs.solve_batch(up, yp, us, ys)
    parts = [cs.solve_batch(up[i:i + 50], yp[i:i + 50], us[i:i + 50], ys[i:i + 50]) for i in range(0, B, 50)]
    import torch
    u_small = torch.cat([p[0] for p in parts]); it_small = torch.cat([p[3] for p in parts])
    cost_small = torch.cat([p[1] for p in parts])
    assert int(st_big.max()) == 0
    assert (it_big == it_small).all()
    assert float((u_big - u_small).abs().max()) <= 1e-9 * max(1.0, float(u_small.abs().max()))
    assert float((cost_big - cost_small).abs().max()) <= 1e-9 * max(1.0, float(cost_small.abs().max()))
    for b in (0, 77, 199):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        assert np.abs(u_big[b].cpu().numpy() - so.optimal_u).max() < 1e-5 * max(1.0, np.abs(so.optimal_u).max())


@pytest.mark.parametrize("c,B", [(1.0, 1024), (0.3, 1500), (0.2, 2077)])
def test_tensor_core_admm_pipeline_matches_scalar_kernels_and_oracle(c, B):
    """Shared CONVEX controller, B >= 1024: the solve runs as k_gemm products around k_admm_dmma (box-row ADMM with
    Phi d on the FP64 tensor cores, csrc/cvx_loop.cu).  Same iterates as the thread-per-solve kernel (option
    "solve_path" = 1) up to FP64 summation order: equal iteration counts, u and cost to 1e-9; oracle on a sample;
    ragged batch (B not a multiple of the 32 problems of a CTA); a non-finite problem stays confined to itself."""
    cs, qp, prm, u_d, y_d = _set(1, True, c, n_mpc=4)
    up, yp, us, ys = _thetas(u_d, y_d, prm, B, seed=5)
    us[7, 0] = np.nan                                              # one poisoned problem
    l0 = _launches()
    u1, c1, s1, i1 = cs.solve_batch(up, yp, us, ys, tol=1e-8)
    assert _launches() - l0 == 8                                   # pack, 2 GEMM, ADMM, GEMM, 2 GEMM + cost rows
    cs.set_option("solve_path", 1)
    u2, c2, s2, i2 = cs.solve_batch(up, yp, us, ys, tol=1e-8)
    cs.set_option("solve_path", 0)
    s1n, s2n = s1.cpu().numpy(), s2.cpu().numpy()
    assert s1n[7] == 3 and s2n[7] == 3 and (np.delete(s1n, 7) == 0).all() and (np.delete(s2n, 7) == 0).all()
    good = np.ones(B, dtype=bool); good[7] = False
    i1n, i2n = i1.cpu().numpy()[good], i2.cpu().numpy()[good]
    assert (i1n > 1).sum() > B // 20                               # the box binds in a good share of the problems
    assert np.abs(i1n - i2n).max() <= 1 and (i1n != i2n).mean() < 0.01   # (a residual may cross the threshold one iteration apart)
    u1n, u2n = u1.cpu().numpy()[good], u2.cpu().numpy()[good]
    assert np.abs(u1n - u2n).max() <= 1e-8 * max(1.0, np.abs(u2n).max())
    c1n, c2n = c1.cpu().numpy()[good], c2.cpu().numpy()[good]
    assert np.abs(c1n - c2n).max() <= 1e-8 * max(1.0, np.abs(c2n).max())
    for b in (0, 1, 31, 32, B // 2, B - 1):
        so = qp.solve(up[b], yp[b], us[b], ys[b])
        assert np.abs(u1[b].cpu().numpy() - so.optimal_u).max() < 1e-5 * max(1.0, np.abs(so.optimal_u).max()), b
        assert abs(float(c1[b]) - so.cost) <= 1e-6 * max(1.0, abs(so.cost))
    # full primal through the same pipeline (sigma within its bound, dynamics constraint holds)
    ub, yb, sg, al = cs.solve_full_batch(up[good][:1100], yp[good][:1100], us[good][:1100], ys[good][:1100])
    assert float(sg[:, 8:].abs().max()) <= c * 0.002 * (1 + 1e-6)
    so = qp.solve(up[good][3], yp[good][3], us[good][3], ys[good][3])
    assert np.abs(sg[3].cpu().numpy() - so.sigma.reshape(-1)).max() < 1e-7


def _launches():
    from direct_data_driven_mpc_b200 import _lib
    return _lib.kernel_launches()


@pytest.mark.parametrize("L", [8, 20, 60])
def test_config5_lambda_horizon_grid_vs_oracle(L):
    """BASELINE config 5 parity sample: corners of the lambda_alpha*eps x lambda_sigma grid at three horizons, one
    controller per grid point built by one ControllerSet(count=...) call, every controller against its own oracle."""
    from direct_data_driven_mpc_b200 import ControllerSet
    plant, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    la = np.array([1e-3, 1e-3, 1e1, 1e1, 0.1]) / prm["eps_max"]
    ls = np.array([1e1, 1e5, 1e1, 1e5, 1e3])
    Q, R = 3.0 * np.eye(2 * L), 1e-4 * np.eye(2 * L)
    cs = ControllerSet(4, 2, 2, u_d, y_d, L, Q, R, prm["eps_max"], la, ls, 1.0, 0, 1, 4, True, count=la.size)
    assert (cs.statuses() == 0).all()
    r = np.random.default_rng(L)
    ks = r.integers(0, 396, la.size)
    up = np.stack([u_d[k:k + 4].reshape(-1) for k in ks])
    yp = np.stack([y_d[k:k + 4].reshape(-1) for k in ks])
    us, ys = np.tile(prm["u_s"].T, (la.size, 1)), np.tile(prm["y_s"].T, (la.size, 1))
    u, cost, st, _ = cs.solve_batch(up, yp, us, ys, ctrl_idx=np.arange(la.size))
    assert int(st.max()) == 0
    for c in range(la.size):
        qp = O.OracleQP(4, 2, 2, u_d, y_d, L, Q, R, prm["eps_max"], la[c], ls[c], 1.0, O.SLACK_NONE, O.ROBUST, True)
        so = qp.solve(up[c], yp[c], us[c], ys[c])
        rel = np.abs(u[c].cpu().numpy() - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max())
        assert rel < 1e-7, (L, c, rel)
        assert abs(float(cost[c]) - so.cost) <= 1e-6 * max(1.0, abs(so.cost))


@pytest.mark.parametrize("N,slack,c,mode", [(113, 0, 1.0, "short"), (150, 0, 1.0, "short"), (150, 1, 0.3, "short"),
                                            (400, 0, 1.0, "exact"), (400, 0, 1.0, "eps0")])
def test_short_data_robust_setup_vs_oracle(N, slack, c, mode):
    """Fewer Hankel columns than rows (N - L - n + 1 < (L + n)(m + p)): the reference accepts N down to N_min = 113 for the
    four-tank configuration (controller.py:275-283), where the Gram matrix W of the stacked Hankel matrix is singular.
    The setup then works with the pseudo-inverse of W and keeps t = [ubar; ybar + sigma] in range(H) through a Schur
    complement (setup.cu build_robust): optimal inputs, cost and the primal alpha against the literal-KKT oracle, for a
    grid of three (lambda_alpha, lambda_sigma) controllers that share the data, then a closed loop.
    "exact": enough columns, but noise-free data under the ROBUST controller - W is singular all the same (rank
    m (L + n) + n_x = 72 of 136), its Cholesky factorisation fails and the same path takes over.
    "eps0": the noise-free configuration the reference's YAML loader anticipates (controller_creation.py:129-136:
    eps_max = 0, lamb_alpha = 1000, so alpha carries no weight at all)."""
    from direct_data_driven_mpc_b200 import ControllerSet, LTIPlant
    prm = O.four_tank_params()
    eps = 0.0 if mode == "eps0" else prm["eps_max"]
    plant = O.four_tank_plant() if mode == "short" else O.Plant(**{**O.FOUR_TANK, "eps_max": 0.0})
    rng = np.random.default_rng(1)
    plant.x = rng.uniform(-1, 1, 4)
    u_d, y_d = O.generate_initial_input_output_data(plant, N, [-1, 1], rng)
    assert mode != "short" or N - 34 + 1 < 136
    la = np.array([1000.0, 1000.0, 1000.0]) if mode == "eps0" else np.array([50.0, 5.0, 500.0])
    ls = np.array([1000.0, 100.0, 1e4])
    cs = ControllerSet(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], eps, la, ls, c, slack, 1, 4, True, count=3)
    assert (cs.statuses() == 0).all()
    r = np.random.default_rng(N)
    ks = r.integers(0, N - 4, 3)
    up = np.stack([u_d[k:k + 4].reshape(-1) for k in ks])
    yp = np.stack([y_d[k:k + 4].reshape(-1) for k in ks])
    us, ys = np.tile(prm["u_s"].T, (3, 1)), np.tile(prm["y_s"].T, (3, 1))
    u, cost, st, it = cs.solve_batch(up, yp, us, ys, ctrl_idx=np.arange(3), tol=1e-9)
    assert int(st.max()) == 0
    if slack:
        assert int(it.max()) > 1                                   # the slack bound binds
    for k in range(3):
        qp = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], eps, la[k], ls[k], c, slack, O.ROBUST, True)
        so = qp.solve(up[k], yp[k], us[k], ys[k])
        rel = np.abs(u[k].cpu().numpy() - so.optimal_u).max() / max(1.0, np.abs(so.optimal_u).max())
        assert rel < (1e-5 if slack else 1e-7), (N, k, rel)
        assert abs(float(cost[k]) - so.cost) <= (1e-5 if slack else 1e-6) * max(1.0, abs(so.cost)), (N, k)
    # full primal: alpha = H^T W^+ t is the minimum-norm alpha - the unique one whenever alpha carries weight
    ub, yb, sg, al = cs.solve_full_batch(up, yp, us, ys, ctrl_idx=np.arange(3), tol=1e-9)
    so = O.OracleQP(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], eps, la[0], ls[0], c, slack, O.ROBUST, True).solve(
        up[0], yp[0], us[0], ys[0])
    tolp = 1e-4 if slack else 1e-6
    if mode != "eps0":
        assert np.abs(al[0].cpu().numpy() - so.alpha.reshape(-1)).max() < tolp * max(1.0, np.abs(so.alpha).max())
    assert np.abs(sg[0].cpu().numpy() - so.sigma.reshape(-1)).max() < tolp * max(1.0, np.abs(so.sigma).max())
    assert np.abs(yb[0].cpu().numpy() - so.ybar.reshape(-1)).max() < tolp * max(1.0, np.abs(so.ybar).max())
    # closed loop with the first controller
    n_steps = 41
    w = eps * rng.uniform(-1, 1, (2, n_steps, 2))
    xs = np.stack([plant.x, plant.x + 0.05])
    pl = LTIPlant(**{k: O.FOUR_TANK[k] for k in "ABCD"}, eps_max=eps)
    uu, yy, st, _ = cs.closed_loop(pl, xs, np.tile(u_d[-4:].reshape(1, -1), (2, 1)), np.tile(y_d[-4:].reshape(1, -1), (2, 1)),
                                   us[:2], ys[:2], n_steps, w=w, ctrl_idx=np.zeros(2, dtype=np.int32))
    assert int(st.max()) == 0
    for b in range(2):
        ctrl = O.make_controller(prm, u_d, y_d, slack_type=slack, c=c, eps_max=eps, lamb_alpha=la[0])
        po = O.four_tank_plant()
        po.x = xs[b].copy()
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[b])
        e = np.abs(uu[b].cpu().numpy() - u_ref).max() / max(1.0, np.abs(u_ref).max())
        assert e < 1e-5, (N, b, e)
