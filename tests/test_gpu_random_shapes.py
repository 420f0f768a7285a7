"""GPU: shapes the reference's shipped configurations never exercise (odd channel counts, m != p, a plant with direct
feed-through, n_mpc_step neither 1 nor n, shared and per-loop controllers) - every controller variant on seeded random stable
plants, batched through the C ABI, against the literal-KKT oracle.  These run on the run-time-sized kernels (the fused
kernels are specialised for the BASELINE shapes), so this is the parity net under everything that is not a BASELINE shape.

Tolerance: 1e-5 relative on u and y (north_star); the equality-only variants are held to 1e-6."""
import numpy as np
import pytest

from oracle import ddmpc_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _random_plant(rng, nx, m, p, feedthrough, eps):
    G = rng.normal(size=(nx, nx))
    A = 0.85 * G / np.abs(np.linalg.eigvals(G)).max()
    Bm = rng.normal(size=(nx, m)) / np.sqrt(nx)
    Cm = rng.normal(size=(p, nx)) / np.sqrt(nx)
    Dm = 0.3 * rng.normal(size=(p, m)) if feedthrough else np.zeros((p, m))
    return O.Plant(A, Bm, Cm, Dm, eps)


# (name, n_x = n, m, p, L, n_mpc, ctrl_type, slack, terminal, feed-through, per-loop controllers, c, tolerance)
CASES = [
    ("robust 3x1x2 n-step", 3, 1, 2, 10, 3, O.ROBUST, O.SLACK_NONE, True, False, False, 1.0, 1e-6),
    ("robust 2x3x1 1-step, D != 0", 2, 3, 1, 9, 1, O.ROBUST, O.SLACK_NONE, True, True, False, 1.0, 1e-6),
    ("robust 5x2x3 2-step of 5, no terminal", 5, 2, 3, 14, 2, O.ROBUST, O.SLACK_NONE, False, False, False, 1.0, 1e-6),
    ("robust convex 3x2x2, per-loop controllers", 3, 2, 2, 10, 3, O.ROBUST, O.SLACK_CONVEX, True, False, True, 0.4, 1e-5),
    ("robust convex 4x1x3 1-step, D != 0", 4, 1, 3, 12, 1, O.ROBUST, O.SLACK_CONVEX, True, True, False, 0.5, 1e-5),
    ("nominal 3x2x1 noise-free", 3, 2, 1, 10, 3, O.NOMINAL, O.SLACK_NONE, True, False, False, None, 1e-5),
    ("nominal 2x1x1 1-step, per-loop controllers", 2, 1, 1, 8, 1, O.NOMINAL, O.SLACK_NONE, True, True, True, None, 1e-5),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_random_shape_closed_loops_vs_oracle(case):
    from direct_data_driven_mpc_b200 import ControllerSet, LTIPlant
    name, nx, m, p, L, n_mpc, ctype, slack, term, ft, per_loop, c, tol = case
    n = nx
    robust = ctype == O.ROBUST
    eps = 0.002 if robust else 0.0                                    # the nominal scheme assumes exact data
    seed = sum(map(ord, name))
    rng = np.random.default_rng(seed)
    plant = _random_plant(rng, nx, m, p, ft, eps)
    B, n_steps = 5, 4 * n_mpc + 1                                     # the last block is a partial one when n_mpc > 1
    N = (m + 1) * (L + 2 * n) + 30
    count = B if per_loop else 1
    ud, yd = [], []
    for _ in range(count):
        plant.x = rng.uniform(-1, 1, nx)
        u_d, y_d = O.generate_initial_input_output_data(plant, N, [-1, 1], rng)
        ud.append(u_d)
        yd.append(y_d)
    Q, R = 3.0 * np.eye(p * L), 1e-2 * np.eye(m * L)
    lam_a, lam_s = (0.1 / eps, 1000.0) if robust else (None, None)
    u_s = rng.uniform(0.3, 0.8, (B, m))
    y_s = u_s @ plant.gain().T
    x0 = rng.uniform(-0.5, 0.5, (B, nx))
    w = eps * rng.uniform(-1, 1, (B, n_steps, p))
    # a consistent past window per loop: n steps of the plant from x0 under random inputs
    up0, yp0, xs = np.zeros((B, n * m)), np.zeros((B, n * p)), np.zeros((B, nx))
    for b in range(B):
        plant.x = x0[b].copy()
        U = rng.uniform(-0.3, 0.3, (n, m))
        Y = plant.simulate(U, eps * rng.uniform(-1, 1, (n, p)), n)
        up0[b], yp0[b], xs[b] = U.reshape(-1), Y.reshape(-1), plant.x
    cs = ControllerSet(n, m, p, np.stack(ud) if per_loop else ud[0], np.stack(yd) if per_loop else yd[0], L, Q, R,
                       eps if robust else None, lam_a, lam_s, c, 1 if slack == O.SLACK_CONVEX else 0, 1 if robust else 0,
                       n_mpc, term)
    assert (cs.statuses() == 0).all(), name
    pl = LTIPlant(A=plant.A, B=plant.B, C=plant.C, D=plant.D, eps_max=eps)
    kw = dict(ctrl_idx=np.arange(B)) if per_loop else {}
    u, y, st, it = cs.closed_loop(pl, xs, up0, yp0, u_s, y_s, n_steps, w=w, **kw)
    u, y = u.cpu().numpy(), y.cpu().numpy()
    assert int(st.max()) <= 1, (name, st)                            # optimal / optimal_inaccurate
    if slack == O.SLACK_CONVEX:
        assert int(it.max()) > (n_steps + n_mpc - 1) // n_mpc, name   # the slack bound binds somewhere
    worst = 0.0
    for b in range(B):
        k = b if per_loop else 0
        ctrl = O.OracleController(n, m, p, ud[k], yd[k], L, Q, R, u_s[b].reshape(-1, 1), y_s[b].reshape(-1, 1), eps if robust else None,
                                  lam_a, lam_s, c, slack, ctype, n_mpc, term, check_pe=False)
        ctrl.set_past_input_output_data(up0[b].reshape(-1, 1), yp0[b].reshape(-1, 1))
        po = O.Plant(plant.A, plant.B, plant.C, plant.D, eps)
        po.x = xs[b].copy()
        u_ref, y_ref = O.closed_loop(po, ctrl, n_steps, w[b])
        assert _rel(u[b], u_ref) < tol and _rel(y[b], y_ref) < tol, (name, b, _rel(u[b], u_ref), _rel(y[b], y_ref))
        worst = max(worst, _rel(u[b], u_ref), _rel(y[b], y_ref))
    print(f"{name}: worst relative deviation from the oracle {worst:.2e} (tolerance {tol:.0e})")
