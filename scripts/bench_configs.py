"""Secondary measurements (BASELINE configs 2, 4, 5): setup rates and closed-loop throughput of the
generic paths.  Not the headline bench; prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S, _lib

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["2", "4", "5"]


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


if "2" in which:
    # config 2: 4096 closed loops over seeds, each loop with its own (u_d, y_d) => its own controller
    B, n_steps = int(os.environ.get("CFG2_B", 4096)), 401
    pl, prm = S.four_tank_plant(), S.four_tank_controller_params()
    # per-seed data on the device with the reference's NumPy streams (stages 1-3 of the example script for --seed b)
    torch.cuda.synchronize(); t = time.perf_counter()
    ds = S.DeviceScenarios(seeds=range(B))
    w_dev = ds.uniform(n_steps * 2, -1.0, 1.0, 0.002).reshape(B, n_steps, 2)       # controller_operation.py:263
    torch.cuda.synchronize(); t_data = time.perf_counter() - t
    ud, yd, x0 = ds.u_d, ds.y_d, ds.x_end
    for name, slack, term, nmpc, ctype in [("ROBUST TEC n-step", 0, True, 4, 1), ("ROBUST TEC 1-step", 0, True, 1, 1),
                                           ("ROBUST UCON 1-step", 0, False, 1, 1), ("ROBUST CONVEX n-step", 1, True, 4, 1),
                                           ("NOMINAL 1-step", 0, True, 1, 0)]:
        torch.cuda.synchronize(); t = time.perf_counter()
        cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                           1.0, slack, ctype, nmpc, term)
        torch.cuda.synchronize(); t_setup = time.perf_counter() - t
        ok = int((cs.statuses() == 0).sum())
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        args = (pl, x0, ud[:, -4:].reshape(B, -1).contiguous(), yd[:, -4:].reshape(B, -1).contiguous(),
                td(np.tile(prm["u_s"].T, (B, 1))), td(np.tile(prm["y_s"].T, (B, 1))), n_steps)
        kw = dict(w=w_dev, ctrl_idx=torch.arange(B, device=dev, dtype=torch.int32))
        dt = timed(lambda: cs.closed_loop(*args, **kw))
        u, y, st, it = cs.closed_loop(*args, **kw)
        print(json.dumps({"config": 2, "variant": name, "loops": B, "controllers_ok": ok, "setup_s": t_setup,
                          "controllers_per_s": B / t_setup, "loop_ms": dt * 1e3, "solves": int(it.sum()),
                          "solves_per_s": float(it.sum()) / dt, "status_max": int(st.max()), "data_gen_s": t_data,
                          "y_final_mean": y[:, -1].mean(0).tolist()}), flush=True)
        del cs

if "4" in which:
    B = int(os.environ.get("CFG4_B", 16384))
    for nmpc in (1, 20):
        sc = S.config4_batch(B, n_mpc_step=nmpc)
        prm, pl = sc["params"], sc["plant"]
        torch.cuda.synchronize(); t = time.perf_counter()
        cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                           prm["lamb_sigma"], prm["c"], 0, 1, nmpc, True)
        torch.cuda.synchronize(); t_setup = time.perf_counter() - t
        tdev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        args = (pl, tdev(sc["x0"]), tdev(sc["u_past0"]), tdev(sc["y_past0"]), tdev(sc["u_s"]), tdev(sc["y_s"]), 401)
        out = (torch.empty(B, 401, 4, dtype=torch.float64, device=dev), torch.empty(B, 401, 4, dtype=torch.float64, device=dev))
        kw = dict(noise_seed=0, noise_eps=0.002, out=out)
        dt = timed(lambda: cs.closed_loop(*args, **kw), n=5)
        u, y, st, it = cs.closed_loop(*args, **kw)
        err = float((y[:, -1] - torch.from_numpy(sc["y_s"]).to(dev)).abs().max())
        print(json.dumps({"config": 4, "n_mpc_step": nmpc, "loops": B, "pe_rank_status": cs.info(0), "setup_s": t_setup,
                          "loop_ms": dt * 1e3, "solves_per_s": float(it.sum()) / dt, "status_max": int(st.max()),
                          "final_tracking_error_max": err}), flush=True)
        del cs

if "5" in which:
    # config 5: lambda_alpha*eps x lambda_sigma x L sweep on four-tank: one controller per grid point
    pl, prm = S.four_tank_plant(), S.four_tank_controller_params()
    rng, x0, u_d, y_d, x_end = S.example_data(0)
    la = np.logspace(-3, 1, 16) / prm["eps_max"]
    ls = np.logspace(1, 5, 16)
    LA, LS = [g.reshape(-1) for g in np.meshgrid(la, ls, indexing="ij")]
    tot_ctrl, tot_t, tot_solves, tot_loop_t = 0, 0.0, 0, 0.0
    per_L = {}
    for L in range(8, 61, 4):
        Q, R = 3.0 * np.eye(2 * L), 1e-4 * np.eye(2 * L)
        mk = lambda: ControllerSet(4, 2, 2, u_d, y_d, L, Q, R, prm["eps_max"], LA, LS, 1.0, 0, 1, 4, True, count=LA.size)
        cs = mk(); del cs                     # warm the stream-ordered memory pool for this size (first-touch cudaMalloc)
        torch.cuda.synchronize(); t = time.perf_counter()
        cs = mk()
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        ok = int((cs.statuses() == 0).sum())
        nl = 64
        B = LA.size * nl
        idx = np.repeat(np.arange(LA.size), nl)
        args = (pl, np.tile(x_end, (B, 1)), np.tile(u_d[-4:].reshape(1, -1), (B, 1)), np.tile(y_d[-4:].reshape(1, -1), (B, 1)),
                np.tile(prm["u_s"].T, (B, 1)), np.tile(prm["y_s"].T, (B, 1)), 401)
        kw = dict(noise_seed=0, noise_eps=0.002, ctrl_idx=idx)
        lt = timed(lambda: cs.closed_loop(*args, **kw), n=1)
        u, y, st, it = cs.closed_loop(*args, **kw)
        per_L[L] = {"setup_s": dt, "ok": ok, "loop_ms": lt * 1e3, "status_max": int(st.max())}
        tot_ctrl += LA.size; tot_t += dt; tot_solves += int(it.sum()); tot_loop_t += lt
        del cs
    print(json.dumps({"config": 5, "controllers": tot_ctrl, "setup_s": tot_t, "controllers_per_s": tot_ctrl / tot_t,
                      "solves": tot_solves, "solves_per_s": tot_solves / tot_loop_t, "per_L": per_L}), flush=True)
