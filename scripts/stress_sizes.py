"""Stress run outside the BASELINE sizes: 262,144 loops x 401 steps (3.4 GB of trajectories) and 2,048 loops x 20,001
steps, NONE and CONVEX - no non-finite value, every loop settled, kernels agree on a sample, shard invariance holds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S

dev = torch.device("cuda", 0)
ok = True
for B, n_steps, slack in ((262144, 401, 0), (262144, 401, 1), (2048, 20001, 0), (2048, 20001, 1)):
    sc = S.config3_batch(min(B, 65536), seed=0)
    rep = -(-B // sc["x0"].shape[0])
    t = lambda k: np.tile(sc[k], (rep, 1))[:B]
    prm, plant = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], 1.0, slack, 1, 4, True, device=dev)
    args = (plant, t("x0"), t("u_past0"), t("y_past0"), t("u_s"), t("y_s"), n_steps)
    u, y, st, it = cs.closed_loop(*args, noise_seed=1, scenario_id0=7, noise_eps=0.002)
    fin = bool(torch.isfinite(u).all() and torch.isfinite(y).all())
    ys = torch.from_numpy(t("y_s")).to(dev)
    settled = float((y[:, -1] - ys).abs().max())
    # a slice through the generic kernel, with the matching global ids
    lo, n = B - 1000, 1000
    cs.set_option("closed_loop_path", "generic")
    a2 = (plant,) + tuple(x[lo:lo + n] for x in args[1:6]) + (n_steps,)
    u2, y2, st2, it2 = cs.closed_loop(*a2, noise_seed=1, scenario_id0=7 + lo, noise_eps=0.002)
    cs.set_option("closed_loop_path", "auto")
    d = float((u[lo:lo + n] - u2).abs().max())
    good = fin and settled < 0.02 and d < 1e-8 and int(st.max()) == 0 and bool((it[lo:lo + n] == it2).all())
    ok = ok and good
    print(f"{B} loops x {n_steps} steps, slack {slack}: finite {fin}, max |y_end - y_s| {settled:.4f}, vs generic kernel on the last "
          f"1000 loops {d:.2e}, iterations equal {bool((it[lo:lo + n] == it2).all())} -> {'ok' if good else 'FAIL'}", flush=True)
    del u, y, u2, y2, cs
    torch.cuda.empty_cache()
print("RESULT:", "PASS" if ok else "FAIL")
