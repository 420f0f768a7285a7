"""Config-3 workload with the CONVEX slack bound (class default of the reference): closed-loop throughput."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
B = int(os.environ.get("B", 65536))
sc = S.config3_batch(B); prm, pl = sc["params"], sc["plant"]
for c in (100.0, 1.0, 0.3):
    cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                       c, 1, 1, 4, True)
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = (pl, td(sc["x0"]), td(sc["u_past0"]), td(sc["y_past0"]), td(sc["u_s"]), td(sc["y_s"]), 401)
    kw = dict(noise_seed=0, noise_eps=0.002)
    u, y, st, it = cs.closed_loop(*args, **kw); torch.cuda.synchronize()
    t = time.perf_counter(); u, y, st, it = cs.closed_loop(*args, **kw); torch.cuda.synchronize(); dt = time.perf_counter() - t
    ys = torch.from_numpy(sc["y_s"]).to(y.device)
    print(json.dumps({"c": c, "loops": B, "ms": dt * 1e3, "solves_per_s": B * 101 / dt, "mean_iters_per_solve": float(it.double().mean()) / 101,
                      "status_max": int(st.max()), "track_err": float((y[:, -1] - ys).abs().max())}), flush=True)
