import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, bench
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
dev = torch.device("cuda", 0)
B = 16384
sc = S.config4_batch(B, n_mpc_step=20)
prm, pl = sc["params"], sc["plant"]
cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, 20, True, device=dev)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cs.set_option("closed_loop_path", "tc")
for ns in (20, 100, 200, 400, 401):
    args = (pl, d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"]), ns)
    bufs = (torch.empty(B, ns, 4, dtype=torch.float64, device=dev), torch.empty(B, ns, 4, dtype=torch.float64, device=dev))
    run = lambda: cs.closed_loop(*args, noise_seed=0, noise_eps=0.002, out=bufs)
    run()
    print(ns, round(bench.median_ms(run, reps=10, warm=2), 4), "ms", flush=True)
