import sys, torch, json, numpy as np
sys.path.insert(0, '.')
import bench
from direct_data_driven_mpc_b200 import scenarios as S, _lib, ControllerSet
dev = torch.device('cuda', 0)
B = 65536
sc = S.config3_batch(B, seed=0)
prm, plant = sc["params"], sc["plant"]
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (plant, d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"]), 401)
bufs = (torch.empty(B, 401, 2, dtype=torch.float64, device=dev), torch.empty(B, 401, 2, dtype=torch.float64, device=dev))
only = sys.argv[1] if len(sys.argv) > 1 else None
for c in (1.0, 0.3, 100.0):
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], c, 1, 1, 4, True, device=dev)
    for ctas in (2, 3):
        if only and f"{c}:{ctas}" != only: continue
        cs.set_option("cvx_ctas_per_sm", ctas)
        run = lambda: cs.closed_loop(*args, noise_seed=0, noise_eps=0.002, out=bufs)
        _, _, st, it = run()
        ms = bench.median_ms(run, reps=5, warm=1)
        print(f"c = {c} ctas/SM {ctas}: {ms:.4f} ms, iters {int(it.sum())}, status {int(st.max())}", flush=True)
