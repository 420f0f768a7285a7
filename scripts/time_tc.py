"""Config 4 (n_mpc_step = 20, 16,384 loops x 401 steps): the FP64 tensor-core kernel against the opt-in tcgen05 / TMEM
kernel (TF32x3 arithmetic), time per pass and the difference between their trajectories."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
sc = S.config4_batch(B, n_mpc_step=20)
prm, pl = sc["params"], sc["plant"]
cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                   prm["lamb_sigma"], prm["c"], 0, 1, 20, True, device=dev)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (pl, d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"]), 401)
res = {}
for path in ("auto", "tc", "tc2", "tc1"):
    cs.set_option("closed_loop_path", path[:2] if path != "auto" else path)
    cs.set_option("tc_passes", int(path[2:]) if path.startswith("tc") and len(path) > 2 else 3)
    bufs = (torch.empty(B, 401, 4, dtype=torch.float64, device=dev), torch.empty(B, 401, 4, dtype=torch.float64, device=dev))
    run = lambda: cs.closed_loop(*args, noise_seed=0, noise_eps=0.002, out=bufs)
    _, _, st, it = run()
    ms = bench.median_ms(run, reps=10, warm=2)
    res[path] = (bufs[0].clone(), bufs[1].clone())
    print(f"{path:5s}: {ms:.4f} ms per pass, {int(it.sum()) / (ms * 1e-3):.3e} solves/s, status {int(st.max())}", flush=True)
for k in ("tc2", "tc1"):
    print(k, "max rel diff u", float((res[k][0] - res["auto"][0]).abs().max() / res["auto"][0].abs().max()))
eu = float((res["tc"][0] - res["auto"][0]).abs().max() / res["auto"][0].abs().max())
ey = float((res["tc"][1] - res["auto"][1]).abs().max() / res["auto"][1].abs().max())
print(f"tcgen05 (TF32x3) vs FP64 kernel over 401 steps: max rel diff u {eu:.2e}, y {ey:.2e}")
