"""Workloads for ncu captures of the round-2 kernels (run under `ncu --set full -k regex:<kernel> -c 1`).

    python scripts/profile_kernels.py ws | cvx | admm | perloop2 | perloop3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S

which = sys.argv[1]
dev = torch.device("cuda", 0)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
if which in ("ws", "cvx", "admm", "perloop3"):
    B = 4096 if which == "perloop3" else 65536
    sc = S.config3_batch(B, seed=0)
    prm, plant = sc["params"], sc["plant"]
    slack, c = (1, float(os.environ.get("CVX_C", "1.0"))) if which == "cvx" else ((1, 0.3) if which == "admm" else (0, 1.0))
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, 1, 4, True, device=dev)
    if which == "cvx" and "CVX_CTAS" in os.environ:
        cs.set_option("cvx_ctas_per_sm", int(os.environ["CVX_CTAS"]))
    if which == "admm":
        r = np.random.default_rng(0)
        ks = r.integers(0, 396, B)
        up = d(np.stack([sc["u_d"][k:k + 4].reshape(-1) for k in ks]))
        yp = d(np.stack([sc["y_d"][k:k + 4].reshape(-1) for k in ks]))
        for _ in range(2):
            out = cs.solve_batch(up, yp, d(sc["u_s"]), d(sc["y_s"]), tol=1e-8, want_cost=False)
        print("iters mean", float(out[3].float().mean()))
    else:
        bufs = (torch.empty(B, 401, 2, dtype=torch.float64, device=dev), torch.empty(B, 401, 2, dtype=torch.float64, device=dev))
        for _ in range(2):
            out = cs.closed_loop(plant, d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"]), 401,
                                 noise_seed=0, noise_eps=0.002, out=bufs)
        print("status", int(out[2].max()), "iters", int(out[3].sum()))
elif which == "perloop2":
    import bench
    r = bench.secondary_config2(dev)
    print({k: round(v["loop_ms"], 4) for k, v in r["variants"].items()})
torch.cuda.synchronize()
