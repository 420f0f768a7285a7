"""Wall time of a single-controller setup (ddmpc_set_create, count = 1): config 3 (four-tank, r = 136) and config 4
(synthetic n = 20, m = p = 4: r = 480), steady state (third construction), plus the CONVEX variants."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S

dev = torch.device("cuda", 0)


def timed(mk, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter()
        cs = mk()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
        del cs
    return ts


sc3 = S.config3_batch(16, seed=0)
p3 = sc3["params"]
sc4 = S.config4_batch(16, n_mpc_step=20)
p4 = sc4["params"]
for name, mk in (
        ("config 3 ROBUST/NONE", lambda: ControllerSet(4, 2, 2, sc3["u_d"], sc3["y_d"], 30, p3["Q"], p3["R"], p3["eps_max"], p3["lamb_alpha"], p3["lamb_sigma"], 1.0, 0, 1, 4, True, device=dev)),
        ("config 3 ROBUST/CONVEX", lambda: ControllerSet(4, 2, 2, sc3["u_d"], sc3["y_d"], 30, p3["Q"], p3["R"], p3["eps_max"], p3["lamb_alpha"], p3["lamb_sigma"], 1.0, 1, 1, 4, True, device=dev)),
        ("config 3 NOMINAL", lambda: ControllerSet(4, 2, 2, sc3["u_d"], sc3["y_d"], 30, p3["Q"], p3["R"], controller_type=0, n_mpc_step=1, device=dev)),
        ("config 4 ROBUST/NONE", lambda: ControllerSet(20, 4, 4, sc4["u_d"], sc4["y_d"], 40, p4["Q"], p4["R"], p4["eps_max"], p4["lamb_alpha"], p4["lamb_sigma"], 1.0, 0, 1, 20, True, device=dev)),
        ("config 4 ROBUST/CONVEX", lambda: ControllerSet(20, 4, 4, sc4["u_d"], sc4["y_d"], 40, p4["Q"], p4["R"], p4["eps_max"], p4["lamb_alpha"], p4["lamb_sigma"], 1.0, 1, 1, 20, True, device=dev))):
    if len(sys.argv) > 1 and sys.argv[1] not in name:
        continue
    ts = timed(mk)
    print(f"{name:24s} setup ms: " + ", ".join(f"{t:.2f}" for t in ts), flush=True)
