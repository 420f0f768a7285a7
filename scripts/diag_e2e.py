import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
dev = torch.device("cuda", 0)
B = 65536
sc = S.config3_batch(B); prm, plant = sc["params"], sc["plant"]
cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0, 0, 1, 4, True)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hx0, hup, hyp, hus, hys = pin(sc["x0"]), pin(sc["u_past0"]), pin(sc["y_past0"]), pin(sc["u_s"]), pin(sc["y_s"])
hu = torch.empty(B, 401, 2, dtype=torch.float64, pin_memory=True); hy = torch.empty_like(hu).pin_memory()
print("pinned?", hu.is_pinned(), hy.is_pinned())
for chunks in (1, 2, 4, 8, 16):
    ts = []
    for i in range(6):
        torch.cuda.synchronize(); t = time.perf_counter()
        cs.closed_loop_host(plant, hx0, hup, hyp, hus, hys, 401, noise_seed=0, noise_eps=0.002, out=(hu, hy), chunks=chunks)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print("chunks", chunks, np.round(ts, 2))
# raw D2H of the two arrays
u = torch.empty(B, 401, 2, dtype=torch.float64, device=dev); y = torch.empty_like(u)
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); hu.copy_(u, non_blocking=True); hy.copy_(y, non_blocking=True); torch.cuda.synchronize()
    print("raw D2H 841MB ms", (time.perf_counter() - t) * 1e3)
os.system("nvidia-smi topo -m 2>/dev/null | head -5; lscpu | grep -i -E 'numa|socket|model name' | head -8")
