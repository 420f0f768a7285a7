"""Per-seed controller setup for 4096 seeds (BASELINE config 2): the Gram products W = H H^T (the 'Hankel GEMM') run as
one batched FP64 tensor-core GEMM of 4096 x (136 x 136 x 367) = 55.6 GFLOP.  For ncu -k regex:k_gemm."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
B = int(os.environ.get("B", 4096))
prm = S.four_tank_controller_params()
ds = S.DeviceScenarios(seeds=range(B))
for i in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    cs = ControllerSet(4, 2, 2, ds.u_d, ds.y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0, 0, 1, 4, True)
    torch.cuda.synchronize(); print("setup s", time.perf_counter() - t, int((cs.statuses() == 0).sum()))
    del cs
