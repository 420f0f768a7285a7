"""Runs the per-controller setup of the synthetic config-4 system (r = 480, c = 1941) a few times so
that ncu can capture the FP64 tensor-core GEMM (Gram matrix W = H H^T: the 'Hankel GEMM')."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
sc = S.config4_batch(8, n_mpc_step=20)
prm = sc["params"]
for i in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], prm["c"], 0, 1, 20, True)
    torch.cuda.synchronize(); print("setup s", time.perf_counter() - t, cs.info(0))
    del cs
