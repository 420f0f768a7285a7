"""Wall time of batched controller setup (config-5 style: 256 controllers, one L) vs its kernel time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S, _lib
pl, prm = S.four_tank_plant(), S.four_tank_controller_params()
rng, x0, u_d, y_d, x_end = S.example_data(0)
la = np.logspace(-3, 1, 16) / prm["eps_max"]; ls = np.logspace(1, 5, 16)
LA, LS = [g.reshape(-1) for g in np.meshgrid(la, ls, indexing="ij")]
for L in (32, 32, 32, 60, 60, 8, 8):
    Q, R = 3.0 * np.eye(2 * L), 1e-4 * np.eye(2 * L)
    torch.cuda.synchronize(); l0 = _lib.kernel_launches(); t = time.perf_counter()
    cs = ControllerSet(4, 2, 2, u_d, y_d, L, Q, R, prm["eps_max"], LA, LS, 1.0, 0, 1, 4, True, count=LA.size)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"L={L} setup {dt*1e3:.1f} ms, {_lib.kernel_launches()-l0} launches, ok {(cs.statuses()==0).sum()}")
    t = time.perf_counter(); del cs; torch.cuda.synchronize(); print(f"   destroy {(time.perf_counter()-t)*1e3:.1f} ms")
