"""Where the B = 1 step latency goes (config 1 through the drop-in class): raw C-ABI call, class solve, plant steps,
window updates.  Run on the GPU box:  python scripts/diag_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from direct_data_driven_mpc_b200 import _lib, scenarios as S
from direct_data_driven_mpc_b200.controller import DirectDataDrivenMPCController, DataDrivenMPCType, SlackVarConstraintTypes

sc = S.config3_batch(1, seed=0)
prm, pl = sc["params"], sc["plant"]
ctrl = DirectDataDrivenMPCController(
    n=prm["n"], m=2, p=2, u_d=sc["u_d"], y_d=sc["y_d"], L=prm["L"], Q=prm["Q"], R=prm["R"],
    u_s=np.asarray(sc["u_s"][0]).reshape(-1, 1), y_s=np.asarray(sc["y_s"][0]).reshape(-1, 1), eps_max=prm["eps_max"],
    lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
    slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST, n_mpc_step=4)
plant = bench._HostPlant(pl, sc["x0"][0])
w = np.zeros(2)


def timeit(f, n=2000):
    for _ in range(200):
        f()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t) * 1e6)
    return np.percentile(ts, 50), np.percentile(ts, 90)


arr, ptr = ctrl._solve_buffers()
raw = lambda: _lib.lib.ddmpc_solve_batch_host(ctrl._set, 1, None, ptr["up"], ptr["yp"], ptr["us"], ptr["ys"], ctrl._solve_tol,
                                              ctrl._solve_max_iter, ptr["out"], ptr["cost"], ptr["status"], ptr["iters"])
print("raw ddmpc_solve_batch_host (B = 1)     p50 %.1f us  p90 %.1f us" % timeit(raw))
print("ctrl.solve_mpc_problem()               p50 %.1f us  p90 %.1f us" % timeit(ctrl.solve_mpc_problem))
print("ctrl.update_and_solve_data_driven_mpc  p50 %.1f us  p90 %.1f us" % timeit(ctrl.update_and_solve_data_driven_mpc))
u = ctrl.get_optimal_control_input_at_step(n_step=0)
print("get_optimal_control_input_at_step      p50 %.1f us  p90 %.1f us" % timeit(lambda: ctrl.get_optimal_control_input_at_step(n_step=1)))
print("plant.simulate_step                    p50 %.1f us  p90 %.1f us" % timeit(lambda: plant.simulate_step(u, w)))
y = plant.simulate_step(u, w)
uu, yy = u.reshape(-1, 1), y.reshape(-1, 1)
print("store_input_output_measurement         p50 %.1f us  p90 %.1f us" % timeit(lambda: ctrl.store_input_output_measurement(uu, yy)))
