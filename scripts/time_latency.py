"""Where the B = 1 step latency goes: the solve call (C ABI, zero-copy staging) vs the host-side window bookkeeping."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import ddmpc_oracle as O
from direct_data_driven_mpc_b200 import DirectDataDrivenMPCController, DataDrivenMPCType, SlackVarConstraintTypes

plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
ctrl = DirectDataDrivenMPCController(n=4, m=2, p=2, u_d=u_d, y_d=y_d, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
                                     eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=1.0,
                                     slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
                                     n_mpc_step=4, use_terminal_constraint=True)
def med(f, n=2000):
    for _ in range(50): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
print("update_and_solve          %.1f us" % med(ctrl.update_and_solve_data_driven_mpc))
u = ctrl.get_optimal_control_input_at_step(n_step=0)
print("get_optimal_input_at_step %.1f us" % med(lambda: ctrl.get_optimal_control_input_at_step(n_step=1)))
y = np.zeros(2)
print("store_measurement         %.1f us" % med(lambda: ctrl.store_input_output_measurement(u.reshape(-1, 1), y.reshape(-1, 1))))
pl = bench._HostPlant(type("P", (), dict(A=O.FOUR_TANK["A"], B=O.FOUR_TANK["B"], C=O.FOUR_TANK["C"], D=O.FOUR_TANK["D"])), np.zeros(4))
w = np.zeros(2)
print("plant step (numpy)        %.1f us" % med(lambda: pl.simulate_step(u, w)))
