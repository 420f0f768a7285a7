import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from direct_data_driven_mpc_b200 import ControllerSet, _lib, scenarios as S
dev = torch.device("cuda", 0)
B = 65536
sc = S.config3_batch(B); prm, plant = sc["params"], sc["plant"]
cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0, 0, 1, 4, True)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x0, up0, yp0, us, ys = d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"])
u_sys = torch.empty(B, 401, 2, dtype=torch.float64, device=dev); y_sys = torch.empty_like(u_sys)
def step(): return cs.closed_loop(plant, x0, up0, yp0, us, ys, 401, noise_seed=0, scenario_id0=0, noise_eps=0.002, out=(u_sys, y_sys))
for _ in range(5): step()
torch.cuda.synchronize()
def run(n, tag=""):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t = time.perf_counter(); e0.record()
    for _ in range(n): r = step()
    e1.record(); th = time.perf_counter() - t; torch.cuda.synchronize()
    print(f"{tag} n={n}: gpu {e0.elapsed_time(e1)/n:.4f} ms/step, host enqueue {th/n*1e3:.4f} ms/step", flush=True)
for n in (1, 10, 50, 200, 500, 1000): run(n)
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = False
def poll():
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h); time.sleep(0.05)
th_ = threading.Thread(target=poll, daemon=True); th_.start()
for n in (50, 500): run(n, "nvml50ms")
stop = True; th_.join()
print("clock now", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "mem", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), "power", pynvml.nvmlDeviceGetPowerUsage(h))
# per-kernel events inside a long back-to-back run
evs = [torch.cuda.Event(enable_timing=True) for _ in range(301)]
evs[0].record()
for i in range(300):
    step(); evs[i+1].record()
torch.cuda.synchronize()
ts = np.array([evs[i].elapsed_time(evs[i+1]) for i in range(300)])
print("per-kernel in b2b run: first10", np.round(ts[:10],3), "median", np.median(ts), "p90", np.percentile(ts,90), "max", ts.max(), "mean", ts.mean())
