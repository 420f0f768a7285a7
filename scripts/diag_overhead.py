import os, sys, time, subprocess, threading
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
def run(tag, n=100):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): step()
    th=time.perf_counter()-t
    torch.cuda.synchronize(); tt=time.perf_counter()-t
    print(f"{tag}: host enqueue {th/n*1e3:.3f} ms/call, total {tt/n*1e3:.3f} ms/call", flush=True)
run("no sampler")
p = subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm,clocks.max.sm,power.draw","--format=csv,noheader,nounits","-lms","100","-i","0"], stdout=subprocess.DEVNULL)
time.sleep(0.5); run("nvidia-smi -lms 100"); p.terminate()
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop=False; rows=[]
def poll():
    while not stop:
        rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
th_=threading.Thread(target=poll, daemon=True); th_.start(); time.sleep(0.2)
run("pynvml 20ms"); stop=True; th_.join(); print(len(rows), rows[-1], pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
# e2e pieces
hu = torch.empty(B, 401, 2, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize(); t=time.perf_counter(); hu.copy_(u_sys, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print(f"D2H 420MB pinned: {dt*1e3:.2f} ms = {hu.numel()*8/dt/1e9:.1f} GB/s")
t=time.perf_counter(); hu.copy_(u_sys, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print(f"D2H again: {dt*1e3:.2f} ms = {hu.numel()*8/dt/1e9:.1f} GB/s")
