"""Where does the e2e step time go with one process per GPU?  Run under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
use_nccl = os.environ.get("DIAG_NCCL", "1") == "1"
if world > 1 and use_nccl:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B = 65536
sc = S.config3_batch(B); prm, plant = sc["params"], sc["plant"]
cs = ControllerSet(4, 2, 2, sc["u_d"], sc["y_d"], 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0, 0, 1, 4, True, device=dev)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hx0, hup, hyp, hus, hys = pin(sc["x0"]), pin(sc["u_past0"]), pin(sc["y_past0"]), pin(sc["u_s"]), pin(sc["y_s"])
hu = torch.empty(B, 401, 2, dtype=torch.float64, pin_memory=True); hy = torch.empty(B, 401, 2, dtype=torch.float64, pin_memory=True)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
def sync_all():
    torch.cuda.synchronize()
    if world > 1 and use_nccl: dist.barrier()
    torch.cuda.synchronize()
u = torch.empty(B, 401, 2, dtype=torch.float64, device=dev); y = torch.empty_like(u)
for tag in ("raw D2H",):
    sync_all()
    ts = []
    for i in range(4):
        t = time.perf_counter(); hu.copy_(u, non_blocking=True); hy.copy_(y, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    say(tag, np.round(ts, 1))
for chunks in (8, 1, 2, 4, 8):
    sync_all()
    ts = []
    for i in range(6):
        t = time.perf_counter()
        cs.closed_loop_host(plant, hx0, hup, hyp, hus, hys, 401, noise_seed=0, noise_eps=0.002, out=(hu, hy), chunks=chunks)
        ts.append((time.perf_counter() - t) * 1e3)
    say("closed_loop_host chunks", chunks, np.round(ts, 1))
# device-side only
sync_all()
t = time.perf_counter()
for i in range(5):
    cs.closed_loop(plant, hx0.to(dev), hup.to(dev), hyp.to(dev), hus.to(dev), hys.to(dev), 401, noise_seed=0, noise_eps=0.002, out=(u, y))
torch.cuda.synchronize(); say("device-only step ms", (time.perf_counter() - t) / 5 * 1e3)
say("affinity", len(os.sched_getaffinity(0)), "OMP", os.environ.get("OMP_NUM_THREADS"))
