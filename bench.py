#!/usr/bin/env python
"""Headline benchmark: batched DD-MPC QP solves/s (BASELINE.json metric).

Workload (BASELINE config 3): four-tank robust n-step DD-MPC
(n_mpc_step = 4, terminal constraint on, slack NONE), shared data (seed 0),
65,536 closed loops per GPU (256 set-points x 256 noise realisations),
n_steps = 401 -> 101 QP solves per loop, 6.62 M solves per step per GPU.
One "step" = one fused closed-loop pass over the whole batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...      (one rank per GPU, weak scaling)

Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched DD-MPC QP solves/sec"
UNIT = "solves/s"
N_STEPS = 401
SOLVES_PER_LOOP = 101          # ceil(401 / 4)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loops", type=int, default=65536, help="closed loops per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time direct launches instead of CUDA-graph replays")
    ap.add_argument("--e2e-steps", type=int, default=10)
    return ap.parse_args()


# --------------------------------------------------------------------------
# CPU baseline (oracle port; the only place outside tests that runs oracle/)
# --------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one host core: closed loops of the bench workload with the oracle."""
    wid, budget_s, flavour = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import ddmpc_oracle as O
    from direct_data_driven_mpc_b200 import scenarios as S
    prm = S.four_tank_controller_params()
    rng, x0, u_d, y_d, x_end = S.example_data(0)
    us_g, ys_g = S.setpoint_grid(S.four_tank_plant(), 16, first=(prm["u_s"], prm["y_s"]))
    cache = flavour == "cached_factor"
    ctrl = O.OracleController(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["u_s"], prm["y_s"], prm["eps_max"],
                              prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], O.SLACK_NONE, O.ROBUST, 4, True,
                              cache_factor=cache, check_pe=False)
    plant = O.four_tank_plant()
    solves, loops = 0, 0
    t0 = time.perf_counter()
    b = wid
    n_steps = N_STEPS if cache else 21
    while time.perf_counter() - t0 < budget_s:
        plant.x = x_end.copy()
        ctrl.u_s, ctrl.y_s = us_g[b % 256].reshape(-1, 1), ys_g[b % 256].reshape(-1, 1)
        ctrl.set_past_input_output_data(u_d[-4:].reshape(-1, 1), y_d[-4:].reshape(-1, 1))
        w = O.philox_noise(0, np.array([b]), n_steps, 2, 0.002)[0]
        O.closed_loop(plant, ctrl, n_steps, w)
        solves += -(-n_steps // 4)
        loops += 1
        b += 1024
    return solves, loops, time.perf_counter() - t0


def cpu_baseline(budget_s: float):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    out = {}
    with ctx.Pool(cores) as pool:
        for flavour, share in (("cached_factor", 0.75), ("rebuild_every_step", 0.25)):
            res = pool.map(_cpu_worker, [(i, budget_s * share, flavour) for i in range(cores)])
            out[flavour] = (sum(r[0] / r[2] for r in res), sum(r[1] for r in res), sum(r[0] for r in res))
    v, loops, solves = out["cached_factor"]
    return {
        "value": v, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{loops} closed loops ({solves} QP solves) of the bench workload, oracle literal-KKT with the LU "
                  f"factor cached per controller, one process per core, BLAS threads = 1, ~{budget_s * 0.75:.0f} s",
        "rebuild_every_step_value": out["rebuild_every_step"][0],
        "rebuild_every_step_note": "assemble + dense KKT solve at every MPC iteration, as the reference re-creates "
                                   "its cvxpy problem each step (controller.py:404-405)",
    }


def single_loop_latency(make_ctrl, plant_factory, w_sys, n_steps=401):
    """BASELINE metric part (ii): p50 single-loop step latency - config 1 (seed 0, t_sim 400) driven through the
    per-step controller API exactly as the reference's loop does (controller_operation.py:269-305); one sample per MPC
    iteration = solve + its n_mpc_step plant steps + window updates."""
    ctrl, plant = make_ctrl(), plant_factory()
    ts = []
    for rep in range(3):                       # rep 0 warms everything up
        ctrl_r, plant_r = (ctrl, plant) if rep == 0 else (make_ctrl(), plant_factory())
        ts = []
        for t in range(0, n_steps, ctrl_r.n_mpc_step):
            t0 = time.perf_counter()
            ctrl_r.update_and_solve_data_driven_mpc()
            for k in range(t, min(t + ctrl_r.n_mpc_step, n_steps)):
                u = ctrl_r.get_optimal_control_input_at_step(n_step=k - t)
                y = plant_r.simulate_step(u, w_sys[k])
                ctrl_r.store_input_output_measurement(u.reshape(-1, 1), y.reshape(-1, 1))
            ts.append((time.perf_counter() - t0) * 1e6)
    ts = np.array(ts)
    return {"p50_us": float(np.percentile(ts, 50)), "p90_us": float(np.percentile(ts, 90)), "iterations": int(ts.size),
            "what": "config 1 (four-tank, seed 0, t_sim 400, n_mpc_step 4): one MPC iteration = solve + 4 plant steps"}


class _HostPlant:
    """utilities/model_simulation.py:93-98 on the host (the plant stays on the host in the per-step API)."""

    def __init__(self, pl, x):
        self.A, self.B, self.C, self.D, self.x = pl.A, pl.B, pl.C, pl.D, np.array(x, dtype=float)

    def simulate_step(self, u, w):
        y = self.C @ self.x + self.D @ u + w
        self.x = self.A @ self.x + self.B @ u
        return y


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores.  cvxpy is not installed in this
    image (no wheel, no network) so the unmodified reference class cannot run; the oracle port runs instead."""
    if rank != 0:
        return
    # each "step" of this arm is a bounded time sample of the same workload; at most 6 samples of ~15 s are
    # executed however large --steps is, so the arm always ends within a few minutes
    n_eff = max(1, min(args.steps, 6))
    per_step = max(4.0, min(15.0, 90.0 / n_eff))
    vals = []
    for i in range(min(args.warmup, 1) + n_eff):
        cb = cpu_baseline(per_step)
        if i >= min(args.warmup, 1):
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals])) if vals else 0.0
    cb = vals[-1]
    cb["value"] = v
    cb["sample"] += f"; {n_eff} such samples executed for --steps {args.steps}"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.loops, world),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "cvxpy unavailable in this image: oracle port of the reference QP (cached-factor flavour) on all host "
                "cores; each step is a bounded time sample of the same workload",
    }
    print(json.dumps(line), flush=True)


def workload_config(loops, world):
    return {"workload": "config 3: four-tank robust n-step DD-MPC (n_mpc_step=4, terminal on, slack NONE), shared "
                        "data seed 0, 256 set-points x noise realisations, n_steps=401 (101 QP solves per loop)",
            "loops_per_gpu": loops, "global_loops": loops * world, "n_steps": N_STEPS,
            "noise": "device Philox4x32-10, seed 0, stream = global scenario id",
            "l2": "each step writes 841 MB of trajectories per GPU (> 126 MB L2); no flush needed",
            "parallelism": f"scenario-sharded x{world}, no data-path collective; one NCCL all_gather of per-loop "
                           "metrics after the timed steps",
            "launch": "each step is one k_closed_loop_ws launch (warp-specialised all-tensor-core kernel) replayed from a CUDA graph"}


# --------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------
class Clocks:
    """Samples SM clock and throttle reasons through NVML (nvidia-ml-py) in a thread while the
    timed region runs; falls back to an `nvidia-smi -lms` child process."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.rows, self.idx, self.stop_flag, self.thread, self.smax = [], gpu_index, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self.stop_flag:
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.05)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.thread.join(timeout=1.0)
        inside = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        reasons = set()
        for _, _, mask in inside:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons),
                "samples": len(inside), "source": "NVML nvmlDeviceGetClockInfo / CurrentClocksEventReasons, 50 ms period"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(dev_index: int):
    """Run this rank on the CPUs local to its GPU, so the pinned staging buffers (first-touch) and the
    copy-launching thread sit on the GPU's NUMA node.  Best effort; returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[dev_index]) if vis and vis.split(",")[dev_index].isdigit() else dev_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            return f"bound to {len(allowed)} GPU-local CPUs"
        return "already local"
    except Exception as exc:
        return f"not bound ({type(exc).__name__})"


def measure_e2e(args, cs, sc, plant, id0, solves_per_step, world, barrier, dev, u_sys):
    """`e2e`: the same metric through the host-buffer API (ControllerSet.closed_loop_host): every step copies its
    inputs from pinned host memory and brings the full trajectories back to pinned host memory."""
    import torch
    import torch.distributed as dist
    B = sc["x0"].shape[0]
    old_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(dev.index)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx0, hup, hyp, hus, hys = pin(sc["x0"]), pin(sc["u_past0"]), pin(sc["y_past0"]), pin(sc["u_s"]), pin(sc["y_s"])
    hu = torch.empty(B, N_STEPS, 2, dtype=torch.float64, pin_memory=True)
    hy = torch.empty(B, N_STEPS, 2, dtype=torch.float64, pin_memory=True)

    def e2e_step():
        return cs.closed_loop_host(plant, hx0, hup, hyp, hus, hys, N_STEPS, w=None, noise_seed=0,
                                   scenario_id0=id0, noise_eps=0.002, out=(hu, hy), chunks=4)

    e2e_step()
    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        _, _, hst = e2e_step()
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = world * solves_per_step * args.e2e_steps / (float(ems.item()) * 1e-3)
    assert int(hst.max()) == 0
    # (chunks of the host path may take a different kernel specialisation: same maths, different FP64 summation order)
    assert torch.allclose(hu[:64], u_sys[:64].cpu(), rtol=1e-9, atol=1e-9), "host-API result differs from the device-resident run"
    h2d = sum(t.numel() * t.element_size() for t in (hx0, hup, hyp, hus, hys))
    d2h = hu.numel() * 8 + hy.numel() * 8 + B * 4
    os.sched_setaffinity(0, old_affinity)      # the CPU baseline must see every core again
    return {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": float(ems.item()) / args.e2e_steps,
           "api": "ControllerSet.closed_loop_host (pinned host in, full trajectories out, 4 chunks on 2 persistent streams)",
           "numa": numa}



def secondary_config4(dev, B=16384):
    """BASELINE config 4 (synthetic n = 20, m = p = 4, N = 2000, L = 40; 16384 closed loops of 401 steps) through the
    fused FP64 tensor-core kernel (dmma_loop.cu), device-resident inputs, CUDA events.  Not the headline: reported next
    to it so the tensor-core-bound configuration has a measured number in the same run."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S
    out = {"workload": f"config 4: synthetic stable LTI n=20 m=p=4 N=2000 L=40 robust, {B} loops x 401 steps",
           "fp64_peak_tflops": 37.2, "fp64_peak_source": "scripts/probes/fp64_pipes.cu: 64 FMA/clk/SM x 148 SMs x 1965 MHz"}
    for nmpc in (1, 20):
        sc = S.config4_batch(B, n_mpc_step=nmpc)
        prm, pl = sc["params"], sc["plant"]
        torch.cuda.synchronize()
        t = time.perf_counter()
        cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                           prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, nmpc, True, device=dev)
        torch.cuda.synchronize()
        setup_ms = (time.perf_counter() - t) * 1e3
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        args = (pl, td(sc["x0"]), td(sc["u_past0"]), td(sc["y_past0"]), td(sc["u_s"]), td(sc["y_s"]), N_STEPS)
        bufs = (torch.empty(B, N_STEPS, 4, dtype=torch.float64, device=dev),
                torch.empty(B, N_STEPS, 4, dtype=torch.float64, device=dev))
        run = lambda: cs.closed_loop(*args, noise_seed=0, noise_eps=0.002, out=bufs)
        for _ in range(3):
            _, _, st, it = run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        solves = int(it.sum().item())
        # executed DMMA per 8 loops and MPC iteration: solve ceil(n_mpc*m/8) x n_theta/4, plant ceil((n_mpc*p+n_x)/8) x (n_x+n_mpc*m)/4
        dmma = -(-nmpc * 4 // 8) * (168 // 4) + -(-(nmpc * 4 + 20) // 8) * ((20 + nmpc * 4) // 4)
        flops = 2.0 * 256 * dmma * (B / 8) * (solves / B)
        err = float((bufs[1][:, -1] - args[5]).abs().max())
        out[f"n_mpc_{nmpc}"] = {"loop_ms": ms, "solves_per_s": solves / (ms * 1e-3), "setup_ms": setup_ms,
                                "status_max": int(st.max().item()), "final_tracking_error_max": err,
                                "executed_tflops": flops / (ms * 1e-3) / 1e12,
                                "frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / 37.2}
        del cs
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from direct_data_driven_mpc_b200 import ControllerSet, _lib
    from direct_data_driven_mpc_b200 import scenarios as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)

    B = args.loops
    sc = S.config3_batch(B, seed=0)
    prm, plant = sc["params"], sc["plant"]
    t_setup0 = time.perf_counter()
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], prm["slack_type"], prm["controller_type"],
                       prm["n_mpc_step"], True, device=dev)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t_setup0) * 1e3
    id0 = rank * B                                   # Philox stream = global scenario id
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0, up0, yp0, us, ys = d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"])
    u_sys = torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev)
    y_sys = torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev)

    def step():
        return cs.closed_loop(plant, x0, up0, yp0, us, ys, N_STEPS, w=None, noise_seed=0, scenario_id0=id0,
                              noise_eps=0.002, out=(u_sys, y_sys))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    _, _, status, iters = step()
    barrier()
    assert int(status.max()) == 0, "non-optimal solve status in the bench workload"
    solves_per_step = int(iters.sum().item())
    assert solves_per_step == B * SOLVES_PER_LOOP
    # One step = one kernel launch; the Python call around it costs 0.1-0.4 ms on a busy host, more than the
    # kernel.  Capture the step once in a CUDA graph and replay it (launch-bound inner loop -> graph).
    graph, launches_per_step = None, 1
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            l0 = _lib.kernel_launches()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                _, _, status, iters = step()
            launches_per_step = _lib.kernel_launches() - l0
        except Exception as exc:  # pragma: no cover
            print(f"CUDA graph capture failed ({exc}); timing direct launches", file=sys.stderr)
            graph = None
    run_step = graph.replay if graph is not None else step
    for _ in range(max(args.warmup, 3)):       # W untimed warm-up steps, no idle gap before the timed region
        run_step()
    # warm the per-loop metric kernels (first use loads their CUDA modules) and the collective
    # (communicator set-up): neither is part of a step
    wtrack = (y_sys[:, -1, :] - ys).abs().amax(dim=1)
    if world > 1:
        dist.all_gather([torch.empty_like(wtrack) for _ in range(world)], wtrack)
    barrier()
    launches0 = _lib.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    dbg = os.environ.get("BENCH_DEBUG_EVENTS") == "1"
    dbg_ev = []
    for _ in range(args.steps):
        run_step()
        if dbg:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            dbg_ev.append(e)
    # the only collective of the job: per-loop metrics (final tracking error + status) to every rank
    track = (y_sys[:, -1, :] - ys).abs().amax(dim=1)
    if world > 1:
        gathered = [torch.empty_like(track) for _ in range(world)]
        dist.all_gather(gathered, track)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    launches = (_lib.kernel_launches() - launches0) if graph is None else launches_per_step * args.steps
    if dbg and rank == 0:
        ts = np.array([dbg_ev[i].elapsed_time(dbg_ev[i + 1]) for i in range(len(dbg_ev) - 1)])
        print("per-step ms: first", np.round(ts[:8], 3), "median", np.median(ts), "p90", np.percentile(ts, 90), "max",
              ts.max(), "host loop s", t1 - t0, file=sys.stderr)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    value = world * solves_per_step * args.steps / (total_ms * 1e-3)
    clk = clocks.stop(t0, t1) if rank == 0 else None

    # ---- dominant kernel alone (k_closed_loop), CUDA events on its stream, for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(min(args.steps, 50))]
    for a, b_ in kev:
        a.record()
        run_step()
        b_.record()
    torch.cuda.synchronize()
    # median, not mean: a host hiccup between two replays (busy multi-rank boxes) must not count as kernel time
    k_ms = float(np.median([a.elapsed_time(b_) for a, b_ in kev]))
    alg_bytes = B * N_STEPS * (2 + 2) * 8          # u_sys + y_sys written once (Philox noise: nothing read)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_closed_loop_ws_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # the default kernel for this batch size is the warp-specialised all-tensor-core one (fast_loop.cu): per n-step
    # block and 64 loops 32 + 48 DMMA m8n8k4 (256 FMA each, zero padding of the 12-row plant block map included)
    roofline = {"bound": "hbm", "kernel": "k_closed_loop_ws", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "flops_per_solve_executed": 2 * 80 * 256 // 64,
                "store_pattern_floor_ms": 0.204,
                "store_pattern_note": "841 MB written as 32-byte sectors 6416 B apart (reference layout (B, n_steps, m)) take "
                                      "0.204 ms on a B200 with no compute at all (scripts/probes/store_pattern.cu, "
                                      "profiles/r1_probes.txt); a fully coalesced write of the same bytes takes 0.138 ms"}

    # ---- end to end through the host-buffer API: pinned inputs H2D, trajectories D2H, every step
    e2e = None
    try:
        e2e = measure_e2e(args, cs, sc, plant, id0, solves_per_step, world, barrier, dev, u_sys)
    except Exception as exc:  # pragma: no cover - never lose the main line over an auxiliary measurement
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(exc)}
    cb, latency, secondary = None, None, None
    if rank == 0 and world == 1:
        try:
            secondary = secondary_config4(dev)
        except Exception as exc:  # pragma: no cover
            secondary = {"error": repr(exc)}
    if rank == 0:
        from direct_data_driven_mpc_b200 import (DataDrivenMPCType, DirectDataDrivenMPCController,
                                                 SlackVarConstraintTypes)
        rng1, x0_1, ud1, yd1, xe1 = S.example_data(0)
        w1 = 0.002 * rng1.uniform(-1.0, 1.0, (N_STEPS, 2))

        def make_gpu_ctrl():
            return DirectDataDrivenMPCController(
                n=4, m=2, p=2, u_d=ud1, y_d=yd1, L=30, Q=prm["Q"], R=prm["R"], u_s=prm["u_s"], y_s=prm["y_s"],
                eps_max=prm["eps_max"], lamb_alpha=prm["lamb_alpha"], lamb_sigma=prm["lamb_sigma"], c=prm["c"],
                slack_var_constraint_type=SlackVarConstraintTypes.NONE, controller_type=DataDrivenMPCType.ROBUST,
                n_mpc_step=4, use_terminal_constraint=True)
        try:
            latency = single_loop_latency(make_gpu_ctrl, lambda: _HostPlant(plant, xe1), w1)
            latency["api"] = "DirectDataDrivenMPCController.update_and_solve_data_driven_mpc (B = 1, host buffers)"
        except Exception as exc:  # pragma: no cover
            latency = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_baseline(args.cpu_seconds)
            from oracle import ddmpc_oracle as O
            cb["single_loop_latency"] = single_loop_latency(
                lambda: O.make_controller(O.four_tank_params(), ud1, yd1), lambda: _HostPlant(plant, xe1), w1)
        except Exception as exc:  # pragma: no cover
            cb = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port", "sample": "failed",
                  "error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(B, world),
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb,
            "single_loop_latency": latency, "secondary": secondary,
            "setup_ms": setup_ms, "solves_per_step_per_gpu": solves_per_step,
            "final_tracking_error_max": float(track.max().item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
