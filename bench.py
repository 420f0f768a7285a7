#!/usr/bin/env python
"""Headline benchmark: batched DD-MPC QP solves/s (BASELINE.json metric).

Workload (BASELINE config 3): four-tank robust n-step DD-MPC
(n_mpc_step = 4, terminal constraint on, slack NONE), shared data (seed 0),
closed loops = 256 set-points x noise realisations, n_steps = 401 -> 101 QP
solves per loop.  One "step" = one fused closed-loop pass over the whole batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
  torchrun ... bench.py --gpus N ...      (one rank per GPU)

--scaling weak  (default): 65,536 loops PER GPU (6.62 M solves per step per GPU).
--scaling strong: BASELINE config 3 as written, 65,536 loops IN TOTAL, cut into N contiguous shards.
A weak-scaling run on N > 1 GPUs also measures the strong-scaling configuration and reports it under "strong".

Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
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
CONFIG3_LOOPS = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--loops", type=int, default=CONFIG3_LOOPS,
                    help="closed loops per GPU (weak scaling) or in total (strong scaling)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config 2 / 4 / 5 / CONVEX measurements")
    ap.add_argument("--no-graph", action="store_true", help="time direct launches instead of CUDA-graph replays")
    ap.add_argument("--e2e-steps", type=int, default=10)
    return ap.parse_args()


WORKLOAD_TEXT = ("config 3: four-tank robust n-step DD-MPC (n_mpc_step=4, terminal on, slack NONE), shared data seed 0, "
                 "256 set-points x noise realisations, n_steps=401 (101 QP solves per loop)")


# --------------------------------------------------------------------------
# CPU arm (oracle port; the only place outside tests that runs oracle/).  It imports NOTHING from the product package:
# no libddmpc.so is mapped into these processes.
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
    from oracle import workloads as W
    sc = W.config3(0)
    prm, u_d, y_d = sc["params"], sc["u_d"], sc["y_d"]
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
        plant.x = sc["x_start"].copy()
        ctrl.u_s, ctrl.y_s = sc["u_s"][b % 256].reshape(-1, 1), sc["y_s"][b % 256].reshape(-1, 1)
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
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT, "n_steps": N_STEPS,
                   "noise": "Philox4x32-10, seed 0, stream = scenario id (oracle/ddmpc_oracle.py philox_noise)",
                   "execution": f"oracle port on {cb['cores']} host cores, one process per core (rank 0 only; no GPU work): "
                                "every process runs whole closed loops of the workload back to back for the sample time",
                   "step": f"one step = a {per_step:.0f} s time sample"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "cvxpy unavailable in this image: oracle port of the reference QP (cached-factor flavour) on all host "
                "cores; each step is a bounded time sample of the same workload",
    }
    print(json.dumps(line), flush=True)


def workload_config(loops, world, scaling, kernel):
    return {"workload": WORKLOAD_TEXT,
            "loops_per_gpu": loops, "global_loops": loops * world, "n_steps": N_STEPS,
            "noise": "device Philox4x32-10, seed 0, stream = global scenario id",
            "l2": f"each step writes {loops * N_STEPS * 32 / 1e6:.0f} MB of trajectories per GPU"
                  + (" (> 126 MB L2); no flush needed" if loops * N_STEPS * 32 > 126e6 else
                     " (< 126 MB L2: the outputs of consecutive steps overwrite each other in L2)"),
            "parallelism": f"scenario-sharded x{world} ({scaling} scaling), no data-path collective; one NCCL all_gather of "
                           "per-loop metrics inside the timed region, the full-trajectory gather timed separately (gather)",
            "launch": f"each step is one {kernel} launch replayed from a CUDA graph"}


def kernel_for(loops: int) -> str:
    """Name of the kernel ddmpc_closed_loop_batch selects for this batch of the bench workload."""
    return "k_closed_loop_ws" if loops >= 6144 else "k_closed_loop_perloop"


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


def graph_of(step, enable=True):
    """Capture `step` (warmed on a side stream first) in a CUDA graph; returns (callable, launches per step, graph)."""
    import torch
    from direct_data_driven_mpc_b200 import _lib
    if not enable:
        return step, None, None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        l0 = _lib.kernel_launches()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        return graph.replay, _lib.kernel_launches() - l0, graph
    except Exception as exc:  # pragma: no cover
        print(f"CUDA graph capture failed ({exc}); timing direct launches", file=sys.stderr)
        return step, None, None


def median_ms(run, reps=30, warm=3):
    """Median device time of one call of `run` (CUDA events around each call, on the current stream)."""
    import torch
    for _ in range(warm):
        run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        run()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def measure_e2e(args, cs, sc, plant, id0, solves_per_step, world, barrier, dev, u_sys, y_sys):
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
    # the floor of this number: the raw device->host copy of the trajectories alone (every rank at the same time)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hu.copy_(u_sys, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    r0.record()
    for _ in range(3):
        hu.copy_(u_sys, non_blocking=True)
        hy.copy_(y_sys, non_blocking=True)
    r1.record()
    barrier()
    rms = torch.tensor([r0.elapsed_time(r1) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rms, op=dist.ReduceOp.MAX)
    h2d = sum(t.numel() * t.element_size() for t in (hx0, hup, hyp, hus, hys))
    d2h = hu.numel() * 8 + hy.numel() * 8 + B * 4
    os.sched_setaffinity(0, old_affinity)      # the CPU baseline must see every core again
    return {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": float(ems.item()) / args.e2e_steps,
            "raw_d2h_ms": float(rms.item()),
            "raw_d2h_note": "cudaMemcpyAsync of the same trajectory bytes into the same pinned buffers, nothing else, all "
                            "ranks concurrently, max over ranks: the floor of ms_per_step on this box",
            "api": "ControllerSet.closed_loop_host (pinned host in, full trajectories out, 4 chunks on 2 persistent streams)",
            "numa": numa}


def measure_gather(u_sys, y_sys, world, dev, barrier, reps=3):
    """K6: the one collective of a job - all ranks receive the full trajectories (NCCL all_gather over NVLink)."""
    import torch
    import torch.distributed as dist
    from direct_data_driven_mpc_b200.sharding import gather_shards
    nbytes = (u_sys.numel() + y_sys.numel()) * 8
    if world == 1:
        return {"gather_ms": 0.0, "bytes_received_per_rank": 0, "note": "single GPU: nothing to gather"}
    total = u_sys.shape[0] * world
    gu = gather_shards(u_sys, total)           # warm-up: communicator buffers, allocator
    gy = gather_shards(y_sys, total)
    ok = bool(torch.equal(gu[dist.get_rank() * u_sys.shape[0]:(dist.get_rank() + 1) * u_sys.shape[0]], u_sys))
    del gu, gy
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gu = gather_shards(u_sys, total)
        gy = gather_shards(y_sys, total)
        del gu, gy
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    g_ms = float(ms.item())
    recv = nbytes * (world - 1)
    return {"gather_ms": g_ms, "bytes_received_per_rank": recv, "bytes_per_rank_shard": nbytes,
            "achieved_gbs_per_gpu": recv / (g_ms * 1e-3) / 1e9, "nvlink5_peak_gbs_per_direction": 900.0,
            "frac_of_nvlink": recv / (g_ms * 1e-3) / 1e9 / 900.0, "own_shard_round_trips": ok,
            "collective": "torch.distributed all_gather_into_tensor (NCCL) of u_sys and y_sys, every rank receives every shard; "
                          "device time, max over ranks, mean of 3"}


def secondary_config4(dev, B=16384, fp64_peak=None, dmma_warps=None):
    """BASELINE config 4 (synthetic n = 20, m = p = 4, N = 2000, L = 40; 16384 closed loops of 401 steps) through the
    fused FP64 tensor-core kernel (dmma_loop.cu), device-resident inputs, CUDA events.  Not the headline: reported next
    to it so the tensor-core-bound configuration has a measured number in the same run."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S
    out = {"workload": f"config 4: synthetic stable LTI n=20 m=p=4 N=2000 L=40 robust, {B} loops x 401 steps",
           "fp64_peak_tflops": fp64_peak, "fp64_peak_source": "ddmpc_probe_fp64_tflops (DMMA m8n8k4), measured in this run"}
    for nmpc in (1, 20):
        sc = S.config4_batch(B, n_mpc_step=nmpc)
        prm, pl = sc["params"], sc["plant"]
        torch.cuda.synchronize()
        t = time.perf_counter()
        cs = ControllerSet(prm["n"], 4, 4, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                           prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], 0, 1, nmpc, True, device=dev)
        torch.cuda.synchronize()
        setup_ms = (time.perf_counter() - t) * 1e3
        if dmma_warps:
            cs.set_option("dmma_warps", dmma_warps)
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
        tf = flops / (ms * 1e-3) / 1e12
        out[f"n_mpc_{nmpc}"] = {"loop_ms": ms, "solves_per_s": solves / (ms * 1e-3), "setup_ms": setup_ms,
                                "status_max": int(st.max().item()), "final_tracking_error_max": err,
                                "executed_tflops": tf, "frac_of_fp64_peak": tf / fp64_peak if fp64_peak else None,
                                "kernel": "k_closed_loop_dmma (FP64 mma.sync m8n8k4)"}
        if nmpc == 20:
            # the same workload on the 5th-generation tensor cores (opt-in path "tc": tcgen05 kind::tf32 with TF32x3 error
            # compensation, loop state resident in TMEM, csrc/tc_loop.cu), and how far its trajectories are from the FP64 ones
            try:
                u64, y64 = bufs[0].clone(), bufs[1].clone()
                cs.set_option("closed_loop_path", "tc")
                for _ in range(3):
                    _, _, st2, it2 = run()
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / reps
                du = float((bufs[0] - u64).abs().max() / u64.abs().max())
                dy = float((bufs[1] - y64).abs().max() / y64.abs().max())
                # tcgen05.mma issued per 128 loops and iteration: 3 passes x (21 k-steps of 128 x 80 x 8 + 13 of 128 x 112 x 8)
                mma_flops = 2.0 * 3 * (21 * 128 * 80 * 8 + 13 * 128 * 112 * 8) * (B / 128) * (N_STEPS // 20)
                out["n_mpc_20_tcgen05"] = {
                    "loop_ms": ms2, "solves_per_s": int(it2.sum().item()) / (ms2 * 1e-3), "status_max": int(st2.max().item()),
                    "speedup_vs_fp64_kernel": ms / ms2, "max_rel_diff_u_vs_fp64_kernel": du, "max_rel_diff_y_vs_fp64_kernel": dy,
                    "tolerance": "north star: 1e-5 relative on u", "tf32_tflops_issued": mma_flops / (ms2 * 1e-3) / 1e12,
                    "hbm_write_gbs": B * N_STEPS * 8 * 8 / (ms2 * 1e-3) / 1e9,
                    "kernel": "k_closed_loop_tc (tcgen05.mma kind::tf32, A operand and accumulators in TMEM)"}
                cs.set_option("closed_loop_path", "auto")
            except Exception as exc:  # pragma: no cover
                out["n_mpc_20_tcgen05"] = {"error": repr(exc)}
        del cs
    return out


def secondary_config2(dev, B=4096):
    """BASELINE config 2: 4096 closed loops over seeds, loop b = the example script with --seed b (its own data, hence its
    own controller), five controller variants; data generated on the device with the reference's NumPy streams.
    Kernel: k_closed_loop_perloop (equality-only variants) / generic k_closed_loop (CONVEX)."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S
    pl, prm = S.four_tank_plant(), S.four_tank_controller_params()
    torch.cuda.synchronize()
    t = time.perf_counter()
    ds = S.DeviceScenarios(seeds=range(B), device=dev)
    w_dev = ds.uniform(N_STEPS * 2, -1.0, 1.0, 0.002).reshape(B, N_STEPS, 2)       # controller_operation.py:263
    torch.cuda.synchronize()
    out = {"workload": f"config 2: {B} closed loops over seeds, per-seed data and controllers, 401 steps, parity-mode noise "
                       "(the seed's own generator)", "data_gen_ms": (time.perf_counter() - t) * 1e3, "variants": {}}
    ud, yd, x0 = ds.u_d, ds.y_d, ds.x_end
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    idx = torch.arange(B, device=dev, dtype=torch.int32)
    for name, slack, term, nmpc, ctype in [("ROBUST TEC n-step", 0, True, 4, 1), ("ROBUST TEC 1-step", 0, True, 1, 1),
                                           ("ROBUST UCON 1-step", 0, False, 1, 1), ("ROBUST CONVEX n-step", 1, True, 4, 1),
                                           ("NOMINAL 1-step", 0, True, 1, 0)]:
        torch.cuda.synchronize()
        t = time.perf_counter()
        cs = ControllerSet(4, 2, 2, ud, yd, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"],
                           1.0, slack, ctype, nmpc, term, device=dev)
        torch.cuda.synchronize()
        setup_ms = (time.perf_counter() - t) * 1e3
        a = (pl, x0, ud[:, -4:].reshape(B, -1).contiguous(), yd[:, -4:].reshape(B, -1).contiguous(),
             td(np.tile(prm["u_s"].T, (B, 1))), td(np.tile(prm["y_s"].T, (B, 1))), N_STEPS)
        bufs = (torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev), torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev))
        run = lambda: cs.closed_loop(*a, w=w_dev, ctrl_idx=idx, out=bufs, check_idx=False)
        _, _, st, it = run()
        ms = median_ms(run, reps=10, warm=2)
        solves = B * (-(-N_STEPS // nmpc))
        out["variants"][name] = {"setup_ms": setup_ms, "controllers_per_s": B / (setup_ms * 1e-3), "controllers_failed": cs.n_failed,
                                 "loop_ms": ms, "qp_solves": solves, "solves_per_s": solves / (ms * 1e-3),
                                 "admm_iterations_mean": float(it.sum().item()) / solves, "status_max": int(st.max().item())}
        del cs
    return out


def secondary_config5(dev):
    """BASELINE config 5: lambda_alpha*eps x lambda_sigma x L sweep on four-tank (16 x 16 weights x 14 horizons = 3584
    controllers on shared data, 64 closed loops each): controllers set up per second and closed-loop throughput."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S
    pl, prm = S.four_tank_plant(), S.four_tank_controller_params()
    rng, x0, u_d, y_d, x_end = S.example_data(0)
    la = np.logspace(-3, 1, 16) / prm["eps_max"]
    ls = np.logspace(1, 5, 16)
    LA, LS = [g.reshape(-1) for g in np.meshgrid(la, ls, indexing="ij")]
    nl = 64
    B = LA.size * nl
    idx = torch.from_numpy(np.repeat(np.arange(LA.size), nl)).to(dev, dtype=torch.int32)
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a_common = (td(np.tile(x_end, (B, 1))), td(np.tile(u_d[-4:].reshape(1, -1), (B, 1))), td(np.tile(y_d[-4:].reshape(1, -1), (B, 1))),
                td(np.tile(prm["u_s"].T, (B, 1))), td(np.tile(prm["y_s"].T, (B, 1))))
    tot_ctrl, tot_t, tot_solves, tot_loop_ms, failed = 0, 0.0, 0, 0.0, 0
    for L in range(8, 61, 4):
        Q, R = 3.0 * np.eye(2 * L), 1e-4 * np.eye(2 * L)
        mk = lambda: ControllerSet(4, 2, 2, u_d, y_d, L, Q, R, prm["eps_max"], LA, LS, 1.0, 0, 1, 4, True, count=LA.size, device=dev)
        cs = mk()
        del cs                                 # warm the stream-ordered memory pool for this size
        torch.cuda.synchronize()
        t = time.perf_counter()
        cs = mk()
        torch.cuda.synchronize()
        tot_t += time.perf_counter() - t
        failed += cs.n_failed
        run = lambda: cs.closed_loop(pl, *a_common, N_STEPS, noise_seed=0, noise_eps=0.002, ctrl_idx=idx, check_idx=False)
        _, _, st, it = run()
        tot_loop_ms += median_ms(run, reps=5, warm=1)
        tot_ctrl += LA.size
        tot_solves += int(it.sum().item())
        del cs
    return {"workload": "config 5: 16 x 16 (lambda_alpha*eps, lambda_sigma) x 14 horizons L = 8..60 on shared four-tank data, "
                        "64 closed loops x 401 steps per controller",
            "controllers": tot_ctrl, "controllers_failed": failed, "setup_s": tot_t, "controllers_per_s": tot_ctrl / tot_t,
            "qp_solves": tot_solves, "loops_ms_total": tot_loop_ms, "solves_per_s": tot_solves / (tot_loop_ms * 1e-3)}


def secondary_convex(dev, sc, fp64_peak, B=CONFIG3_LOOPS):
    """Config 3 with the CONVEX slack bound ||sigma_pred||_inf <= c eps_max (controller.py:659-675), c = 1 and c = 0.3:
    the iterative part of the solver (box-row ADMM) inside the fused closed loop."""
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    prm, plant = sc["params"], sc["plant"]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a[:B])).to(dev)
    args = (plant, d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"]), N_STEPS)
    bufs = (torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev), torch.empty(B, N_STEPS, 2, dtype=torch.float64, device=dev))
    out = {"workload": f"config 3 with slack CONVEX, {B} loops x 401 steps, tol 1e-8", "fp64_peak_tflops": fp64_peak}
    nb, nth = 60, 20
    for c in (1.0, 0.3, 100.0):      # (c = 100: the bound never binds - what the screens and the loop cost by themselves)
        cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                           prm["lamb_alpha"], prm["lamb_sigma"], c, 1, 1, 4, True, device=dev)
        run = lambda: cs.closed_loop(*args, noise_seed=0, noise_eps=0.002, out=bufs)
        _, _, st, it = run()
        ms = median_ms(run, reps=5, warm=1)
        solves = B * SOLVES_PER_LOOP
        iters = float(it.sum().item())
        active_loops = int((it > SOLVES_PER_LOOP).sum().item())
        # executed FP64 work: every solve pays the gain product + plant block (640 flop); every ADMM iteration beyond the
        # check one Phi d product (2 nb^2).  The slack check Ks theta (2 nb n_theta per solve) runs as a one-pass TF32
        # screen plus its modulus product on the legacy tensor path and is reported separately - it is not FP64 work; the
        # exact FP64 re-check of the suspicious blocks, the final t = Phi d and the Psi correction of the active solves are
        # not counted (their number is not reported by the kernel).
        flops = solves * 640 + (iters - solves) * 2 * nb * nb
        screen = solves * 2 * (2 * nb * nth)
        tf = flops / (ms * 1e-3) / 1e12
        out[f"c_{c}"] = {"loop_ms": ms, "solves_per_s": solves / (ms * 1e-3), "status_max": int(st.max().item()),
                         "iterations_mean_per_solve": iters / solves, "admm_iterations_total": iters - solves,
                         "loops_with_an_active_bound": active_loops, "executed_fp64_tflops": tf,
                         "frac_of_fp64_peak": tf / fp64_peak if fp64_peak else None,
                         "tf32_screen_tflops": screen / (ms * 1e-3) / 1e12}
        del cs
    # ---- the batched solve itself (ddmpc_solve_batch, no plant): B QPs of one shared CONVEX controller through the
    #      tensor-core pipeline (k_gemm products around k_admm_dmma, csrc/cvx_loop.cu), every problem with a binding box
    try:
        c = 0.3
        cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                           prm["lamb_alpha"], prm["lamb_sigma"], c, 1, 1, 4, True, device=dev)
        r = np.random.default_rng(0)
        ks = r.integers(0, sc["u_d"].shape[0] - 4, B)
        up = torch.from_numpy(np.stack([sc["u_d"][k:k + 4].reshape(-1) for k in ks])).to(dev)
        yp = torch.from_numpy(np.stack([sc["y_d"][k:k + 4].reshape(-1) for k in ks])).to(dev)
        us, ys = d(sc["u_s"]), d(sc["y_s"])
        run = lambda: cs.solve_batch(up, yp, us, ys, tol=1e-8, want_cost=False)
        _, _, st, it = run()
        ms = median_ms(run, reps=5, warm=1)
        cs.set_option("solve_path", 1)
        run1 = lambda: cs.solve_batch(up, yp, us, ys, tol=1e-8, want_cost=False)
        ms_scalar = median_ms(run1, reps=3, warm=1)
        iters = float(it.sum().item())
        active = int((it > 1).sum().item())
        # executed: u0 = Ku theta (2 Lm n_theta), s = Ks theta (2 nb n_theta), u -= Psi t (2 Lm nb) per problem as GEMMs;
        # per ADMM iteration of an ACTIVE problem one Phi d (2 nb^2); the lock-step iterations a CTA spends on problems
        # that have already converged are executed too but are not counted here
        flops = B * (2 * 60 * nth + 2 * nb * nth + 2 * 60 * nb) + (iters - (B - active)) * 2 * nb * nb
        tf = flops / (ms * 1e-3) / 1e12
        out["solve_batch"] = {"problems": B, "c": c, "pipeline_ms": ms, "solves_per_s": B / (ms * 1e-3),
                              "thread_per_solve_kernel_ms": ms_scalar, "speedup_vs_thread_per_solve": ms_scalar / ms,
                              "problems_with_an_active_bound": active, "iterations_mean_per_active_problem":
                              (iters - (B - active)) / max(active, 1), "status_max": int(st.max().item()),
                              "executed_tflops_useful": tf, "frac_of_fp64_peak": tf / fp64_peak if fp64_peak else None,
                              "kernels": "k_pack_theta, k_gemm x2, k_admm_dmma, k_gemm"}
        del cs
    except Exception as exc:  # pragma: no cover
        out["solve_batch"] = {"error": repr(exc)}
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

    strong = args.scaling == "strong"
    if strong and args.loops % world:
        raise SystemExit("--scaling strong needs --loops divisible by the number of GPUs")
    B = args.loops // world if strong else args.loops            # loops on this GPU
    total = B * world
    sc = S.config3_batch(total if strong else B, seed=0)
    if strong:                                       # this rank's contiguous shard of the 65,536 scenarios
        for k in ("x0", "u_past0", "y_past0", "u_s", "y_s"):
            sc[k] = sc[k][rank * B:(rank + 1) * B]
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
    run_step, launches_per_step, graph = graph_of(step, not args.no_graph)
    if launches_per_step is None:
        launches_per_step = 1
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
    for _ in range(args.steps):
        run_step()
    # per-loop metrics (final tracking error) to every rank: the collective a Monte-Carlo job needs every pass
    track = (y_sys[:, -1, :] - ys).abs().amax(dim=1)
    if world > 1:
        gathered = [torch.empty_like(track) for _ in range(world)]
        dist.all_gather(gathered, track)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    launches = (_lib.kernel_launches() - launches0) if graph is None else launches_per_step * args.steps
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    value = world * solves_per_step * args.steps / (total_ms * 1e-3)
    clk = clocks.stop(t0, t1) if rank == 0 else None

    # ---- dominant kernel alone, CUDA events on its stream, for the roofline
    # median, not mean: a host hiccup between two replays (busy multi-rank boxes) must not count as kernel time
    k_ms = median_ms(run_step, reps=min(max(args.steps, 10), 50), warm=0)
    alg_bytes = B * N_STEPS * (2 + 2) * 8          # u_sys + y_sys written once (Philox noise: nothing read)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    kernel = kernel_for(B)
    # roofline denominators measured in this run, on this device (csrc/probes.cu)
    fp64_dmma = fp64_dfma = floor_owner = floor_coal = None
    try:
        stream = torch.cuda.current_stream().cuda_stream
        fp64_dmma = _lib.probe_fp64_tflops(True, stream)
        fp64_dfma = _lib.probe_fp64_tflops(False, stream)
        floor_owner = _lib.probe_store_ms(B, N_STEPS, False, u_sys.data_ptr(), y_sys.data_ptr(), stream)
        floor_coal = _lib.probe_store_ms(B, N_STEPS, True, u_sys.data_ptr(), y_sys.data_ptr(), stream)
        run_step()                                 # the probes overwrote the trajectories
        torch.cuda.synchronize()
    except Exception as exc:  # pragma: no cover
        print(f"probes failed: {exc!r}", file=sys.stderr)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get(kernel)
            if ent and ent.get("loops") == B:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            traffic = None
    # executed flops per solve: ws kernel = 32 + 48 DMMA m8n8k4 (256 FMA each, zero padding of the 12-row plant block
    # map included) per 64 loops and n-step block; per-loop kernel = (8 x 16 gain + 8 x 4 set-point fold once) + 12 x 12 map
    flops_per_solve = 2 * 80 * 256 // 64 if kernel == "k_closed_loop_ws" else 2 * (8 * 16 + 12 * 12)
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "flops_per_solve_executed": flops_per_solve,
                "fp64": {"executed_tflops": flops_per_solve * solves_per_step / (k_ms * 1e-3) / 1e12,
                         "peak_tflops_dmma": fp64_dmma, "peak_tflops_dfma": fp64_dfma,
                         "frac": (flops_per_solve * solves_per_step / (k_ms * 1e-3) / 1e12 / fp64_dmma) if fp64_dmma else None,
                         "peak_source": "ddmpc_probe_fp64_tflops, this run, this device"},
                "store_pattern_floor_ms": floor_owner, "coalesced_store_floor_ms": floor_coal,
                "frac_of_store_pattern_floor": (floor_owner / k_ms) if floor_owner else None,
                "store_pattern_note": "time to write the same trajectory bytes with NO compute, measured in this run "
                                      "(ddmpc_probe_store_ms): in the reference layout (B, n_steps, m) a loop owns a contiguous "
                                      "run and produces 32 bytes of it per step pair, so a warp store touches 32 sectors "
                                      "16*n_steps bytes apart (store_pattern_floor_ms); fully coalesced stores of the same "
                                      "bytes (coalesced_store_floor_ms) are the HBM write peak"}

    # ---- K6: gather of the full trajectories (the only data-path collective of a job), timed on NCCL
    gather = None
    try:
        gather = measure_gather(u_sys, y_sys, world, dev, barrier)
        gather["value_incl_gather"] = world * solves_per_step / ((ms_per_step + gather["gather_ms"]) * 1e-3)
        gather["value_incl_gather_note"] = "one job = one closed-loop pass + the gather of its trajectories"
    except Exception as exc:  # pragma: no cover
        gather = {"error": repr(exc)}

    # ---- strong scaling: BASELINE config 3 as written (65,536 loops in total, sharded), measured in the same run
    strong_res = None
    if not strong and world > 1 and CONFIG3_LOOPS % world == 0:
        try:
            Bs = CONFIG3_LOOPS // world
            sl = slice(rank * Bs, (rank + 1) * Bs)
            sc3 = S.config3_batch(CONFIG3_LOOPS, seed=0)
            sx0, sup, syp, sus, sys_ = (d(sc3[k][sl]) for k in ("x0", "u_past0", "y_past0", "u_s", "y_s"))
            su, sy = u_sys[:Bs], y_sys[:Bs]
            sstep = lambda: cs.closed_loop(plant, sx0, sup, syp, sus, sys_, N_STEPS, w=None, noise_seed=0,
                                           scenario_id0=rank * Bs, noise_eps=0.002, out=(su, sy))
            _, _, sst, sit = sstep()
            srun, _, sgraph = graph_of(sstep, not args.no_graph)
            for _ in range(5):
                srun()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.steps):
                srun()
            s1.record()
            barrier()
            sms = torch.tensor([s0.elapsed_time(s1) / args.steps], dtype=torch.float64, device=dev)
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
            sg = measure_gather(su.contiguous(), sy.contiguous(), world, dev, barrier)
            s_ms = float(sms.item())
            strong_res = {"loops_total": CONFIG3_LOOPS, "loops_per_gpu": Bs, "kernel": kernel_for(Bs), "ms_per_step": s_ms,
                          "value": CONFIG3_LOOPS * SOLVES_PER_LOOP / (s_ms * 1e-3), "unit": UNIT, "status_max": int(sst.max().item()),
                          "gather_ms": sg["gather_ms"],
                          "value_incl_gather": CONFIG3_LOOPS * SOLVES_PER_LOOP / ((s_ms + sg["gather_ms"]) * 1e-3),
                          "note": "BASELINE config 3 as written: 65,536 closed loops in total, contiguous shards, device time, "
                                  "max over ranks"}
            step()                                 # restore the weak-scaling outputs for the checks below
            torch.cuda.synchronize()
        except Exception as exc:  # pragma: no cover
            strong_res = {"error": repr(exc)}

    # ---- end to end through the host-buffer API: pinned inputs H2D, trajectories D2H, every step
    e2e = None
    try:
        e2e = measure_e2e(args, cs, sc, plant, id0, solves_per_step, world, barrier, dev, u_sys, y_sys)
    except Exception as exc:  # pragma: no cover - never lose the main line over an auxiliary measurement
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(exc)}
    cb, latency, secondary, shard_times = None, None, None, None
    if rank == 0 and world == 1:
        # what one GPU of a strong-scaled job runs: the shard sizes of 65,536 loops over 2 / 4 / 8 GPUs
        try:
            shard_times = {}
            for Bs in (32768, 16384, 8192):
                if Bs >= B:
                    continue
                sstep = lambda: cs.closed_loop(plant, x0[:Bs], up0[:Bs], yp0[:Bs], us[:Bs], ys[:Bs], N_STEPS, w=None, noise_seed=0,
                                               scenario_id0=0, noise_eps=0.002, out=(u_sys[:Bs], y_sys[:Bs]))
                srun, _, sgraph = graph_of(sstep, not args.no_graph)
                shard_times[str(Bs)] = {"ms": median_ms(srun, reps=20), "kernel": kernel_for(Bs)}
            step()
            torch.cuda.synchronize()
        except Exception as exc:  # pragma: no cover
            shard_times = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = {}

        def secondary_step_major():
            """The same workload with the trajectories stored step-major (n_steps, B, m): an OPTION for consumers that stay on
            the device (the headline keeps the reference's per-loop layout).  Same numbers, other strides."""
            us_, ys_ = u_sys.view(N_STEPS, B, 2), y_sys.view(N_STEPS, B, 2)
            run = lambda: cs.closed_loop(plant, x0, up0, yp0, us, ys, N_STEPS, w=None, noise_seed=0, scenario_id0=id0,
                                         noise_eps=0.002, out=(us_, ys_), layout="step_major")
            run()
            ms = median_ms(run, reps=20)
            alg = B * N_STEPS * 4 * 8
            out = {"workload": "config 3, trajectories stored (n_steps, B, m)", "loop_ms": ms,
                   "solves_per_s": B * SOLVES_PER_LOOP / (ms * 1e-3), "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                   "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak if peak else None,
                   "coalesced_store_floor_ms": floor_coal,
                   "kernel": "k_closed_loop_ws, trajectory_layout = 1", "loop_major_ms": k_ms}
            step()                                        # the headline buffers hold the loop-major result again
            return out

        for name, fn in (("step_major_layout", secondary_step_major),
                         ("config4", lambda: secondary_config4(dev, fp64_peak=fp64_dmma)),
                         ("config2", lambda: secondary_config2(dev)),
                         ("config5", lambda: secondary_config5(dev)),
                         ("convex", lambda: secondary_convex(dev, sc, fp64_dmma, min(B, CONFIG3_LOOPS)))):
            try:
                secondary[name] = fn()
            except Exception as exc:  # pragma: no cover
                secondary[name] = {"error": repr(exc)}
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
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(B, world, args.scaling, kernel),
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb,
            "gather": gather, "strong": strong_res, "single_gpu_shard_times": shard_times,
            "single_loop_latency": latency, "secondary": secondary,
            "setup_ms": setup_ms, "solves_per_step_per_gpu": solves_per_step,
            "final_tracking_error_max": float(track.max().item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
