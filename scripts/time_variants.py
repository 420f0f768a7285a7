"""Time the closed-loop kernel variants on the bench workload (config 3: 65,536 loops x 101 solves).

    python scripts/time_variants.py [--loops 65536] [--reps 30]

The product kernels are selected with ControllerSet.set_option("closed_loop_path", ...); the measured-and-dropped
designs live in experiments/closed_loop_variants.cu, which this script builds into experiments/libddmpc_experiments.so
(nvcc, sm_100a) and calls through ddmpc_exp_closed_loop().  Prints the median launch time (CUDA events around 10 graph
replays, output buffers larger than L2) and, where supported, the time without the trajectory stores (compute only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, product path or None, experiment variant or None, nostore)
VARIANTS = [
    ("ws (default)", "ws", None, 0),
    ("ws, no stores", None, "ws2", 1),
    ("ws, 4 math warps", None, "ws4", 0),
    ("ws, 1 math warp", None, "ws1", 0),
    ("ws, math warps draw", None, "ws2md", 0),
    ("ws, 32-loop CTAs (1 + 1 warps)", None, "ws1l1", 0),
    ("ws, 32-loop CTAs, no stores", None, "ws1l1", 1),
    ("ws, 2 + 2 warps (draw | record)", None, "ws2io2", 0),
    ("ws, 2 + 2 warps, no stores", None, "ws2io2", 1),
    ("ws, drawing shared", None, "ws2mh", 0),
    ("ws, drawing shared, no stores", None, "ws2mh", 1),
    ("rws", None, "rws", 0),
    ("rws, no stores", None, "rws", 1),
    ("regx", None, "regx", 0),
    ("reg NT=4", None, "reg", 0),
    ("reg NT=4, no stores", None, "reg", 1),
    ("single-warp mma", None, "mma", 0),
    ("hybrid", "fast", None, 0),
    ("8 lanes per loop", "perloop", None, 0),
]


def build_experiments() -> C.CDLL:
    from direct_data_driven_mpc_b200 import _lib
    src = os.path.join(ROOT, "experiments", "closed_loop_variants.cu")
    out = os.path.join(ROOT, "experiments", "libddmpc_experiments.so")
    csrc = os.path.dirname(_lib.LIB_PATH)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                        "-fPIC", "-shared", "-o", out, src, "-L" + csrc, "-lddmpc", "-Xlinker", "-rpath", "-Xlinker",
                        "$ORIGIN/../direct_data_driven_mpc_b200/csrc"], check=True)
    C.CDLL(_lib.LIB_PATH, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(out)
    vp, i32, f64, u64 = C.c_void_p, C.c_int, C.c_double, C.c_uint64
    lib.ddmpc_exp_closed_loop.restype = i32
    lib.ddmpc_exp_closed_loop.argtypes = [vp, C.POINTER(_lib.Plant), C.c_char_p, i32, i32, vp, vp, vp, vp, vp, vp, u64, u64,
                                          f64, i32, vp, vp, vp, vp, vp, vp]
    return lib


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--loops", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--only", default=None, help="substring of the variant name")
    args = ap.parse_args()
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S

    dev = torch.device("cuda", 0)
    B, n_steps = args.loops, 401
    sc = S.config3_batch(B, seed=0)
    prm, plant = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"],
                       prm["lamb_alpha"], prm["lamb_sigma"], prm["c"], prm["slack_type"], prm["controller_type"],
                       prm["n_mpc_step"], True, device=dev)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0, up0, yp0, us, ys = d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"])
    u_sys = torch.empty(B, n_steps, 2, dtype=torch.float64, device=dev)
    y_sys = torch.empty(B, n_steps, 2, dtype=torch.float64, device=dev)

    exp = build_experiments()
    from direct_data_driven_mpc_b200 import _lib
    ps = plant.c_struct()
    status = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)

    ref = None
    for name, path, variant, nostore in VARIANTS:
        if args.only and args.only not in name:
            continue
        if path is not None:
            cs.set_option("closed_loop_path", path)

            def step():
                return cs.closed_loop(plant, x0, up0, yp0, us, ys, n_steps, w=None, noise_seed=0, scenario_id0=0,
                                      noise_eps=0.002, out=(u_sys, y_sys))
        else:
            def step(variant=variant, nostore=nostore):
                _lib.check(exp.ddmpc_exp_closed_loop(cs._h, C.byref(ps), variant.encode(), nostore, B, x0.data_ptr(),
                                                     up0.data_ptr(), yp0.data_ptr(), us.data_ptr(), ys.data_ptr(), None, 0, 0,
                                                     0.002, n_steps, u_sys.data_ptr(), y_sys.data_ptr(), status.data_ptr(),
                                                     iters.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        # one launch per CUDA graph replay: the Python call around a launch costs about as much as the kernel
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                graph.replay()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        line = f"{name:28s} median {np.median(ts):.4f} ms  min {np.min(ts):.4f} ms"
        if not nostore:
            cur = (u_sys.clone(), y_sys.clone())
            if ref is None:
                ref = cur
            else:
                du = float((cur[0] - ref[0]).abs().max() / ref[0].abs().max())
                dy = float((cur[1] - ref[1]).abs().max() / ref[1].abs().max())
                line += f"  max rel diff vs first variant: u {du:.2e} y {dy:.2e}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
