"""Time the closed-loop kernel variants on the bench workload (config 3: 65,536 loops x 101 solves).

    python scripts/time_variants.py [--loops 65536] [--reps 30]

Each variant is selected with the environment switches DESIGN.md section 6a lists; the library reads them at every call.
Prints the median launch time (CUDA events around 10 graph replays, output buffers larger than L2) and, for the
variants that support it, the time without the trajectory stores (DDMPC_DEBUG_NOSTORE=1: compute only).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = [
    ("ws (default)", {}),
    ("ws, no stores", {"DDMPC_DEBUG_NOSTORE": "1"}),
    ("ws, 4 math warps", {"DDMPC_WS_MATH_WARPS": "4"}),
    ("ws, 4 math warps, no stores", {"DDMPC_WS_MATH_WARPS": "4", "DDMPC_DEBUG_NOSTORE": "1"}),
    ("ws, math warps draw", {"DDMPC_WS_MATH_DRAWS": "1"}),
    ("ws, math warps draw, no stores", {"DDMPC_WS_MATH_DRAWS": "1", "DDMPC_DEBUG_NOSTORE": "1"}),
    ("rws (DDMPC_REG=2)", {"DDMPC_REG": "2"}),
    ("rws, no stores", {"DDMPC_REG": "2", "DDMPC_DEBUG_NOSTORE": "1"}),
    ("regx (DDMPC_REG=3)", {"DDMPC_REG": "3"}),
    ("regx, no stores", {"DDMPC_REG": "3", "DDMPC_DEBUG_NOSTORE": "1"}),
    ("reg NT=4 (DDMPC_REG=1)", {"DDMPC_REG": "1"}),
    ("hybrid (DDMPC_WS=0)", {"DDMPC_WS": "0"}),
]
SWITCHES = ("DDMPC_WS", "DDMPC_WS_MATH_WARPS", "DDMPC_WS_MATH_DRAWS", "DDMPC_REG", "DDMPC_REG_NT", "DDMPC_DEBUG_NOSTORE", "DDMPC_PLANT_MMA")


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

    def step():
        return cs.closed_loop(plant, x0, up0, yp0, us, ys, n_steps, w=None, noise_seed=0, scenario_id0=0,
                              noise_eps=0.002, out=(u_sys, y_sys))

    ref = None
    for name, env in VARIANTS:
        if args.only and args.only not in name:
            continue
        for k in SWITCHES:
            os.environ.pop(k, None)
        os.environ.update(env)
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
        if "DDMPC_DEBUG_NOSTORE" not in env:
            cur = (u_sys.clone(), y_sys.clone())
            if ref is None:
                ref = cur
            else:
                du = float((cur[0] - ref[0]).abs().max() / ref[0].abs().max())
                dy = float((cur[1] - ref[1]).abs().max() / ref[1].abs().max())
                line += f"  max rel diff vs first variant: u {du:.2e} y {dy:.2e}"
        print(line, flush=True)
    for k in SWITCHES:
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
