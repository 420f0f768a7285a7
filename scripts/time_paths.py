"""Closed-loop kernel selection data: config 3 (four-tank robust n-step, 401 steps) at several batch sizes through each
kernel that applies (ControllerSet.set_option("closed_loop_path", ...)), CUDA-graph replays timed with CUDA events.

    python scripts/time_paths.py [--sizes 1024,4096,8192,16384,32768,65536] [--n-mpc 4]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,4096,8192,16384,32768,65536")
    ap.add_argument("--n-mpc", type=int, default=4)
    ap.add_argument("--paths", default="perloop,fast,ws,generic")
    args = ap.parse_args()
    import torch
    from direct_data_driven_mpc_b200 import ControllerSet
    from direct_data_driven_mpc_b200 import scenarios as S
    dev = torch.device("cuda", 0)
    sizes = [int(s) for s in args.sizes.split(",")]
    sc = S.config3_batch(max(sizes), seed=0)
    prm, plant = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], prm["c"], 0, 1, args.n_mpc, True, device=dev)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0, up0, yp0, us, ys = d(sc["x0"]), d(sc["u_past0"]), d(sc["y_past0"]), d(sc["u_s"]), d(sc["y_s"])
    n_steps = 401
    u = torch.empty(max(sizes), n_steps, 2, dtype=torch.float64, device=dev)
    y = torch.empty_like(u)
    for B in sizes:
        row = []
        for path in args.paths.split(","):
            if path == "ws" and args.n_mpc != 4:
                continue
            if path == "generic" and B > 16384:
                continue
            cs.set_option("closed_loop_path", path)
            step = lambda: cs.closed_loop(plant, x0[:B], up0[:B], yp0[:B], us[:B], ys[:B], n_steps, noise_seed=0, noise_eps=0.002,
                                          out=(u[:B], y[:B]))
            _, _, st, it = step()
            run, _, g = bench.graph_of(step)
            ms = bench.median_ms(run, reps=20)
            row.append(f"{path} {ms:.4f} ms ({int(it.sum()) / (ms * 1e-3):.3e} solves/s)")
        print(f"B = {B:6d} n_mpc = {args.n_mpc}: " + " | ".join(row), flush=True)
    cs.set_option("closed_loop_path", "auto")


if __name__ == "__main__":
    main()
