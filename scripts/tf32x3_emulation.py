"""Bounded experiment on the tcgen05 route: would a TF32x3 split of the config-4 gain product keep the
closed loop within the north-star tolerance (1e-5 relative on u over 401 steps)?

tcgen05.mma has no FP64 kind; kind::tf32 multiplies 10-bit-mantissa operands and accumulates in FP32 (TMEM).  The classic
3xTF32 scheme writes each FP32 operand as hi + lo (both TF32) and sums hi*hi + hi*lo + lo*hi.  Two things bound its accuracy
here, and neither can be fixed by adding the three partial products in FP64 afterwards:
  (1) the operands themselves are FP64: hi + lo carries 21-22 mantissa bits, so each operand is rounded to ~2^-22 relative;
  (2) every partial product is a K = 168 term dot product accumulated in FP32 inside the tensor core.
This script emulates exactly that arithmetic in NumPy (round-to-nearest TF32 splitting, FP32 accumulation of each partial
dot product, FP64 sum of the three partials) inside the oracle's config-4 closed loop and reports the error on u against
the FP64 loop.  No GPU needed; the result decides whether a tcgen05 kernel is worth writing.

    python scripts/tf32x3_emulation.py [--steps 401] [--loops 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ddmpc_oracle as O  # noqa: E402
from oracle import workloads as W  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import condensed_numpy as CN  # noqa: E402


def tf32(x):
    """Round FP32 values to TF32 (10 explicit mantissa bits), round to nearest even."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & ~np.uint64(0x1FFF)
    return b.astype(np.uint32).view(np.float32)


def split(x64):
    hi = tf32(x64.astype(np.float32))
    lo = tf32((x64 - hi.astype(np.float64)).astype(np.float32))
    return hi, lo


def matvec_tf32x3(K_hi, K_lo, v64, extra=False):
    v_hi, v_lo = split(v64)
    parts = [K_hi @ v_hi, K_hi @ v_lo, K_lo @ v_hi]           # float32 @ float32: FP32 accumulation
    if extra:
        parts.append(K_lo @ v_lo)
    return sum(p.astype(np.float64) for p in parts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=401)
    ap.add_argument("--loops", type=int, default=4)
    args = ap.parse_args()
    for nmpc in (20, 1):
        sc = W.config4(n_mpc_step=nmpc)
        prm, pl = sc["params"], sc["plant"]
        plan = CN.build_plan(20, 4, 4, sc["u_d"], sc["y_d"], 40, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                             prm["lamb_sigma"], prm["c"], 0, 1, True)
        Ku = plan.Ku[: nmpc * 4]                              # applied gain rows (n_mpc m x n_theta = 168)
        K_hi, K_lo = split(Ku)
        # n_mpc-step block map of the plant: [Y (n_mpc p); x+ (n_x)] = Mb [x (n_x); U (n_mpc m)]
        nx = 20
        Mb = np.zeros((nmpc * 4 + nx, nx + nmpc * 4))
        for j in range(nx + nmpc * 4):
            e = np.zeros(nx + nmpc * 4); e[j] = 1.0
            xx = e[:nx].copy()
            for k in range(nmpc):
                uk = e[nx + 4 * k: nx + 4 * k + 4]
                Mb[4 * k:4 * k + 4, j] = pl.C @ xx + pl.D @ uk
                xx = pl.A @ xx + pl.B @ uk
            Mb[nmpc * 4:, j] = xx
        M_hi, M_lo = split(Mb)
        r = np.random.default_rng(0)
        worst = {"tf32x3": 0.0, "tf32x4": 0.0, "fp32": 0.0, "tf32x3 gain+plant": 0.0}
        for loop in range(args.loops):
            x0 = sc["x_end"] + 0.1 * r.normal(size=20)
            w = pl.eps_max * r.uniform(-1, 1, (args.steps, 4))
            runs = {}
            for mode in ("fp64", "tf32x3", "tf32x4", "fp32", "tf32x3 gain+plant"):
                x = x0.copy()
                up, yp = sc["u_d"][-20:].reshape(-1).copy(), sc["y_d"][-20:].reshape(-1).copy()
                us = np.zeros((args.steps, 4))
                for t in range(0, args.steps, nmpc):
                    th = np.concatenate([up, yp, prm["u_s"].reshape(-1), prm["y_s"].reshape(-1)])
                    if mode == "fp64":
                        u = Ku @ th
                    elif mode == "fp32":
                        u = (Ku.astype(np.float32) @ th.astype(np.float32)).astype(np.float64)
                    else:
                        u = matvec_tf32x3(K_hi, K_lo, th, extra=mode == "tf32x4")
                    if mode == "tf32x3 gain+plant" and t + nmpc <= args.steps:
                        # both products of the block on the tensor cores: the plant through its block map, outputs and next
                        # state come back as FP32 accumulators; the noise is added in FP64
                        out = matvec_tf32x3(M_hi, M_lo, np.concatenate([x, u]))
                        for k in range(t, t + nmpc):
                            uk = u[(k - t) * 4:(k - t + 1) * 4]
                            y = out[(k - t) * 4:(k - t + 1) * 4] + w[k]
                            us[k] = uk
                            up = np.concatenate([up[4:], uk])
                            yp = np.concatenate([yp[4:], y])
                        x = out[nmpc * 4:]
                        continue
                    for k in range(t, min(t + nmpc, args.steps)):
                        uk = u[(k - t) * 4:(k - t + 1) * 4]
                        y = pl.C @ x + pl.D @ uk + w[k]
                        x = pl.A @ x + pl.B @ uk
                        us[k] = uk
                        up = np.concatenate([up[4:], uk])
                        yp = np.concatenate([yp[4:], y])
                runs[mode] = us
            ref = runs["fp64"]
            for mode in worst:
                worst[mode] = max(worst[mode], float(np.abs(runs[mode] - ref).max() / np.abs(ref).max()))
        print(f"config 4, n_mpc_step = {nmpc}, {args.loops} loops x {args.steps} steps: max relative error on u vs the FP64 loop: "
              + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()) + "   (tolerance 1e-5)", flush=True)
        print(f"   gain block {Ku.shape}: max |K| = {np.abs(Ku).max():.3g}, max row sum |K_r| . |theta| ~ "
              f"{(np.abs(Ku) @ np.abs(np.concatenate([sc['u_d'][-20:].reshape(-1), sc['y_d'][-20:].reshape(-1), prm['u_s'].reshape(-1), prm['y_s'].reshape(-1)]))).max():.3g}")


if __name__ == "__main__":
    main()
