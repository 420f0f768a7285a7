"""Hardware evidence that the reference's example and reproduction scripts run UNCHANGED on the CUDA class.

    python scripts/check_reference_scripts.py --reference <checkout> [--log profiles/r2_reference_scripts_on_b200.log]

Runs, through scripts/run_reference_script.py --backend cuda (shadow package ahead of the reference on sys.path),
  examples/direct_data_driven_mpc_example.py --seed 0 --t_sim 400          (BASELINE config 1)
  examples/robust_data_driven_mpc_reproduction.py                          (seed 4, t_sim 600: TEC, TEC n-step, UCON)
and compares the trajectories each script hands to its plotting functions with the fixtures the UNMODIFIED reference
class produced (tests/golden/refclass_example_seed0.npz, refclass_reproduction_seed4.npz): <= 1e-5 relative on u
(north_star tolerance).  The GPU pool has no reference checkout, so a gpurun call stages one in a git-ignored scratch
directory for the duration of the call; nothing of the reference is committed.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def run(ref, script, script_args, lines):
    out = os.path.join(tempfile.mkdtemp(), "capture.npz")
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "run_reference_script.py"), "--reference", ref, "--backend", "cuda",
           "--capture", out, "--", script, *script_args]
    lines.append("$ " + " ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    lines.extend(l for l in res.stdout.splitlines() if l.startswith("[harness]"))
    if res.returncode != 0:
        lines.append(res.stderr[-3000:])
        raise SystemExit("\n".join(lines))
    shadow = os.path.join(ROOT, "direct_data_driven_mpc")
    assert any(shadow in l for l in lines if "controller class came from" in l), "the shadow package did not supply the class"
    return np.load(out)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--log", default=None)
    args = ap.parse_args()
    import torch
    lines = [f"device: {torch.cuda.get_device_name(0)}; reference checkout staged at {os.path.abspath(args.reference)}"]
    ok = True
    g = np.load(os.path.join(GOLDEN, "refclass_example_seed0.npz"))
    c = run(args.reference, "examples/direct_data_driven_mpc_example.py", ["--seed", "0", "--t_sim", "400", "--verbose", "0"], lines)
    eu, ey = rel(c["0_plot_input_output_u_k"], g["u_sys"]), rel(c["0_plot_input_output_y_k"], g["y_sys"])
    lines.append(f"example script (seed 0, t_sim 400, robust n-step): u rel err {eu:.2e}, y rel err {ey:.2e} vs the reference class "
                 f"(401 steps, 101 QP solves)")
    ok &= eu <= 1e-5 and ey <= 1e-5
    g = np.load(os.path.join(GOLDEN, "refclass_reproduction_seed4.npz"))
    c = run(args.reference, "examples/robust_data_driven_mpc_reproduction.py", ["--verbose", "0"], lines)
    for i, name in ((1, "TEC"), (2, "TEC_N_STEP"), (3, "UCON")):
        n = g[f"u_{name}"].shape[0]                      # the UCON fixture holds the first 150 steps (it diverges by design)
        # the script plots [U_n; u_sys]: the n = 4 warm-up steps come first (robust_data_driven_mpc_reproduction.py:294-295)
        assert np.array_equal(c[f"{i}_plot_input_output_u_k"][:4], g["U_n"])
        eu = rel(c[f"{i}_plot_input_output_u_k"][4:4 + n], g[f"u_{name}"])
        ey = rel(c[f"{i}_plot_input_output_y_k"][4:4 + n], g[f"y_{name}"])
        lines.append(f"reproduction script (seed 4, t_sim 600) {name}: u rel err {eu:.2e}, y rel err {ey:.2e} over {n} steps "
                     f"(script ran {c[f'{i}_plot_input_output_u_k'].shape[0]})")
        ok &= eu <= 1e-5 and ey <= 1e-5
    lines.append("RESULT: " + ("PASS (<= 1e-5 on u and y for all schemes)" if ok else "FAIL"))
    text = "\n".join(lines)
    print(text)
    if args.log:
        os.makedirs(os.path.dirname(os.path.abspath(args.log)), exist_ok=True)
        open(args.log, "w").write(text + "\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
