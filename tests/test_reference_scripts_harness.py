"""The reference's own example script, UNMODIFIED, through scripts/run_reference_script.py.

CPU leg (this container has /root/reference but no GPU): backend `refclass` = the reference class with the cvxpy
stand-in; the trajectories the script hands to its plotting functions must equal the oracle's closed loop for the same
seed.  GPU leg: backend `cuda` = the shadow package (our class) under the same unmodified script; it needs a machine
with both a GPU and the reference checkout (the GPU pool boxes have no /root/reference, where it is skipped)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ddmpc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DDMPC_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")


def _run(backend, tmp_path, t_sim):
    out = str(tmp_path / "capture.npz")
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "run_reference_script.py"), "--reference", REF, "--backend", backend,
           "--capture", out, "--", "examples/direct_data_driven_mpc_example.py", "--seed", "0", "--t_sim", str(t_sim),
           "--verbose", "0"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    c = np.load(out)
    return c["0_plot_input_output_u_k"], c["0_plot_input_output_y_k"], res.stdout


@needs_ref
def test_example_script_unmodified_with_reference_class(tmp_path):
    u, y, log = _run("refclass", tmp_path, 60)
    assert os.path.join(REF, "direct_data_driven_mpc") in log            # the reference's own controller class ran
    u_ref, y_ref, _, _ = O.run_example(seed=0, t_sim=60)
    assert u.shape == (61, 2)
    assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max() and np.abs(y - y_ref).max() <= 1e-9


@needs_ref
@pytest.mark.gpu
def test_example_script_unmodified_with_cuda_class(tmp_path):
    u, y, log = _run("cuda", tmp_path, 400)
    assert os.path.join(ROOT, "direct_data_driven_mpc") in log           # the shadow package supplied the class
    u_ref, y_ref, _, _ = O.run_example(seed=0, t_sim=400)
    assert np.abs(u - u_ref).max() <= 1e-5 * np.abs(u_ref).max() and np.abs(y - y_ref).max() <= 1e-5
