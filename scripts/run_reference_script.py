"""Run one of the reference's own scripts UNMODIFIED on top of this repository.

    python scripts/run_reference_script.py [--reference /root/reference] [--backend cuda|refclass]
                                           [--capture out.npz] -- examples/direct_data_driven_mpc_example.py --seed 0 --t_sim 400

--backend cuda      (default) the shadow package `direct_data_driven_mpc/` of this repository is put ahead of the reference
                    checkout on sys.path, so the script's `from direct_data_driven_mpc.direct_data_driven_mpc_controller
                    import ...` resolves to the CUDA-backed class (needs a GPU and the built libddmpc.so);
--backend refclass  the reference's own class, with tests/golden/mini_cvxpy.py standing in for cvxpy (CPU only; this is how
                    the refclass_* fixtures are made and how the harness itself is tested without a GPU).

The scripts import matplotlib at module top and end in a blocking plt.show(); neither matplotlib nor a display exists on the
build / GPU boxes.  Nothing in the reference is edited: if matplotlib cannot be imported, a stub is placed in sys.modules,
and the reference's plotting module (utilities.visualization.data_visualization) is replaced by a recorder that keeps the
arrays the script hands to it.  `--capture` writes those arrays (the closed-loop trajectories the script would plot) to an
.npz, which is how a run is compared with the oracle or with another backend.
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import types
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_plot_stubs(captured: list) -> None:
    try:
        import matplotlib  # noqa: F401
        have_mpl = True
    except Exception:  # noqa: BLE001
        have_mpl = False
    if not have_mpl:
        top = mock.MagicMock(name="matplotlib")
        sys.modules["matplotlib"] = top
        for sub in ("pyplot", "animation", "gridspec", "patches", "lines", "axes", "figure", "legend", "legend_handler",
                    "text", "transforms", "collections", "ticker"):
            m = mock.MagicMock(name=f"matplotlib.{sub}")
            setattr(top, sub, m)
            sys.modules[f"matplotlib.{sub}"] = m
    else:
        os.environ.setdefault("MPLBACKEND", "Agg")
    # recorder for the reference's plotting front end (its signature: data_visualization.py plot_input_output(u_k, y_k, ...))
    viz = types.ModuleType("utilities.visualization.data_visualization")

    def recorder(kind):
        def fn(*args, **kw):
            captured.append((kind, {k: np.array(v) for k, v in kw.items() if k in ("u_k", "y_k", "u_s", "y_s")}))
            ret = mock.MagicMock(name=kind)              # figure / axes stand-ins; unpackable as (fig, axs_u, axs_y)
            ret.__iter__.side_effect = lambda: iter([mock.MagicMock(), mock.MagicMock(), mock.MagicMock()])
            return ret
        return fn

    for name in ("plot_input_output", "plot_input_output_animation", "save_animation", "create_input_output_figure",
                 "plot_data"):
        setattr(viz, name, recorder(name))
    def other(name):                                         # any other plotting helper a script may import
        if name.startswith("__"):                            # (inspect.getmodule probes __file__ etc. of every module)
            raise AttributeError(name)
        return recorder(name)
    viz.__getattr__ = other
    sys.modules["utilities.visualization.data_visualization"] = viz


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--backend", choices=["cuda", "refclass"], default="cuda")
    ap.add_argument("--capture", default=None)
    ap.add_argument("script", help="path of the reference script, relative to the reference checkout")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    ref = os.path.abspath(args.reference)
    if not os.path.isdir(ref):
        raise SystemExit(f"reference checkout not found: {ref}")
    captured: list = []
    if args.backend == "refclass":
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import mini_cvxpy
        sys.modules["cvxpy"] = mini_cvxpy
        sys.path.insert(0, ref)
    else:
        import torch  # noqa: F401  (before the matplotlib stand-ins exist: torch inspects sys.modules while it loads)
        sys.path.insert(0, ref)
        sys.path.insert(0, ROOT)                             # shadow package wins over the reference's own package
    install_plot_stubs(captured)
    script = os.path.join(ref, args.script)
    sys.argv = [script] + [a for a in args.script_args if a != "--"]
    cwd = os.getcwd()
    os.chdir(ref)                                            # the scripts locate their YAML files relative to the checkout
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        os.chdir(cwd)
    mod = sys.modules.get("direct_data_driven_mpc.direct_data_driven_mpc_controller")
    print(f"[harness] controller class came from: {getattr(mod, '__file__', '?')}")
    print(f"[harness] plotting calls recorded: {[k for k, _ in captured]}")
    if args.capture:
        out = {}
        for i, (kind, arrs) in enumerate(captured):
            for k, v in arrs.items():
                out[f"{i}_{kind}_{k}"] = v
        np.savez_compressed(args.capture, **out)
        print(f"[harness] wrote {args.capture}")


if __name__ == "__main__":
    main()
