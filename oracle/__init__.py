"""CPU oracle for the DD-MPC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``direct_data_driven_mpc_b200``) never does and has no CPU fallback.

PARITY UNPINNED: the arithmetic of the reference's QP solve lives in cvxpy
(unpinned in the reference's ``setup.py:21``; not installed in this image, no
network), and the reference ships no tests or golden vectors for the solve.
The oracle therefore restates the *formulation* of
``direct_data_driven_mpc/direct_data_driven_mpc_controller.py:433-445, 533-545,
577-581, 612-627, 659-675, 703-722`` and solves its KKT system exactly in FP64 -
the unique optimum every cvxpy backend converges to.  What IS pinned: the
Hankel builder, the LTI plant step, the observer/equilibrium helpers and the
scenario generators are checked against the live reference functions (golden
fixtures under ``tests/golden`` made by ``tests/golden/make_golden.py``) and
against the two docstring known-answer examples the reference holds.
"""
