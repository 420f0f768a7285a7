"""CPU oracle for the DD-MPC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``direct_data_driven_mpc_b200``) never does and has no CPU fallback.

Pinning status.  The arithmetic of the reference's QP solve lives in cvxpy
(unpinned in the reference's ``setup.py:21``; not installed in this image, no
network), and the reference ships no tests or golden vectors for the solve, so
PARITY IS UNPINNED AT THE LEVEL OF cvxpy's BACKEND (its solver tolerance).
Everything above that level is pinned against the reference's own code:

* the QP itself.  ``tests/golden/make_golden_refclass.py`` runs the UNMODIFIED
  reference class ``DirectDataDrivenMPCController`` - constructor, Hankel
  matrices, ``define_mpc_constraints`` / ``define_cost_function`` /
  ``define_mpc_problem`` rebuilt every step, ``get_optimal_control_input``, the
  window methods, ``set_input_output_setpoints``, its exceptions - inside the
  reference's own loop driver, with ``tests/golden/mini_cvxpy.py`` standing in
  for cvxpy (an affine-expression tracker + generic dense QP solve that knows
  nothing about MPC).  The oracle's restatement of the formulation
  (``direct_data_driven_mpc_controller.py:433-445, 533-545, 577-581, 612-627,
  659-675, 703-722``) equals the optimum of the problem the reference builds to
  1e-13 (ROBUST: NONE / CONVEX with active bounds / UCON / general Q, R;
  NOMINAL noise-free; whole 401- and 597-step closed loops) and 4e-10 (NOMINAL
  on noisy data): ``tests/golden/refclass_*.npz``, ``tests/test_refclass_golden.py``.
* the Hankel builder, the LTI plant step, the observer/equilibrium helpers, the
  scenario generators, the YAML parameter derivation and the loop driver:
  checked against the live reference functions (``tests/golden/make_golden.py``)
  and the two docstring known-answer examples the reference holds.
"""
