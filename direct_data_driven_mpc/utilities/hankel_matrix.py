"""Module path of the reference's Hankel utilities
(direct_data_driven_mpc/utilities/hankel_matrix.py), backed by the CUDA kernels."""
from direct_data_driven_mpc_b200.hankel import (  # noqa: F401
    evaluate_persistent_excitation, hankel_matrix)
