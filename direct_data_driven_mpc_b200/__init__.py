"""B200-native DD-MPC hot path (sm_100a).  See DESIGN.md.

The CUDA extension is mandatory: importing ``_lib`` fails loudly when
``csrc/libddmpc.so`` has not been built.  There is no CPU fallback.
"""
from .controller import (DataDrivenMPCType, DirectDataDrivenMPCController,  # noqa: F401
                         SlackVarConstraintTypes)
from .hankel import evaluate_persistent_excitation, hankel_matrix  # noqa: F401
from .batch import ControllerSet, LTIPlant  # noqa: F401

__all__ = ["DirectDataDrivenMPCController", "DataDrivenMPCType", "SlackVarConstraintTypes", "hankel_matrix",
           "evaluate_persistent_excitation", "ControllerSet", "LTIPlant"]
