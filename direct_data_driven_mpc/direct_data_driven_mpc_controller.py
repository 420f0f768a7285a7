"""Module path of the reference controller
(direct_data_driven_mpc/direct_data_driven_mpc_controller.py), re-exporting the
B200 implementation under the reference's names."""
from direct_data_driven_mpc_b200.controller import (  # noqa: F401
    DataDrivenMPCType, DirectDataDrivenMPCController, SlackVarConstraintTypes)
