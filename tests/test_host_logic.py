"""CPU: host-side mirror of the reference interface - argument validation that
happens before any device work, enum quirks, module paths."""
import numpy as np
import pytest

from direct_data_driven_mpc_b200 import (DataDrivenMPCType, DirectDataDrivenMPCController,
                                         SlackVarConstraintTypes)


def _args(**over):
    n, m, p, L, N = 2, 1, 1, 4, 40
    rng = np.random.default_rng(0)
    kw = dict(n=n, m=m, p=p, u_d=rng.uniform(-1, 1, (N, m)), y_d=rng.uniform(-1, 1, (N, p)), L=L,
              Q=np.eye(p * L), R=np.eye(m * L), u_s=np.zeros((m, 1)), y_s=np.zeros((p, 1)), eps_max=0.01,
              lamb_alpha=1.0, lamb_sigma=10.0, c=1.0, slack_var_constraint_type=SlackVarConstraintTypes.NONE,
              controller_type=DataDrivenMPCType.ROBUST)
    kw.update(over)
    return kw


def test_enum_quirks_match_reference():
    # controller.py:11-20: trailing commas make these tuple-valued
    assert DataDrivenMPCType.NOMINAL.value == (0,)
    assert DataDrivenMPCType.ROBUST.value == 1
    assert SlackVarConstraintTypes.NON_CONVEX.value == (0,)
    assert SlackVarConstraintTypes.CONVEX.value == (1,)
    assert SlackVarConstraintTypes.NONE.value == 2


def test_shadow_module_paths():
    from direct_data_driven_mpc.direct_data_driven_mpc_controller import (  # noqa: F401
        DataDrivenMPCType as T2, DirectDataDrivenMPCController as C2, SlackVarConstraintTypes as S2)
    from direct_data_driven_mpc.utilities.hankel_matrix import (  # noqa: F401
        evaluate_persistent_excitation, hankel_matrix)
    assert C2 is DirectDataDrivenMPCController and T2 is DataDrivenMPCType and S2 is SlackVarConstraintTypes


def test_bad_controller_and_slack_type():
    with pytest.raises(ValueError, match="Unsupported controller type"):
        DirectDataDrivenMPCController(**_args(controller_type="robust"))
    with pytest.raises(ValueError, match="Unsupported slack variable constraint type"):
        DirectDataDrivenMPCController(**_args(slack_var_constraint_type=1))


def test_missing_robust_params():
    with pytest.raises(ValueError, match="All robust MPC parameters"):
        DirectDataDrivenMPCController(**_args(lamb_sigma=None))


def test_channel_mismatch_and_short_data():
    with pytest.raises(ValueError, match="should match the number of inputs"):
        DirectDataDrivenMPCController(**_args(m=2, R=np.eye(8), u_s=np.zeros((2, 1))))
    a = _args()
    a["u_d"], a["y_d"] = a["u_d"][:10], a["y_d"][:10]
    with pytest.raises(ValueError, match="required minimum N is 15"):
        DirectDataDrivenMPCController(**a)


def test_hankel_window_error_is_raised_on_host():
    from direct_data_driven_mpc_b200 import hankel_matrix
    with pytest.raises(ValueError, match="N must be greater than or equal to L"):
        hankel_matrix(np.zeros((3, 2)), 4)
