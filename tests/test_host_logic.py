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


def test_workload_recipes_agree():
    """oracle/workloads.py (lib-free NumPy recipes used by bench.py's CPU arm and the fixture generators) and the
    product-side direct_data_driven_mpc_b200/scenarios.py produce identical arrays for configs 3 and 4."""
    import numpy as np
    from oracle import workloads as W
    from direct_data_driven_mpc_b200 import scenarios as S
    w3, s3 = W.config3(0), S.config3_batch(512, seed=0)
    assert np.array_equal(w3["u_d"], s3["u_d"]) and np.array_equal(w3["y_d"], s3["y_d"])
    assert np.array_equal(w3["x_start"], s3["x0"][0])
    assert np.array_equal(w3["u_s"], s3["u_s"][:256]) and np.array_equal(w3["y_s"], s3["y_s"][:256])
    assert np.array_equal(s3["u_s"][256:512], s3["u_s"][:256])
    for nmpc in (1, 20):
        w4, s4 = W.config4(n_mpc_step=nmpc), S.config4_batch(2, n_mpc_step=nmpc)
        for k in ("u_d", "y_d"):
            assert np.array_equal(w4[k], s4[k])
        assert np.array_equal(w4["x_end"], s4["x0"][0]) and np.array_equal(w4["y_s"], s4["y_s"][0])
        for k in "ABCD":
            assert np.array_equal(getattr(w4["plant"], k), getattr(s4["plant"], k))
        assert w4["params"]["n_mpc_step"] == nmpc and np.array_equal(w4["params"]["Q"], s4["params"]["Q"])


def test_tf32_screen_margin_bounds_the_rounding_error():
    """The CONVEX closed-loop kernel screens the slack rows with a one-pass TF32 product and skips the exact FP64 check when
    |s~| + 2^-8 * (|Ks| |theta|) <= bound (csrc/cvx_loop.cu, kScreenRel).  Emulation of that arithmetic (Ks rounded to TF32,
    theta truncated to TF32 as the tensor core does with FP32 bit patterns, FP32 accumulation): the margin must dominate the
    error on the four-tank gain rows for windows of every scale, with at least a factor two to spare."""
    from oracle import ddmpc_oracle as O
    import condensed_numpy as C
    plant_o, prm, rng, x0, u_d, y_d = O.example_scenario(0)
    pl = C.build_plan(4, 2, 2, u_d, y_d, 30, prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"], prm["lamb_sigma"], 1.0,
                      C.SLACK_CONVEX, C.ROBUST, True)
    Ks = pl.Ks

    def rna(x):       # cvt.rna.tf32.f32: round to nearest, ties away, 10 mantissa bits
        b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
        return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)

    def trunc(x):     # what mma.sync .tf32 reads of an FP32 register
        return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    r = np.random.default_rng(0)
    a_op = rna(Ks.astype(np.float32))
    worst = 0.0
    for scale in (1e-3, 1.0, 37.0):
        th = scale * r.uniform(-1.5, 1.5, (Ks.shape[1], 4096))
        thf = th.astype(np.float32)
        b_op = trunc(thf)
        s_t = np.zeros((Ks.shape[0], th.shape[1]), np.float32)
        a_t = np.zeros_like(s_t)
        for k in range(Ks.shape[1]):                               # sequential FP32 accumulation (worse than the tensor core's)
            s_t += a_op[:, k:k + 1] * b_op[k:k + 1, :]
            a_t += np.abs(a_op[:, k:k + 1]) * np.abs(b_op[k:k + 1, :])
        err = np.abs(Ks @ th - s_t.astype(np.float64))
        margin = a_t.astype(np.float64) / 256.0
        ok = margin > 0
        worst = max(worst, float((err[ok] / margin[ok]).max()))
        assert (err <= margin).all()
    assert worst < 0.5, worst
