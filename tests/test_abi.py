"""CPU: the C-ABI library loads and exports every symbol include/ddmpc.h declares
(no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddmpc.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddmpc_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("ddmpc_hankel", "ddmpc_set_create", "ddmpc_solve_batch", "ddmpc_closed_loop_batch",
                 "ddmpc_set_destroy", "ddmpc_pe_rank_host"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from direct_data_driven_mpc_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in include/ddmpc.h but not exported"
    assert sorted(_lib.EXPORTED) == declared_symbols()
    assert b"sm_100a" in _lib.lib.ddmpc_version()
    assert _lib.lib.ddmpc_strerror(_lib.ERR_NOT_PE) == b"input data not persistently exciting"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from direct_data_driven_mpc_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib._load()


def test_error_code_to_exception_mapping():
    from direct_data_driven_mpc_b200 import _lib
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_NOT_IMPLEMENTED)
    for code in (_lib.ERR_NOT_PE, _lib.ERR_N_TOO_SMALL, _lib.ERR_HORIZON, _lib.ERR_ROBUST_PARAMS,
                 _lib.ERR_CONTROLLER_TYPE, _lib.ERR_SLACK_TYPE, _lib.ERR_HANKEL_WINDOW):
        with pytest.raises(ValueError):
            _lib.check(code)
    with pytest.raises(_lib.DDMPCError):
        _lib.check(_lib.ERR_CUDA)
    _lib.check(_lib.OK)
