"""ctypes binding of libddmpc.so (include/ddmpc.h).

This is the reference-side stub a maintainer would add: the
reference is pure Python, so the "FFI" is ctypes.  There is NO fallback: if the
library has not been built the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libddmpc.so")

# status codes (include/ddmpc.h)
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_CONTROLLER_TYPE, ERR_SLACK_TYPE, ERR_ROBUST_PARAMS = range(6)
ERR_N_TOO_SMALL, ERR_NOT_PE, ERR_HORIZON, ERR_NOT_IMPLEMENTED, ERR_FACTORIZATION, ERR_HANKEL_WINDOW = range(6, 12)
NOMINAL, ROBUST = 0, 1
SLACK_NONE, SLACK_CONVEX, SLACK_NON_CONVEX = 0, 1, 2
SOLVE_OPTIMAL, SOLVE_OPTIMAL_INACCURATE, SOLVE_INFEASIBLE, SOLVE_NONFINITE = range(4)
# kernel selection of ddmpc_closed_loop_batch (ddmpc_set_option "closed_loop_path")
PATHS = {"auto": 0, "generic": 1, "fast": 2, "ws": 3, "perloop": 4, "dmma": 5, "gemm": 6, "cvx": 7, "tc": 8}
STATUS_STRINGS = {SOLVE_OPTIMAL: "optimal", SOLVE_OPTIMAL_INACCURATE: "optimal_inaccurate",
                  SOLVE_INFEASIBLE: "infeasible", SOLVE_NONFINITE: "solver_error"}


class Params(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("p", C.c_int32), ("N", C.c_int32), ("L", C.c_int32),
                ("controller_type", C.c_int32), ("slack_type", C.c_int32), ("use_terminal", C.c_int32),
                ("n_mpc_step", C.c_int32), ("check_pe", C.c_int32),
                ("eps_max", C.c_double), ("lamb_alpha", C.c_double), ("lamb_sigma", C.c_double), ("c", C.c_double),
                ("u_min", C.c_void_p), ("u_max", C.c_void_p),   # optional input box (host pointers, m values each)
                ("y_min", C.c_void_p), ("y_max", C.c_void_p)]   # optional output box (host pointers, p values each)


class Plant(C.Structure):
    _fields_ = [("n_x", C.c_int32), ("m", C.c_int32), ("p", C.c_int32),
                ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("D", C.c_void_p)]


class DDMPCError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libddmpc error {code}: {message}")
        self.code, self.message = code, message


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f64, u64, sz = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_size_t
    PP = C.POINTER(Params)
    sigs = {
        "ddmpc_version": (C.c_char_p, []),
        "ddmpc_strerror": (C.c_char_p, [i32]),
        "ddmpc_last_error": (C.c_char_p, []),
        "ddmpc_kernel_launches": (u64, []),
        "ddmpc_trim_memory": (i32, []),
        "ddmpc_probe_fp64_tflops": (i32, [i32, C.POINTER(f64), vp]),
        "ddmpc_probe_store_ms": (i32, [i32, i32, i32, vp, vp, C.POINTER(f64), vp]),
        "ddmpc_hankel": (i32, [vp, i32, i32, i32, vp, vp]),
        "ddmpc_hankel_host": (i32, [vp, i32, i32, i32, vp]),
        "ddmpc_pe_rank_host": (i32, [vp, i32, i32, i32, C.POINTER(C.c_int)]),
        "ddmpc_set_create": (i32, [PP, i32, vp, sz, vp, sz, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "ddmpc_set_create_host": (i32, [PP, i32, vp, sz, vp, sz, vp, vp, vp, vp, C.POINTER(vp)]),
        "ddmpc_set_destroy": (None, [vp]),
        "ddmpc_set_count": (i32, [vp]),
        "ddmpc_set_failed_count": (i32, [vp]),
        "ddmpc_set_option": (i32, [vp, C.c_char_p, i32]),
        "ddmpc_set_info": (i32, [vp, i32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ddmpc_set_get": (i32, [vp, C.c_char_p, i32, vp, sz, C.POINTER(sz)]),
        "ddmpc_solve_batch": (i32, [vp, i32, vp, vp, vp, vp, vp, f64, i32, vp, vp, vp, vp, vp]),
        "ddmpc_solve_batch_host": (i32, [vp, i32, vp, vp, vp, vp, vp, f64, i32, vp, vp, vp, vp]),
        "ddmpc_solve_full_batch": (i32, [vp, i32, vp, vp, vp, vp, vp, f64, i32, vp, vp, vp, vp, vp]),
        "ddmpc_closed_loop_batch": (i32, [vp, C.POINTER(Plant), i32, vp, vp, vp, vp, vp, vp, vp, u64, u64, f64,
                                          i32, f64, i32, vp, vp, vp, vp, vp, vp]),
        "ddmpc_generate_example_data": (i32, [C.POINTER(Plant), vp, vp, i32, vp, i32, f64, f64, f64, vp, vp, vp, vp, vp, vp]),
        "ddmpc_pcg64_uniform": (i32, [vp, i32, i32, f64, f64, f64, vp, vp]),
        "ddmpc_closed_loop_batch_host": (i32, [vp, C.POINTER(Plant), i32, vp, vp, vp, vp, vp, vp, vp, u64, u64, f64,
                                               i32, f64, i32, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError = ABI mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
EXPORTED = ["ddmpc_version", "ddmpc_strerror", "ddmpc_last_error", "ddmpc_kernel_launches", "ddmpc_hankel",
            "ddmpc_hankel_host", "ddmpc_pe_rank_host", "ddmpc_set_create", "ddmpc_set_create_host",
            "ddmpc_set_destroy", "ddmpc_set_count", "ddmpc_set_failed_count", "ddmpc_set_option", "ddmpc_set_info", "ddmpc_set_get", "ddmpc_solve_batch",
            "ddmpc_solve_batch_host", "ddmpc_solve_full_batch", "ddmpc_closed_loop_batch",
            "ddmpc_closed_loop_batch_host", "ddmpc_generate_example_data", "ddmpc_pcg64_uniform", "ddmpc_trim_memory", "ddmpc_probe_fp64_tflops",
            "ddmpc_probe_store_ms"]


def last_error() -> str:
    return lib.ddmpc_last_error().decode("utf-8", "replace")


def check(code: int) -> None:
    """Raise the exception the reference raises for this condition."""
    if code == OK:
        return
    msg = last_error() or lib.ddmpc_strerror(code).decode()
    if code == ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(msg)
    if code in (ERR_CONTROLLER_TYPE, ERR_SLACK_TYPE, ERR_ROBUST_PARAMS, ERR_N_TOO_SMALL, ERR_NOT_PE, ERR_HORIZON,
                ERR_HANKEL_WINDOW, ERR_INVALID_ARG, ERR_FACTORIZATION):
        raise ValueError(msg)
    raise DDMPCError(code, msg)


def kernel_launches() -> int:
    return int(lib.ddmpc_kernel_launches())


def probe_fp64_tflops(use_dmma: bool, stream: int = 0) -> float:
    """Sustained FP64 TFLOP/s of the current device issued as DFMA or as DMMA m8n8k4 (csrc/probes.cu)."""
    out = C.c_double()
    check(lib.ddmpc_probe_fp64_tflops(1 if use_dmma else 0, C.byref(out), stream))
    return out.value


def probe_store_ms(B: int, n_steps: int, coalesced: bool, u_ptr: int, y_ptr: int, stream: int = 0) -> float:
    """Time (ms) to write the (B, n_steps, 2) x 2 trajectory arrays with no compute (csrc/probes.cu)."""
    out = C.c_double()
    check(lib.ddmpc_probe_store_ms(B, n_steps, 1 if coalesced else 0, u_ptr, y_ptr, C.byref(out), stream))
    return out.value


def trim_memory() -> None:
    """Return the unused part of the library's device-memory pool to the driver (see ddmpc_trim_memory)."""
    check(lib.ddmpc_trim_memory())
