"""Hankel matrix and persistency-of-excitation test on the GPU.

Host-facing mirror of the reference's
``direct_data_driven_mpc/utilities/hankel_matrix.py`` (same names, arguments
and errors); the work is done by ``ddmpc_hankel_host`` / ``ddmpc_pe_rank_host``.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib


def _as_f64(X: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(X, dtype=np.float64)


def hankel_matrix(X: np.ndarray, L: int) -> np.ndarray:
    """hankel_matrix.py:5-53: (N, n) data -> (L*n, N-L+1) matrix whose column i
    is the flattened window X[i:i+L].  Bit-exact copy kernel (K1)."""
    N, n = X.shape
    if N < L:
        raise ValueError("N must be greater than or equal to L.")     # hankel_matrix.py:43-44
    Xc = _as_f64(X)
    HL = np.empty((L * n, N - L + 1), dtype=np.float64)
    _lib.check(_lib.lib.ddmpc_hankel_host(Xc.ctypes.data, N, n, L, HL.ctypes.data))
    return HL


def evaluate_persistent_excitation(X: np.ndarray, order: int) -> Tuple[int, bool]:
    """hankel_matrix.py:55-87: (rank of H_order(X), rank == n*order)."""
    N, n = X.shape
    if N < order:
        raise ValueError("N must be greater than or equal to L.")
    Xc = _as_f64(X)
    rank = C.c_int(-1)
    _lib.check(_lib.lib.ddmpc_pe_rank_host(Xc.ctypes.data, N, n, order, C.byref(rank)))
    return int(rank.value), bool(rank.value == n * order)
