"""ORACLE (test infrastructure) -- optional multi-threaded sparse x dense product for the CPU-baseline timing legs of
bench.py.  Falls back to SciPy (single-threaded) when the C piece (oracle/spmm_omp.c) has not been built."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_spmm.so")
        if not os.path.exists(path):
            _LIB = False
        else:
            lib = C.CDLL(path)
            lib.oracle_csr_spmm.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_void_p]
            lib.oracle_csr_spmm.restype = None
            lib.oracle_spmm_threads.restype = C.c_int
            _LIB = lib
    return _LIB


def threads() -> int:
    lib = _load()
    return int(lib.oracle_spmm_threads()) if lib else 1


class FastCSR:
    """C @ U for a scipy CSR matrix C with int32 indices, rows spread over OpenMP threads."""

    def __init__(self, Ccsr):
        self.n = Ccsr.shape[0]
        self.rowptr = np.ascontiguousarray(Ccsr.indptr, dtype=np.int32)
        self.col = np.ascontiguousarray(Ccsr.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(Ccsr.data, dtype=np.float64)
        self._C = Ccsr
        self.shape = Ccsr.shape

    def __matmul__(self, U):
        lib = _load()
        if not lib or U.ndim != 2:
            return self._C @ U
        U = np.ascontiguousarray(U, dtype=np.float64)
        out = np.empty((self.n, U.shape[1]))
        lib.oracle_csr_spmm(self.n, self.rowptr.ctypes.data, self.col.ctypes.data, self.val.ctypes.data, U.ctypes.data,
                            U.shape[1], out.ctypes.data)
        return out
