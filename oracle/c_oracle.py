"""ctypes access to oracle/flat_scan.c (built by oracle/Makefile).  TEST / BASELINE
INFRASTRUCTURE ONLY -- see the header of flat_scan.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class OracleFilter(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("season", C.c_uint32), ("period", C.c_uint32), ("year", C.c_uint32),
                ("month", C.c_uint32), ("reserved", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "flat_scan.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        # -march=native: the library is rebuilt on the machine that runs it when the ISA differs
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            build()
            _lib = C.CDLL(LIB)
            _lib.oracle_max_threads()
            _lib.oracle_set_threads
        except Exception:
            build(force=True)
            _lib = C.CDLL(LIB)
        _lib.oracle_flat_search.restype = C.c_int
        _lib.oracle_flat_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_fill_unit_rows.restype = None
        _lib.oracle_fill_unit_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64]
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_set_threads.restype = None
        _lib.oracle_set_threads.argtypes = [C.c_int]
    return _lib


def host_cores() -> int:
    """Cores this process may run on -- NOT omp_get_max_threads(), which follows OMP_NUM_THREADS (torchrun sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def search(x: np.ndarray, q: np.ndarray, k: int, metric: int = 0, attrs: Optional[np.ndarray] = None,
           flt=None, nthreads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    x = np.ascontiguousarray(x, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    if q.ndim == 1:
        q = q[None]
    nq, d = q.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    f = None
    if flt is not None:
        f = OracleFilter(*(getattr(flt, name) for name, _ in OracleFilter._fields_))
    if attrs is not None:
        attrs = np.ascontiguousarray(attrs, np.uint64)
    rc = lib().oracle_flat_search(x.ctypes.data, x.shape[0], d, q.ctypes.data, nq, int(k), int(metric),
                                  attrs.ctypes.data if attrs is not None else None,
                                  C.addressof(f) if f is not None else None, D.ctypes.data, I.ctypes.data, int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_flat_search failed")
    return D, I


def fill_unit_rows(n: int, d: int, seed: int = 1) -> np.ndarray:
    x = np.empty((n, d), np.float32)
    lib().oracle_fill_unit_rows(x.ctypes.data, n, d, seed)
    return x
