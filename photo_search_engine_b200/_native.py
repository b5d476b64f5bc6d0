"""ctypes binding of ``libpsx.so`` (C ABI in ``include/psx.h``).

There is deliberately no fallback: if the shared library is missing the import of the
product modules fails with ``ImportError`` (as the reference does when faiss is missing,
utils/vector_store.py:9-12), and if no B200 is usable ``psx_create`` fails and the wrapper
raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSX_LIB", os.path.join(_HERE, "libpsx.so"))

PSX_OK = 0
PSX_ERR_INVALID, PSX_ERR_CUDA, PSX_ERR_OOM, PSX_ERR_RANGE, PSX_ERR_STATE = -1, -2, -3, -4, -5
METRIC_IP, METRIC_L2 = 0, 1
STORE_F32, STORE_BF16, STORE_BF16_MASTER = 0, 1, 2
K_PASS_MAX = 2048

F_SEASON, F_PERIOD, F_YEAR, F_MONTH, F_NEED_DT, F_START, F_END = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40


class PsxFilter(C.Structure):
    """``psx_filter`` of include/psx.h."""

    _fields_ = [
        ("flags", C.c_uint32),
        ("season", C.c_uint32),
        ("period", C.c_uint32),
        ("year", C.c_uint32),
        ("month", C.c_uint32),
        ("reserved", C.c_uint32),
        ("start", C.c_uint64),
        ("end", C.c_uint64),
    ]


# every symbol include/psx.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGNATURES = [
    ("psx_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    ("psx_create_sharded", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(_P)]),
    ("psx_device_count", C.c_int, [_P]),
    ("psx_shard_rows", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_int]),
    ("psx_group_stats", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("psx_destroy", C.c_int, [_P]),
    ("psx_reset", C.c_int, [_P]),
    ("psx_ntotal", C.c_int64, [_P]),
    ("psx_dim", C.c_int, [_P]),
    ("psx_metric", C.c_int, [_P]),
    ("psx_last_error", C.c_char_p, []),
    ("psx_abi_version", C.c_int, []),
    ("psx_add", C.c_int, [_P, _P, C.c_int64]),
    ("psx_add_device", C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    ("psx_upload_gbps", C.c_double, [_P]),
    ("psx_reserve", C.c_int, [_P, C.c_int64]),
    ("psx_sync", C.c_int, [_P]),
    ("psx_set_attrs", C.c_int, [_P, C.c_int64, _P, C.c_int64]),
    ("psx_set_attrs_device", C.c_int, [_P, C.c_int64, _P, C.c_int64, _P]),
    ("psx_search", C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(PsxFilter), _P, _P]),
    ("psx_search_device", C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(PsxFilter), C.c_uint32, _P, _P, _P, _P]),
    ("psx_kpad", C.c_int64, [C.c_int64]),
    ("psx_search_batch_device", C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(PsxFilter), C.c_float, C.c_uint32, _P, _P, _P, _P, _P]),
    ("psx_batch_supported", C.c_int, [_P, C.c_int64]),
    ("psx_batch_stats", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("psx_merge_keys_device", C.c_int, [C.c_int, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P, _P, _P]),
    ("psx_exchange_bytes", C.c_int64, []),
    ("psx_search_exchange_device", C.c_int, [_P, _P, C.c_int64, C.POINTER(PsxFilter), C.c_uint32, C.c_int, C.c_int, _P, C.c_uint32,
                                              C.c_int, _P, _P, _P]),
    ("psx_exchange_status", C.c_int, [_P, C.POINTER(C.c_int)]),
    ("psx_hybrid_fuse_device", C.c_int, [C.c_int, C.c_int64, C.c_int64, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_double, C.c_double,
                                          C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    ("psx_finalize_device", C.c_int, [C.c_int, C.c_int64, C.c_int64, _P, _P, C.c_int, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P]),
    ("psx_reconstruct", C.c_int, [_P, C.c_int64, _P]),
    ("psx_read_rows", C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    ("psx_storage_device", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("psx_set_tunable", C.c_int, [_P, C.c_char_p, C.c_int]),
    ("psx_set_trace_device", C.c_int, [_P, _P]),
    ("psx_launch_count", C.c_int64, []),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen ``libpsx.so`` and type every entry point.  Raises ImportError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"native library {LIB_PATH} not found: build it with `python -m photo_search_engine_b200.build` "
            "(this engine has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)  # AttributeError -> the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.psx_abi_version() != 2:
        raise ImportError("libpsx.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return (load_library().psx_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a psx_status to the exception classes the reference surface uses."""
    if rc == PSX_OK:
        return
    msg = last_error()
    if rc in (PSX_ERR_INVALID, PSX_ERR_RANGE):
        raise ValueError(msg)
    if rc == PSX_ERR_OOM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class NativeIndex:
    """Thin RAII wrapper of one ``psx_index`` handle: one GPU (``device`` an int), or several GPUs of the box behind
    the one handle (``device`` a sequence of ordinals, ``psx_create_sharded``; ``device[0]`` merges)."""

    def __init__(self, d: int, metric: int = METRIC_IP, store_dtype: int = STORE_F32, device=0) -> None:
        self._lib = load_library()
        self._h = _P()
        devices = [int(device)] if isinstance(device, (int, np.integer)) else [int(x) for x in device]
        if not devices:
            raise ValueError("at least one device is needed")
        if len(devices) == 1:
            check(self._lib.psx_create(int(d), int(metric), int(store_dtype), devices[0], C.byref(self._h)))
        else:
            arr = (C.c_int * len(devices))(*devices)
            check(self._lib.psx_create_sharded(int(d), int(metric), int(store_dtype), len(devices), arr, C.byref(self._h)))
        self.d = int(d)
        self.metric = int(metric)
        self.store_dtype = int(store_dtype)
        self.device = devices[0]
        self.devices = tuple(devices)

    # -- lifecycle ---------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.psx_destroy(self._h)
            self._h = _P()

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        return int(self._lib.psx_ntotal(self._h))

    def reset(self) -> None:
        check(self._lib.psx_reset(self._h))

    # -- write side --------------------------------------------------------------------
    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"向量维度不匹配: {x.shape[-1]} != {self.d}")
        check(self._lib.psx_add(self._h, _ptr(x), x.shape[0]))

    def add_device(self, ptr: int, n: int, normalize: bool = False, stream: int = 0) -> None:
        check(self._lib.psx_add_device(self._h, ptr, int(n), int(bool(normalize)), stream or None))

    def upload_gbps(self) -> float:
        return float(self._lib.psx_upload_gbps(self._h))

    def reserve(self, n: int) -> None:
        check(self._lib.psx_reserve(self._h, int(n)))

    def sync(self) -> None:
        check(self._lib.psx_sync(self._h))

    def set_attrs(self, row0: int, attrs: np.ndarray) -> None:
        attrs = np.ascontiguousarray(attrs, dtype=np.uint64)
        check(self._lib.psx_set_attrs(self._h, int(row0), _ptr(attrs), attrs.shape[0]))

    def set_attrs_device(self, row0: int, ptr: int, n: int, stream: int = 0) -> None:
        check(self._lib.psx_set_attrs_device(self._h, int(row0), ptr, int(n), stream or None))

    # -- read side ---------------------------------------------------------------------
    def search(self, q: np.ndarray, k: int, flt: Optional[PsxFilter] = None):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"向量维度不匹配: {q.shape[-1]} != {self.d}")
        nq = q.shape[0]
        scores = np.empty((nq, k), np.float32)
        ids = np.empty((nq, k), np.int64)
        fp = C.byref(flt) if flt is not None else None
        check(self._lib.psx_search(self._h, _ptr(q), nq, int(k), fp, _ptr(scores), _ptr(ids)))
        return scores, ids

    def search_device(self, q_ptr: int, nq: int, k: int, out_scores_ptr: int, out_ids_ptr: int, out_keys_ptr: int = 0,
                      flt: Optional[PsxFilter] = None, id_base: int = 0, stream: int = 0) -> None:
        fp = C.byref(flt) if flt is not None else None
        check(self._lib.psx_search_device(self._h, q_ptr, int(nq), int(k), fp, int(id_base), out_scores_ptr or None,
                                          out_ids_ptr or None, out_keys_ptr or None, stream or None))

    def search_exchange_device(self, q_ptr: int, k: int, rank: int, world: int, peer_bases: np.ndarray, seq: int,
                               out_scores_ptr: int, out_ids_ptr: int, flt: Optional[PsxFilter] = None, id_base: int = 0,
                               stream: int = 0, phases: int = 3) -> None:
        fp = C.byref(flt) if flt is not None else None
        check(self._lib.psx_search_exchange_device(self._h, q_ptr or None, int(k), fp, int(id_base), int(rank), int(world),
                                                   peer_bases.ctypes.data, int(seq), int(phases), out_scores_ptr or None,
                                                   out_ids_ptr or None, stream or None))

    def exchange_status(self) -> int:
        """0, or 1 + the rank whose keys never arrived in a fused exchange since the last call (clears the word)."""
        v = C.c_int()
        check(self._lib.psx_exchange_status(self._h, C.byref(v)))
        return v.value

    def search_batch_device(self, q_ptr: int, nq: int, k: int, out_scores_ptr: int, out_ids_ptr: int, flags_ptr: int,
                            out_keys_ptr: int = 0, qnorm_max: float = 1.0, id_base: int = 0, stream: int = 0,
                            flt: Optional[PsxFilter] = None) -> None:
        fp = C.byref(flt) if flt is not None else None
        check(self._lib.psx_search_batch_device(self._h, q_ptr, int(nq), int(k), fp, float(qnorm_max), int(id_base), out_scores_ptr,
                                                out_ids_ptr, out_keys_ptr or None, flags_ptr, stream or None))

    def shard_rows(self):
        """``[(device, rows), ...]`` of the shards behind the handle."""
        n = int(self._lib.psx_device_count(self._h))
        rows, devs = (C.c_int64 * n)(), (C.c_int * n)()
        check(self._lib.psx_shard_rows(self._h, rows, devs, n))
        return [(int(devs[i]), int(rows[i])) for i in range(n)]

    def group_stats(self):
        """multi-device handles: (queries through the fused exchange, through key lists, fused time-outs)."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.psx_group_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def batch_supported(self, k: int) -> bool:
        return bool(self._lib.psx_batch_supported(self._h, int(k)))

    def batch_stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(self._lib.psx_batch_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reconstruct(self, i: int) -> np.ndarray:
        out = np.empty(self.d, np.float32)
        check(self._lib.psx_reconstruct(self._h, int(i), _ptr(out)))
        return out

    def read_rows(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((int(n), self.d), np.float32)
        check(self._lib.psx_read_rows(self._h, int(row0), int(n), _ptr(out)))
        return out

    def storage_device(self):
        p, ld, dt = _P(), C.c_int64(), C.c_int()
        check(self._lib.psx_storage_device(self._h, C.byref(p), C.byref(ld), C.byref(dt)))
        return p.value or 0, ld.value, dt.value

    def set_trace_device(self, ptr: int) -> None:
        check(self._lib.psx_set_trace_device(self._h, ptr or None))

    def set_tunable(self, key: str, value: int) -> None:
        check(self._lib.psx_set_tunable(self._h, key.encode(), int(value)))


def kpad(k: int) -> int:
    return int(load_library().psx_kpad(int(k)))


def merge_keys_device(device: int, keys_ptr: int, nq: int, nlists: int, k: int, metric: int, out_scores_ptr: int,
                      out_ids_ptr: int, stream: int = 0) -> None:
    check(load_library().psx_merge_keys_device(int(device), keys_ptr, int(nq), int(nlists), int(k), int(metric),
                                               out_scores_ptr, out_ids_ptr, stream or None))


def exchange_bytes() -> int:
    return int(load_library().psx_exchange_bytes())


def launch_count() -> int:
    return int(load_library().psx_launch_count())
