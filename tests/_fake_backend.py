"""Oracle-backed stand-in for ``_native.NativeIndex`` -- test infrastructure only.

It lets the ``-m "not gpu"`` suite drive the *host* logic of the drop-in ``VectorStore``
(metadata bookkeeping, persistence, filter packing, error behaviour) on a machine without a
GPU.  The product never imports this; its scores come from ``oracle.flat_ip``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from oracle.flat_ip import OracleIndexFlat
from photo_search_engine_b200 import _native as N


def attr_pass_np(words: np.ndarray, f: N.PsxFilter) -> np.ndarray:
    """numpy restatement of ``attr_pass`` (csrc/psx_scan.cuh) over packed attribute words."""
    from photo_search_engine_b200.exif_attrs import words_pass

    return words_pass(words, f)


class FakeIndex:
    def __init__(self, d: int, metric: int = 0, store_dtype: int = 0, device=0) -> None:
        self.devices = (int(device),) if isinstance(device, (int, np.integer)) else tuple(int(x) for x in device)
        self.d, self.metric, self.store_dtype, self.device = int(d), int(metric), int(store_dtype), self.devices[0]
        self._ix = OracleIndexFlat(self.d, self.metric)
        self._attrs = np.zeros(0, np.uint64)

    @property
    def ntotal(self) -> int:
        return self._ix.ntotal

    def close(self) -> None:
        pass

    def reset(self) -> None:
        self._ix = OracleIndexFlat(self.d, self.metric)
        self._attrs = np.zeros(0, np.uint64)

    def reserve(self, n: int) -> None:
        pass

    def sync(self) -> None:
        pass

    def set_tunable(self, key: str, value: int) -> None:
        pass

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 1:
            x = x[None]
        self._ix.add(x)

    def set_attrs(self, row0: int, attrs: np.ndarray) -> None:
        need = row0 + len(attrs)
        if need > self.ntotal:
            raise ValueError("attribute rows exceed ntotal")
        if len(self._attrs) < need:
            self._attrs = np.concatenate([self._attrs, np.zeros(need - len(self._attrs), np.uint64)])
        self._attrs[row0:need] = attrs

    def search(self, q: np.ndarray, k: int, flt: Optional[N.PsxFilter] = None):
        mask = None
        if flt is not None and flt.flags:
            words = np.zeros(self.ntotal, np.uint64)
            words[: len(self._attrs)] = self._attrs
            mask = attr_pass_np(words, flt)
        return self._ix.search(np.asarray(q, np.float32), int(k), mask=mask)

    def reconstruct(self, i: int) -> np.ndarray:
        if not 0 <= i < self.ntotal:
            raise ValueError("id out of range")
        return self._ix.reconstruct(i)

    def read_rows(self, row0: int, n: int) -> np.ndarray:
        return self._ix._matrix()[row0 : row0 + n].copy()
