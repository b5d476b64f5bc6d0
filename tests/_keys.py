"""Host-side view of the sortable 64-bit hit keys the kernels exchange (csrc/psx_common.cuh).

``key = orderable(score) << 32 | ~uint32(id)``; 0 = empty slot.  A larger key is a better hit
(higher score, then lower id), so shards only need to exchange and integer-merge keys.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def encode(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """(fp32 'larger is better' scores, ids; id -1 = empty) -> uint64 keys."""
    s = np.ascontiguousarray(scores, np.float32).copy()
    s[np.isnan(s)] = -np.inf
    s = s + np.float32(0.0)
    u = s.view(np.uint32).astype(np.uint64)
    neg = (u >> np.uint64(31)).astype(bool)
    o = np.where(neg, (~u) & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    low = (~np.asarray(ids).astype(np.int64).astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    keys = (o << np.uint64(32)) | low
    return np.where(np.asarray(ids) < 0, np.uint64(0), keys)


def decode(keys: np.ndarray, metric_l2: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """uint64 keys -> (scores fp32 as the API reports them, ids int64)."""
    k = np.asarray(keys).astype(np.uint64)
    o = (k >> np.uint64(32)).astype(np.uint32)
    pos = (o >> np.uint32(31)).astype(bool)
    u = np.where(pos, o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    s = u.view(np.float32).copy()
    ids = ((~k) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    empty = k == 0
    s[empty] = -np.inf
    ids[empty] = -1
    if metric_l2:
        s = -s
    return s, ids
