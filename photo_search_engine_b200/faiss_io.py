"""FAISS-compatible persistence of the flat index (``photo_search.index``).

The reference persists with ``faiss.write_index`` / ``faiss.read_index``
(utils/vector_store.py:234, :249).  FAISS is not a dependency of this package, so the
container format is implemented here:

flat (``IxFI`` inner product, ``IxF2`` L2)::

    fourcc | int32 d | int64 ntotal | int64 1<<20 | int64 1<<20 | uint8 is_trained |
    int32 metric_type | uint64 d*ntotal | float32[d*ntotal]

HNSW (``IHNf``, what the shipped env templates select, .env.example:82): index header as
above, then the graph vectors (each ``uint64 n`` + payload), five int32 scalars and a nested
flat block.  The graph is skipped on load -- rows are served by the exact GPU scan, a recall
superset of the HNSW walk -- and never written (``save()`` always writes the flat block).
"""
from __future__ import annotations

import os
import struct
from typing import Callable, Dict, Iterator, Tuple

import numpy as np

HEADER = struct.Struct("<iqqqBi")
FOURCC_BY_METRIC = {0: b"IxFI", 1: b"IxF2"}
_GRAPH_VECTORS = (8, 4, 4, 8, 4)  # assign_probas f64, cum_nneighbor i32, levels i32, offsets u64, neighbors i32


def write_flat_index(path: str, d: int, metric: int, ntotal: int,
                     read_rows: Callable[[int, int], np.ndarray], chunk_rows: int = 1 << 16) -> None:
    """Stream ``ntotal`` rows (``read_rows(row0, n) -> float32 [n,d]``) into a FAISS flat file."""
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(FOURCC_BY_METRIC[metric])
        f.write(HEADER.pack(d, ntotal, 1 << 20, 1 << 20, 1, metric))
        f.write(struct.pack("<Q", d * ntotal))
        for row0 in range(0, ntotal, chunk_rows):
            n = min(chunk_rows, ntotal - row0)
            f.write(np.ascontiguousarray(read_rows(row0, n), dtype="<f4").tobytes())
    os.replace(tmp, path)


def _header(f) -> Tuple[int, int, int]:
    d, ntotal, _a, _b, _trained, metric = HEADER.unpack(f.read(HEADER.size))
    if metric > 1:
        f.read(4)  # metric_arg, only written for the exotic metrics
    return d, ntotal, metric


def open_index(path: str) -> Tuple[Dict[str, int], Iterator[np.ndarray]]:
    """Parse the container and return ``(info, chunks)``: ``info`` has d / ntotal / metric /
    is_hnsw, ``chunks`` yields float32 ``[n,d]`` blocks in id order (memory mapped)."""
    with open(path, "rb") as f:
        fourcc = f.read(4)
        is_hnsw = fourcc == b"IHNf"
        if is_hnsw:
            _header(f)
            for itemsize in _GRAPH_VECTORS:
                (n,) = struct.unpack("<Q", f.read(8))
                f.seek(n * itemsize, os.SEEK_CUR)
            f.seek(20, os.SEEK_CUR)  # entry_point, max_level, efConstruction, efSearch, upper_beam
            fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2"):
            raise ValueError(f"不支持的索引文件格式: {fourcc!r}")
        d, ntotal, metric = _header(f)
        (count,) = struct.unpack("<Q", f.read(8))
        offset = f.tell()
    if count != d * ntotal or os.path.getsize(path) < offset + 4 * count:
        raise ValueError("索引文件损坏，请重新构建索引")
    info = {"d": d, "ntotal": ntotal, "metric": metric, "is_hnsw": int(is_hnsw)}

    def chunks(chunk_rows: int = 0) -> Iterator[np.ndarray]:
        if ntotal == 0:
            return
        if chunk_rows <= 0:  # ~1 GB per block: the native upload pipelines inside a block
            chunk_rows = max(1, (1 << 30) // (4 * max(d, 1)))
        mm = np.memmap(path, dtype="<f4", mode="r", offset=offset, shape=(ntotal, d))
        for row0 in range(0, ntotal, chunk_rows):
            yield np.asarray(mm[row0 : row0 + chunk_rows])

    return info, chunks()
