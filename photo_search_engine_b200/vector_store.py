"""Drop-in ``VectorStore`` whose arithmetic runs on a B200.

Surface contract: the constructor keywords, public attributes, method
names, return shapes and exception types of the reference class
``utils/vector_store.py::VectorStore`` (a FAISS CPU ``IndexFlatIP`` / ``IndexFlatL2`` wrapper),
so ``main.py:59-68``, ``core/searcher.py`` and ``core/indexer.py`` run unchanged once
``utils/vector_store.py`` re-exports this class.  Below the class nothing is
shared with the reference: rows live in HBM behind ``libpsx.so`` (include/psx.h, ctypes) and a
search is one launch of the TMA-fed streaming scan with the EXIF predicate and the top-k
selection fused in.  There is no CPU fallback; importing this module without the native
library raises ``ImportError`` exactly as the reference does without faiss
(utils/vector_store.py:9-12).

Additive, optional surface (defaults reproduce the reference):

* keyword-only ctor arguments ``device`` / ``store_dtype`` (env ``PSX_DEVICE`` /
  ``PSX_STORE_DTYPE``): ``fp32`` (default), ``bf16`` (half the HBM, approximate), ``bf16+fp32`` (bf16
  rows for the scan plus an fp32 master: results bit-identical to ``fp32`` at about half the bytes
  streamed per query);
* ``devices=[0, 1, ...]`` (env ``PSX_DEVICES=0,1,...``): the corpus is row-sharded over several GPUs of the
  box BEHIND this one instance -- the reference constructs a single store in a single process
  (main.py:59-68) -- with results bit-identical to one device (``psx_create_sharded``, include/psx.h);
* ``search(..., constraints=None)``: fused pre-filter with the semantics of
  ``Searcher._check_time_match_v2`` (core/searcher.py:1884-1950);
* ``search_batch`` / ``add_batch``: arrays in, arrays out, no per-hit Python objects;
* ``coalesce=True`` (env ``PSX_COALESCE=1``): concurrent ``search`` calls of the threaded server share one
  batched search while the GPU is busy (``coalesce.py``); same results, no added latency when idle.
"""
from __future__ import annotations

import json
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native, faiss_io
from .coalesce import SearchCoalescer
from .exif_attrs import attr_words, build_filter

_native.load_library()

_METRICS = ("l2", "cosine")
_INDEX_TYPES = ("flat", "hnsw")
_DTYPES = {"fp32": "fp32", "float32": "fp32", "bf16": "bf16", "bfloat16": "bf16", "bf16+fp32": "bf16+fp32", "mixed": "bf16+fp32"}
_DTYPE_CODES = {"fp32": _native.STORE_F32, "bf16": _native.STORE_BF16, "bf16+fp32": _native.STORE_BF16_MASTER}

# messages are part of the observable behaviour (they surface in HTTP 500 bodies, api/routes.py:196-207)
_E_METRIC = "metric仅支持l2或cosine"
_E_INDEX_TYPE = "index_type仅支持flat或hnsw"
_E_EMPTY_VECTOR = "向量不能为空"
_E_DIM = "向量维度不匹配: {} != {}"
_E_NO_INDEX = "索引未初始化"
_E_META_MISSING = "索引元信息缺失，请重新构建索引"
_E_META_BROKEN = "索引元信息损坏，请重新构建索引"
_E_TYPE_MISMATCH = "索引类型与配置不一致，请重新构建索引"
_E_METRIC_MISMATCH = "索引度量与配置不一致，请重新构建索引"
_E_COUNT_MISMATCH = "索引与元数据数量不一致，请重新构建索引"


def _native_index(dimension: int, metric: int, store_dtype: int, device: int):
    return _native.NativeIndex(dimension, metric, store_dtype, device)


def _unit_rows(rows: np.ndarray) -> np.ndarray:
    """Row-wise L2 normalisation in fp32; zero rows stay as they are."""
    norms = np.linalg.norm(rows, axis=1, keepdims=True).astype(np.float32)
    norms[norms == 0] = 1.0
    return (rows / norms).astype(np.float32)


class VectorStore:
    """Vector storage and exact retrieval on the GPU.

    Attributes mirrored from the reference: ``dimension``, ``index_path``, ``metadata_path``,
    ``meta_path``, ``metric``, ``index_type``, ``hnsw_*``, ``metadata`` (the very dict objects
    handed to ``add_item``, in insertion order) and ``index`` (the backend handle or ``None``).
    """

    # tests replace this with an oracle-backed fake to exercise the host logic without a GPU
    _index_factory: Callable[[int, int, int, int], Any] = staticmethod(_native_index)

    def __init__(
        self,
        dimension: Optional[int],
        index_path: str,
        metadata_path: str,
        metric: str = "cosine",
        index_type: str = "flat",
        hnsw_m: int = 32,
        hnsw_ef_construction: int = 200,
        hnsw_ef_search: int = 96,
        *,
        device: Optional[int] = None,
        devices: Optional[Sequence[int]] = None,
        store_dtype: Optional[str] = None,
        coalesce: Optional[bool] = None,
    ) -> None:
        # argument handling of utils/vector_store.py:44-62
        metric_name = metric.lower().strip() if metric else "l2"
        if metric_name not in _METRICS:
            raise ValueError(_E_METRIC)
        type_name = (index_type or "flat").strip().lower()
        if type_name not in _INDEX_TYPES:
            raise ValueError(_E_INDEX_TYPE)
        dtype_name = _DTYPES.get((store_dtype or os.environ.get("PSX_STORE_DTYPE", "fp32")).strip().lower())
        if dtype_name is None:
            raise ValueError("store_dtype仅支持fp32、bf16或bf16+fp32")

        self.dimension = dimension
        self.index_path = index_path
        self.metadata_path = metadata_path
        self.meta_path = index_path + ".meta.json"
        self.metric = metric_name
        self.index_type = type_name
        self.hnsw_m = max(4, int(hnsw_m))
        self.hnsw_ef_construction = max(8, int(hnsw_ef_construction))
        self.hnsw_ef_search = max(8, int(hnsw_ef_search))
        if devices is None and device is None and os.environ.get("PSX_DEVICES", "").strip():
            devices = [int(tok) for tok in os.environ["PSX_DEVICES"].replace(";", ",").split(",") if tok.strip()]
        if devices is not None and len(devices) == 0:
            raise ValueError("devices不能为空")
        self.devices = tuple(int(x) for x in devices) if devices is not None else None
        if self.devices is not None:
            self.device = self.devices[0]
        else:
            self.device = int(os.environ.get("PSX_DEVICE", "0")) if device is None else int(device)
        self.store_dtype = dtype_name

        self._normalize = metric_name == "cosine"
        # load() / clear() swap index + metadata as a pair under this lock; readers take the pair under it too.  The
        # old backend handle is never destroyed explicitly: a search running on another request thread still holds a
        # reference, and the handle is freed when the last reference goes (as with the reference's faiss object).
        self._swap_lock = threading.RLock()
        self._attrs_lock = threading.Lock()  # the lazy EXIF sidecar is built once, whichever request thread needs it first
        self.metadata: List[Dict] = []
        self._embeddings: List[Optional[List[float]]] = []
        self._path_to_index: Dict[str, int] = {}
        self._attrs_built = 0  # rows whose attribute word is already on the device
        self._attr_words = np.zeros(0, np.uint64)  # host copy of those words (persisted by save())
        self.index = self._create_index(dimension) if dimension else None
        # opt-in: concurrent search() calls (Flask's request threads) share one batched search
        if coalesce is None:
            coalesce = os.environ.get("PSX_COALESCE", "0").strip() not in ("", "0", "false", "no")
        self._coalescer = SearchCoalescer(self._run_search) if coalesce else None

    # ------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------
    @property
    def _metric_code(self) -> int:
        return _native.METRIC_IP if self._normalize else _native.METRIC_L2

    def _create_index(self, dimension: int):
        """Backend for ``dimension`` (utils/vector_store.py:72-81); ``hnsw`` is served exactly."""
        dtype = _DTYPE_CODES[self.store_dtype]
        where = self.devices if self.devices is not None and len(self.devices) > 1 else self.device
        index = type(self)._index_factory(int(dimension), self._metric_code, dtype, where)
        min_rows = os.environ.get("PSX_SHARD_MIN_ROWS", "").strip()
        if min_rows and hasattr(index, "set_tunable") and self.devices is not None and len(self.devices) > 1:
            index.set_tunable("shard_min_rows", int(min_rows))
        return index

    def _require_dimension(self, vector: Sequence[float]) -> None:
        if len(vector) != self.dimension:
            raise ValueError(_E_DIM.format(len(vector), self.dimension))

    def _normalize_vector(self, vector: Sequence[float]):
        """utils/vector_store.py:83-90 -- the same numpy operations in the same order, so the
        stored bits equal what the reference hands to FAISS."""
        if self._normalize:
            as_f32 = np.array(vector, dtype="float32")
            length = np.linalg.norm(as_f32)
            if length != 0:
                return (as_f32 / length).astype("float32").tolist()
        return vector

    def _query_row(self, vector: Sequence[float]) -> np.ndarray:
        """The ``(1, d)`` float32 query ``search`` hands to the backend: the same numpy operations as
        ``_normalize_vector`` (utils/vector_store.py:83-90) without the detour through a Python list -- a float32 value
        survives list -> float32 unchanged, so the bits are the ones the reference would send."""
        as_f32 = np.array(vector, dtype="float32")
        if self._normalize:
            length = np.linalg.norm(as_f32)
            if length != 0:
                as_f32 = (as_f32 / length).astype("float32")
        return as_f32[None, :]

    def _remember_path(self, metadata: Dict, row: int) -> None:
        photo_path = metadata.get("photo_path")
        if isinstance(photo_path, str) and photo_path:
            self._path_to_index[photo_path] = row  # the last duplicate wins, as in the reference

    def _filter_for(self, constraints: Optional[Dict[str, Any]]):
        """-> (psx_filter | None, never_matches)."""
        if not constraints:
            return None, False
        flt, never = build_filter(constraints)
        if flt is not None and not never:
            self._upload_attrs()
        return flt, never

    @property
    def attrs_path(self) -> str:
        """Columnar EXIF sidecar next to the index file: one little-endian uint64 per row (layout:
        include/psx.h).  Written by ``save()`` once the words exist, picked up by ``load()``; purely a
        cache of what ``metadata.json`` says -- when it is missing or stale the words are re-packed."""
        return self.index_path + ".attrs"

    def _upload_attrs(self) -> None:
        """Lazy EXIF sidecar: pack and upload attribute words for rows that lack one."""
        with self._attrs_lock:
            total = len(self.metadata)
            if self.index is None or self._attrs_built >= total:
                return
            words = attr_words(self.metadata[self._attrs_built : total])
            self.index.set_attrs(self._attrs_built, words)
            self._attr_words = np.concatenate([self._attr_words[: self._attrs_built], words])
            self._attrs_built = total

    def _run_search(self, queries: np.ndarray, k: int, flt) -> Tuple[np.ndarray, np.ndarray]:
        if flt is None:
            return self.index.search(queries, k)
        return self.index.search(queries, k, flt)

    # ------------------------------------------------------------------------------------
    # write side
    # ------------------------------------------------------------------------------------
    def add_item(self, embedding: List[float], metadata: Dict) -> None:
        """Append one vector with its metadata record (utils/vector_store.py:143-169).

        Raises ``ValueError`` for a ``None`` embedding or a dimension mismatch.  The first call
        on a store created with ``dimension=None`` fixes the dimension.
        """
        if embedding is None:
            raise ValueError(_E_EMPTY_VECTOR)
        if self.index is None:
            self.dimension = len(embedding)
            self.index = self._create_index(self.dimension)
        self._require_dimension(embedding)

        stored = self._normalize_vector(embedding)
        self.index.add(np.array([stored], dtype="float32"))
        row = len(self.metadata)
        self.metadata.append(metadata)
        # Normalised rows are bit-identical to what sits in HBM, so they are re-read on demand
        # instead of holding a Python list per row; raw (l2) inputs are kept like the reference.
        self._embeddings.append(None if self._normalize else stored)
        self._remember_path(metadata, row)

    def add_batch(self, embeddings: np.ndarray, metadatas: Sequence[Dict]) -> None:
        """Array form of ``add_item``: float32 ``[n, d]`` plus one metadata dict per row."""
        rows = np.ascontiguousarray(embeddings, dtype=np.float32)
        if rows.ndim != 2 or len(metadatas) != rows.shape[0]:
            raise ValueError("embeddings 与 metadatas 数量不一致")
        if self.index is None:
            self.dimension = int(rows.shape[1])
            self.index = self._create_index(self.dimension)
        if rows.shape[1] != self.dimension:
            raise ValueError(_E_DIM.format(rows.shape[1], self.dimension))
        self.index.add(_unit_rows(rows) if self._normalize else rows)
        first = len(self.metadata)
        self.metadata.extend(metadatas)
        self._embeddings.extend([None] * len(metadatas))
        for offset, metadata in enumerate(metadatas):
            self._remember_path(metadata, first + offset)

    # ------------------------------------------------------------------------------------
    # read side
    # ------------------------------------------------------------------------------------
    def search(self, query_embedding: List[float], top_k: int, constraints: Optional[Dict[str, Any]] = None) -> List[Dict]:
        """Top-k retrieval (utils/vector_store.py:172-198).

        Returns ``[{"metadata": <the stored dict>, "distance": float}, ...]`` best first -- inner
        product for cosine (higher is better), squared L2 for l2 -- with ``k = min(top_k,
        ntotal)``; ``[]`` on an empty store.  ``constraints`` (not part of the reference
        signature) limits the scan to rows passing the EXIF predicate.
        """
        with self._swap_lock:
            index, records = self.index, self.metadata
        if index is None or index.ntotal == 0:
            return []
        self._require_dimension(query_embedding)
        k = min(int(top_k), index.ntotal)
        if k <= 0:
            return []
        flt, never = self._filter_for(constraints)
        if never:
            return []
        query = self._query_row(query_embedding)
        if self._coalescer is not None:
            row_scores, row_labels = self._coalescer.submit(query[0], k, None if flt is None else bytes(flt), flt)
        else:
            distances, labels = index.search(query, k) if flt is None else index.search(query, k, flt)
            row_scores, row_labels = distances[0], labels[0]
        return [
            {"metadata": records[label], "distance": float(distance)}
            for distance, label in zip(row_scores.tolist(), row_labels.tolist())
            if label != -1 and label < len(records)
        ]

    def search_batch(self, queries: np.ndarray, top_k: int, constraints: Optional[Dict[str, Any]] = None,
                     normalize: Optional[bool] = None) -> Tuple[np.ndarray, np.ndarray]:
        """FAISS-shaped batch search: ``(scores float32 [nq,k], ids int64 [nq,k])``, unfilled
        slots ``(-inf | +inf, -1)``."""
        batch = np.ascontiguousarray(queries, dtype=np.float32)
        if batch.ndim == 1:
            batch = batch[None, :]
        k = max(int(top_k), 0)
        blank = (np.full((batch.shape[0], k), np.inf if self.metric == "l2" else -np.inf, np.float32),
                 np.full((batch.shape[0], k), -1, np.int64))
        if self.index is None or self.index.ntotal == 0 or k == 0:
            return blank
        if batch.shape[1] != self.dimension:
            raise ValueError(_E_DIM.format(batch.shape[1], self.dimension))
        flt, never = self._filter_for(constraints)
        if never:
            return blank
        if self._normalize if normalize is None else normalize:
            batch = _unit_rows(batch)
        return self._run_search(batch, k, flt)

    def get_embedding_by_photo_path(self, photo_path: str) -> Optional[List[float]]:
        """Stored (normalised) vector of a photo as a fresh list, or ``None``
        (utils/vector_store.py:200-212; the cache miss path is ``index.reconstruct``)."""
        row = self._path_to_index.get(photo_path)
        if row is None or row >= len(self._embeddings):
            return None
        kept = self._embeddings[row]
        if kept is not None:
            return list(kept)
        if self.index is None:
            return None
        return self.index.reconstruct(row).astype("float32").tolist()

    def has_photo_path(self, photo_path: str) -> bool:
        return photo_path in self._path_to_index

    def get_total_items(self) -> int:
        """Number of stored vectors (utils/vector_store.py:262-271)."""
        return 0 if self.index is None else int(self.index.ntotal)

    # ------------------------------------------------------------------------------------
    # persistence: <index_path> (FAISS flat container) + <index_path>.meta.json + metadata.json
    # ------------------------------------------------------------------------------------
    def save(self) -> None:
        """Write the three files the reference writes (utils/vector_store.py:217-237, :104-114).
        ``ValueError`` if nothing was ever indexed."""
        if self.index is None:
            raise ValueError(_E_NO_INDEX)
        for target in (self.index_path, self.metadata_path):
            folder = os.path.dirname(target)
            if folder:
                os.makedirs(folder, exist_ok=True)
        faiss_io.write_flat_index(self.index_path, int(self.dimension), self._metric_code, self.index.ntotal,
                                  self.index.read_rows)
        settings = dict(
            index_type=self.index_type,
            metric=self.metric,
            dimension=self.dimension,
            hnsw_m=self.hnsw_m,
            hnsw_ef_construction=self.hnsw_ef_construction,
            hnsw_ef_search=self.hnsw_ef_search,
        )
        for target, payload in ((self.meta_path, settings), (self.metadata_path, self.metadata)):
            with open(target, "w", encoding="utf-8") as handle:
                json.dump(payload, handle, ensure_ascii=False, indent=2)
        if self._attrs_built == len(self.metadata) and self._attrs_built > 0:
            self._attr_words.astype("<u8").tofile(self.attrs_path)
        elif os.path.exists(self.attrs_path):
            os.remove(self.attrs_path)  # never leave a sidecar that disagrees with metadata.json

    def _read_settings(self) -> Dict[str, Any]:
        """``.meta.json`` checks of utils/vector_store.py:116-140."""
        if not os.path.exists(self.meta_path):
            raise ValueError(_E_META_MISSING)
        with open(self.meta_path, "r", encoding="utf-8") as handle:
            settings = json.load(handle)
        if not isinstance(settings, dict):
            raise ValueError(_E_META_BROKEN)
        if str(settings.get("index_type") or "").strip().lower() != self.index_type:
            raise ValueError(_E_TYPE_MISMATCH)
        if str(settings.get("metric") or "").strip().lower() != self.metric:
            raise ValueError(_E_METRIC_MISMATCH)
        return settings

    def load(self) -> bool:
        """Load index + metadata (utils/vector_store.py:239-260): ``False`` when either file is
        absent, ``ValueError`` for missing/corrupt ``.meta.json``, a type/metric mismatch or a
        row-count mismatch.  ``IxFI`` / ``IxF2`` files and FAISS ``IHNf`` containers are read;
        the whole matrix is uploaded to HBM."""
        if not (os.path.exists(self.index_path) and os.path.exists(self.metadata_path)):
            return False
        info, blocks = faiss_io.open_index(self.index_path)
        self._read_settings()
        if info["metric"] != self._metric_code:
            raise ValueError(_E_METRIC_MISMATCH)
        with open(self.metadata_path, "r", encoding="utf-8") as handle:
            records = json.load(handle)
        if info["ntotal"] != len(records):
            raise ValueError(_E_COUNT_MISMATCH)

        # Build the new backend completely before anything is swapped: request threads keep searching the old one
        # meanwhile (core/searcher.py:1579 calls load_index lazily from request threads, core/indexer.py:680 reloads
        # during serving), and the old handle is released by reference count, never destroyed under a running search.
        dimension = int(info["d"])
        saved_dimension, self.dimension = self.dimension, dimension
        try:
            fresh = self._create_index(dimension)
        except MemoryError:
            # both copies do not fit: give the old one up first (searches in flight keep it alive until they return)
            with self._swap_lock:
                self.index = None
            fresh = self._create_index(dimension)
        finally:
            self.dimension = saved_dimension
        if hasattr(fresh, "reserve"):
            fresh.reserve(info["ntotal"])
        for block in blocks:
            fresh.add(block)
        attr_words, attrs_built = np.zeros(0, np.uint64), 0
        if os.path.exists(self.attrs_path) and os.path.getsize(self.attrs_path) == 8 * info["ntotal"] and info["ntotal"] > 0 \
                and os.path.getmtime(self.attrs_path) >= os.path.getmtime(self.metadata_path):
            words = np.fromfile(self.attrs_path, dtype="<u8").astype(np.uint64)
            fresh.set_attrs(0, words)
            attr_words, attrs_built = words, info["ntotal"]
        paths: Dict[str, int] = {}
        for row, metadata in enumerate(records):
            photo_path = metadata.get("photo_path") if isinstance(metadata, dict) else None
            if isinstance(photo_path, str) and photo_path:
                paths[photo_path] = row
        with self._swap_lock:
            self.dimension = dimension
            self.index = fresh
            self.metadata = records
            self._embeddings = [None] * info["ntotal"]
            self._attr_words, self._attrs_built = attr_words, attrs_built
            self._path_to_index = paths
        return True

    def clear(self) -> None:
        """Drop all vectors and metadata, keep the dimension (utils/vector_store.py:273-280)."""
        with self._swap_lock:
            if self.index is not None and hasattr(self.index, "reset"):
                self.index.reset()
            elif self.dimension:
                self.index = self._create_index(self.dimension)
            else:
                self.index = None
            self.metadata = []
            self._embeddings = []
            self._path_to_index = {}
            self._attrs_built = 0
            self._attr_words = np.zeros(0, np.uint64)
