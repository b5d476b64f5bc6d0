"""CPU oracle for the dense-recall hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of what the reference does on the path
``utils/vector_store.py::VectorStore`` -> FAISS CPU ``IndexFlatIP`` / ``IndexFlatL2``
and of the score / predicate / fusion arithmetic at its ``core/searcher.py`` call
sites.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it; nothing under ``photo_search_engine_b200/`` does (the product
path has no CPU fallback).

Where the arithmetic comes from
-------------------------------
The reference delegates the scan to the third-party ``faiss-cpu`` wheel
(``/root/reference/requirements.txt:5``, ``faiss-cpu>=1.7.0``, unpinned, source not
vendored, not installable in this image).  What is restated here is therefore FAISS'
*published* IndexFlat semantics -- exact fp32 inner product (or squared L2), the k
best returned best-first, int64 labels, unfilled slots ``(-inf | +inf, -1)`` -- anchored
on the reference's own call sites:

* ``faiss.IndexFlatIP(d)`` / ``IndexFlatL2(d)``  utils/vector_store.py:72-81
* ``index.add((1,d) float32)``                   utils/vector_store.py:163-164
* ``index.search((1,d) float32, k)``             utils/vector_store.py:188-197
* ``index.reconstruct(i)``                       utils/vector_store.py:207
* ``faiss.write_index`` / ``read_index``         utils/vector_store.py:234, :249

Pinning status
--------------
PINNED by the reference's fixtures (tests/test_oracle_golden.py):
  * file format + normalisation: ``pytest-tmp/build-smoke/data/idx`` (a FAISS-written
    IndexFlatIP, d=8, one row) is reproduced byte-for-byte by ``normalize_vector`` +
    ``write_index``;
  * ``data/photo_search.index`` (FAISS ``IHNf`` container, 77 x 4096 real embeddings) is
    parsed by ``read_index`` and its nested flat block is served by the exact scan;
  * the tie / ordering / normalisation behaviour asserted by the reference's
    ``tests/test_vector_store.py`` (first inserted id wins an exact tie at k=1,
    ``1/sqrt(d)`` to 6 places, ``k = min(top_k, ntotal)``).
UNPINNED at the arithmetic level: no test or fixture in the reference holds FAISS'
numeric output on non-trivial data, and FAISS itself cannot run
here.  Scores are therefore compared with a tolerance (1e-5 relative, the north-star
bar) and ids are required to be equal except where the oracle's own scores tie within
that tolerance.  Tie order here is deterministic: higher score first, then lower id.
"""
from __future__ import annotations

import json
import math
import os
import struct
from datetime import datetime
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

METRIC_INNER_PRODUCT = 0  # faiss.METRIC_INNER_PRODUCT
METRIC_L2 = 1  # faiss.METRIC_L2


# --------------------------------------------------------------------------------------
# normalisation -- utils/vector_store.py:83-90
# --------------------------------------------------------------------------------------
def normalize_vector(vector: Sequence[float], normalize: bool = True) -> List[float]:
    """``VectorStore._normalize_vector`` (utils/vector_store.py:83-90).

    float32 array -> fp32 ``np.linalg.norm`` -> zero norm returns the input unchanged ->
    ``(array / norm).astype(float32).tolist()``.
    """
    if not normalize:
        return list(vector)
    array = np.array(vector, dtype="float32")
    norm = np.linalg.norm(array)
    if norm == 0:
        return list(vector)
    return (array / norm).astype("float32").tolist()


# --------------------------------------------------------------------------------------
# exact flat index -- FAISS IndexFlatIP / IndexFlatL2 semantics
# --------------------------------------------------------------------------------------
def _topk_desc(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k largest of a 1-D fp32 array, best first, ties -> lower index first."""
    n = scores.shape[0]
    k = min(k, n)
    if k <= 0:
        return np.empty(0, np.float32), np.empty(0, np.int64)
    if k < n:
        # everything >= the k-th value, then an exact (score desc, id asc) order
        kth = np.partition(scores, n - k)[n - k]
        cand = np.nonzero(scores >= kth)[0]
    else:
        cand = np.arange(n)
    order = np.lexsort((cand, -scores[cand].astype(np.float64)))
    sel = cand[order[:k]].astype(np.int64)
    return scores[sel].astype(np.float32), sel


class OracleIndexFlat:
    """Exact brute-force index with FAISS ``IndexFlat`` semantics (fp32)."""

    def __init__(self, d: int, metric_type: int = METRIC_INNER_PRODUCT) -> None:
        self.d = int(d)
        self.metric_type = int(metric_type)
        self._chunks: List[np.ndarray] = []
        self._x: Optional[np.ndarray] = np.empty((0, self.d), np.float32)

    # -- storage -----------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(sum(c.shape[0] for c in self._chunks) + (self._x.shape[0] if self._x is not None else 0))

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError("bad shape")
        self._chunks.append(x.copy())

    def _matrix(self) -> np.ndarray:
        if self._chunks:
            parts = ([self._x] if self._x is not None and self._x.shape[0] else []) + self._chunks
            self._x = np.ascontiguousarray(np.concatenate(parts, axis=0))
            self._chunks = []
        return self._x

    def reconstruct(self, i: int) -> np.ndarray:
        return self._matrix()[int(i)].copy()

    # -- search ------------------------------------------------------------------
    def scores(self, q: np.ndarray, chunk_rows: int = 1 << 16) -> np.ndarray:
        """All N 'larger is better' scores for ONE query (IP, or minus squared L2)."""
        x = self._matrix()
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(self.d)
        out = np.empty(x.shape[0], np.float32)
        for s in range(0, x.shape[0], chunk_rows):
            blk = x[s : s + chunk_rows]
            if self.metric_type == METRIC_INNER_PRODUCT:
                out[s : s + chunk_rows] = blk @ q
            else:
                diff = blk - q[None, :]
                out[s : s + chunk_rows] = -np.einsum("ij,ij->i", diff, diff)
        return out

    def search(
        self, q: np.ndarray, k: int, mask: Optional[np.ndarray] = None
    ) -> Tuple[np.ndarray, np.ndarray]:
        """``index.search`` (utils/vector_store.py:191): D float32 (nq,k), I int64 (nq,k).

        ``mask`` (bool[N], optional) restricts the scan to rows where it is True: the
        oracle of the fused EXIF predicate.  Unfilled slots are ``(-inf, -1)`` for IP and
        ``(+inf, -1)`` for L2, as FAISS leaves them.
        """
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        ip = self.metric_type == METRIC_INNER_PRODUCT
        D = np.full((nq, k), -np.inf if ip else np.inf, np.float32)
        I = np.full((nq, k), -1, np.int64)
        if self.ntotal == 0 or k <= 0:
            return D, I
        rows = None if mask is None else np.nonzero(np.asarray(mask, bool))[0]
        for qi in range(nq):
            s = self.scores(q[qi])
            if rows is not None:
                s_sel, i_sel = _topk_desc(s[rows], k)
                i_sel = rows[i_sel]
            else:
                s_sel, i_sel = _topk_desc(s, k)
            m = s_sel.shape[0]
            D[qi, :m] = s_sel if ip else -s_sel
            I[qi, :m] = i_sel
        return D, I


# --------------------------------------------------------------------------------------
# FAISS on-disk format (write_index / read_index) -- utils/vector_store.py:234, :249
# --------------------------------------------------------------------------------------
_FLAT_FOURCC = {METRIC_INNER_PRODUCT: b"IxFI", METRIC_L2: b"IxF2"}
_DUMMY = 1 << 20


def write_index(index: OracleIndexFlat, path: str) -> None:
    """FAISS ``write_index`` for IndexFlatIP / IndexFlatL2.

    Layout (verified against ``pytest-tmp/build-smoke/data/idx``):
    fourcc, int32 d, int64 ntotal, int64 1<<20, int64 1<<20, uint8 is_trained=1,
    int32 metric_type, uint64 count=d*ntotal, count x float32 little-endian.
    """
    x = index._matrix()
    with open(path, "wb") as f:
        f.write(_FLAT_FOURCC[index.metric_type])
        f.write(struct.pack("<iqqqBi", index.d, x.shape[0], _DUMMY, _DUMMY, 1, index.metric_type))
        f.write(struct.pack("<Q", x.shape[0] * index.d))
        f.write(np.ascontiguousarray(x, dtype="<f4").tobytes())


def _read_header(buf: bytes, off: int) -> Tuple[int, int, int, int]:
    d, ntotal, _d1, _d2, _trained, metric = struct.unpack_from("<iqqqBi", buf, off)
    off += struct.calcsize("<iqqqBi")
    if metric > 1:  # FAISS writes metric_arg only for the exotic metrics
        off += 4
    return d, ntotal, metric, off


def _skip_vector(buf: bytes, off: int, itemsize: int) -> Tuple[int, int]:
    (n,) = struct.unpack_from("<Q", buf, off)
    return n, off + 8 + n * itemsize


def read_index(path: str) -> Tuple[OracleIndexFlat, Dict[str, Any]]:
    """FAISS ``read_index`` for ``IxFI`` / ``IxF2`` and for the ``IHNf`` HNSW container.

    For ``IHNf`` the graph (assign_probas f64, cum_nneighbor_per_level i32, levels i32,
    offsets u64, neighbors i32, then entry_point / max_level / efConstruction / efSearch /
    upper_beam as int32) is skipped and the nested flat storage block is returned: exact
    search over it is a recall superset of the HNSW walk.
    """
    with open(path, "rb") as f:
        buf = f.read()
    info: Dict[str, Any] = {"fourcc": buf[:4].decode("ascii", "replace")}
    off = 4
    if buf[:4] == b"IHNf":
        d, ntotal, metric, off = _read_header(buf, off)
        info.update(hnsw=True, hnsw_d=d, hnsw_ntotal=ntotal, hnsw_metric=metric)
        for name, size in (("assign_probas", 8), ("cum_nneighbor_per_level", 4), ("levels", 4), ("offsets", 8), ("neighbors", 4)):
            n, off = _skip_vector(buf, off, size)
            info[f"n_{name}"] = n
        entry_point, max_level, ef_c, ef_s, upper_beam = struct.unpack_from("<iiiii", buf, off)
        off += 20
        info.update(entry_point=entry_point, max_level=max_level, efConstruction=ef_c, efSearch=ef_s, upper_beam=upper_beam)
        info["storage_offset"] = off
        fourcc = buf[off : off + 4]
        off += 4
    else:
        fourcc = buf[:4]
    if fourcc not in (b"IxFI", b"IxF2"):
        raise ValueError(f"unsupported index fourcc {fourcc!r}")
    d, ntotal, metric, off = _read_header(buf, off)
    (count,) = struct.unpack_from("<Q", buf, off)
    off += 8
    if count != d * ntotal:
        raise ValueError("corrupt flat block")
    x = np.frombuffer(buf, dtype="<f4", count=count, offset=off).reshape(ntotal, d)
    index = OracleIndexFlat(d, metric)
    index.add(x)
    info.update(d=d, ntotal=ntotal, metric=metric, flat_fourcc=fourcc.decode())
    return index, info


# --------------------------------------------------------------------------------------
# VectorStore restatement -- utils/vector_store.py (whole file)
# --------------------------------------------------------------------------------------
class OracleVectorStore:
    """Line-by-line behavioural restatement of ``utils/vector_store.py::VectorStore``
    over ``OracleIndexFlat``.  ``index_type='hnsw'`` is accepted and served exactly."""

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
    ) -> None:
        self.dimension = dimension
        self.index_path = index_path
        self.metadata_path = metadata_path
        self.meta_path = f"{self.index_path}.meta.json"
        self.metric = metric.lower().strip() if metric else "l2"
        if self.metric not in {"l2", "cosine"}:
            raise ValueError("metric仅支持l2或cosine")
        self.index_type = (index_type or "flat").strip().lower()
        if self.index_type not in {"flat", "hnsw"}:
            raise ValueError("index_type仅支持flat或hnsw")
        self.hnsw_m = max(4, int(hnsw_m))
        self.hnsw_ef_construction = max(8, int(hnsw_ef_construction))
        self.hnsw_ef_search = max(8, int(hnsw_ef_search))
        self.index = self._create_index(dimension) if dimension else None
        self.metadata: List[Dict] = []
        self._normalize = self.metric == "cosine"
        self._embeddings: List[Optional[List[float]]] = []
        self._path_to_index: Dict[str, int] = {}

    def _create_index(self, dimension: int) -> OracleIndexFlat:
        return OracleIndexFlat(dimension, METRIC_INNER_PRODUCT if self.metric == "cosine" else METRIC_L2)

    def add_item(self, embedding: List[float], metadata: Dict) -> None:
        if embedding is None:
            raise ValueError("向量不能为空")
        if self.index is None:
            self.dimension = len(embedding)
            self.index = self._create_index(self.dimension)
        if len(embedding) != self.dimension:
            raise ValueError(f"向量维度不匹配: {len(embedding)} != {self.dimension}")
        normalized = normalize_vector(embedding, self._normalize)
        self.index.add(np.array([normalized], dtype="float32"))
        self.metadata.append(metadata)
        self._embeddings.append(normalized)
        photo_path = metadata.get("photo_path")
        if isinstance(photo_path, str) and photo_path:
            self._path_to_index[photo_path] = len(self.metadata) - 1

    def search(self, query_embedding: List[float], top_k: int) -> List[Dict]:
        if self.index is None or self.index.ntotal == 0:
            return []
        if len(query_embedding) != self.dimension:
            raise ValueError(f"向量维度不匹配: {len(query_embedding)} != {self.dimension}")
        k = min(top_k, self.index.ntotal)
        normalized = normalize_vector(query_embedding, self._normalize)
        distances, indices = self.index.search(np.array([normalized], dtype="float32"), k)
        results: List[Dict] = []
        for distance, index in zip(distances[0].tolist(), indices[0].tolist()):
            if index == -1:
                continue
            results.append({"metadata": self.metadata[index], "distance": float(distance)})
        return results

    def get_embedding_by_photo_path(self, photo_path: str) -> Optional[List[float]]:
        index = self._path_to_index.get(photo_path)
        if index is None:
            return None
        if index < len(self._embeddings):
            cached = self._embeddings[index]
            if cached is None and self.index is not None:
                cached = self.index.reconstruct(index).astype("float32").tolist()
                self._embeddings[index] = cached
            if cached is not None:
                return list(cached)
        return None

    def has_photo_path(self, photo_path: str) -> bool:
        return photo_path in self._path_to_index

    def save(self) -> None:
        if self.index is None:
            raise ValueError("索引未初始化")
        for directory in (os.path.dirname(self.index_path), os.path.dirname(self.metadata_path)):
            if directory:
                os.makedirs(directory, exist_ok=True)
        write_index(self.index, self.index_path)
        payload = {
            "index_type": self.index_type,
            "metric": self.metric,
            "dimension": self.dimension,
            "hnsw_m": self.hnsw_m,
            "hnsw_ef_construction": self.hnsw_ef_construction,
            "hnsw_ef_search": self.hnsw_ef_search,
        }
        with open(self.meta_path, "w", encoding="utf-8") as file:
            json.dump(payload, file, ensure_ascii=False, indent=2)
        with open(self.metadata_path, "w", encoding="utf-8") as file:
            json.dump(self.metadata, file, ensure_ascii=False, indent=2)

    def load(self) -> bool:
        if not os.path.exists(self.index_path) or not os.path.exists(self.metadata_path):
            return False
        self.index, _info = read_index(self.index_path)
        if not os.path.exists(self.meta_path):
            raise ValueError("索引元信息缺失，请重新构建索引")
        with open(self.meta_path, "r", encoding="utf-8") as file:
            payload = json.load(file)
        if not isinstance(payload, dict):
            raise ValueError("索引元信息损坏，请重新构建索引")
        if str(payload.get("index_type") or "").strip().lower() != self.index_type:
            raise ValueError("索引类型与配置不一致，请重新构建索引")
        if str(payload.get("metric") or "").strip().lower() != self.metric:
            raise ValueError("索引度量与配置不一致，请重新构建索引")
        want = METRIC_INNER_PRODUCT if self.metric == "cosine" else METRIC_L2
        if self.index.metric_type != want:
            raise ValueError("索引度量与配置不一致，请重新构建索引")
        with open(self.metadata_path, "r", encoding="utf-8") as file:
            self.metadata = json.load(file)
        if self.index.ntotal != len(self.metadata):
            raise ValueError("索引与元数据数量不一致，请重新构建索引")
        self.dimension = self.index.d
        self._embeddings = [None] * self.index.ntotal
        self._path_to_index = {}
        for index, metadata in enumerate(self.metadata):
            photo_path = metadata.get("photo_path")
            if isinstance(photo_path, str) and photo_path:
                self._path_to_index[photo_path] = index
        return True

    def get_total_items(self) -> int:
        return 0 if self.index is None else int(self.index.ntotal)

    def clear(self) -> None:
        self.index = self._create_index(self.dimension) if self.dimension else None
        self.metadata = []
        self._embeddings = []
        self._path_to_index = {}


# --------------------------------------------------------------------------------------
# call-site arithmetic -- core/searcher.py
# --------------------------------------------------------------------------------------
def distance_to_score(distance: float, metric: str = "cosine") -> float:
    """``Searcher._distance_to_score`` (core/searcher.py:605-625), Python float64."""
    if metric == "cosine":
        similarity = max(-1.0, min(1.0, distance))
        score = (similarity + 1.0) / 2.0
        if score > 0.7:
            score = 0.7 + (score - 0.7) * 1.3
        elif score < 0.3:
            score = score * 0.8
        return round(max(0.0, min(1.0, score)), 6)
    if distance < 0:
        distance = 0
    return round(round(float(np.exp(-0.5 * distance)), 6), 6)


def calculate_candidate_k(total_items: int, top_k: int, has_time_filter: bool, relaxation_level: int = 0) -> int:
    """``Searcher._calculate_candidate_k`` (core/searcher.py:771-820)."""
    base = 10 if has_time_filter else 5
    if total_items <= 50:
        candidate_k = total_items
    elif total_items <= 500:
        candidate_k = top_k * base
    elif total_items <= 5000:
        candidate_k = max(top_k * (base - 2), 100)
    else:
        candidate_k = max(top_k * 3, min(int(total_items * 0.01), 500))
    if relaxation_level > 0:
        candidate_k = max(candidate_k, top_k * (base + relaxation_level))
        candidate_k = math.ceil(candidate_k * (1 + min(relaxation_level, 3) * 0.35))
    return min(candidate_k, total_items)


_DATE_FORMATS = [
    "%Y-%m-%d",
    "%Y-%m-%dT%H:%M:%S",
    "%Y-%m-%d %H:%M:%S",
    "%Y:%m:%d %H:%M:%S",
    "%Y/%m/%d %H:%M:%S",
    "%Y/%m/%d",
    "%Y%m%d",
]


def parse_date(value: Any, is_end_date: bool = False) -> Optional[datetime]:
    """``Searcher._parse_date`` (core/searcher.py:1963-2001)."""
    if not value or not isinstance(value, str):
        return None
    cleaned = value.strip().rstrip("\x00")
    for fmt in _DATE_FORMATS:
        try:
            parsed = datetime.strptime(cleaned, fmt)
            if fmt in ("%Y-%m-%d", "%Y/%m/%d", "%Y%m%d") and is_end_date:
                return datetime(parsed.year, parsed.month, parsed.day, 23, 59, 59)
            return parsed
        except ValueError:
            continue
    try:
        return datetime.fromisoformat(cleaned)
    except Exception:
        return None


def check_time_match_v2(metadata: Dict[str, Any], constraints: Dict[str, Any]) -> bool:
    """``Searcher._check_time_match_v2`` (core/searcher.py:1884-1950): the EXIF predicate."""
    time_info = metadata.get("time_info") or {}
    exif_data = metadata.get("exif_data") or {}
    exif_datetime = exif_data.get("datetime")
    for key in ("season", "time_period", "year", "month"):
        if constraints.get(key):
            if not exif_datetime:
                return False
            if time_info.get(key) != constraints[key]:
                return False
    start_date = constraints.get("start_date")
    end_date = constraints.get("end_date")
    if start_date or end_date:
        photo_datetime_str = time_info.get("datetime_str") or exif_datetime
        if not photo_datetime_str:
            return False
        photo_date = parse_date(photo_datetime_str)
        if not photo_date:
            return False
        if start_date:
            start = parse_date(start_date)
            if start and photo_date < start:
                return False
        if end_date:
            end = parse_date(end_date, is_end_date=True)
            if end and photo_date > end:
                return False
    return True


def month_to_season(month: int) -> Optional[str]:
    """``Indexer._month_to_season`` (core/indexer.py:619-629)."""
    if month in (3, 4, 5):
        return "春天"
    if month in (6, 7, 8):
        return "夏天"
    if month in (9, 10, 11):
        return "秋天"
    if month in (12, 1, 2):
        return "冬天"
    return None


def hour_to_period(hour: int) -> str:
    """7-way part-of-day buckets of ``Indexer._extract_time_info`` (core/indexer.py:583-598)."""
    if 0 <= hour < 5:
        return "凌晨"
    if 5 <= hour < 8:
        return "早晨"
    if 8 <= hour < 12:
        return "上午"
    if 12 <= hour < 14:
        return "中午"
    if 14 <= hour < 17:
        return "下午"
    if 17 <= hour < 19:
        return "傍晚"
    return "夜晚"


def time_info_from_exif(exif_datetime: Optional[str]) -> Dict[str, Any]:
    """``Indexer._extract_time_info`` (core/indexer.py:535-609) for an ISO EXIF datetime."""
    info: Dict[str, Any] = dict.fromkeys(
        ("year", "month", "day", "hour", "season", "time_period", "weekday", "datetime_str")
    )
    if not exif_datetime:
        return info
    try:
        t = datetime.fromisoformat(exif_datetime)
    except Exception:
        return info
    names = ["星期一", "星期二", "星期三", "星期四", "星期五", "星期六", "星期日"]
    info.update(
        year=t.year, month=t.month, day=t.day, hour=t.hour, datetime_str=t.isoformat(),
        season=month_to_season(t.month), time_period=hour_to_period(t.hour), weekday=names[t.weekday()],
    )
    return info


def hybrid_fuse(
    vector_hits: Iterable[Tuple[int, float]],
    keyword_hits: Iterable[Tuple[int, float]],
    *,
    vector_weight: float = 0.8,
    keyword_weight: float = 0.2,
    metric: str = "cosine",
    allow_keyword_only: bool = True,
    keyword_filtered: bool = False,
    boosts: Optional[Dict[int, float]] = None,
) -> List[Tuple[int, float, float, float]]:
    """The numeric core of ``Searcher._hybrid_search`` (core/searcher.py:893-986) keyed by
    row id instead of photo path (paths are unique per row in the synthetic workload).

    vector_hits: (id, raw distance).  keyword_hits: (id, score in [0,1]).  Returns
    [(id, fused, vector_score, keyword_score)] sorted by fused score descending
    (stable: first-seen order breaks ties, vector hits first, as the dict/set union does
    not define an order the sort here is made deterministic by id).
    ``keyword_filtered`` = the ES filter branch was taken (es_filtered_paths is not None).
    """
    v: Dict[int, float] = {}
    for i, dist in vector_hits:
        v[int(i)] = distance_to_score(float(dist), metric)
    kw: Dict[int, float] = {int(i): float(s) for i, s in keyword_hits}
    ids = set(v)
    if allow_keyword_only:
        ids |= set(kw)
    out = []
    for i in ids:
        has_v, has_k = i in v, i in kw
        vs, ks = v.get(i, 0.0), kw.get(i, 0.0)
        avail = (vector_weight if has_v else 0.0) + (keyword_weight if has_k else 0.0)
        if avail <= 0:
            continue
        weighted = (vector_weight * vs if has_v else 0.0) + (keyword_weight * ks if has_k else 0.0)
        score = weighted / avail
        score *= (boosts or {}).get(i, 1.0)
        if has_k and not has_v:
            score *= 0.65
            if not keyword_filtered and ks < 0.45:
                continue
        out.append((i, round(score, 6), round(vs, 6), round(ks, 6)))
    out.sort(key=lambda t: (-t[1], t[0]))
    return out


def calculate_dynamic_threshold(scores: Sequence[float], top_k: int, threshold_floor: float = 0.05) -> float:
    """``Searcher._calculate_dynamic_threshold`` (core/searcher.py:627-674) restated: ``scores`` in candidate order
    (best first), numpy percentiles / median exactly as the reference calls them."""
    if not scores:
        return 0.1
    n = len(scores)
    if n <= top_k * 2:
        return max(scores[-1] * 0.9, threshold_floor)
    q25 = np.percentile(scores, 25)
    q75 = np.percentile(scores, 75)
    median = np.median(scores)
    cv = (q75 - q25) / median if median > 0 else 1.0
    if cv < 0.2:
        threshold = max(median * 0.85, q25 * 0.9)
    elif cv < 0.5:
        threshold = q25
    else:
        threshold = max(q25 * 0.7, median * 0.7)
    if n >= top_k:
        threshold = max(threshold, scores[top_k - 1] * 0.8)
    return round(max(threshold, threshold_floor), 6)


def finalize_thresholds(scores: Sequence[float], top_k: int, strict_floor: float, broad_floor: float,
                        threshold_floor: float = 0.05) -> Tuple[float, float, List[int]]:
    """The numeric part of ``Searcher._finalize_results`` (core/searcher.py:1497-1526) and the score-only part of
    ``_assign_confidence_bucket`` (:828-840): ``(strict_threshold, broad_threshold, bucket per score)``."""
    scores = [float(s) for s in scores]
    if scores:
        dynamic_threshold = calculate_dynamic_threshold(scores, top_k, threshold_floor)
        strict_threshold = max(dynamic_threshold, strict_floor)
        broad_threshold = min(strict_threshold - 0.05, max(broad_floor, strict_threshold * 0.84))
        broad_threshold = round(max(broad_floor, broad_threshold), 6)
    else:
        strict_threshold, broad_threshold = strict_floor, broad_floor
    buckets = [3 if s >= strict_threshold else 2 if s >= broad_threshold else 1 for s in scores]
    return float(strict_threshold), float(broad_threshold), buckets
