"""Opt-in batching of the Searcher's expansion fan-out.

The reference runs one embedding call and one ``vector_store.search`` per expansion alternative,
one after the other (core/searcher.py:1392-1412 -> ``_run_single_search_round`` :1157-1217 ->
``_hybrid_search`` :887 / ``vector_store.search`` :1195), although ``expand_query_intents`` returns
every alternative in ONE call (:1382) and the embedding service has a batch entry point that the
Searcher never uses (utils/embedding_service.py:85-114).  On the B200 backend nq searches cost
about as much as one (the corpus is streamed once for the whole batch), so::

    from core.searcher import Searcher
    from photo_search_engine_b200.searcher_ext import BatchedExpansionMixin

    class BatchedSearcher(BatchedExpansionMixin, Searcher):
        pass

keeps the reference's control flow byte for byte -- the loop, the contract check, the merge, the
thresholds all run unmodified -- and only answers the loop's ``vector_store.search`` calls from one
``search_batch`` issued the moment the alternatives are known.  Results are identical to the
unbatched Searcher: the top-k of a query is a prefix of its top-k_max, and both come from the same
exact scan arithmetic.

Nothing here imports the reference: the mixin only relies on the method names cited above.
"""
from __future__ import annotations

import threading
from collections.abc import Sequence
from typing import Any, Dict, List, Optional

import numpy as np


def _embedding_key(embedding: Sequence[float]) -> bytes:
    return np.asarray(embedding, dtype=np.float32).tobytes()


class _PrefetchingStore:
    """What the Searcher sees as ``self.vector_store``: the real store, except that ``search`` is
    answered from the current request's prefetched batch when the same embedding was part of it."""

    def __init__(self, store: Any, local: threading.local) -> None:
        object.__setattr__(self, "_store", store)
        object.__setattr__(self, "_local", local)

    def __getattr__(self, name: str) -> Any:
        return getattr(self._store, name)

    def __setattr__(self, name: str, value: Any) -> None:
        setattr(self._store, name, value)

    def search(self, query_embedding: List[float], top_k: int, *args: Any, **kwargs: Any) -> List[Dict]:
        # FusedPrefilterMixin: the round's EXIF constraints ride along, so that the scan itself only considers
        # rows passing Searcher._check_time_match_v2 (core/searcher.py:1884-1950)
        fused = getattr(self._local, "constraints", None)
        if fused and not args and "constraints" not in kwargs:
            return self._store.search(query_embedding, top_k, constraints=fused)
        # FusedRecallMixin: the hits stay arrays (dicts only on demand)
        if getattr(self._local, "lazy", False) and not args and not kwargs and query_embedding is not None:
            store = self._store
            if getattr(store, "index", None) is None or store.get_total_items() == 0:
                return []
            k = min(int(top_k), int(store.get_total_items()))
            if k <= 0:
                return []
            if len(query_embedding) != store.dimension:
                return store.search(query_embedding, top_k)  # raises the reference's ValueError
            row = np.array([store._normalize_vector(query_embedding)], dtype="float32")
            scores, ids = store.search_batch(row, k, normalize=False)
            return _LazyHits(store.metadata, scores[0], ids[0])
        prefetched = getattr(self._local, "prefetch", None)
        if prefetched and not args and not kwargs and query_embedding is not None:
            got = prefetched.get(_embedding_key(query_embedding))
            if got is not None:
                k_have, hits = got
                k_want = min(int(top_k), int(self._store.get_total_items()))
                if k_want <= k_have:
                    stats = getattr(self._local, "stats", None)
                    if stats is not None:
                        stats["served_from_batch"] += 1
                    return [dict(h) for h in hits[: max(k_want, 0)]]
        return self._store.search(query_embedding, top_k, *args, **kwargs)


class _PrefetchingFormatter:
    """``self.query_formatter`` with one hook: as soon as ``expand_query_intents`` has returned the
    alternatives (core/searcher.py:1382), the owning Searcher embeds and searches them as a batch."""

    def __init__(self, formatter: Any, owner: "BatchedExpansionMixin") -> None:
        object.__setattr__(self, "_formatter", formatter)
        object.__setattr__(self, "_owner", owner)

    def __getattr__(self, name: str) -> Any:
        return getattr(self._formatter, name)

    def __setattr__(self, name: str, value: Any) -> None:
        setattr(self._formatter, name, value)

    def expand_query_intents(self, *args: Any, **kwargs: Any) -> Any:
        alternatives = self._formatter.expand_query_intents(*args, **kwargs)
        self._owner._psx_prefetch_alternatives(alternatives)
        return alternatives


class BatchedExpansionMixin:
    """Put in front of ``core.searcher.Searcher`` in the MRO.  Thread-safe in the way the reference is
    used (Flask's threaded server, main.py:353): all per-request state is thread-local."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if not hasattr(self, "_psx_local"):
            self._psx_local = threading.local()
        self.psx_batch_stats = {"batches": 0, "batched_queries": 0, "served_from_batch": 0}
        if getattr(self, "query_formatter", None) is not None:
            self.query_formatter = _PrefetchingFormatter(self.query_formatter, self)
        if not isinstance(self.vector_store, _PrefetchingStore):
            self.vector_store = _PrefetchingStore(self.vector_store, self._psx_local)

    # -- request scope ---------------------------------------------------------------------------
    def _maybe_expand_query_results(self, *, query, base_intent, normalized_top_k, has_filter, **kwargs):
        local = self._psx_local
        local.ctx = {"query": query, "base_intent": base_intent, "top_k": normalized_top_k, "has_filter": has_filter}
        local.prefetch, local.embeddings, local.stats = {}, {}, self.psx_batch_stats
        try:
            return super()._maybe_expand_query_results(query=query, base_intent=base_intent, normalized_top_k=normalized_top_k,
                                                       has_filter=has_filter, **kwargs)
        finally:
            local.ctx = local.prefetch = local.embeddings = None

    def _generate_embedding(self, embedding_query: str) -> List[float]:
        known = getattr(self._psx_local, "embeddings", None)
        if known:
            hit = known.get(str(embedding_query or "").strip())
            if hit is not None:
                return list(hit)
        return super()._generate_embedding(embedding_query)

    # -- the batch -------------------------------------------------------------------------------
    def _psx_embed_many(self, texts: List[str]) -> Optional[List[List[float]]]:
        """Embeddings of ``texts`` with ONE service call for the ones the Searcher's cache does not hold
        (utils/embedding_service.py:85-114), falling back to per-text calls."""
        out: Dict[str, List[float]] = {}
        missing: List[str] = []
        for text in texts:
            cached = self._cache_get(self._embedding_cache, text) if self.embedding_cache_enabled else None
            if cached is not None:
                out[text] = list(cached)
            else:
                missing.append(text)
        if missing:
            fresh = None
            batch_call = getattr(self.embedding_service, "generate_embedding_batch", None)
            if len(missing) > 1 and callable(batch_call):
                try:
                    fresh = batch_call(list(missing))
                except Exception:
                    fresh = None
                if not isinstance(fresh, (list, tuple)) or len(fresh) != len(missing):
                    fresh = None
            if fresh is None:
                fresh = [self.embedding_service.generate_embedding(text) for text in missing]
            for text, emb in zip(missing, fresh):
                out[text] = list(emb)
                if self.embedding_cache_enabled:
                    self._cache_put(self._embedding_cache, text, list(emb), self.embedding_cache_size)
        return [out[text] for text in texts]

    def _psx_prefetch_alternatives(self, alternatives: Any) -> None:
        local = self._psx_local
        ctx = getattr(local, "ctx", None)
        if not ctx or not alternatives:
            return
        try:
            limit = self.query_expansion_max_alternatives
            if limit == 0:
                limit = self._max_relaxation_rounds_until_floor(1)
            plan: Dict[str, int] = {}  # embedding text -> largest candidate_k asked for it
            for alt_index, alt in enumerate(list(alternatives)[: max(limit, 0)], start=1):
                if not self._intent_contract_is_satisfied(ctx["base_intent"], alt):
                    continue  # the reference skips it too (core/searcher.py:1393-1394)
                text = self._build_query_text(
                    search_text=str(alt.get("search_text") or ""),
                    media_terms=list(alt.get("media_terms") or []),
                    identity_terms=list(alt.get("identity_terms") or []),
                    original_query=ctx["query"],
                )
                text = str(text or "").strip()
                if not text:
                    continue
                k = self._calculate_candidate_k(ctx["top_k"], ctx["has_filter"], relaxation_level=alt_index)
                plan[text] = max(plan.get(text, 0), int(k))
            store = self.vector_store._store
            if len(plan) < 2 or not hasattr(store, "search_batch"):
                return  # nothing to batch
            texts = list(plan)
            embeddings = self._psx_embed_many(texts)
            k_max = min(max(plan.values()), int(store.get_total_items()))
            if k_max <= 0:
                return
            # normalise each query exactly as VectorStore.search does (utils/vector_store.py:83-90, one vector at a
            # time) so that the batch sees the very bits the unbatched calls would have sent
            normalise = getattr(store, "_normalize_vector", None)
            rows = [normalise(list(e)) for e in embeddings] if callable(normalise) else embeddings
            scores, ids = store.search_batch(np.asarray(rows, dtype=np.float32), k_max, normalize=False)
            records = store.metadata
            for text, emb, row_scores, row_ids in zip(texts, embeddings, scores.tolist(), ids.tolist()):
                local.embeddings[text] = emb
                hits = [{"metadata": records[i], "distance": float(s)} for s, i in zip(row_scores, row_ids) if i != -1]
                local.prefetch[_embedding_key(emb)] = (k_max, hits)
            self.psx_batch_stats["batches"] += 1
            self.psx_batch_stats["batched_queries"] += len(texts)
        except Exception:
            # batching is an optimisation: on any surprise the unmodified loop simply runs unbatched
            local.prefetch = {}
            local.embeddings = {}


class FusedPrefilterMixin:
    """Opt-in: hand the round's EXIF constraints to the vector store so that the predicate
    runs on the device BEFORE the top-k selection.  The reference post-filters the top ``candidate_k`` hits
    (core/searcher.py:1477-1480) and can come back with fewer than ``top_k`` photos although more match; with
    the pre-filter the k best *matching* rows are recalled.  This changes results (strictly more complete for
    filtered queries), hence a separate opt-in::

        class MySearcher(FusedPrefilterMixin, Searcher): ...

    Only the pure vector branch is affected (``keyword_store is None``: the branch in which the reference applies
    ``_check_time_match_v2`` itself); with Elasticsearch the filter already runs there.  The reference's own
    post-filter still runs afterwards and is a no-op on rows that passed.  May be combined with
    ``BatchedExpansionMixin`` (list this one first); expansion rounds under a predicate then search unbatched.
    """

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if not hasattr(self, "_psx_local"):
            self._psx_local = threading.local()
        if not isinstance(self.vector_store, _PrefetchingStore):
            self.vector_store = _PrefetchingStore(self.vector_store, self._psx_local)

    def _run_single_search_round(self, *, constraints, has_filter, **kwargs):
        use = bool(has_filter) and bool(constraints) and self.keyword_store is None
        self._psx_local.constraints = dict(constraints) if use else None
        try:
            return super()._run_single_search_round(constraints=constraints, has_filter=has_filter, **kwargs)
        finally:
            self._psx_local.constraints = None


# ---------------------------------------------------------------------------------------------------------------------
# the Python tail of a search round
# ---------------------------------------------------------------------------------------------------------------------
class _LazyHits(Sequence):
    """What ``vector_store.search`` returns under ``FusedRecallMixin``: the FAISS-shaped arrays of the search, behaving
    as the reference's ``List[{"metadata", "distance"}]`` for any code that iterates it (dicts are built on demand)."""

    def __init__(self, records: List[Dict], scores: np.ndarray, ids: np.ndarray) -> None:
        keep = ids >= 0
        self.records, self.ids, self.distances = records, ids[keep], scores[keep]

    def __len__(self) -> int:
        return int(self.ids.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return {"metadata": self.records[int(self.ids[i])], "distance": float(self.distances[i])}


class _LazyCombined(Sequence):
    """``_vector_results_to_combined`` under ``FusedRecallMixin``: candidate rows, their scores (the reference's own
    ``_distance_to_score``) in recall order, de-duplicated by path key.  Iterating it yields the reference's dicts."""

    def __init__(self, owner: Any, records: List[Dict], rows: List[int], scores: List[float]) -> None:
        self.owner, self.records, self.rows, self.scores = owner, records, rows, scores

    def __len__(self) -> int:
        return len(self.rows)

    def item(self, i: int) -> Dict[str, Any]:
        metadata = self.records[self.rows[i]] or {}
        return {
            "photo_path": metadata.get("photo_path"),
            "description": metadata.get("description"),
            "retrieval_text": metadata.get("retrieval_text"),
            "score": self.scores[i],
            "metadata": metadata,
            "match_summary": self.owner._psx_match_summary(metadata),
        }

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.item(j) for j in range(*i.indices(len(self)))]
        return self.item(i)


class _LazyFused(Sequence):
    """``_hybrid_search`` under ``FusedRecallMixin``: fused entries ``(score, vector_score, keyword_score, photo_path,
    metadata)`` in result order; the reference's dicts are built on demand."""

    def __init__(self, owner: Any, entries: List[tuple]) -> None:
        self.owner, self.entries = owner, entries

    def __len__(self) -> int:
        return len(self.entries)

    def item(self, i: int) -> Dict[str, Any]:
        score, v_score, k_score, photo_path, metadata = self.entries[i]
        return {
            "photo_path": photo_path,
            "description": metadata.get("description", ""),
            "score": score,
            "vector_score": v_score,
            "keyword_score": k_score,
            "rank": 0,
            "metadata": metadata,
            "match_summary": self.owner._psx_match_summary(metadata),
        }

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.item(j) for j in range(*i.indices(len(self)))]
        return self.item(i)


class FusedRecallMixin:
    """Opt-in: once the scan takes a few milliseconds, a search round is dominated by
    O(candidate_k) Python -- one dict, one ``normalize_local_path``, one ``build_match_summary`` and (under a time
    filter) several ``strptime`` calls per candidate (core/searcher.py:1131-1156, :1460-1565, :1884-2001) although only
    ``top_k`` of the 500-1333 candidates are returned.  This mixin keeps the candidates as arrays from the store
    (``search_batch``) to the end of ``_finalize_results`` and builds dicts for the returned photos only::

        class MySearcher(FusedRecallMixin, Searcher): ...

    Results are identical to the unmodified Searcher: scores come from its own ``_distance_to_score``, the thresholds
    from its own ``_calculate_dynamic_threshold`` / ``_get_round_score_floors``, path keys from its own ``_path_key``
    (cached per row), the time filter from the packed EXIF words the fused predicate uses (same conjunction as
    ``_check_time_match_v2``; constraints that cannot be packed go through the reference function).  Rounds without
    media / identity terms take the array path (with a keyword store: ``_hybrid_search`` below joins on tuples); every
    other case (term matching, file-existence validation) runs the reference code on lazily materialised dicts.
    Combine with the other mixins by listing this one first.
    """

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if not hasattr(self, "_psx_local"):
            self._psx_local = threading.local()
        if not isinstance(self.vector_store, _PrefetchingStore):
            self.vector_store = _PrefetchingStore(self.vector_store, self._psx_local)
        self._psx_key_cache: Dict[int, str] = {}
        self._psx_key_owner: Any = None
        self._psx_vector_scores_ok: Optional[bool] = None  # None = not probed yet
        self.psx_recall_stats = {"array_rounds": 0, "reference_rounds": 0}

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _psx_match_summary(self, metadata: Dict[str, Any]) -> Any:
        import sys

        for klass in type(self).__mro__:
            mod = sys.modules.get(klass.__module__)
            fn = getattr(mod, "build_match_summary", None) if mod is not None else None
            if callable(fn):  # the name core.searcher imported (utils/structured_analysis.py)
                return fn(metadata)
        return None

    def _psx_row_keys(self, store: Any, rows: Sequence[int]) -> List[str]:
        """``_path_key`` of the photo of every row, cached (the key of a row never changes while the store lives)."""
        records = store.metadata
        if self._psx_key_owner is not records:  # load() / clear() replaced the list
            self._psx_key_cache, self._psx_key_owner = {}, records
        cache, out = self._psx_key_cache, []
        for row in rows:
            key = cache.get(row)
            if key is None:
                photo_path = (records[row] or {}).get("photo_path")
                key = self._path_key(photo_path) if photo_path else ""
                cache[row] = key
            out.append(key)
        return out

    @staticmethod
    def _psx_cosine_scores(distances: np.ndarray) -> List[float]:
        """``_distance_to_score`` of the cosine metric (core/searcher.py:611-620) over an array: the same IEEE double
        operations in the same order (``min`` / ``max`` as Python evaluates them, NaN included), Python's own ``round``."""
        d = distances.astype(np.float64)
        sim = np.where(d < 1.0, d, 1.0)            # min(1.0, distance)
        sim = np.where(sim > -1.0, sim, -1.0)      # max(-1.0, .)
        score = (sim + 1.0) / 2.0
        score = np.where(score > 0.7, 0.7 + (score - 0.7) * 1.3, np.where(score < 0.3, score * 0.8, score))
        score = np.where(score < 1.0, score, 1.0)  # min(1.0, score)
        score = np.where(score > 0.0, score, 0.0)  # max(0.0, .)
        return [round(v, 6) for v in score.tolist()]

    def _psx_scores(self, distances: np.ndarray) -> List[float]:
        """Scores of a whole candidate list.  The array form is used only while it reproduces THIS Searcher's own
        ``_distance_to_score`` bit for bit on a probe of 4k distances (checked once per instance: an overridden or changed
        method simply keeps being called candidate by candidate)."""
        ok = self._psx_vector_scores_ok
        if ok is None:
            ok = False
            if getattr(self, "metric", None) == "cosine":
                probe = np.concatenate([np.linspace(-1.25, 1.25, 4001), np.array([0.4, -0.4, 1.0, -1.0, 0.0, np.nan, np.inf, -np.inf]),
                                        np.nextafter(np.float32(0.4), np.float32(1), dtype=np.float32).reshape(1)]).astype(np.float32)
                try:
                    ok = self._psx_cosine_scores(probe) == [self._distance_to_score(float(v)) for v in probe.tolist()]
                except Exception:
                    ok = False
            self._psx_vector_scores_ok = ok
        if ok:
            return self._psx_cosine_scores(distances)
        return [self._distance_to_score(float(v)) for v in distances.tolist()]

    # -- the three hooks --------------------------------------------------------------------------------------------
    def _run_single_search_round(self, **kwargs: Any):
        fast = (not kwargs.get("media_terms") and not kwargs.get("identity_terms")
                and not getattr(self, "validate_file_exists", False) and hasattr(self.vector_store._store, "search_batch")
                and not getattr(self._psx_local, "constraints", None))
        self._psx_local.lazy = bool(fast)
        self.psx_recall_stats["array_rounds" if fast else "reference_rounds"] += 1
        try:
            return super()._run_single_search_round(**kwargs)
        finally:
            self._psx_local.lazy = False

    def _hybrid_search(self, query, query_embedding, candidate_k, filters=None, allow_keyword_only_results=False, media_terms=None,
                       identity_terms=None, strict_identity_filter=False):
        """The Elasticsearch branch (core/searcher.py:855-988) with the per-candidate dicts deferred: the same dict / set
        operations on the same path strings (so that even the iteration order of ``all_paths``, which breaks score ties, is
        the reference's), the same Python-float arithmetic, the same stable sort -- but ``build_match_summary`` and the
        result dict are produced only for the entries somebody looks at."""
        if not getattr(self._psx_local, "lazy", False) or media_terms or identity_terms or self.keyword_store is None:
            return super()._hybrid_search(query, query_embedding, candidate_k, filters=filters,
                                          allow_keyword_only_results=allow_keyword_only_results, media_terms=media_terms,
                                          identity_terms=identity_terms, strict_identity_filter=strict_identity_filter)
        vector_results = self.vector_store.search(query_embedding, candidate_k)
        vector_scores: Dict[str, float] = {}
        if isinstance(vector_results, _LazyHits):
            records = vector_results.records
            for row, score in zip(vector_results.ids.tolist(), self._psx_scores(vector_results.distances)):
                metadata = records[row] or {}
                vector_scores[metadata.get("photo_path", "")] = score
        else:
            for item in vector_results:
                metadata = item.get("metadata") or {}
                vector_scores[metadata.get("photo_path", "")] = self._distance_to_score(float(item.get("distance", 0.0)))
        keyword_scores: Dict[str, float] = {}
        es_filtered_paths = None
        keyword_candidate_k = max(1, min(candidate_k, max(self.top_k * 3, 15)))
        es_filters = self._build_es_filters(filters) if filters else {}
        if es_filters:
            keyword_results = self.keyword_store.search_with_filters(query, es_filters, keyword_candidate_k)
            es_filtered_paths = set()
            for item in keyword_results:
                keyword_scores[item["photo_path"]] = item["score"]
                es_filtered_paths.add(item["photo_path"])
        else:
            for item in self.keyword_store.search(query, keyword_candidate_k):
                keyword_scores[item["photo_path"]] = item["score"]
        all_paths = set(vector_scores.keys())
        if allow_keyword_only_results:
            all_paths |= set(keyword_scores.keys())
        strict = bool(filters) and self._has_strict_filters(filters)
        entries = []  # (combined, v_score, k_score, photo_path, metadata) in the iteration order of all_paths
        for photo_path in all_paths:
            if es_filtered_paths is not None and photo_path not in es_filtered_paths and strict:
                continue
            has_vector = photo_path in vector_scores
            has_keyword = photo_path in keyword_scores
            v_score = vector_scores.get(photo_path, 0.0)
            k_score = keyword_scores.get(photo_path, 0.0)
            metadata = self._get_metadata_by_path(photo_path)
            if metadata is None:
                continue
            available_weight = 0.0
            weighted_score = 0.0
            if has_vector:
                available_weight += self.vector_weight
                weighted_score += self.vector_weight * v_score
            if has_keyword:
                available_weight += self.keyword_weight
                weighted_score += self.keyword_weight * k_score
            if available_weight <= 0:
                continue
            combined_score = weighted_score / available_weight
            combined_score *= 1.0  # _compute_metadata_boost without media / identity terms
            if has_keyword and not has_vector:
                combined_score *= 0.65
            if has_keyword and not has_vector and es_filtered_paths is None and k_score < 0.45:
                continue
            entries.append((round(combined_score, 6), round(v_score, 6), round(k_score, 6), photo_path, metadata))
        entries.sort(key=lambda e: e[0], reverse=True)  # stable, as list.sort in the reference
        return _LazyFused(self, entries)

    def _vector_results_to_combined(self, raw_results):
        if not isinstance(raw_results, _LazyHits):
            return super()._vector_results_to_combined(raw_results)
        store = self.vector_store._store
        rows = raw_results.ids.tolist()
        keys = self._psx_row_keys(store, rows)
        scores = self._psx_scores(raw_results.distances)
        kept_rows: List[int] = []
        kept_scores: List[float] = []
        seen = set()
        for row, score, key in zip(rows, scores, keys):
            # the reference drops hits without a usable path, then keeps the first hit per path key: hits arrive best
            # first and _distance_to_score is monotone, so a later duplicate never has a strictly higher score
            if not key or key in seen:
                continue
            seen.add(key)
            kept_rows.append(row)
            kept_scores.append(score)
        return _LazyCombined(self, raw_results.records, kept_rows, kept_scores)

    def _psx_candidate_words(self, records, rows, attr_words):
        """Packed EXIF words of the candidate rows when the store has not built its sidecar: packed once per row and
        remembered for as long as the store keeps the same metadata list (rows are only ever appended to it)."""
        memo = self.__dict__.get("_psx_word_memo")
        if memo is None or memo[0] is not records:
            memo = self._psx_word_memo = (records, {})
        known = memo[1]
        missing = [r for r in rows if r not in known]
        if missing:
            known.update(zip(missing, attr_words([records[r] or {} for r in missing]).tolist()))
        return np.fromiter((known[r] for r in rows), dtype=np.uint64, count=len(rows))

    def _finalize_results(self, combined_results, normalized_top_k, has_filter, constraints, search_text="", media_terms=None,
                          identity_terms=None, strict_identity_filter=False, relaxation_level=0, strip_internal=True):
        if isinstance(combined_results, _LazyFused) and not media_terms and not identity_terms and self.keyword_store is not None:
            return self._psx_finalize_fused(combined_results, normalized_top_k, relaxation_level, strip_internal)
        if not isinstance(combined_results, _LazyCombined) or media_terms or identity_terms or self.keyword_store is not None:
            if isinstance(combined_results, (_LazyCombined, _LazyFused)):
                combined_results = list(combined_results)
            return super()._finalize_results(combined_results=combined_results, normalized_top_k=normalized_top_k, has_filter=has_filter,
                                             constraints=constraints, search_text=search_text, media_terms=media_terms,
                                             identity_terms=identity_terms, strict_identity_filter=strict_identity_filter,
                                             relaxation_level=relaxation_level, strip_internal=strip_internal)
        from .exif_attrs import attr_words, build_filter, words_pass

        records, rows, scores = combined_results.records, combined_results.rows, combined_results.scores
        if has_filter:  # post-filter of the pure vector branch (core/searcher.py:1477-1480)
            flt, never = build_filter(constraints) if constraints else (None, False)
            if never:  # not representable as packed words: the reference function, candidate by candidate
                keep = [self._check_time_match_v2(records[r] or {}, constraints) for r in rows]
            elif flt is None:
                keep = [True] * len(rows)
            else:
                store = self.vector_store._store
                words = getattr(store, "_attr_words", None)
                if words is not None and getattr(store, "_attrs_built", 0) == len(records) and len(words) == len(records):
                    cand_words = np.asarray(words)[np.asarray(rows, dtype=np.int64)] if rows else np.zeros(0, np.uint64)
                else:
                    cand_words = self._psx_candidate_words(records, rows, attr_words)
                keep = words_pass(cand_words, flt).tolist()
            rows = [r for r, ok in zip(rows, keep) if ok]
            scores = [s for s, ok in zip(scores, keep) if ok]
        # thresholds and buckets: the reference's own arithmetic on the same list of scores
        strict_floor, broad_floor = self._get_round_score_floors(relaxation_level)
        if scores:
            dynamic_threshold = self._calculate_dynamic_threshold(scores, normalized_top_k)
            strict_threshold = max(dynamic_threshold, strict_floor)
            broad_threshold = min(strict_threshold - 0.05, max(broad_floor, strict_threshold * 0.84))
            broad_threshold = round(max(broad_floor, broad_threshold), 6)
        else:
            strict_threshold, broad_threshold = strict_floor, broad_floor
        buckets = [3 if s >= strict_threshold else 2 if s >= broad_threshold else 1 for s in scores]
        reliable = [i for i, b in enumerate(buckets) if b >= 3]
        generalized = [i for i, b in enumerate(buckets) if b == 2]
        prioritized = reliable + generalized
        # _fill_results_to_top_k: prioritised first, then the remaining candidates in recall order (path keys are unique here)
        chosen = prioritized[:normalized_top_k]
        if len(chosen) < normalized_top_k:
            taken = set(chosen)
            for i in range(len(rows)):
                if i not in taken:
                    chosen.append(i)
                    if len(chosen) >= normalized_top_k:
                        break
        prioritized_set = set(prioritized)
        level = max(0, int(relaxation_level))
        self._last_round_quality = {
            "raw_count": len(rows),
            "returned_count": len(chosen),
            "reliable_count": len(reliable),
            "generalized_count": len(prioritized),
            "fallback_used_count": sum(1 for i in chosen if i not in prioritized_set),
            "strict_threshold": round(strict_threshold, 6),
            "broad_threshold": round(broad_threshold, 6),
            "relaxation_level": level,
            "top_score": round(float(scores[0]), 6) if scores else 0.0,
        }
        view = _LazyCombined(self, records, rows, scores)
        final_results = []
        for rank, i in enumerate(chosen, start=1):
            item = view.item(i)
            item["_confidence_bucket"] = buckets[i]
            item["_relaxation_level"] = level
            item["rank"] = rank
            final_results.append(item)
        return self._sanitize_results(final_results) if strip_internal else final_results

    def _psx_path_keys(self, paths: Sequence[Any]) -> List[str]:
        """``_path_key`` per path string, cached (path strings are few and repeat from query to query)."""
        cache = self.__dict__.setdefault("_psx_pathkey_cache", {})
        if len(cache) > 2_000_000:
            cache.clear()
        out = []
        for path in paths:
            key = cache.get(path)
            if key is None:
                key = self._path_key(path)
                cache[path] = key
            out.append(key)
        return out

    def _psx_finalize_fused(self, fused: "_LazyFused", normalized_top_k: int, relaxation_level: int, strip_internal: bool):
        """``_finalize_results`` for the Elasticsearch branch (no post-filter there, core/searcher.py:1474-1481): de-duplicate
        by path key (first occurrence stays; the list is sorted by score, so a later duplicate is never strictly better),
        thresholds, buckets, fill -- dicts for the returned photos only."""
        entries = fused.entries
        keys = self._psx_path_keys([e[3] for e in entries])
        keep, seen = [], set()
        for i, key in enumerate(keys):
            if not key or key in seen:
                continue
            seen.add(key)
            keep.append(i)
        scores = [entries[i][0] for i in keep]
        strict_floor, broad_floor = self._get_round_score_floors(relaxation_level)
        if scores:
            dynamic_threshold = self._calculate_dynamic_threshold(scores, normalized_top_k)
            strict_threshold = max(dynamic_threshold, strict_floor)
            broad_threshold = min(strict_threshold - 0.05, max(broad_floor, strict_threshold * 0.84))
            broad_threshold = round(max(broad_floor, broad_threshold), 6)
        else:
            strict_threshold, broad_threshold = strict_floor, broad_floor
        buckets = [3 if s >= strict_threshold else 2 if s >= broad_threshold else 1 for s in scores]
        reliable = [i for i, b in enumerate(buckets) if b >= 3]
        generalized = [i for i, b in enumerate(buckets) if b == 2]
        prioritized = reliable + generalized
        chosen = prioritized[:normalized_top_k]
        if len(chosen) < normalized_top_k:
            taken = set(chosen)
            for i in range(len(keep)):
                if i not in taken:
                    chosen.append(i)
                    if len(chosen) >= normalized_top_k:
                        break
        prioritized_set = set(prioritized)
        level = max(0, int(relaxation_level))
        self._last_round_quality = {
            "raw_count": len(keep),
            "returned_count": len(chosen),
            "reliable_count": len(reliable),
            "generalized_count": len(prioritized),
            "fallback_used_count": sum(1 for i in chosen if i not in prioritized_set),
            "strict_threshold": round(strict_threshold, 6),
            "broad_threshold": round(broad_threshold, 6),
            "relaxation_level": level,
            "top_score": round(float(scores[0]), 6) if scores else 0.0,
        }
        final_results = []
        for rank, i in enumerate(chosen, start=1):
            item = fused.item(keep[i])
            item["_confidence_bucket"] = buckets[i]
            item["_relaxation_level"] = level
            item["rank"] = rank
            final_results.append(item)
        return self._sanitize_results(final_results) if strip_internal else final_results
