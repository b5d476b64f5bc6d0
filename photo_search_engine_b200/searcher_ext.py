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
from typing import Any, Dict, List, Optional, Sequence

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
