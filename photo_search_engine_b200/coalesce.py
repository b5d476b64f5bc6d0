"""Opt-in coalescing of concurrent single-query searches into one batched search.

The reference's Flask server is threaded (main.py:353): several request threads call
``VectorStore.search`` on the one shared instance at the same time, and FAISS serves them one after
the other (each an exhaustive scan on one core).  On the B200 backend a batch of nq queries costs
about as much as one (the corpus is streamed once), so while one search is on the GPU the newcomers
queue up, and the thread that finds the GPU free takes everything queued with the same predicate as
ONE ``search_batch``.  An idle server is unaffected: a lone request runs at once, no timer, no added
latency.  Results are those of the unbatched calls (the batched path is bit-identical to the scan).

Enabled with ``VectorStore(..., coalesce=True)`` or ``PSX_COALESCE=1``.
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Hashable, List, Optional, Tuple

import numpy as np

MAX_BATCH = 256


class _Request:
    __slots__ = ("query", "k", "key", "flt", "done", "result", "error")

    def __init__(self, query: np.ndarray, k: int, key: Hashable, flt: Any) -> None:
        self.query, self.k, self.key, self.flt = query, k, key, flt
        self.done = False
        self.result: Optional[Tuple[np.ndarray, np.ndarray]] = None
        self.error: Optional[BaseException] = None


class SearchCoalescer:
    """``run(queries [m, d] float32, k, flt) -> (scores [m, k], ids [m, k])`` is the batched backend call."""

    def __init__(self, run: Callable[[np.ndarray, int, Any], Tuple[np.ndarray, np.ndarray]]) -> None:
        self._run = run
        self._cv = threading.Condition()
        self._pending: List[_Request] = []
        self._busy = False
        self.batches = 0          # backend calls made
        self.requests = 0         # searches served
        self.largest_batch = 0

    def submit(self, query: np.ndarray, k: int, key: Hashable, flt: Any) -> Tuple[np.ndarray, np.ndarray]:
        """One query row in, its ``(scores [k], ids [k])`` out.  ``key`` identifies the predicate: only
        requests with equal keys share a batch (the batched kernels take one predicate per batch)."""
        req = _Request(query, int(k), key, flt)
        with self._cv:
            self._pending.append(req)
            while not req.done and self._busy:
                self._cv.wait()
            if req.done:
                return self._unwrap(req)
            # the GPU is free and this request is still waiting: lead a batch of everything queued with its key
            self._busy = True
            batch = [req] + [r for r in self._pending if r is not req and r.key == req.key][: MAX_BATCH - 1]
            taken = set(map(id, batch))
            self._pending = [r for r in self._pending if id(r) not in taken]
        try:
            k_max = max(r.k for r in batch)
            scores, ids = self._run(np.stack([r.query for r in batch]).astype(np.float32, copy=False), k_max, req.flt)
            for row, r in enumerate(batch):
                r.result = (scores[row, : r.k], ids[row, : r.k])  # top-k is a prefix of top-k_max
        except BaseException as exc:  # every waiter of this batch sees the failure
            for r in batch:
                r.error = exc
        finally:
            with self._cv:
                for r in batch:
                    r.done = True
                self._busy = False
                self.batches += 1
                self.requests += len(batch)
                self.largest_batch = max(self.largest_batch, len(batch))
                self._cv.notify_all()
        return self._unwrap(req)

    @staticmethod
    def _unwrap(req: _Request) -> Tuple[np.ndarray, np.ndarray]:
        if req.error is not None:
            raise req.error
        assert req.result is not None
        return req.result
