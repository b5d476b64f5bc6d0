"""Row-sharded exact search across the GPUs of one box: one process per GPU
(``torch.distributed``), each rank scans its contiguous row range with the single-GPU kernel,
the per-rank k best travel as 64-bit keys in ONE all-gather (k * 8 bytes per rank per query --
latency-bound over NVLink), and every rank runs the same integer merge, so all ranks return the
identical global top-k.  SURVEY.md section 8(e); there is nothing to compare with in the
reference, which is single-process.

Rows are partitioned contiguously: rank r owns ``[bounds[r], bounds[r+1])``; ids reported are
global (``id_base`` is added inside the kernel), so the result equals the single-GPU result bit
for bit -- the per-row reduction tree does not depend on where a row is scanned.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _native


def shard_bounds(n_total: int, world_size: int) -> List[int]:
    """Contiguous row ranges, ceil(n/world) rows per rank (the last ranks may be short/empty)."""
    per = -(-int(n_total) // int(world_size)) if n_total > 0 else 0
    return [min(r * per, n_total) for r in range(world_size + 1)]


class ShardedIndex:
    """One rank's handle on a row-sharded corpus.

    ``local`` is this rank's single-GPU index (``_native.NativeIndex``) holding rows
    ``[row0, row0 + local.ntotal)`` of the global corpus.  ``group`` is a ``torch.distributed``
    process group (NCCL on the GPU box; the gloo tests inject ``local_search`` / ``merge``).
    """

    def __init__(self, local, row0: int, k_max: int = 128, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None, exchange: str = "auto",
                 bounds: Optional[List[int]] = None) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.local = local
        self.row0 = int(row0)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_search = local_search
        self._merge = merge
        self.device = torch.device("cuda", local.device) if local_search is None else torch.device("cpu")
        self._bufs = {}
        # first row of every shard (+ the total), when the caller knows it (``shard_bounds``): lets
        # ``search_by_id`` name the owner of a row without a collective
        self.bounds = list(bounds) if bounds is not None else None
        # exchange: "p2p" = fused into the kernels over NVLink peer mappings (torch symmetric memory
        # provides the mappings), "nccl" = all_gather_into_tensor + merge kernel, "auto" = p2p if it can
        # be set up.  The CPU (gloo) test path always uses the collective.
        self.exchange = "nccl"
        self._seq = 0
        mixed = getattr(local, "store_dtype", 0) == _native.STORE_BF16_MASTER  # exchanges the re-scored keys, not the scan's
        if local_search is None and self.world > 1 and exchange in ("auto", "p2p") and not mixed:
            try:
                self._setup_p2p()
                self.exchange = "p2p"
            except Exception as exc:  # pragma: no cover - depends on the box
                if exchange == "p2p":
                    raise
                self._p2p_error = repr(exc)

    def _setup_p2p(self) -> None:
        import torch.distributed._symmetric_memory as symm

        t, dist = self.torch, self.dist
        nwords = (_native.exchange_bytes() + 7) // 8
        buf = symm.empty(nwords, dtype=t.int64, device=self.device)
        handle = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        buf.zero_()
        t.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._xchg_buf, self._xchg_handle = buf, handle
        self._xchg_bases = np.array([int(p) for p in handle.buffer_ptrs], dtype=np.uint64)

    # -- buffers are cached per (nq, k): the steady-state query allocates nothing ----------------
    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            t = self.torch
            kp = _native.kpad(k)
            d = self.local.d
            pin = self.device.type == "cuda"
            self._bufs[key] = dict(
                kp=kp,
                q=t.empty((nq, d), dtype=t.float32, device=self.device),
                q_host=t.empty((nq, d), dtype=t.float32, pin_memory=pin),
                mine=t.zeros((nq, kp), dtype=t.int64, device=self.device),
                gathered=t.zeros((self.world * nq, kp), dtype=t.int64, device=self.device),
                lists=t.zeros((nq, self.world, kp), dtype=t.int64, device=self.device),
                scores=t.empty((nq, k), dtype=t.float32, device=self.device),
                ids=t.empty((nq, k), dtype=t.int64, device=self.device),
                scores_host=t.empty((nq, k), dtype=t.float32, pin_memory=pin),
                ids_host=t.empty((nq, k), dtype=t.int64, pin_memory=pin),
                flags=t.zeros((nq,), dtype=t.int32, device=self.device),
                flags_host=t.zeros((nq,), dtype=t.int32, pin_memory=pin),
            )
        return self._bufs[key]

    def search_device(self, q_dev, k: int, flt=None):
        """``q_dev``: float32 ``[nq, d]`` tensor on this rank's device (same on every rank).
        Returns device tensors ``(scores [nq,k], ids [nq,k])`` -- identical on every rank.
        Everything is enqueued on the current torch stream; single queries need no host synchronisation,
        batches of ``BATCH_MIN`` or more read their certificate flags back once."""
        t = self.torch
        nq = int(q_dev.shape[0])
        b = self._buffers(nq, k)
        if self._local_search is not None:  # CPU test path: keys from the injected scanner
            b["mine"].copy_(self._local_search(q_dev, k, self.row0, flt))
        else:
            stream = t.cuda.current_stream(self.device).cuda_stream
            if self.world == 1 and self._merge is None:
                # one shard: the scan's fused cross-CTA merge already emits the final result
                self.local.search_device(q_dev.data_ptr(), nq, k, b["scores"].data_ptr(), b["ids"].data_ptr(), 0,
                                         flt=flt, id_base=self.row0, stream=stream)
                return b["scores"], b["ids"]
            if nq >= self.BATCH_MIN:
                # query batches: every rank produces its shard's keys (tensor-core GEMM + exact re-score where the
                # shard qualifies, scans otherwise), ONE all-gather of nq x k keys, one merge CTA per query
                self._local_batch_keys(q_dev, nq, k, flt, b, stream)
            elif self.exchange == "p2p":
                d = self.local.d
                for qi in range(nq):  # one fused scan+publish / wait+merge pair per query
                    self._seq += 1
                    self.local.search_exchange_device(q_dev.data_ptr() + qi * d * 4, k, self.rank, self.world, self._xchg_bases,
                                                      self._seq, b["scores"].data_ptr() + qi * k * 4, b["ids"].data_ptr() + qi * k * 8,
                                                      flt=flt, id_base=self.row0, stream=stream)
                return b["scores"], b["ids"]
            else:
                self.local.search_device(q_dev.data_ptr(), nq, k, 0, 0, b["mine"].data_ptr(), flt=flt,
                                         id_base=self.row0, stream=stream)
        if self.world > 1:
            self.dist.all_gather_into_tensor(b["gathered"], b["mine"], group=self.group)
            if nq == 1:
                lists = b["gathered"].view(1, self.world, b["kp"])
            else:
                lists = b["lists"]
                lists.copy_(b["gathered"].view(self.world, nq, b["kp"]).permute(1, 0, 2))
        else:
            lists = b["mine"].view(nq, 1, b["kp"])
        if self._merge is not None:
            s, i = self._merge(lists, k)
            b["scores"].copy_(s)
            b["ids"].copy_(i)
        else:
            stream = t.cuda.current_stream(self.device).cuda_stream
            _native.merge_keys_device(self.local.device, lists.data_ptr(), nq, self.world, k, self.local.metric,
                                      b["scores"].data_ptr(), b["ids"].data_ptr(), stream)
        return b["scores"], b["ids"]

    BATCH_MIN = 4  # batches of at least this many queries take the all-gather path on every rank

    def _local_batch_keys(self, q_dev, nq: int, k: int, flt, b, stream: int) -> None:
        """This shard's sorted key lists for a batch into ``b["mine"]``: the tensor-core path
        (``psx_search_batch_device``) when the shard qualifies, with the unproven queries re-run on the scan
        (one host synchronisation to read the certificate flags), else one scan per query."""
        local = self.local
        mine, d = b["mine"], local.d
        eligible = local.batch_supported(k) if hasattr(local, "batch_supported") else False
        if not eligible:
            local.search_device(q_dev.data_ptr(), nq, k, 0, 0, mine.data_ptr(), flt=flt, id_base=self.row0, stream=stream)
            return
        local.search_batch_device(q_dev.data_ptr(), nq, k, b["scores"].data_ptr(), b["ids"].data_ptr(), b["flags"].data_ptr(),
                                  out_keys_ptr=mine.data_ptr(), id_base=self.row0, stream=stream, flt=flt)
        b["flags_host"].copy_(b["flags"], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        for qi in b["flags_host"].nonzero().flatten().tolist():
            local.search_device(q_dev.data_ptr() + qi * d * 4, 1, k, 0, 0, mine.data_ptr() + qi * b["kp"] * 8, flt=flt,
                                id_base=self.row0, stream=stream)

    def search_by_id(self, global_id: int, k: int, flt=None, n_total: Optional[int] = None):
        """Image -> image (core/searcher.py:1751-1814): the stored vector of row ``global_id`` is the
        query.  The owning rank re-reads its row (K6 = ``index.reconstruct``, utils/vector_store.py:207),
        broadcasts it, every rank scans for k+1 and the row itself is dropped -- the reference drops it
        by path (``same_file_path``), here by id.  Collective.  Returns device tensors ``(scores [k], ids [k])``."""
        t = self.torch
        d = self.local.d
        owner_mine = self.row0 <= global_id < self.row0 + self.local.ntotal
        q = t.zeros((1, d), dtype=t.float32, device=self.device)
        if owner_mine:
            q.copy_(t.from_numpy(self.local.reconstruct(global_id - self.row0))[None, :])
        if self.world > 1:
            if self.bounds is not None:  # contiguous shards: the owner follows from the id, no collective, no host sync
                owner_rank = int(np.searchsorted(np.asarray(self.bounds[1:]), global_id, side="right"))
            else:
                owner = t.tensor([self.rank if owner_mine else -1], device=self.device)
                self.dist.all_reduce(owner, op=self.dist.ReduceOp.MAX, group=self.group)
                owner_rank = int(owner.item())
            src = owner_rank if self.group is None else self.dist.get_global_rank(self.group, owner_rank)
            self.dist.broadcast(q, src=src, group=self.group)
        s, i = self.search_device(q, k + 1, flt)
        keep = i[0] != global_id
        # the row itself is the best hit unless an exact duplicate with a lower id exists: drop it wherever it is
        idx = t.nonzero(keep, as_tuple=False)[:k, 0]
        return s[0][idx], i[0][idx]

    # -- throughput form of `search`: several queries in flight -------------------------------------------------------
    PIPELINE_DEPTH = 2

    def submit(self, q: np.ndarray, k: int, flt=None) -> int:
        """Enqueue one host query (pinned H2D, sharded search, D2H of the merged result) WITHOUT waiting for it and
        return a ticket for :meth:`collect`.  Up to ``PIPELINE_DEPTH`` tickets may be outstanding: the host-side work
        of query i+1 (copies, launches) then overlaps the scan of query i, which is what a server with several request
        threads gets.  Collective: every rank submits the same queries in the same order.  Single queries only."""
        t = self.torch
        q = np.ascontiguousarray(q, np.float32).reshape(1, -1)
        if self.device.type != "cuda":
            raise RuntimeError("submit/collect is the GPU throughput path")
        slots = self._bufs.setdefault(("pipe", k), [])
        n = getattr(self, "_pipe_next", 0)
        self._pipe_next = n + 1
        if len(slots) < self.PIPELINE_DEPTH:
            kp, d = _native.kpad(k), self.local.d
            slots.append(dict(q=t.empty((1, d), dtype=t.float32, device=self.device), q_host=t.empty((1, d), dtype=t.float32, pin_memory=True),
                              scores=t.empty((1, k), dtype=t.float32, device=self.device), ids=t.empty((1, k), dtype=t.int64, device=self.device),
                              scores_host=t.empty((1, k), dtype=t.float32, pin_memory=True), ids_host=t.empty((1, k), dtype=t.int64, pin_memory=True),
                              mine=t.zeros((1, kp), dtype=t.int64, device=self.device), gathered=t.zeros((self.world, kp), dtype=t.int64, device=self.device),
                              done=t.cuda.Event(), busy=False))
        slot = slots[n % self.PIPELINE_DEPTH]
        if slot["busy"]:
            raise RuntimeError("collect() the oldest ticket before submitting another query")
        slot["q_host"].copy_(t.from_numpy(q))
        slot["q"].copy_(slot["q_host"], non_blocking=True)
        stream = t.cuda.current_stream(self.device).cuda_stream
        if self.world == 1:
            self.local.search_device(slot["q"].data_ptr(), 1, k, slot["scores"].data_ptr(), slot["ids"].data_ptr(), 0, flt=flt,
                                     id_base=self.row0, stream=stream)
        elif self.exchange == "p2p":
            self._seq += 1
            self.local.search_exchange_device(slot["q"].data_ptr(), k, self.rank, self.world, self._xchg_bases, self._seq,
                                              slot["scores"].data_ptr(), slot["ids"].data_ptr(), flt=flt, id_base=self.row0, stream=stream)
        else:
            self.local.search_device(slot["q"].data_ptr(), 1, k, 0, 0, slot["mine"].data_ptr(), flt=flt, id_base=self.row0, stream=stream)
            self.dist.all_gather_into_tensor(slot["gathered"], slot["mine"], group=self.group)
            _native.merge_keys_device(self.local.device, slot["gathered"].data_ptr(), 1, self.world, k, self.local.metric,
                                      slot["scores"].data_ptr(), slot["ids"].data_ptr(), stream)
        slot["scores_host"].copy_(slot["scores"], non_blocking=True)
        slot["ids_host"].copy_(slot["ids"], non_blocking=True)
        slot["done"].record(t.cuda.current_stream(self.device))
        slot["busy"] = True
        return n

    def collect(self, ticket: int, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Wait for the query behind ``ticket`` and return ``(scores [1,k], ids [1,k])`` on the host."""
        slot = self._bufs[("pipe", k)][ticket % self.PIPELINE_DEPTH]
        slot["done"].synchronize()
        slot["busy"] = False
        return slot["scores_host"].numpy().copy(), slot["ids_host"].numpy().copy()

    def search(self, q: np.ndarray, k: int, flt=None) -> Tuple[np.ndarray, np.ndarray]:
        """Host in, host out (the call a user makes): pinned H2D of the query, sharded search,
        D2H of the merged result.  Collective: every rank must call it with the same query."""
        t = self.torch
        q = np.ascontiguousarray(q, np.float32)
        if q.ndim == 1:
            q = q[None]
        nq = q.shape[0]
        b = self._buffers(nq, k)
        b["q_host"].copy_(t.from_numpy(q))
        b["q"].copy_(b["q_host"], non_blocking=True)
        s, i = self.search_device(b["q"], k, flt)
        b["scores_host"].copy_(s, non_blocking=True)
        b["ids_host"].copy_(i, non_blocking=True)
        if self.device.type == "cuda":
            t.cuda.current_stream(self.device).synchronize()
            if self.exchange == "p2p" and self._local_search is None and hasattr(self.local, "exchange_status"):
                # the fused wait is bounded: a rank that never published leaves a status word instead of a trap.  Every
                # rank waits for every rank, so all of them see it; answer this query over the collective path.
                silent = self.local.exchange_status()
                if silent:
                    self.exchange_timeouts = getattr(self, "exchange_timeouts", 0) + 1
                    self.exchange = "nccl"
                    try:
                        s, i = self.search_device(b["q"], k, flt)
                        b["scores_host"].copy_(s, non_blocking=True)
                        b["ids_host"].copy_(i, non_blocking=True)
                        t.cuda.current_stream(self.device).synchronize()
                    finally:
                        self.exchange = "p2p"
        return b["scores_host"].numpy().copy(), b["ids_host"].numpy().copy()
