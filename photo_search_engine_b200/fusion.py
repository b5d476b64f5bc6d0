"""On-device hybrid recall fusion (K5): vector candidates + keyword hits -> fused ranking.

Array form of the numeric core of ``Searcher._hybrid_search`` (core/searcher.py:893-986) with the
score map of ``Searcher._distance_to_score`` (core/searcher.py:605-625); string work (path
normalisation, ``build_match_summary``, metadata-term matching for the boost) stays on the host
and enters as the optional per-hit ``*_boost`` multipliers.
"""
from __future__ import annotations

from typing import Optional, Tuple

from . import _native


def hybrid_fuse(vec_dist, vec_ids, kw_ids, kw_scores, *, vector_weight: float = 0.8, keyword_weight: float = 0.2,
                metric: str = "cosine", allow_keyword_only: bool = True, keyword_filtered: bool = False,
                vec_boost=None, kw_boost=None) -> Tuple["torch.Tensor", "torch.Tensor", "torch.Tensor", "torch.Tensor", "torch.Tensor"]:
    """All inputs are CUDA tensors on one device: ``vec_dist`` float32 ``[nq,kv]``, ``vec_ids`` int64
    ``[nq,kv]`` (-1 = empty), ``kw_ids`` int64 ``[nq,kw]`` (-1 = empty), ``kw_scores`` float64 ``[nq,kw]``.
    Returns ``(ids, fused, vector_score, keyword_score, count)``; per query the first ``count`` slots
    are valid, sorted by fused score descending then id ascending.  Runs on the current stream."""
    import torch

    nq, kv = vec_ids.shape
    kw = kw_ids.shape[1]
    dev = vec_ids.device
    E = kv + kw
    vec_dist = vec_dist.contiguous().float()
    vec_ids = vec_ids.contiguous().long()
    kw_ids = kw_ids.contiguous().long()
    kw_scores = kw_scores.contiguous().double()
    vb = vec_boost.contiguous().double() if vec_boost is not None else None
    kb = kw_boost.contiguous().double() if kw_boost is not None else None
    out_ids = torch.empty((nq, E), dtype=torch.int64, device=dev)
    out_fused = torch.empty((nq, E), dtype=torch.float64, device=dev)
    out_v = torch.empty((nq, E), dtype=torch.float64, device=dev)
    out_k = torch.empty((nq, E), dtype=torch.float64, device=dev)
    count = torch.empty((nq,), dtype=torch.int32, device=dev)
    _native.check(_native.load_library().psx_hybrid_fuse_device(
        dev.index or 0, nq, kv, vec_dist.data_ptr(), vec_ids.data_ptr(), vb.data_ptr() if vb is not None else None, kw,
        kw_ids.data_ptr() if kw else None, kw_scores.data_ptr() if kw else None, kb.data_ptr() if kb is not None else None,
        float(vector_weight), float(keyword_weight), _native.METRIC_IP if metric == "cosine" else _native.METRIC_L2,
        int(bool(allow_keyword_only)), int(bool(keyword_filtered)), out_ids.data_ptr(), out_fused.data_ptr(), out_v.data_ptr(),
        out_k.data_ptr(), count.data_ptr(), torch.cuda.current_stream(dev).cuda_stream or None))
    return out_ids, out_fused, out_v, out_k, count


def finalize(fused, count, top_k: int, strict_floor: float, broad_floor: float, threshold_floor: float = 0.05):
    """The numeric part of ``Searcher._finalize_results`` (core/searcher.py:1497-1526) for a batch of queries on the
    device: ``fused`` float64 ``[nq, m]`` (descending per query) and ``count`` int32 ``[nq]`` as ``hybrid_fuse`` returns
    them; the floors are the round's ``Searcher._get_round_score_floors(relaxation_level)``.  Returns
    ``(strict_threshold [nq], broad_threshold [nq], bucket [nq, m], counts [nq, 2])`` -- thresholds bit-identical to the
    reference's Python arithmetic (``_calculate_dynamic_threshold`` included), buckets 3 / 2 / 1 by score."""
    import torch

    fused = fused.contiguous().double()
    count = count.contiguous().int()
    nq, m = fused.shape
    dev = fused.device
    strict = torch.empty((nq,), dtype=torch.float64, device=dev)
    broad = torch.empty((nq,), dtype=torch.float64, device=dev)
    bucket = torch.empty((nq, m), dtype=torch.int32, device=dev)
    counts = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    _native.check(_native.load_library().psx_finalize_device(
        dev.index or 0, nq, m, fused.data_ptr(), count.data_ptr(), int(top_k), float(strict_floor), float(broad_floor),
        float(threshold_floor), strict.data_ptr(), broad.data_ptr(), bucket.data_ptr(), counts.data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream or None))
    return strict, broad, bucket, counts
