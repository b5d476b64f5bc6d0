"""K5 on-device hybrid fusion == the restated ``Searcher._hybrid_search`` / ``_distance_to_score``
arithmetic (oracle.flat_ip.hybrid_fuse, Python floats + round(x, 6)), bit for bit."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import flat_ip as O
from tests.conftest import has_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_gpu():
        pytest.skip("no GPU")


def _case(rng, nq, kv, kw, n=100_000, overlap=0.5, l2=False):
    vec_ids = np.stack([rng.choice(n, kv, replace=False) for _ in range(nq)]).astype(np.int64)
    if l2:
        vec_dist = np.sort(rng.random((nq, kv)).astype(np.float32) * 3, axis=1)
    else:
        vec_dist = -np.sort(-(rng.random((nq, kv)).astype(np.float32) * 2.2 - 1.1), axis=1)  # includes |ip| > 1
    kw_ids = np.full((nq, kw), -1, np.int64)
    kw_scores = np.zeros((nq, kw), np.float64)
    for qi in range(nq):
        m = int(rng.integers(min(36, kw), kw + 1))
        from_vec = rng.random(m) < overlap
        pool = rng.choice(n, m, replace=False)
        ids = np.where(from_vec, rng.permutation(vec_ids[qi])[:m] if m <= kv else pool, pool)
        _, first = np.unique(ids, return_index=True)
        ids = ids[np.sort(first)]
        u = 1.0 - rng.random(len(ids))
        kw_ids[qi, : len(ids)] = ids
        kw_scores[qi, : len(ids)] = u / u.max()  # max exactly 1.0, utils/keyword_store.py:270-279
    return vec_dist, vec_ids, kw_ids, kw_scores


@pytest.mark.parametrize("nq,kv,kw,l2", [(8, 100, 150, False), (3, 1333, 150, False), (5, 500, 45, False), (4, 1, 1, False),
                                         (6, 100, 150, True), (2, 1898, 150, False)])
def test_fusion_matches_reference_arithmetic(nq, kv, kw, l2):
    import torch

    from photo_search_engine_b200.fusion import hybrid_fuse

    rng = np.random.default_rng(nq * 1000 + kv + kw)
    vec_dist, vec_ids, kw_ids, kw_scores = _case(rng, nq, kv, kw, l2=l2)
    vec_ids[0, -1] = -1  # an unfilled slot
    vboost = 1.0 + 0.18 * (rng.random((nq, kv)) < 0.2) + 0.12 * (rng.random((nq, kv)) < 0.1)
    kboost = 1.0 + 0.18 * (rng.random((nq, kw)) < 0.2)
    for allow, filtered, wv, wk in [(True, False, 0.8, 0.2), (True, True, 0.78, 0.22), (False, False, 0.8, 0.2)]:
        out = hybrid_fuse(torch.from_numpy(vec_dist).cuda(), torch.from_numpy(vec_ids).cuda(), torch.from_numpy(kw_ids).cuda(),
                          torch.from_numpy(kw_scores).cuda(), vector_weight=wv, keyword_weight=wk,
                          metric="l2" if l2 else "cosine", allow_keyword_only=allow, keyword_filtered=filtered,
                          vec_boost=torch.from_numpy(vboost).cuda(), kw_boost=torch.from_numpy(kboost).cuda())
        ids, fused, vs, ks, count = [t.cpu().numpy() for t in out]
        for qi in range(nq):
            vhits = [(int(i), float(dd)) for i, dd in zip(vec_ids[qi], vec_dist[qi]) if i >= 0]
            khits = [(int(i), float(s)) for i, s in zip(kw_ids[qi], kw_scores[qi]) if i >= 0]
            boosts = {int(i): float(b) for i, b in zip(kw_ids[qi], kboost[qi]) if i >= 0}
            boosts.update({int(i): float(b) for i, b in zip(vec_ids[qi], vboost[qi]) if i >= 0})
            want = O.hybrid_fuse(vhits, khits, vector_weight=wv, keyword_weight=wk, metric="l2" if l2 else "cosine",
                                 allow_keyword_only=allow, keyword_filtered=filtered, boosts=boosts)
            c = int(count[qi])
            assert c == len(want)
            assert ids[qi, :c].tolist() == [w[0] for w in want]
            assert (ids[qi, c:] == -1).all()
            if l2:  # exp() differs from libm by an ulp at most -> equal after the 6-digit rounding except at boundaries
                assert np.allclose(fused[qi, :c], [w[1] for w in want], atol=1.1e-6, rtol=0)
            else:   # bit-identical doubles
                assert fused[qi, :c].tolist() == [w[1] for w in want]
                assert vs[qi, :c].tolist() == [w[2] for w in want]
                assert ks[qi, :c].tolist() == [w[3] for w in want]


def test_distance_to_score_bit_exact_on_a_dense_grid():
    """_distance_to_score over every fp32 in a dense sweep incl. the 0.3 / 0.7 knees and round-half cases."""
    import torch

    from photo_search_engine_b200.fusion import hybrid_fuse

    grid = np.concatenate([np.linspace(-1.2, 1.2, 40001), np.array([-1, -0.4, 0.4, 1, 0.0078125, -0.0078125, 0.39999998, 0.4000001])]).astype(np.float32)
    kv = 1024
    pad = (-len(grid)) % kv
    dist = np.concatenate([grid, np.zeros(pad, np.float32)]).reshape(-1, kv)
    nq = dist.shape[0]
    ids = np.tile(np.arange(kv, dtype=np.int64), (nq, 1))
    out = hybrid_fuse(torch.from_numpy(dist).cuda(), torch.from_numpy(ids).cuda(), torch.full((nq, 1), -1, dtype=torch.int64).cuda(),
                      torch.zeros((nq, 1), dtype=torch.float64).cuda(), allow_keyword_only=False)
    oid, fused, vs, ks, count = [t.cpu().numpy() for t in out]
    for qi in range(nq):
        got = {int(i): float(v) for i, v in zip(oid[qi, :kv], vs[qi, :kv])}
        for j in range(kv):
            assert got[j] == O.distance_to_score(float(dist[qi, j]))


def test_finalize_thresholds_and_buckets_match_the_reference_arithmetic():
    """psx_finalize_device vs the oracle restatement of Searcher._calculate_dynamic_threshold / _finalize_results
    (core/searcher.py:627-674, :1497-1526): thresholds equal as doubles (np.percentile interpolation, np.median,
    round(x, 6)), buckets and bucket counts equal -- over score lists of every length class (empty, <= 2 top_k, odd /
    even, concentrated / dispersed distributions, scores on the 6-digit grid the fusion produces)."""
    import torch

    from oracle import flat_ip as O
    from photo_search_engine_b200.fusion import finalize

    rng = np.random.default_rng(77)
    m = 700
    lists = [[], [0.5], [0.9, 0.1]]
    for t in range(400):
        n = int(rng.integers(1, m + 1))
        kind = t % 4
        if kind == 0:
            x = rng.random(n)
        elif kind == 1:
            x = 0.55 + 0.05 * rng.random(n)          # concentrated: cv < 0.2
        elif kind == 2:
            x = 0.3 + 0.5 * rng.random(n) ** 3        # skewed
        else:
            x = np.clip(rng.normal(0.4, 0.25, n), 0, 1)
        lists.append(sorted(np.round(x, 6).tolist(), reverse=True))
    for top_k, (sf, bf), floor in ((10, (0.4, 0.28), 0.05), (50, (0.24, 0.12), 0.05), (1, (0.32, 0.2), 0.3), (12, (0.22, 0.12), 0.0)):
        fused = torch.zeros((len(lists), m), dtype=torch.float64)
        count = torch.zeros((len(lists),), dtype=torch.int32)
        for i, l in enumerate(lists):
            fused[i, : len(l)] = torch.tensor(l, dtype=torch.float64)
            count[i] = len(l)
        strict, broad, bucket, counts = finalize(fused.cuda(), count.cuda(), top_k, sf, bf, floor)
        strict, broad, bucket, counts = strict.cpu().numpy(), broad.cpu().numpy(), bucket.cpu().numpy(), counts.cpu().numpy()
        for i, l in enumerate(lists):
            ws, wb, wbuckets = O.finalize_thresholds(l, top_k, sf, bf, floor)
            assert strict[i] == ws and broad[i] == wb, (top_k, i, len(l), strict[i], ws, broad[i], wb)
            assert bucket[i, : len(l)].tolist() == wbuckets and (bucket[i, len(l):] == 0).all()
            assert counts[i].tolist() == [wbuckets.count(3), wbuckets.count(2)]
