"""Pin the oracle: golden vectors / fixtures the reference itself ships."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

from oracle import flat_ip as O
from tests.conftest import GOLDEN, REFERENCE


def _fake_embedding(text: str, d: int = 8):
    # tests/helpers.py:6-15 FakeEmbeddingService, restated
    seed = float(sum(ord(c) for c in text) % 13)
    return [seed + float(i) for i in range(d)]


def test_flat_file_reproduced_byte_for_byte(tmp_path):
    """add_item + save of the row the reference indexed == the FAISS-written fixture bytes."""
    want = open(os.path.join(GOLDEN, "build_smoke.idx"), "rb").read()
    store = O.OracleVectorStore(8, str(tmp_path / "idx"), str(tmp_path / "meta.json"))
    store.add_item(_fake_embedding("photo 图片 32x24"), {"photo_path": "x"})
    store.save()
    assert open(tmp_path / "idx", "rb").read() == want
    assert json.load(open(tmp_path / "idx.meta.json")) == json.load(open(os.path.join(GOLDEN, "build_smoke.idx.meta.json")))


def test_flat_file_round_trip():
    index, info = O.read_index(os.path.join(GOLDEN, "build_smoke.idx"))
    assert (info["d"], info["ntotal"], info["metric"]) == (8, 1, 0)
    row = index.reconstruct(0)
    want = np.array(O.normalize_vector(_fake_embedding("photo 图片 32x24")), np.float32)
    assert np.array_equal(row, want)


def test_ihnf_container_parsed(tmp_path):
    """FAISS IHNf container (graph header + nested flat block) -> the 77x4096 real rows."""
    gold = json.load(open(os.path.join(GOLDEN, "real77_topk.json")))
    path = tmp_path / "photo_search.index"
    with open(path, "wb") as f:
        f.write(open(os.path.join(GOLDEN, "real77_hnsw_header.bin"), "rb").read())
        f.write(open(os.path.join(GOLDEN, "real77.index"), "rb").read())
    index, info = O.read_index(str(path))
    assert info["fourcc"] == "IHNf" and info["d"] == 4096 and info["ntotal"] == 77
    for key, val in gold["ihnf"].items():
        assert info[key] == val
    flat, _ = O.read_index(os.path.join(GOLDEN, "real77.index"))
    assert np.array_equal(flat._matrix(), index._matrix())
    norms = np.linalg.norm(index._matrix(), axis=1)
    assert np.all(np.abs(norms - 1) < 1e-6)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_fixtures_match_reference_checkout():
    assert open(os.path.join(GOLDEN, "build_smoke.idx"), "rb").read() == open(
        os.path.join(REFERENCE, "pytest-tmp", "build-smoke", "data", "idx"), "rb").read()
    raw = open(os.path.join(REFERENCE, "data", "photo_search.index"), "rb").read()
    assert raw[30865:] == open(os.path.join(GOLDEN, "real77.index"), "rb").read()


def test_real77_known_answers():
    gold = json.load(open(os.path.join(GOLDEN, "real77_topk.json")))
    index, _ = O.read_index(os.path.join(GOLDEN, "real77.index"))
    x = index._matrix()
    D, I = index.search(x, gold["k"])
    assert I.tolist() == gold["ids"]
    assert np.allclose(D, np.array(gold["scores"], np.float32), rtol=1e-6, atol=1e-7)
    assert I[:, 0].tolist() == list(range(77))  # every row is its own best hit
    # float64 ground truth agrees within the north-star tolerance
    D64 = np.sort(x.astype(np.float64) @ x.astype(np.float64).T, axis=1)[:, ::-1][:, : gold["k"]]
    assert np.allclose(D, D64, rtol=1e-5, atol=1e-6)


def test_real77_predicate_known_answers():
    gold = json.load(open(os.path.join(GOLDEN, "real77_topk.json")))
    meta = json.load(open(os.path.join(GOLDEN, "real77_time.json"), encoding="utf-8"))
    index, _ = O.read_index(os.path.join(GOLDEN, "real77.index"))
    x = index._matrix()
    for case in gold["predicates"]:
        mask = np.array([O.check_time_match_v2(m, case["constraints"]) for m in meta])
        assert np.nonzero(mask)[0].tolist() == case["pass_rows"]
        _, I = index.search(x[:4], 10, mask=mask)
        assert I.tolist() == case["ids"]


def test_reference_tie_and_normalisation_semantics(tmp_path):
    """tests/test_vector_store.py:35-51 and :163-175 of the reference, on the oracle."""
    for d in (8, 768, 1024, 4096):
        s = O.OracleVectorStore(d, str(tmp_path / "i"), str(tmp_path / "m"))
        s.add_item([0.1] * d, {"id": 1})
        s.add_item([0.5] * d, {"id": 2, "photo_path": "/b.jpg"})
        hit = s.search([0.1] * d, 1)
        assert len(hit) == 1 and hit[0]["metadata"]["id"] == 1 and hit[0]["distance"] >= 0.0
        assert abs(s.get_embedding_by_photo_path("/b.jpg")[0] - 1 / d**0.5) < 5e-7


def test_l2_and_unfilled_slots():
    ix = O.OracleIndexFlat(4, O.METRIC_L2)
    ix.add(np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0]], np.float32))
    D, I = ix.search(np.array([[0, 0, 0, 0]], np.float32), 5)
    assert I[0].tolist() == [0, 1, 2, -1, -1]
    assert D[0, :3].tolist() == [0.0, 1.0, 4.0] and np.isinf(D[0, 3])


def test_call_site_arithmetic():
    # core/searcher.py:605-625 pins from tests/test_searcher.py:33-46
    assert O.distance_to_score(1.0) > 0.9 and O.distance_to_score(-1.0) < 0.1
    assert O.distance_to_score(0.0, "l2") == 1.0
    # core/searcher.py:771-820: the k the kernel has to serve
    assert O.calculate_candidate_k(10_000_000, 12, False) == 500
    assert O.calculate_candidate_k(10_000_000, 50, True, 3) == 1333
    assert O.calculate_candidate_k(40, 12, False) == 40


def test_oracle_thresholds_equal_the_reference_searcher():
    """oracle.calculate_dynamic_threshold / finalize_thresholds restate core/searcher.py:627-674, :822-853, :1497-1526:
    pinned here against the reference's own methods (checkout or staged copy) on random score lists."""
    import sys
    import types

    import pytest

    from oracle import flat_ip as O
    from oracle import stage_reference

    ref = stage_reference.locate()
    if ref is None:
        pytest.skip("reference Searcher neither checked out nor staged")
    saved = list(sys.path)
    sys.path.insert(0, ref)
    try:
        if "utils.vector_store" not in sys.modules:
            shim = types.ModuleType("utils.vector_store")
            shim.VectorStore = object
            sys.modules["utils.vector_store"] = shim
        from core.searcher import Searcher
    finally:
        sys.path[:] = saved
    s = Searcher.__new__(Searcher)
    s.query_dynamic_threshold_floor, s.query_strict_floor_min, s.query_broad_floor_min = 0.05, 0.22, 0.12
    rng = np.random.default_rng(3)
    for t in range(300):
        n = int(rng.integers(0, 600))
        x = rng.random(n) if t % 2 else 0.5 + 0.1 * rng.random(n)
        scores = sorted(np.round(x, 6).tolist(), reverse=True)
        top_k = int(rng.integers(1, 51))
        level = int(rng.integers(0, 4))
        assert O.calculate_dynamic_threshold(scores, top_k, 0.05) == s._calculate_dynamic_threshold(scores, top_k)
        sf, bf = s._get_round_score_floors(level)
        strict, broad, buckets = O.finalize_thresholds(scores, top_k, sf, bf, 0.05)
        want = [s._assign_confidence_bucket(item={"score": v}, strict_threshold=strict, broad_threshold=broad, media_terms=[],
                                            identity_terms=[], strict_identity_filter=False) for v in scores]
        assert buckets == want


def _reference_searcher_class():
    import sys
    import types

    from oracle import stage_reference

    ref = stage_reference.locate()
    if ref is None:
        pytest.skip("reference Searcher neither checked out nor staged")
    saved = list(sys.path)
    sys.path.insert(0, ref)
    try:
        if "utils.vector_store" not in sys.modules:
            shim = types.ModuleType("utils.vector_store")
            shim.VectorStore = object
            sys.modules["utils.vector_store"] = shim
        from core.searcher import Searcher
    finally:
        sys.path[:] = saved
    return Searcher


def test_oracle_call_site_numerics_equal_the_reference_searcher():
    """oracle.distance_to_score / calculate_candidate_k / hybrid_fuse restate core/searcher.py:605-625, :771-820, :855-988:
    pinned against the reference's own methods on swept and random inputs (fakes stand in for the stores)."""
    Searcher = _reference_searcher_class()

    class Store:
        metric, index_path, metadata_path = "cosine", "/tmp/_o.index", "/tmp/_o.json"

        def __init__(self, n):
            self.metadata = [{"photo_path": f"/p/{i}.jpg", "description": f"d{i}"} for i in range(n)]
            self.hits = []

        def get_total_items(self):
            return len(self.metadata)

        def search(self, q, k):
            return self.hits[:k]

        def load(self):
            return True

    class Keywords:
        hits = []

        def search(self, query, k):
            return self.hits[:k]

        def search_with_filters(self, query, filters, k):
            return self.hits[:k]

    class Emb:
        def generate_embedding(self, text):
            return [0.0]

    class NoTime:
        def extract_time_constraints(self, query):
            return {}

    rng = np.random.default_rng(8)
    vs, ks = Store(400), Keywords()
    s = Searcher(embedding=Emb(), time_parser=NoTime(), vector_store=vs, keyword_store=ks, query_formatter=None)
    s.index_loaded = True
    # _distance_to_score, both metrics, the knees, the clamps, NaN
    sweep = np.concatenate([rng.uniform(-1.3, 1.3, 20_000), [0.4, -0.4, 1.0, -1.0, 0.0, np.nan, np.inf, -np.inf]]).astype(np.float32)
    for metric, values in (("cosine", sweep), ("l2", np.abs(sweep) * 6 - 0.5)):
        s.metric = metric
        for v in values.tolist():
            a, b = O.distance_to_score(v, metric), s._distance_to_score(v)
            assert a == b or (a != a and b != b), (metric, v, a, b)
    s.metric = "cosine"
    # _calculate_candidate_k over every corpus-size regime
    for n in (0, 1, 50, 51, 500, 501, 5000, 5001, 49_999, 60_000, 1_000_000):
        vs.metadata = [None] * n
        for top_k in (1, 5, 12, 50):
            for has_filter in (False, True):
                for level in range(6):
                    assert s._calculate_candidate_k(top_k, has_filter, level) == O.calculate_candidate_k(n, top_k, has_filter, level)
    # _hybrid_search fusion: random overlaps of vector candidates and keyword hits, keyword-only hits allowed or not
    vs.metadata = [{"photo_path": f"/p/{i}.jpg", "description": f"d{i}"} for i in range(400)]
    for trial in range(40):
        nv, nk = int(rng.integers(1, 200)), int(rng.integers(0, 60))
        ids = rng.permutation(400)[:nv]
        dist = np.sort(rng.uniform(-0.3, 1.0, nv).astype(np.float32))[::-1]
        vs.hits = [{"metadata": vs.metadata[i], "distance": float(x)} for i, x in zip(ids, dist)]
        kid = rng.permutation(400)[:nk]
        ks.hits = [{"photo_path": f"/p/{i}.jpg", "score": float(x)} for i, x in zip(kid, np.sort(rng.uniform(0, 1, nk))[::-1])]
        allow = bool(trial % 2)
        out = s._hybrid_search("q", [0.0], nv, filters=None, allow_keyword_only_results=allow)
        got = [(int(r["photo_path"][3:-4]), r["score"], r["vector_score"], r["keyword_score"]) for r in out]
        kk = max(1, min(nv, max(s.top_k * 3, 15)))  # core/searcher.py:905
        want = O.hybrid_fuse([(int(i), float(x)) for i, x in zip(ids, dist)], [(int(i), h["score"]) for i, h in zip(kid, ks.hits)][:kk],
                             vector_weight=s.vector_weight, keyword_weight=s.keyword_weight, allow_keyword_only=allow)
        assert sorted(got) == sorted(want), trial                       # same entries, same three scores each
        assert [g[1] for g in got] == [w[1] for w in want], trial       # same descending order of fused scores
