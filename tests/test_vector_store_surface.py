"""Surface + host logic of the drop-in ``VectorStore``: every case runs over the oracle-backed
fake backend (CPU suite) and, under ``-m gpu``, over the real CUDA backend.

The first block restates the reference's own ``tests/test_vector_store.py`` case by case;
the rest covers persistence compatibility, the EXIF sidecar and the additive batch API.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

from oracle import flat_ip as O
from tests.conftest import GOLDEN, has_gpu

DIMS = (8, 768, 4096)


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def VS(request, monkeypatch):
    """The drop-in class over the oracle-backed fake (CPU suite) or the real CUDA backend
    (``-m gpu``): the same cases must hold for both."""
    from photo_search_engine_b200.vector_store import VectorStore

    if request.param == "fake":
        from tests._fake_backend import FakeIndex

        monkeypatch.setattr(VectorStore, "_index_factory", staticmethod(FakeIndex))
    elif not has_gpu():
        pytest.skip("no GPU")
    return VectorStore


def _paths(tmp_path):
    return str(tmp_path / "index.bin"), str(tmp_path / "metadata.json")


# ---- reference tests/test_vector_store.py, restated ----------------------------------------
@pytest.mark.parametrize("d", DIMS)
def test_init(VS, tmp_path, d):
    ip, mp = _paths(tmp_path)
    s = VS(dimension=d, index_path=ip, metadata_path=mp)
    assert (s.dimension, s.index_path, s.metadata_path, s.get_total_items()) == (d, ip, mp, 0)
    assert s.meta_path == ip + ".meta.json"


@pytest.mark.parametrize("d", DIMS)
def test_add_and_search_item(VS, tmp_path, d):
    s = VS(d, *_paths(tmp_path))
    s.add_item([0.1] * d, {"id": 1, "photo": "test1.jpg"})
    s.add_item([0.5] * d, {"id": 2, "photo": "test2.jpg"})
    r = s.search([0.1] * d, top_k=1)
    assert len(r) == 1 and r[0]["metadata"]["id"] == 1 and r[0]["distance"] >= 0.0


@pytest.mark.parametrize("index_type", ["flat", "hnsw"])
def test_save_and_load(VS, tmp_path, index_type):
    ip, mp = _paths(tmp_path)
    kw = dict(index_type=index_type, hnsw_m=16, hnsw_ef_construction=80, hnsw_ef_search=48)
    s = VS(16, ip, mp, **kw)
    s.add_item([0.1] * 16, {"photo_path": "/a.jpg", "id": 1})
    s.add_item([0.2] * 16, {"photo_path": "/b.jpg", "id": 2})
    s.save()
    assert os.path.exists(ip) and os.path.exists(mp) and os.path.exists(ip + ".meta.json")
    t = VS(16, ip, mp, **kw)
    assert t.load() is True
    assert t.get_total_items() == 2 and t.has_photo_path("/b.jpg")
    meta = json.load(open(ip + ".meta.json"))
    assert meta == {"index_type": index_type, "metric": "cosine", "dimension": 16, "hnsw_m": 16,
                    "hnsw_ef_construction": 80, "hnsw_ef_search": 48}


def test_load_nonexistent(VS, tmp_path):
    assert VS(8, str(tmp_path / "no.bin"), str(tmp_path / "no.json")).load() is False


def test_load_metadata_mismatch(VS, tmp_path):
    ip, mp = _paths(tmp_path)
    s = VS(8, ip, mp)
    s.add_item([0.1] * 8, {"id": 1})
    s.save()
    open(mp, "w").write("[]")
    with pytest.raises(ValueError):
        VS(8, ip, mp).load()


def test_dimension_mismatch(VS, tmp_path):
    s = VS(8, *_paths(tmp_path))
    s.add_item([0.1] * 8, {"id": 1})
    with pytest.raises(ValueError):
        s.add_item([0.1] * 9, {"id": 2})
    with pytest.raises(ValueError):
        s.search([0.1] * 9, 1)
    with pytest.raises(ValueError):
        s.add_item(None, {})


def test_search_empty_and_top_k_limit(VS, tmp_path):
    s = VS(8, *_paths(tmp_path))
    assert s.search([0.1] * 8, top_k=10) == []
    for i in range(10):
        s.add_item([i * 0.1] * 8, {"id": i})
    assert len(s.search([0.1] * 8, top_k=5)) == 5
    assert len(s.search([0.1] * 8, top_k=50)) == 10  # k = min(top_k, ntotal)


@pytest.mark.parametrize("d", DIMS)
def test_get_embedding_by_photo_path(VS, tmp_path, d):
    s = VS(d, *_paths(tmp_path))
    s.add_item([0.1] * d, {"photo_path": "/a.jpg"})
    s.add_item([0.2] * d, {"photo_path": "/b.jpg"})
    e = s.get_embedding_by_photo_path("/b.jpg")
    assert len(e) == d and abs(e[0] - 1.0 / d**0.5) < 5e-7
    assert s.get_embedding_by_photo_path("/missing.jpg") is None
    e[0] = 123.0  # a fresh list every time
    assert s.get_embedding_by_photo_path("/b.jpg")[0] != 123.0


# ---- constructor contract (main.py:59-68, utils/vector_store.py:44-56) -------------------------
def test_constructor_contract(VS, tmp_path):
    ip, mp = _paths(tmp_path)
    with pytest.raises(ValueError):
        VS(8, ip, mp, metric="dot")
    with pytest.raises(ValueError):
        VS(8, ip, mp, index_type="ivf")
    s = VS(None, ip, mp, metric=" COSINE ", index_type=None, hnsw_m=1, hnsw_ef_construction=1, hnsw_ef_search=1)
    assert (s.metric, s.index_type, s.hnsw_m, s.hnsw_ef_construction, s.hnsw_ef_search) == ("cosine", "flat", 4, 8, 8)
    assert s.index is None and s.get_total_items() == 0 and s.search([1.0], 3) == []
    with pytest.raises(ValueError):
        s.save()
    s.add_item([3.0, 4.0], {"photo_path": "p"})  # lazy dimension
    assert s.dimension == 2 and s.get_embedding_by_photo_path("p") == pytest.approx([0.6, 0.8])
    assert VS(4, ip, mp, metric=None).metric == "l2"


def test_metadata_objects_are_shared(VS, tmp_path):
    s = VS(4, *_paths(tmp_path))
    record = {"photo_path": "a", "x": 1}
    s.add_item([1, 0, 0, 0], record)
    assert s.metadata[0] is record and s.search([1, 0, 0, 0], 1)[0]["metadata"] is record


def test_duplicate_path_last_wins_and_clear(VS, tmp_path):
    s = VS(4, *_paths(tmp_path))
    s.add_item([1, 0, 0, 0], {"photo_path": "dup"})
    s.add_item([0, 1, 0, 0], {"photo_path": "dup"})
    assert s.get_embedding_by_photo_path("dup") == [0.0, 1.0, 0.0, 0.0]
    s.clear()
    assert s.get_total_items() == 0 and s.metadata == [] and not s.has_photo_path("dup") and s.dimension == 4
    s.add_item([0, 0, 1, 0], {"photo_path": "z"})
    assert s.get_total_items() == 1


def test_l2_metric(VS, tmp_path):
    s = VS(3, *_paths(tmp_path), metric="l2")
    s.add_item([0.0, 0.0, 0.0], {"photo_path": "o"})
    s.add_item([3.0, 4.0, 0.0], {"photo_path": "far"})
    r = s.search([0.0, 0.0, 0.0], 2)
    assert [h["metadata"]["photo_path"] for h in r] == ["o", "far"] and [h["distance"] for h in r] == [0.0, 25.0]
    assert s.get_embedding_by_photo_path("far") == [3.0, 4.0, 0.0]  # not normalised for l2


# ---- persistence compatibility ----------------------------------------------------------------
def test_save_is_byte_identical_to_faiss_fixture(VS, tmp_path):
    ip, mp = _paths(tmp_path)
    s = VS(8, ip, mp)
    s.add_item([11.0 + i for i in range(8)], {"photo_path": "x"})
    s.save()
    assert open(ip, "rb").read() == open(os.path.join(GOLDEN, "build_smoke.idx"), "rb").read()
    assert open(ip + ".meta.json").read() == open(os.path.join(GOLDEN, "build_smoke.idx.meta.json")).read()
    # and the oracle's reader (independent restatement of the format) accepts it
    ix, info = O.read_index(ip)
    assert info["ntotal"] == 1


def test_loads_faiss_hnsw_container(VS, tmp_path):
    ip, mp = _paths(tmp_path)
    with open(ip, "wb") as f:
        f.write(open(os.path.join(GOLDEN, "real77_hnsw_header.bin"), "rb").read())
        f.write(open(os.path.join(GOLDEN, "real77.index"), "rb").read())
    json.dump({"index_type": "hnsw", "metric": "cosine", "dimension": 4096, "hnsw_m": 48,
               "hnsw_ef_construction": 320, "hnsw_ef_search": 192}, open(ip + ".meta.json", "w"))
    meta = json.load(open(os.path.join(GOLDEN, "real77_time.json"), encoding="utf-8"))
    for i, m in enumerate(meta):
        m["photo_path"] = f"/p/{i}.jpg"
    json.dump(meta, open(mp, "w"))
    s = VS(4096, ip, mp, index_type="hnsw", hnsw_m=48, hnsw_ef_construction=320, hnsw_ef_search=192)
    assert s.load() and s.get_total_items() == 77
    gold = json.load(open(os.path.join(GOLDEN, "real77_topk.json")))
    q = s.get_embedding_by_photo_path("/p/5.jpg")
    hits = s.search(q, 10)
    assert [s.metadata.index(h["metadata"]) for h in hits] == gold["ids"][5]
    # config/type mismatches are ValueErrors (utils/vector_store.py:125-140)
    with pytest.raises(ValueError):
        VS(4096, ip, mp, index_type="flat").load()
    with pytest.raises(ValueError):
        VS(4096, ip, mp, index_type="hnsw", metric="l2").load()
    os.remove(ip + ".meta.json")
    with pytest.raises(ValueError):
        VS(4096, ip, mp, index_type="hnsw").load()


def test_corrupt_meta_json(VS, tmp_path):
    ip, mp = _paths(tmp_path)
    s = VS(4, ip, mp)
    s.add_item([1, 2, 3, 4], {})
    s.save()
    open(ip + ".meta.json", "w").write("[1, 2]")
    with pytest.raises(ValueError):
        VS(4, ip, mp).load()


# ---- EXIF sidecar ---------------------------------------------------------------------------------
def _real_meta():
    return json.load(open(os.path.join(GOLDEN, "real77_time.json"), encoding="utf-8"))


def test_attr_words_reproduce_reference_predicate():
    """Packed word + psx_filter == Searcher._check_time_match_v2 on the reference's real metadata
    and on adversarial records."""
    from photo_search_engine_b200.exif_attrs import attr_words, build_filter
    from tests._fake_backend import attr_pass_np

    meta = _real_meta()
    meta += [
        {},
        {"exif_data": None, "time_info": None},
        {"exif_data": {"datetime": "2024:02:29 23:59:59"}, "time_info": {"season": "冬天", "year": 2024, "month": 2}},
        {"exif_data": {"datetime": ""}, "time_info": {"season": "夏天", "time_period": "下午", "year": 2023, "month": 7,
                                                     "datetime_str": "2023-07-01T14:00:00"}},
        {"exif_data": {"datetime": "garbage"}, "time_info": {"season": "夏天", "year": 2023, "month": 7}},
        {"exif_data": {"datetime": "2023-12-31T23:59:59"}, "time_info": O.time_info_from_exif("2023-12-31T23:59:59")},
        {"exif_data": {"datetime": "2024-01-01T00:00:00"}, "time_info": O.time_info_from_exif("2024-01-01T00:00:00")},
        {"exif_data": {"datetime": "0001-01-01T00:00:00"}, "time_info": O.time_info_from_exif("0001-01-01T00:00:00")},
        {"exif_data": {"datetime": "9999-12-31T23:59:59"}, "time_info": O.time_info_from_exif("9999-12-31T23:59:59")},
    ]
    words = attr_words(meta)
    cases = [
        {"season": "夏天"}, {"season": "春天"}, {"time_period": "下午"}, {"time_period": "夜晚", "season": "冬天"},
        {"year": 2023}, {"year": 2024, "month": 2}, {"month": 12},
        {"start_date": "2023-01-01", "end_date": "2023-12-31"}, {"start_date": "2024-01-01"},
        {"end_date": "2023-12-31"}, {"end_date": "2023-12-31T23:59:58"}, {"start_date": "2023/07/01 14:00:00"},
        {"start_date": "not a date"}, {"start_date": "20230101", "end_date": "20231231", "season": "夏天"},
        {"start_date": "0001-01-01", "end_date": "9999-12-31"},
        {"season": None, "year": 0, "start_date": None},
    ]
    for c in cases:
        want = np.array([O.check_time_match_v2(m, c) for m in meta])
        flt, never = build_filter(c)
        if flt is None:
            assert want.all(), c
            continue
        got = np.zeros(len(meta), bool) if never else attr_pass_np(words, flt)
        assert got.tolist() == want.tolist(), c
    # unrepresentable constraints match nothing, as `!=` does in the reference
    for c in ({"season": "雨季"}, {"year": "2023"}, {"month": 13}):
        flt, never = build_filter(c)
        assert never and not any(O.check_time_match_v2(m, c) for m in meta)


def test_array_form_of_distance_to_score_is_bit_identical():
    """FusedRecallMixin computes a candidate list's scores as one array; every value equals the oracle's restatement of
    Searcher._distance_to_score (core/searcher.py:605-625) on 200k fp32 distances, the kinks, the clamps and NaN / inf."""
    from photo_search_engine_b200.searcher_ext import FusedRecallMixin

    rng = np.random.default_rng(17)
    d = np.concatenate([rng.uniform(-1.3, 1.3, 200_000), [0.4, -0.4, 1.0, -1.0, 0.0, -0.0, np.nan, np.inf, -np.inf],
                        np.nextafter(np.float32([0.4, 0.4, -0.4, -0.4]), np.float32([1, -1, 1, -1]))]).astype(np.float32)
    got = FusedRecallMixin._psx_cosine_scores(d)
    want = [O.distance_to_score(float(v), "cosine") for v in d.tolist()]
    assert got == want


def test_parse_date_fast_path_equals_the_format_loop():
    """exif_attrs.parse_date short-cuts the shape the indexer writes (YYYY-MM-DDTHH:MM:SS); it must accept, refuse and
    return exactly what the oracle's restatement of Searcher._parse_date (core/searcher.py:1963-2001) does."""
    from photo_search_engine_b200.exif_attrs import parse_date

    texts = ["2023-07-01T14:00:00", "0001-01-01T00:00:00", "9999-12-31T23:59:59", "2024-02-29T23:59:59", "2023-02-29T00:00:00",
             "2023-13-01T00:00:00", "2023-00-10T00:00:00", "2023-06-31T12:00:00", "2023-06-15T24:00:00", "2023-06-15T23:60:00",
             "2023-06-15T23:59:60", "2023-06-15T23:59:61", "0000-01-01T00:00:00", " 2023-07-01T14:00:00 ", "2023-07-01T14:00:00\x00\x00",
             "2023-07-01 14:00:00", "2023:07:01 14:00:00", "2023/07/01 14:00:00", "2023-07-01", "2023/07/01", "20230701",
             "2023-7-1T4:5:6", "2023-07-01T14:00", "2023-07-01T14:00:00.250", "2023-07-01T14:00:00+08:00", "２０２３-07-01T14:00:00",
             "2023-07-01t14:00:00", "2023-07-01T14-00-00", "", "garbage", None, 20230701]
    for text in texts:
        for is_end in (False, True):
            assert parse_date(text, is_end) == O.parse_date(text, is_end), (text, is_end)


def test_filtered_search_equals_reference_post_filter_superset(VS, tmp_path):
    ix, _ = O.read_index(os.path.join(GOLDEN, "real77.index"))
    x = ix._matrix()
    meta = _real_meta()
    s = VS(4096, *_paths(tmp_path))
    for i, m in enumerate(meta):
        m["photo_path"] = f"/p/{i}.jpg"
        s.add_item(x[i].tolist(), m)
    gold = json.load(open(os.path.join(GOLDEN, "real77_topk.json")))
    for case in gold["predicates"]:
        for qi in range(4):
            hits = s.search(x[qi].tolist(), 10, constraints=case["constraints"])
            got = [int(h["metadata"]["photo_path"].split("/")[-1][:-4]) for h in hits]
            assert got == [i for i in case["ids"][qi] if i != -1]
            # identical to filtering the reference's full ranking afterwards
            full = s.search(x[qi].tolist(), 77)
            post = [h for h in full if O.check_time_match_v2(h["metadata"], case["constraints"])][:10]
            assert [h["metadata"] for h in hits] == [h["metadata"] for h in post]
    assert s.search(x[0].tolist(), 5, constraints={"season": "雨季"}) == []
    # rows appended later get their word lazily
    s.add_item(x[0].tolist(), {"photo_path": "/late.jpg", "exif_data": {"datetime": "2030-07-01T10:00:00"},
                               "time_info": O.time_info_from_exif("2030-07-01T10:00:00")})
    hits = s.search(x[0].tolist(), 3, constraints={"year": 2030})
    assert [h["metadata"]["photo_path"] for h in hits] == ["/late.jpg"]


# ---- batch API ---------------------------------------------------------------------------------------
def test_batch_api(VS, tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((300, 32)).astype(np.float32)
    s = VS(32, *_paths(tmp_path))
    s.add_batch(x, [{"photo_path": f"{i}"} for i in range(300)])
    t = VS(32, str(tmp_path / "b"), str(tmp_path / "c"))
    for i in range(300):
        t.add_item(x[i].tolist(), {"photo_path": f"{i}"})
    q = rng.standard_normal((5, 32)).astype(np.float32)
    D, I = s.search_batch(q, 7)
    for qi in range(5):
        single = t.search(q[qi].tolist(), 7)
        assert [int(h["metadata"]["photo_path"]) for h in single] == I[qi].tolist()
        assert np.allclose([h["distance"] for h in single], D[qi], rtol=1e-6, atol=1e-7)
    D, I = s.search_batch(q, 400)
    assert D.shape == (5, 400) and (I[:, 300:] == -1).all() and np.isinf(D[:, 300:]).all()


# ---- the three reference tests that drive a real store with data through Searcher ------------------
# (tests/test_searcher.py:293-321, 323-350, 352-406), restated at the VectorStore level: what Searcher
# does around the store there is get_embedding_by_photo_path -> search(k = max(top_k+1, 5*top_k) clamped
# to N) -> drop the query photo -> dedupe by path (core/searcher.py:1759-1786, 1839-1850).
def _image_search(store, photo_path, top_k):
    emb = store.get_embedding_by_photo_path(photo_path)
    assert emb is not None
    n = store.get_total_items()
    hits = store.search(emb, min(n, max(top_k + 1, top_k * 5)))
    seen, out = set(), []
    for h in hits:
        p = h["metadata"]["photo_path"]
        if p == photo_path or p in seen:
            continue
        seen.add(p)
        out.append(h)
    return out[:top_k]


def test_search_by_image_path_excludes_self(VS, tmp_path):
    s = VS(8, *_paths(tmp_path))
    paths = [f"/photos/photo_{i}.jpg" for i in range(3)]
    for i, p in enumerate(paths):
        s.add_item([float(i + o) for o in range(8)], {"photo_path": p, "description": f"图片 {i}"})
    s.save()
    t = VS(8, s.index_path, s.metadata_path)
    assert t.load()
    res = _image_search(t, paths[0], 2)
    assert len(res) == 2 and all(h["metadata"]["photo_path"] != paths[0] for h in res)
    assert [h["metadata"]["photo_path"] for h in res] == [paths[1], paths[2]]  # nearest first


def test_search_by_image_path_deduplicates_same_photo(VS, tmp_path):
    s = VS(8, *_paths(tmp_path))
    s.add_item([1.0] * 8, {"photo_path": "/q.jpg", "description": "query"})
    s.add_item([0.9] * 8, {"photo_path": "/dup.jpg", "description": "dup-a"})
    s.add_item([0.9] * 8, {"photo_path": "/dup.jpg", "description": "dup-b"})
    s.add_item([0.8] * 8, {"photo_path": "/other.jpg", "description": "other"})
    res = _image_search(s, "/q.jpg", 3)
    got = [h["metadata"]["photo_path"] for h in res]
    assert got.count("/dup.jpg") == 1 and set(got) == {"/dup.jpg", "/other.jpg"}
    # all four rows normalise to the same direction: exact ties come back in insertion order
    full = s.search([1.0] * 8, 4)
    assert [h["metadata"]["description"] for h in full] == ["query", "dup-a", "dup-b", "other"]


def test_search_with_fake_embedding_service_vectors(VS, tmp_path):
    """tests/helpers.py FakeEmbeddingService vectors (seed + i): nearly parallel rows, tiny score gaps."""
    def fake(text, d=8):
        seed = float(sum(ord(c) for c in text) % 13)
        return [seed + float(i) for i in range(d)]

    s = VS(8, *_paths(tmp_path))
    texts = [f"photo 图片 {i}" for i in range(3)]
    for i, tx in enumerate(texts):
        s.add_item([float(i + o) for o in range(8)], {"photo_path": f"/p{i}.jpg", "retrieval_text": tx})
    res = s.search(fake("photo 图片 1 上传图片"), max(2 * 5, 2 + 5))
    assert len(res) == 3 and {h["metadata"]["photo_path"] for h in res} == {"/p0.jpg", "/p1.jpg", "/p2.jpg"}
    want = O.OracleVectorStore(8, str(tmp_path / "o"), str(tmp_path / "om"))
    for i, tx in enumerate(texts):
        want.add_item([float(i + o) for o in range(8)], {"photo_path": f"/p{i}.jpg"})
    ref = want.search(fake("photo 图片 1 上传图片"), 7)
    assert [h["metadata"]["photo_path"] for h in res] == [h["metadata"]["photo_path"] for h in ref]
    assert np.allclose([h["distance"] for h in res], [h["distance"] for h in ref], rtol=1e-5, atol=1e-6)


def test_exif_sidecar_round_trip(VS, tmp_path):
    """<index>.attrs: written by save() once the words exist, reused by load(), ignored when stale."""
    ip, mp = _paths(tmp_path)
    s = VS(4, ip, mp)
    for i in range(6):
        t = f"202{i}-07-0{i + 1}T10:00:00"
        s.add_item([1.0, i, 0, 0], {"photo_path": f"/{i}.jpg", "exif_data": {"datetime": t}, "time_info": O.time_info_from_exif(t)})
    s.save()
    assert not os.path.exists(ip + ".attrs")            # nothing packed yet: no sidecar
    assert len(s.search([1, 0, 0, 0], 6, constraints={"year": 2023})) == 1
    s.save()
    assert os.path.getsize(ip + ".attrs") == 6 * 8
    t2 = VS(4, ip, mp)
    assert t2.load() and t2._attrs_built == 6             # words came from the sidecar
    assert [h["metadata"]["photo_path"] for h in t2.search([1, 0, 0, 0], 6, constraints={"season": "夏天", "start_date": "2022-01-01"})] \
        == [f"/{i}.jpg" for i in (2, 3, 4, 5)]
    with open(ip + ".attrs", "ab") as f:                  # wrong size -> ignored, rebuilt lazily from metadata
        f.write(b"x")
    t3 = VS(4, ip, mp)
    assert t3.load() and t3._attrs_built == 0
    assert len(t3.search([1, 0, 0, 0], 6, constraints={"year": 2023})) == 1


def test_coalesced_concurrent_searches_equal_sequential(fake_backend, tmp_path, monkeypatch):
    """coalesce=True: request threads that arrive while a search is running are served by ONE batched backend
    call; every caller gets exactly what the plain store returns, filters are never mixed in a batch, and a
    backend failure reaches every waiter of that batch."""
    import threading
    import time

    from photo_search_engine_b200.vector_store import VectorStore

    d, n = 16, 300
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    metas = [{"photo_path": f"/p/{i}.jpg", "exif_data": {"datetime": f"20{20 + i % 4}-0{1 + i % 9}-10T10:00:00"},
              "time_info": {"year": 2020 + i % 4, "month": 1 + i % 9, "season": None, "time_period": None,
                            "datetime_str": f"20{20 + i % 4}-0{1 + i % 9}-10T10:00:00"}} for i in range(n)]
    plain = VectorStore(d, str(tmp_path / "a.index"), str(tmp_path / "a.json"))
    shared = VectorStore(d, str(tmp_path / "b.index"), str(tmp_path / "b.json"), coalesce=True)
    for store in (plain, shared):
        store.add_batch(rows, metas)
    calls = []
    real = fake_backend.search

    def slow_search(self, q, k, flt=None):
        calls.append(np.asarray(q).shape[0] if np.asarray(q).ndim == 2 else 1)
        time.sleep(0.02)  # long enough for the other threads to queue up behind the running search
        return real(self, q, k, flt)

    monkeypatch.setattr(fake_backend, "search", slow_search)
    queries = [rng.standard_normal(d).astype(np.float32).tolist() for _ in range(24)]
    cons = [None, {"year": 2021}, None, {"start_date": "2022-01-01", "end_date": "2023-12-31"}]
    want = [plain.search(q, 5 + i % 7, constraints=cons[i % 4]) for i, q in enumerate(queries)]
    calls.clear()
    got = [None] * len(queries)

    def worker(i):
        got[i] = shared.search(queries[i], 5 + i % 7, constraints=cons[i % 4])

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(queries))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(30)
    assert all(not t.is_alive() for t in threads)
    for g, w in zip(got, want):
        assert [h["metadata"]["photo_path"] for h in g] == [h["metadata"]["photo_path"] for h in w]
        assert [h["distance"] for h in g] == [h["distance"] for h in w]
    co = shared._coalescer
    assert co.requests == len(queries) and co.batches == len(calls) < len(queries) and co.largest_batch > 1
    # a failing backend call fails every request of its batch, and the coalescer keeps working afterwards
    def broken(self, q, k, flt=None):
        raise RuntimeError("boom")

    monkeypatch.setattr(fake_backend, "search", broken)
    with pytest.raises(RuntimeError):
        shared.search(queries[0], 3)
    monkeypatch.setattr(fake_backend, "search", real)
    assert len(shared.search(queries[0], 3)) == 3


def test_query_row_has_the_bits_of_the_reference_normalisation(fake_backend, tmp_path):
    """search() converts the query without the list round trip of _normalize_vector (utils/vector_store.py:83-90): the
    float32 bits handed to the backend must be the ones the reference would send, for cosine and l2, zero vectors too."""
    from photo_search_engine_b200.vector_store import VectorStore

    rng = np.random.default_rng(8)
    for metric in ("cosine", "l2"):
        store = VectorStore(33, str(tmp_path / f"{metric}.index"), str(tmp_path / f"{metric}.json"), metric=metric)
        for trial in range(50):
            v = (rng.standard_normal(33) * 10.0 ** rng.integers(-3, 4)).tolist()
            if trial == 0:
                v = [0.0] * 33
            want = np.array([store._normalize_vector(v)], dtype="float32")
            got = store._query_row(v)
            assert got.dtype == np.float32 and got.shape == (1, 33)
            assert got.tobytes() == want.tobytes()
