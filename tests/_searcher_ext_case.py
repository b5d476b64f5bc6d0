"""Subprocess body of tests/test_searcher_ext.py: the reference's UNMODIFIED Searcher with and without
``BatchedExpansionMixin`` over the same drop-in store; prints one JSON object.  Run with the reference
checkout first on PYTHONPATH (its ``tests``/``core``/``utils`` packages must win)."""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

import ref_inject_plugin  # noqa: F401  (installs utils.vector_store -> the drop-in class; fake or gpu backend)
from core.searcher import Searcher  # the reference, unmodified
from tests.helpers import FakeQueryFormatter, FakeTimeParser  # the reference's own fakes

from photo_search_engine_b200.searcher_ext import BatchedExpansionMixin, FusedPrefilterMixin, FusedRecallMixin
from photo_search_engine_b200.vector_store import VectorStore

D = 16


class CountingEmbedding:
    """Deterministic text -> vector; counts single and batch calls."""

    def __init__(self):
        self.single = 0
        self.batch = 0

    def _vec(self, text):
        rng = np.random.default_rng(sum(ord(c) * (i + 1) for i, c in enumerate(text)) % (2 ** 32))
        return rng.standard_normal(D).astype(np.float32).tolist()

    def generate_embedding(self, text):
        self.single += 1
        return self._vec(text)

    def generate_embedding_batch(self, texts):
        self.batch += 1
        return [self._vec(t) for t in texts]


class CountingStore(VectorStore):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.n_search = 0
        self.n_batch = 0

    def search(self, *a, **k):
        self.n_search += 1
        return super().search(*a, **k)

    def search_batch(self, *a, **k):
        self.n_batch += 1
        return super().search_batch(*a, **k)


class BatchedSearcher(BatchedExpansionMixin, Searcher):
    pass


class PrefilterSearcher(FusedPrefilterMixin, BatchedExpansionMixin, Searcher):
    pass


class RecallSearcher(FusedRecallMixin, Searcher):
    pass


def recall_case(tmp):
    """FusedRecallMixin: candidates stay arrays through _vector_results_to_combined / _finalize_results; every round
    (plain, time-filtered, relaxed, tiny top_k, duplicates by path, rows without a path) must return exactly what the
    unmodified Searcher returns, including _last_round_quality, and materialise far fewer dicts."""
    import time as _time

    rng = np.random.default_rng(21)
    n = 3000
    rows = rng.standard_normal((n, D)).astype(np.float32)
    rows[1500] = rows[7]                                  # same vector under another path
    metas = []
    for i in range(n):
        stamp = f"20{10 + i % 15:02d}-{1 + i % 12:02d}-{1 + i % 27:02d}T{i % 24:02d}:30:00"
        has_exif = i % 5 != 0
        metas.append({
            "photo_path": f"/photos/album{i % 31}/IMG_{i}.JPG" if i % 97 != 3 else ("" if i % 2 else None),
            "description": f"photo {i}", "retrieval_text": f"text {i}",
            "exif_data": {"datetime": stamp} if has_exif else {},
            "time_info": ({"year": 2010 + i % 15, "month": 1 + i % 12, "season": ["春天", "夏天", "秋天", "冬天"][i % 4],
                           "time_period": ["凌晨", "早晨", "上午", "中午", "下午", "傍晚", "夜晚"][i % 7], "datetime_str": stamp}
                          if has_exif else {}),
        })
    metas[40]["photo_path"] = metas[41]["photo_path"].lower()   # normcase-equal on Windows only; distinct here
    metas[60]["photo_path"] = metas[61]["photo_path"]           # an exact duplicate path
    rounds = [
        dict(constraints={}, has_filter=False, normalized_top_k=10, relaxation_level=0),
        dict(constraints={}, has_filter=False, normalized_top_k=50, relaxation_level=2),
        dict(constraints={"start_date": "2015-01-01", "end_date": "2018-12-31", "precision": "year"}, has_filter=True,
             normalized_top_k=10, relaxation_level=0),
        dict(constraints={"season": "夏天", "time_period": "下午"}, has_filter=True, normalized_top_k=5, relaxation_level=1),
        dict(constraints={"year": 2013, "month": 4}, has_filter=True, normalized_top_k=12, relaxation_level=3),
        dict(constraints={"season": "雨季"}, has_filter=True, normalized_top_k=5, relaxation_level=0),   # not representable
        dict(constraints={"start_date": "not a date"}, has_filter=True, normalized_top_k=5, relaxation_level=0),
        dict(constraints={}, has_filter=False, normalized_top_k=1, relaxation_level=0),
    ]
    out = {"rounds": []}
    timing = {}
    for tag, cls in (("plain", Searcher), ("recall", RecallSearcher)):
        store = CountingStore(D, os.path.join(tmp, tag + "_rc.index"), os.path.join(tmp, tag + "_rc.json"))
        store.add_batch(rows, metas)
        qs = np.random.default_rng(5).standard_normal((len(rounds) + 4, D)).astype(np.float32)
        qs[0] = rows[7]
        emb = CountingEmbedding()
        s = cls(embedding=emb, time_parser=FakeTimeParser(), vector_store=store, keyword_store=None, query_formatter=None)
        s.index_loaded = True
        res = []
        t0 = _time.perf_counter()
        for rep in range(3):
            for qi, kw in enumerate(rounds):
                emb._vec = lambda text, _q=qs[qi]: _q.tolist()
                r = s._run_single_search_round(query="q", intent={"search_text": "q"}, embedding_query="q", media_terms=[],
                                               identity_terms=[], strict_identity_filter=False, **kw)
                if rep == 0:
                    res.append({"results": [[x.get("photo_path"), x.get("score"), x.get("rank"), x.get("_confidence_bucket"),
                                             x.get("_relaxation_level"), x.get("description"), x.get("retrieval_text"),
                                             x.get("match_summary"), x.get("metadata", {}).get("description")] for x in r],
                                "quality": s._get_last_round_quality()})
        timing[tag] = (_time.perf_counter() - t0) / (3 * len(rounds)) * 1e3
        # a round with media terms takes the reference path under the mixin too
        emb._vec = lambda text, _q=qs[-1]: _q.tolist()
        r = s._run_single_search_round(query="q", intent={"search_text": "q"}, embedding_query="q", media_terms=["photo"],
                                       identity_terms=[], strict_identity_filter=False, constraints={}, has_filter=False, normalized_top_k=10)
        res.append({"results": [[x.get("photo_path"), x.get("score"), x.get("rank"), x.get("_confidence_bucket")] for x in r],
                    "quality": s._get_last_round_quality()})
        out[tag] = res
        out[tag + "_stats"] = getattr(s, "psx_recall_stats", None)
    out["ms_per_round"] = timing
    out["vector_scores"] = s._psx_vector_scores_ok  # the array form of _distance_to_score passed its probe

    # a Searcher whose _distance_to_score is not the reference's: the probe fails, the method keeps being called, results agree
    class OddScore(Searcher):
        def _distance_to_score(self, distance):
            return round(max(0.0, min(1.0, 0.5 + 0.25 * distance)), 6)

    class OddRecall(FusedRecallMixin, OddScore):
        pass

    odd = {}
    for tag, cls in (("plain", OddScore), ("recall", OddRecall)):
        store = CountingStore(D, os.path.join(tmp, tag + "_odd.index"), os.path.join(tmp, tag + "_odd.json"))
        store.add_batch(rows, metas)
        emb = CountingEmbedding()
        emb._vec = lambda text, _q=qs[1]: _q.tolist()
        s = cls(embedding=emb, time_parser=FakeTimeParser(), vector_store=store, keyword_store=None, query_formatter=None)
        s.index_loaded = True
        r = s._run_single_search_round(query="q", intent={"search_text": "q"}, embedding_query="q", media_terms=[], identity_terms=[],
                                       strict_identity_filter=False, constraints={}, has_filter=False, normalized_top_k=10)
        odd[tag] = [[x.get("photo_path"), x.get("score"), x.get("rank"), x.get("_confidence_bucket")] for x in r]
        odd[tag + "_vector_scores"] = getattr(s, "_psx_vector_scores_ok", None)
    out["odd"] = odd
    return out


def prefilter_case(tmp):
    """A year filter that only 1 photo in 12 passes: the reference recalls candidate_k hits and post-filters them,
    the fused pre-filter recalls the best *passing* rows."""
    out = {}
    rng = np.random.default_rng(11)
    rows = rng.standard_normal((600, D)).astype(np.float32)
    metas = []
    for i in range(600):
        year = 2019 if i % 12 == 0 else 2021
        stamp = f"{year}-06-15T12:00:00"
        metas.append({"photo_path": f"/photos/{i}.jpg", "description": f"photo {i}", "exif_data": {"datetime": stamp},
                      "time_info": {"year": year, "month": 6, "season": "夏天", "time_period": "中午", "datetime_str": stamp}})
    constraints = {"start_date": "2019-01-01", "end_date": "2019-12-31", "precision": "year"}
    qvec = rng.standard_normal(D).astype(np.float32).tolist()
    for tag, cls in (("plain", Searcher), ("prefilter", PrefilterSearcher)):
        store = CountingStore(D, os.path.join(tmp, tag + "_pf.index"), os.path.join(tmp, tag + "_pf.json"))
        for r, m in zip(rows, metas):
            store.add_item(r.tolist(), m)
        emb = CountingEmbedding()
        emb._vec = lambda text: list(qvec)
        s = cls(embedding=emb, time_parser=FakeTimeParser(), vector_store=store, keyword_store=None, query_formatter=None)
        s.index_loaded = True
        res = s._run_single_search_round(query="q", intent={"search_text": "q"}, embedding_query="q", media_terms=[],
                                         identity_terms=[], strict_identity_filter=False, constraints=constraints,
                                         normalized_top_k=10, has_filter=True)
        out[tag] = [[r["photo_path"], r.get("score")] for r in res]
    # ground truth: the 10 best rows of year 2019 by cosine
    unit = rows / np.linalg.norm(rows, axis=1, keepdims=True)
    qn = np.asarray(qvec, np.float32) / np.linalg.norm(qvec)
    s_all = unit @ qn
    passing = [i for i in range(600) if i % 12 == 0]
    best = sorted(passing, key=lambda i: -s_all[i])[:10]
    out["truth"] = [f"/photos/{i}.jpg" for i in best]
    return out


class FakeKeywordStore:
    """What Searcher needs of utils/keyword_store.py: BM25-like hits as [{"photo_path", "score"}], scores in (0, 1]."""

    def __init__(self, metas):
        self.paths = [m["photo_path"] for m in metas if m.get("photo_path")]
        self.by_year = {}
        for m in metas:
            year = (m.get("time_info") or {}).get("year")
            if m.get("photo_path"):
                self.by_year.setdefault(year, []).append(m["photo_path"])

    def _hits(self, pool, query, k):
        rng = np.random.default_rng(abs(hash(("kw", len(pool), k))) % (2 ** 32) if False else len(query) * 7919 + k)
        picks = rng.permutation(len(pool))[: min(k, len(pool))]
        scores = np.sort(rng.random(len(picks)))[::-1]
        if len(scores):
            scores = scores / scores[0]
        return [{"photo_path": pool[int(i)], "score": float(round(s, 4))} for i, s in zip(picks, scores)] + \
               [{"photo_path": "/gone/stale.jpg", "score": 0.9}]  # a stale ES document: no metadata in the local index

    def search(self, query, k):
        return self._hits(self.paths, query, k)

    def search_with_filters(self, query, filters, k):
        pool = self.by_year.get(filters.get("year"), self.paths) if isinstance(filters, dict) else self.paths
        return self._hits(pool, query, k)


def hybrid_case(tmp):
    """FusedRecallMixin on the Elasticsearch branch (_hybrid_search + _finalize_results): identical rounds, fewer dicts."""
    rng = np.random.default_rng(33)
    n = 2500
    rows = rng.standard_normal((n, D)).astype(np.float32)
    metas = [{"photo_path": f"/photos/y{2010 + i % 6}/IMG_{i}.jpg" if i % 53 else f"/photos/y{2010 + (i + 1) % 6}/IMG_{i + 1}.jpg",
              "description": f"photo {i}", "exif_data": {"datetime": f"{2010 + i % 6}-03-0{1 + i % 9}T10:00:00"},
              "time_info": {"year": 2010 + i % 6, "month": 3, "season": "春天", "time_period": "上午",
                            "datetime_str": f"{2010 + i % 6}-03-0{1 + i % 9}T10:00:00"}} for i in range(n)]
    rounds = [dict(constraints={}, has_filter=False, normalized_top_k=10, relaxation_level=0),
              dict(constraints={"year": 2013}, has_filter=True, normalized_top_k=10, relaxation_level=1),
              dict(constraints={}, has_filter=False, normalized_top_k=40, relaxation_level=2),
              dict(constraints={"year": 2011}, has_filter=True, normalized_top_k=3, relaxation_level=0)]
    out = {}
    for tag, cls in (("plain", Searcher), ("recall", RecallSearcher)):
        store = CountingStore(D, os.path.join(tmp, tag + "_hy.index"), os.path.join(tmp, tag + "_hy.json"))
        store.add_batch(rows, metas)
        emb = CountingEmbedding()
        s = cls(embedding=emb, time_parser=FakeTimeParser(), vector_store=store, keyword_store=FakeKeywordStore(metas), query_formatter=None)
        s.index_loaded = True
        qs = np.random.default_rng(6).standard_normal((len(rounds), D)).astype(np.float32)
        res = []
        for qi, kw in enumerate(rounds):
            emb._vec = lambda text, _q=qs[qi]: _q.tolist()
            r = s._run_single_search_round(query="海边 日落 photo", intent={"search_text": "q"}, embedding_query="q", media_terms=[],
                                           identity_terms=[], strict_identity_filter=False, **kw)
            res.append({"results": [[x.get("photo_path"), x.get("score"), x.get("vector_score"), x.get("keyword_score"), x.get("rank"),
                                     x.get("_confidence_bucket"), x.get("match_summary"), x.get("description")] for x in r],
                        "quality": s._get_last_round_quality()})
        out[tag] = res
        out[tag + "_stats"] = getattr(s, "psx_recall_stats", None)
    return out


def alt(text, terms=()):
    return {"search_text": text, "media_terms": list(terms), "identity_terms": [], "strict_identity_filter": False,
            "intent_mode": "open", "time_hint": None, "season": None, "time_period": None, "original_query": QUERY,
            "reason": "test"}


QUERY = "海边 日落"
ALTS = [alt("海滩 黄昏 天空"), alt("沙滩 夕阳", ["landscape"]), alt("ocean sunset photo"), alt("海滩 黄昏 天空")]


def build(cls, tmp, tag):
    store = CountingStore(D, os.path.join(tmp, tag + ".index"), os.path.join(tmp, tag + ".json"))
    rng = np.random.default_rng(5)
    for i in range(400):
        store.add_item(rng.standard_normal(D).astype(np.float32).tolist(),
                       {"photo_path": f"/photos/{i}.jpg", "description": f"photo {i}", "exif_data": {}, "time_info": {}})
    fmt = FakeQueryFormatter()
    fmt.expansion_mapping[QUERY] = ALTS
    emb = CountingEmbedding()
    s = cls(embedding=emb, time_parser=FakeTimeParser(), vector_store=store, keyword_store=None, query_formatter=fmt,
            query_expansion_max_alternatives=4, query_multi_round_enabled=True, query_reflection_enabled=False,
            embedding_cache_enabled=False)
    s.index_loaded = True
    return s, store, emb


def slim(results):
    return [[r.get("photo_path"), r.get("score"), r.get("rank"), r.get("vector_score")] for r in results]


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, cls in (("plain", Searcher), ("batched", BatchedSearcher)):
            s, store, emb = build(cls, tmp, tag)
            debug = s._empty_search_debug()
            base_intent = s.query_formatter.format_query(QUERY)
            direct = s._maybe_expand_query_results(query=QUERY, base_intent=base_intent, base_results=[], base_round_quality=None,
                                                   normalized_top_k=10, constraints={}, has_filter=False, debug=debug)
            after_direct = (store.n_search, store.n_batch, emb.single, emb.batch)
            full = s.search(QUERY, top_k=10, search_mode="high_recall")
            out[tag] = {"direct": slim(direct), "full": slim(full), "alternatives_run": len(debug["alternatives"]),
                        "after_direct": after_direct, "total": (store.n_search, store.n_batch, emb.single, emb.batch),
                        "stats": getattr(s, "psx_batch_stats", None),
                        "expansion_triggered_full": bool(s._last_search_debug.get("expansion_triggered"))}
        out["prefilter_case"] = prefilter_case(tmp)
        out["recall_case"] = recall_case(tmp)
        out["hybrid_case"] = hybrid_case(tmp)
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
