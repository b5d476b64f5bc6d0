"""Property test of the packed EXIF predicate: for generated metadata records and constraint
dicts, ``attr_words`` + ``build_filter`` + ``words_pass`` (the numpy form of the device predicate ``attr_pass``,
csrc/psx_scan.cuh) decide exactly what the oracle's restatement of ``Searcher._check_time_match_v2``
(core/searcher.py:1884-1950, ``_parse_date`` :1963-2001) decides.

The generators stay inside the documented contract of photo_search_engine_b200/exif_attrs.py: no value that is
unrepresentable on BOTH sides at once (the same unknown season string in a record and in the constraint), no fractions
of a second, no bools where the reference expects integers.
"""
from __future__ import annotations

import functools
from datetime import datetime

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import flat_ip as O
from photo_search_engine_b200.exif_attrs import attr_words, build_filter, words_pass

SEASONS = ["春天", "夏天", "秋天", "冬天"]
PERIODS = ["凌晨", "早晨", "上午", "中午", "下午", "傍晚", "夜晚"]
FORMATS = ["%Y-%m-%dT%H:%M:%S", "%Y-%m-%d %H:%M:%S", "%Y:%m:%d %H:%M:%S", "%Y/%m/%d %H:%M:%S", "%Y-%m-%d", "%Y/%m/%d", "%Y%m%d"]

moments = st.datetimes(min_value=datetime(1, 1, 1), max_value=datetime(9999, 12, 31, 23, 59, 59)).map(lambda t: t.replace(microsecond=0))
# a window of a few years makes start / end bounds actually cut through the records
near = st.datetimes(min_value=datetime(2019, 1, 1), max_value=datetime(2024, 12, 31, 23, 59, 59)).map(lambda t: t.replace(microsecond=0))


def _fmt(t: datetime, fmt: str) -> str:
    # strftime does not zero-pad years below 1000 on every platform: build the year by hand
    return fmt.replace("%Y", f"{t.year:04d}").replace("%m", f"{t.month:02d}").replace("%d", f"{t.day:02d}") \
              .replace("%H", f"{t.hour:02d}").replace("%M", f"{t.minute:02d}").replace("%S", f"{t.second:02d}")


stamps = st.one_of(
    st.builds(_fmt, st.one_of(moments, near, near), st.sampled_from(FORMATS)),
    st.builds(lambda s: " " + s + "\x00", st.builds(_fmt, near, st.sampled_from(FORMATS[:4]))),
    st.sampled_from(["", "garbage", "2023-13-01T00:00:00", "2023-02-30", "0000-01-01T00:00:00", "2023-7-1 4:5:6", "2023-07-01T14:00"]),
    st.none(),
)

record_time_info = st.one_of(
    st.none(),
    st.fixed_dictionaries({}, optional={
        "season": st.one_of(st.none(), st.sampled_from(SEASONS + ["旱季", ""])),
        "time_period": st.one_of(st.none(), st.sampled_from(PERIODS + ["半夜", ""])),
        "year": st.one_of(st.none(), st.integers(0, 9999), st.sampled_from([2023.0, 2023.5, "2023", 20000, -3])),
        "month": st.one_of(st.none(), st.integers(0, 13), st.sampled_from([7.0, "7"])),
        "datetime_str": stamps,
    }),
)
record_exif = st.one_of(st.none(), st.just({}), st.fixed_dictionaries({"datetime": stamps}))
records = st.fixed_dictionaries({}, optional={"time_info": record_time_info, "exif_data": record_exif, "photo_path": st.just("/p/x.jpg")})

constraints = st.fixed_dictionaries({}, optional={
    "season": st.one_of(st.none(), st.sampled_from(SEASONS + ["雨季", ""])),
    "time_period": st.one_of(st.none(), st.sampled_from(PERIODS + ["深夜", ""])),
    "year": st.one_of(st.none(), st.integers(0, 2030), st.sampled_from([2023, 2021, 2020, "2022", 2023.0, 70000])),
    "month": st.one_of(st.none(), st.integers(0, 13)),
    "start_date": st.one_of(st.none(), st.builds(_fmt, near, st.sampled_from(FORMATS)), st.sampled_from(["", "not a date", "0001-01-01", "2023-02-30"])),
    "end_date": st.one_of(st.none(), st.builds(_fmt, near, st.sampled_from(FORMATS)), st.sampled_from(["", "not a date", "9999-12-31", "20231231"])),
    "precision": st.sampled_from(["year", "month", "day"]),
})


@functools.lru_cache(maxsize=1)
def _reference_searcher():
    """The reference's own Searcher (checkout or the staged copy under oracle/_ref), or None."""
    import sys
    import types

    from oracle import stage_reference

    ref = stage_reference.locate()
    if ref is None:
        return None
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
    return Searcher.__new__(Searcher)


@settings(max_examples=400, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(records, min_size=1, max_size=12), constraints)
def test_packed_predicate_equals_reference_rule(metas, cons):
    want = [O.check_time_match_v2(m, cons) for m in metas]
    ref = _reference_searcher()
    if ref is not None:  # the oracle's restatement against the function it restates, on the same generated inputs
        assert want == [ref._check_time_match_v2(m, cons) for m in metas], (metas, cons)
    flt, never = build_filter(cons)
    if flt is None:
        got = [True] * len(metas)
    elif never:
        got = [False] * len(metas)
    else:
        got = words_pass(attr_words(metas), flt).tolist()
    assert got == want, (metas, cons, got, want)


def test_words_are_stable_under_the_formats_the_reference_accepts():
    """One moment written in each accepted date-time format packs to the same word; a date-only string packs to midnight."""
    t = datetime(2023, 7, 1, 14, 5, 6)
    words = attr_words([{"exif_data": {"datetime": _fmt(t, f)}} for f in FORMATS[:4]])
    assert len(set(words.tolist())) == 1
    day = attr_words([{"exif_data": {"datetime": _fmt(t, f)}} for f in FORMATS[4:]])
    midnight = attr_words([{"exif_data": {"datetime": "2023-07-01T00:00:00"}}])[0]
    assert set(day.tolist()) == {int(midnight)}
    assert int(words[0]) - int(midnight) == 14 * 3600 + 5 * 60 + 6
    assert np.uint64(words[0]) >> np.uint64(63) == 1
