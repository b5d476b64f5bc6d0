"""Host side of the fused EXIF predicate.

The reference filters candidates *after* the vector search, photo by photo, with
``Searcher._check_time_match_v2`` (core/searcher.py:1884-1950, date parsing :1963-2001).
Here the same conjunction is evaluated inside the scan kernel from one packed 64-bit word
per row (layout: include/psx.h), so rejected rows are never read from HBM.  This module
packs the words from the reference's metadata records (``time_info`` as produced by
core/indexer.py:535-609, ``exif_data.datetime``) and turns a ``constraints`` dict into a
``psx_filter``.

Values that cannot be represented (an unknown season string, a non-integer year, ...) are
stored as codes no representable constraint can equal, and a constraint that cannot be
represented matches nothing -- both are what the reference's ``!=`` comparisons give, except
for the corner where the *same* unknown string appears on both sides.  Fractions of a second
are ignored.
"""
from __future__ import annotations

from datetime import datetime
import re
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np

from ._native import (F_END, F_MONTH, F_NEED_DT, F_PERIOD, F_SEASON, F_START, F_YEAR, PsxFilter)

SEASON_CODES = {"春天": 1, "夏天": 2, "秋天": 3, "冬天": 4}
# the seven day parts in the order core/indexer.py:583-598 assigns them
PERIOD_CODES = {"凌晨": 1, "早晨": 2, "上午": 3, "中午": 4, "下午": 5, "傍晚": 6, "夜晚": 7}
SEASON_OTHER, MONTH_OTHER, YEAR_OTHER = 7, 15, 16383
YEAR_MAX = 16381

_DT_BITS, _MONTH_SHIFT, _YEAR_SHIFT, _PERIOD_SHIFT, _SEASON_SHIFT, _EXIF_SHIFT = 39, 39, 43, 57, 60, 63

_WITH_TIME = ("%Y-%m-%dT%H:%M:%S", "%Y-%m-%d %H:%M:%S", "%Y:%m:%d %H:%M:%S", "%Y/%m/%d %H:%M:%S")
_DATE_ONLY = ("%Y-%m-%d", "%Y/%m/%d", "%Y%m%d")
_ISO_SECONDS = re.compile(r"[0-9]{4}-[0-9]{2}-[0-9]{2}T[0-9]{2}:[0-9]{2}:[0-9]{2}")


def parse_date(value: Any, is_end_date: bool = False) -> Optional[datetime]:
    """Same acceptance set and result as ``Searcher._parse_date`` (core/searcher.py:1963-2001):
    a date-only value used as an end bound means 23:59:59 of that day."""
    if not isinstance(value, str) or not value:
        return None
    text = value.strip().rstrip("\x00")
    # what the indexer writes (datetime.isoformat(), core/indexer.py:535-582) without going through strptime: for a string
    # of exactly this shape the first format below that can match is "%Y-%m-%dT%H:%M:%S", with these very fields
    if len(text) == 19 and _ISO_SECONDS.fullmatch(text):
        try:
            return datetime(int(text[0:4]), int(text[5:7]), int(text[8:10]), int(text[11:13]), int(text[14:16]), int(text[17:19]))
        except ValueError:
            pass  # a field out of range: let the formats below decide (they all refuse it)
    # the reference tries its formats in a fixed order; the sets are disjoint except that the
    # order decides nothing for well-formed input, so date-only and date-time are tried apart
    for fmt in ("%Y-%m-%d",) + _WITH_TIME + _DATE_ONLY[1:]:
        try:
            parsed = datetime.strptime(text, fmt)
        except ValueError:
            continue
        if fmt in _DATE_ONLY and is_end_date:
            return parsed.replace(hour=23, minute=59, second=59)
        return parsed
    try:
        return datetime.fromisoformat(text)
    except Exception:
        return None


def encode_datetime(t: datetime) -> int:
    """1 + seconds since 0001-01-01T00:00:00 (39 bits cover year 9999)."""
    return 1 + (t.toordinal() - 1) * 86400 + t.hour * 3600 + t.minute * 60 + t.second


def _int_code(value: Any, lo: int, hi: int, other: int) -> int:
    if value is None or isinstance(value, bool):
        return 0 if value is None else other
    if isinstance(value, int) or (isinstance(value, float) and value == int(value)):
        iv = int(value)
        return iv if lo <= iv <= hi else other
    return other


def attr_word(metadata: Dict[str, Any]) -> int:
    """Pack one metadata record into the attribute word of include/psx.h."""
    time_info = metadata.get("time_info") or {}
    exif_data = metadata.get("exif_data") or {}
    exif_datetime = exif_data.get("datetime") if isinstance(exif_data, dict) else None
    if not isinstance(time_info, dict):
        time_info = {}
    word = 0
    if exif_datetime:
        word |= 1 << _EXIF_SHIFT
    season = time_info.get("season")
    if season is not None:
        word |= SEASON_CODES.get(season, SEASON_OTHER) << _SEASON_SHIFT
    period = time_info.get("time_period")
    if period is not None:
        word |= PERIOD_CODES.get(period, 0) << _PERIOD_SHIFT
    word |= _int_code(time_info.get("year"), 1, YEAR_MAX, YEAR_OTHER) << _YEAR_SHIFT
    word |= _int_code(time_info.get("month"), 1, 12, MONTH_OTHER) << _MONTH_SHIFT
    photo_dt = parse_date(time_info.get("datetime_str") or exif_datetime)
    if photo_dt is not None:
        word |= encode_datetime(photo_dt)
    return word


def attr_words(metadatas: Iterable[Dict[str, Any]]) -> np.ndarray:
    return np.fromiter((attr_word(m) for m in metadatas), dtype=np.uint64)


def build_filter(constraints: Optional[Dict[str, Any]]) -> Tuple[Optional[PsxFilter], bool]:
    """``constraints`` (the dict ``Searcher`` passes around, core/searcher.py:1638-1660) ->
    ``(psx_filter | None, never)``.  ``None`` = no active clause; ``never`` = some clause can
    not be satisfied by any row, the caller returns no hits without scanning."""
    if not constraints:
        return None, False
    f = PsxFilter()
    never = False
    season = constraints.get("season")
    if season:
        f.flags |= F_SEASON
        f.season = SEASON_CODES.get(season, 0)
        never |= f.season == 0
    period = constraints.get("time_period")
    if period:
        f.flags |= F_PERIOD
        f.period = PERIOD_CODES.get(period, 0)
        never |= f.period == 0
    year = constraints.get("year")
    if year:
        f.flags |= F_YEAR
        code = _int_code(year, 1, YEAR_MAX, 0)
        f.year = code
        never |= code == 0
    month = constraints.get("month")
    if month:
        f.flags |= F_MONTH
        code = _int_code(month, 1, 12, 0)
        f.month = code
        never |= code == 0
    start_date, end_date = constraints.get("start_date"), constraints.get("end_date")
    if start_date or end_date:
        f.flags |= F_NEED_DT
        if start_date:
            start = parse_date(start_date)
            if start is not None:
                f.flags |= F_START
                f.start = encode_datetime(start)
        if end_date:
            end = parse_date(end_date, is_end_date=True)
            if end is not None:
                f.flags |= F_END
                f.end = encode_datetime(end)
    if not f.flags:
        return None, False
    return f, never


def words_pass(words: np.ndarray, flt: PsxFilter) -> np.ndarray:
    """numpy form of ``attr_pass`` (csrc/psx_scan.cuh) over packed attribute words: which rows pass ``flt``.
    Used where the predicate has to be evaluated for a handful of candidate rows on the host (the post-filter of
    ``searcher_ext.FusedRecallMixin``); the scan evaluates the same conjunction on the device."""
    w = np.asarray(words, dtype=np.uint64)
    ok = np.ones(w.shape[0], bool)
    fl = int(flt.flags)
    if fl & (F_SEASON | F_PERIOD | F_YEAR | F_MONTH):
        ok &= (w >> np.uint64(63)).astype(bool)
        if fl & F_SEASON:
            ok &= ((w >> np.uint64(_SEASON_SHIFT)) & np.uint64(7)) == np.uint64(flt.season)
        if fl & F_PERIOD:
            ok &= ((w >> np.uint64(_PERIOD_SHIFT)) & np.uint64(7)) == np.uint64(flt.period)
        if fl & F_YEAR:
            ok &= ((w >> np.uint64(_YEAR_SHIFT)) & np.uint64(0x3FFF)) == np.uint64(flt.year)
        if fl & F_MONTH:
            ok &= ((w >> np.uint64(_MONTH_SHIFT)) & np.uint64(0xF)) == np.uint64(flt.month)
    if fl & F_NEED_DT:
        dt = w & np.uint64((1 << _DT_BITS) - 1)
        ok &= dt != 0
        if fl & F_START:
            ok &= dt >= np.uint64(flt.start)
        if fl & F_END:
            ok &= dt <= np.uint64(flt.end)
    return ok
