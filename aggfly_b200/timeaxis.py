"""Time axes and resample-group bounds for the CUDA engine.

``group_bounds(tindex, freq)`` returns exactly what the reference's ``resample_groups`` returns
(aggfly/aggregate/nb_kernels.py:80-115): contiguous group bounds over the (monotonic) time axis,
one group per resample bin from the first to the last stamp, **empty interior bins kept as
zero-width groups**, plus the bin labels (``1D`` -> day start, ``ME`` -> month-end date, ``YE`` ->
Dec 31, ``W`` -> week-ending Sunday).  The reference gets them from
``pd.Series(1, index=t).resample(freq).count()``; here they are computed with integer calendar
arithmetic on ``datetime64`` so a 40-year hourly axis costs microseconds, and the pandas route
stays in the oracle as the checker (tests/test_timeaxis.py).

Non-standard CF calendars (``noleap`` / ``360_day``; xarray.CFTimeIndex in the reference,
nb_kernels.py:100-110) are represented by :class:`CalendarIndex`, since neither xarray nor cftime
is a dependency of this package.
"""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import pandas as pd

FREQ_OF_GROUPBY = {"date": "1D", "month": "ME", "year": "YE", "week": "W"}   # temporal.py:441-456

_NOLEAP = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=np.int64)


def translate_groupby(groupby: str) -> str:
    """Same mapping (and the same KeyError on unknown names) as aggfly/aggregate/temporal.py:441."""
    return FREQ_OF_GROUPBY[groupby]


class CFDate:
    """One stamp of a non-standard calendar (stand-in for a cftime.datetime in panel output)."""
    __slots__ = ("calendar", "year", "month", "day", "hour")

    def __init__(self, calendar, year, month, day, hour=0):
        self.calendar, self.year, self.month, self.day, self.hour = calendar, int(year), int(month), int(day), int(hour)

    def _key(self):
        return (self.year, self.month, self.day, self.hour)

    def __eq__(self, o):
        return isinstance(o, CFDate) and self.calendar == o.calendar and self._key() == o._key()

    def __lt__(self, o):
        return self._key() < o._key()

    def __hash__(self):
        return hash((self.calendar,) + self._key())

    def __repr__(self):
        return f"{self.year:04d}-{self.month:02d}-{self.day:02d} {self.hour:02d}:00:00 ({self.calendar})"

    def isoformat(self):
        return f"{self.year:04d}-{self.month:02d}-{self.day:02d}T{self.hour:02d}:00:00"


class CalendarIndex:
    """A monotonic time axis on a ``noleap`` or ``360_day`` calendar (per-step y/m/d/h arrays)."""

    def __init__(self, calendar: str, year, month, day, hour=None):
        if calendar in ("365_day",):
            calendar = "noleap"
        if calendar not in ("noleap", "360_day"):
            raise ValueError(f"unsupported calendar {calendar!r} (noleap / 360_day)")
        self.calendar = calendar
        self.year = np.asarray(year, dtype=np.int64)
        self.month = np.asarray(month, dtype=np.int64)
        self.day = np.asarray(day, dtype=np.int64)
        self.hour = np.zeros_like(self.year) if hour is None else np.asarray(hour, dtype=np.int64)

    def __len__(self):
        return len(self.year)

    @property
    def month_lengths(self) -> np.ndarray:
        return _NOLEAP if self.calendar == "noleap" else np.full(12, 30, dtype=np.int64)

    @property
    def year_length(self) -> int:
        return int(self.month_lengths.sum())

    def ordinal_hours(self) -> np.ndarray:
        cum = np.concatenate([[0], np.cumsum(self.month_lengths)])
        days = self.year * self.year_length + cum[self.month - 1] + (self.day - 1)
        return days * 24 + self.hour

    @property
    def is_monotonic_increasing(self) -> bool:
        return bool(np.all(np.diff(self.ordinal_hours()) >= 0))

    def to_objects(self) -> np.ndarray:
        out = np.empty(len(self), dtype=object)
        for i in range(len(self)):
            out[i] = CFDate(self.calendar, self.year[i], self.month[i], self.day[i], self.hour[i])
        return out

    def __getitem__(self, sl):
        return CalendarIndex(self.calendar, self.year[sl], self.month[sl], self.day[sl], self.hour[sl])

    @classmethod
    def range(cls, calendar: str, start_year: int, periods: int, freq: str = "D") -> "CalendarIndex":
        """``periods`` daily ("D") or hourly ("h") stamps starting Jan 1 of ``start_year``."""
        tmp = cls(calendar, [start_year], [1], [1])
        md, ylen = tmp.month_lengths, tmp.year_length
        per_day = {"D": 1, "h": 24}[freq]
        n = np.arange(periods, dtype=np.int64)
        days = n // per_day
        hour = (n % per_day) * (24 // per_day) if per_day > 1 else np.zeros_like(n)
        year = start_year + days // ylen
        doy = days % ylen
        cum = np.concatenate([[0], np.cumsum(md)])
        month = np.searchsorted(cum, doy, side="right")
        day = doy - cum[month - 1] + 1
        return cls(calendar, year, month, day, hour)


TimeIndex = Union[pd.DatetimeIndex, CalendarIndex]

_EPOCH_WEEKDAY_SHIFT = 3       # 1970-01-01 is a Thursday; (days + 3) // 7 counts Monday-based weeks


def _bounds_from_ids(ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    first, last = int(ids[0]), int(ids[-1])
    counts = np.bincount(ids - first, minlength=last - first + 1)
    bounds = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=bounds[1:])
    return bounds, np.arange(first, last + 1, dtype=np.int64)


def group_bounds(tindex: TimeIndex, freq: str):
    """(bounds int64[G+1], labels) for ``freq`` in {"1D", "ME", "YE", "W"}."""
    if freq not in ("1D", "ME", "YE", "W"):
        raise ValueError(f"unsupported resample frequency {freq!r}")
    if len(tindex) == 0:
        raise ValueError("empty time axis")
    if not tindex.is_monotonic_increasing:
        raise ValueError("numba engine requires a monotonic-increasing time index "
                         "(xarray's resample path enforces the same).")
    if isinstance(tindex, CalendarIndex):
        return _calendar_bounds(tindex, freq)
    t = pd.DatetimeIndex(tindex).values.astype("datetime64[ns]")
    if np.isnat(t).any():
        raise ValueError("time axis contains NaT")
    days = t.astype("datetime64[D]").astype(np.int64)
    if freq == "1D":
        bounds, bins = _bounds_from_ids(days)
        labels = bins.astype("datetime64[D]")
    elif freq == "W":
        bounds, bins = _bounds_from_ids((days + _EPOCH_WEEKDAY_SHIFT) // 7)
        labels = (bins * 7 - _EPOCH_WEEKDAY_SHIFT + 6).astype("datetime64[D]")          # the Sunday
    elif freq == "ME":
        bounds, bins = _bounds_from_ids(t.astype("datetime64[M]").astype(np.int64))
        labels = ((bins + 1).astype("datetime64[M]").astype("datetime64[D]") - np.timedelta64(1, "D"))
    else:
        bounds, bins = _bounds_from_ids(t.astype("datetime64[Y]").astype(np.int64))
        labels = ((bins + 1).astype("datetime64[Y]").astype("datetime64[D]") - np.timedelta64(1, "D"))
    return bounds, pd.DatetimeIndex(labels.astype("datetime64[ns]"))


def _calendar_bounds(t: CalendarIndex, freq: str):
    if freq == "W":
        # aggfly/aggregate/temporal.py:221-227
        raise NotImplementedError(
            "groupby='week' is not supported on non-standard CF calendars "
            "(noleap/360_day/etc.): xarray/cftime has no weekly offset. Use "
            "'date', 'month', or 'year', or convert to a standard calendar first "
            "with DataArray.convert_calendar('standard').")
    md, ylen = t.month_lengths, t.year_length
    cum = np.concatenate([[0], np.cumsum(md)])
    if freq == "1D":
        bounds, bins = _bounds_from_ids(t.year * ylen + cum[t.month - 1] + (t.day - 1))
        yy, doy = bins // ylen, bins % ylen
        mm = np.searchsorted(cum, doy, side="right")
        labels = CalendarIndex(t.calendar, yy, mm, doy - cum[mm - 1] + 1)
    elif freq == "ME":
        bounds, bins = _bounds_from_ids(t.year * 12 + (t.month - 1))
        mm = bins % 12 + 1
        labels = CalendarIndex(t.calendar, bins // 12, mm, md[mm - 1])
    else:
        bounds, bins = _bounds_from_ids(t.year)
        labels = CalendarIndex(t.calendar, bins, np.full_like(bins, 12), np.full_like(bins, md[11]))
    return bounds, labels


def labels_equal(a, b) -> bool:
    if isinstance(a, CalendarIndex) != isinstance(b, CalendarIndex) or len(a) != len(b):
        return False
    if isinstance(a, CalendarIndex):
        return a.calendar == b.calendar and bool(np.array_equal(a.ordinal_hours(), b.ordinal_hours()))
    return bool((pd.DatetimeIndex(a) == pd.DatetimeIndex(b)).all())


def label_values(labels) -> np.ndarray:
    """Values for the panel's ``time`` column."""
    if isinstance(labels, CalendarIndex):
        return labels.to_objects()
    return pd.DatetimeIndex(labels).values
