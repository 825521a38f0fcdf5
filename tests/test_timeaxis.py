"""Product group-bounds builder (integer calendar arithmetic) vs the reference's pandas route
(restated in the oracle) and vs the fixtures made from the reference itself."""
import numpy as np
import pandas as pd
import pytest

from aggfly_b200.timeaxis import CalendarIndex, group_bounds, translate_groupby
from oracle import oracle as orc


def _same(a, b):
    ba, la = a
    bb, lb = b
    assert np.array_equal(ba, bb)
    assert np.array_equal(pd.DatetimeIndex(la).values, pd.DatetimeIndex(lb).values)


@pytest.mark.parametrize("tag", ["full", "gap", "daily"])
def test_bounds_match_reference_fixtures(golden, tag):
    t = pd.DatetimeIndex(golden[f"time_{tag}"].astype("datetime64[ns]"))
    for freq in ("1D", "ME", "YE", "W"):
        b, lab = group_bounds(t, freq)
        assert np.array_equal(b, golden[f"bounds_{tag}_{freq}"])
        assert np.array_equal(lab.values.astype("datetime64[ns]").astype(np.int64), golden[f"labels_{tag}_{freq}"])
    lab1 = group_bounds(t, "1D")[1]
    for f2 in ("ME", "YE", "W"):
        b2, lab2 = group_bounds(lab1, f2)
        assert np.array_equal(b2, golden[f"bounds2_{tag}_{f2}"])
        assert np.array_equal(lab2.values.astype("datetime64[ns]").astype(np.int64), golden[f"labels2_{tag}_{f2}"])


@pytest.mark.parametrize("start,periods,freq", [
    ("2001-01-01", 8760, "h"), ("1999-12-31 23:00", 24 * 800, "h"), ("2000-02-28 13:00", 100, "6h"),
    ("1969-12-28", 40, "D"), ("1900-01-01", 3000, "D"), ("2023-12-31 12:00", 5, "12h"),
    ("2000-07-01", 4, "12h"), ("2024-02-29", 1, "D")])
def test_bounds_match_pandas(start, periods, freq):
    t = pd.date_range(start, periods=periods, freq=freq)
    for f in ("1D", "ME", "YE", "W"):
        _same(group_bounds(t, f), orc.resample_groups(t, f))
    # ragged axis with holes
    rng = np.random.default_rng(5)
    keep = np.sort(rng.choice(periods, size=max(1, periods // 3), replace=False))
    for f in ("1D", "ME", "YE", "W"):
        _same(group_bounds(t[keep], f), orc.resample_groups(t[keep], f))


def test_non_monotonic_raises():
    t = pd.DatetimeIndex(["2000-01-02", "2000-01-01"])
    with pytest.raises(ValueError, match="monotonic"):
        group_bounds(t, "1D")


def test_calendar_bounds_reference_expectations():
    """aggfly/tests/test_aggregate.py:454-466."""
    t360 = CalendarIndex.range("360_day", 2000, 720)
    b_m, lab_m = group_bounds(t360, "ME")
    assert set(np.diff(b_m).tolist()) == {30} and len(lab_m) == 24 and isinstance(lab_m, CalendarIndex)
    assert group_bounds(t360, "YE")[0].tolist() == [0, 360, 720]
    b_nl, _ = group_bounds(CalendarIndex.range("noleap", 2000, 365), "ME")
    assert np.diff(b_nl)[:3].tolist() == [31, 28, 31]
    with pytest.raises(NotImplementedError, match="week"):
        group_bounds(t360, "W")


@pytest.mark.parametrize("cal", ["noleap", "360_day"])
def test_calendar_bounds_match_oracle(cal):
    n = 800
    t = CalendarIndex.range(cal, 1999, n)
    o = orc.cal_range(cal, 1999, n)
    assert np.array_equal(t.year, o.year) and np.array_equal(t.month, o.month) and np.array_equal(t.day, o.day)
    keep = np.ones(n, bool); keep[40:75] = False
    for idx in (slice(None), keep):
        tt = t[idx]
        oo = orc.CalTime(cal, o.year[idx], o.month[idx], o.day[idx])
        for f in ("1D", "ME", "YE"):
            b, lab = group_bounds(tt, f)
            bo, labo = orc.resample_groups(oo, f)
            assert np.array_equal(b, bo)
            assert np.array_equal(lab.year, labo.year) and np.array_equal(lab.month, labo.month)
            assert np.array_equal(lab.day, labo.day)
    hourly = CalendarIndex.range(cal, 2001, 24 * 400, freq="h")
    b, lab = group_bounds(hourly, "1D")
    assert set(np.diff(b).tolist()) == {24} and len(lab) == 400


def test_translate_groupby():
    assert [translate_groupby(g) for g in ("date", "month", "year", "week")] == ["1D", "ME", "YE", "W"]
    with pytest.raises(KeyError):
        translate_groupby("decade")
