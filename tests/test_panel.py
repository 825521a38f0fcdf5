"""Panel assembly (SURVEY f-4): ``aggregate._panel_frame`` -- row selection, gather and transposition with torch ops on
whatever device the panel lives on, columns handed to pandas without copies -- against the literal route it replaces,
``_assemble_panel`` (aggfly/aggregate/spatial.py:136-153) followed by the region join of aggfly/aggregate/aggregate.py:276-280.
Runs on CPU tensors here; the GPU tests exercise the same code on the device through every ``aggregate_dataset`` call."""
from types import SimpleNamespace as NS

import numpy as np
import pandas as pd
import pytest
import torch

import aggfly_b200 as af
from aggfly_b200.aggregate import _assemble_panel, _panel_frame


def _literal(panel, names, labels, rids, w):
    df = _assemble_panel(panel, names, labels, rids, w)
    rid = w.georegions.regionid
    return w.georegions.shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")


def _weights(rids, shp_ids, shp_vals, zero_weight, zero_frac, rng):
    shp = pd.DataFrame({"geoid": shp_vals}, index=shp_ids)
    zero = rng.random(len(rids)) < zero_frac
    frame = pd.DataFrame({"index_right": np.repeat(rids, 2), "weight": np.where(np.repeat(zero, 2), 0.0, 1.0)})
    return NS(georegions=NS(shp=shp, regionid="geoid"), zero_weight=zero_weight, weights=frame)


@pytest.mark.parametrize("zero_weight", ["nan", "area", "drop"])
@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("extra", [0, 3])
@pytest.mark.parametrize("string_ids", [True, False])
def test_panel_frame_equals_the_literal_route(zero_weight, shuffle, extra, string_ids):
    rng = np.random.default_rng(len(zero_weight) * 100 + shuffle * 10 + extra + 50 * string_ids)
    R, G, NC = 40, 7, 3
    panel = rng.random((R, G, NC))
    panel[rng.random(R) < 0.1] = np.nan                                  # regions without a valid cell
    panel[rng.random((R, G)) < 0.03, rng.integers(0, NC)] = np.nan       # single NaN entries drop their row
    rids = np.sort(rng.choice(np.arange(5 * R), R, replace=False))
    shp_ids = np.concatenate([rids[: R - R // 7], rng.choice(np.arange(5 * R, 6 * R), extra, replace=False)])   # both sides miss some
    if shuffle:
        shp_ids = rng.permutation(shp_ids)                                # the join follows the region frame's order
    vals = [f"g{v}" for v in shp_ids] if string_ids else (shp_ids * 10).astype(np.int64)
    w = _weights(rids, shp_ids, vals, zero_weight, 0.2, rng)
    names, labels = [f"c{i}" for i in range(NC)], pd.date_range("2001-01-01", periods=G, freq="D")
    want = _literal(panel, names, labels, rids, w)
    got = _panel_frame(torch.from_numpy(panel), names, labels, rids, w)
    pd.testing.assert_frame_equal(got, want)
    assert got.index.equals(want.index) and list(got.columns) == ["geoid", "time"] + names


def test_panel_frame_edge_cases():
    rng = np.random.default_rng(3)
    R, G, NC = 20, 4, 3
    rids = np.arange(R) * 3 + 10
    names, labels = [f"c{i}" for i in range(NC)], pd.date_range("2001-01-01", periods=G, freq="D")
    w = _weights(rids, rids, [f"g{v}" for v in rids], "area", 0.0, rng)
    full = rng.random((R, G, NC))
    holes = full.copy()
    holes[3] = np.nan
    holes[7, 2, 1] = np.nan
    for panel in (full, holes, np.full((R, G, NC), np.nan)):            # everything kept (fast path) / some / nothing
        pd.testing.assert_frame_equal(_panel_frame(torch.from_numpy(panel), names, labels, rids, w), _literal(panel, names, labels, rids, w))
    cal = af.CalendarIndex.range("noleap", 2001, G, "D")               # object-valued time column
    for panel in (full, holes):
        pd.testing.assert_frame_equal(_panel_frame(torch.from_numpy(panel), names, cal, rids, w), _literal(panel, names, cal, rids, w))
    dup = _weights(rids, np.array([10, 13, 13, 16]), ["a", "b", "b2", "c"], "area", 0.0, rng)      # repeated region rows: literal route
    pd.testing.assert_frame_equal(_panel_frame(torch.from_numpy(holes), names, labels, rids, dup), _literal(holes, names, labels, rids, dup))
    one = _panel_frame(torch.from_numpy(full[:, :1]), names, labels[:1], rids, w)                   # yearly panel: one period
    assert len(one) == R and one["time"].nunique() == 1


@pytest.mark.parametrize("zero_weight", ["nan", "area"])
@pytest.mark.parametrize("string_ids", [True, False])
@pytest.mark.parametrize("calendar", [False, True])
def test_panel_table_holds_the_frames_rows_and_round_trips_through_the_writers(tmp_path, zero_weight, string_ids, calendar):
    """``_panel_table`` (the Arrow route ``aggfly run`` writes from) == ``_panel_frame`` row for row; parquet / feather / csv
    written by ``io.write_table`` read back like the files ``io.write_output`` writes from the frame."""
    from aggfly_b200 import io as aio
    from aggfly_b200.aggregate import _panel_table
    from aggfly_b200.timeaxis import CalendarIndex
    rng = np.random.default_rng(7 + string_ids + 2 * calendar)
    R, G, NC = 25, 6, 2
    panel = rng.random((R, G, NC))
    panel[rng.random(R) < 0.15] = np.nan
    rids = np.arange(R) * 3
    shp_ids = rng.permutation(rids[:-2])
    vals = [f"g{v}" for v in shp_ids] if string_ids else (shp_ids * 10).astype(np.int64)
    w = _weights(rids, shp_ids, vals, zero_weight, 0.2, rng)
    names = ["a", "b"]
    labels = CalendarIndex.range("noleap", 2001, G) if calendar else pd.date_range("2001-01-01", periods=G, freq="D")
    frame = _panel_frame(torch.from_numpy(panel), names, labels, rids, w)
    table = _panel_table(torch.from_numpy(panel), names, labels, rids, w)
    assert table.column_names == list(frame.columns) and table.num_rows == len(frame)
    got = table.to_pandas()
    want = frame.reset_index(drop=True)
    if calendar:
        want = want.assign(time=want["time"].map(lambda t: t.isoformat()))
    pd.testing.assert_frame_equal(got, want, check_dtype=False)
    read_csv = pd.read_csv if calendar else (lambda p: pd.read_csv(p, parse_dates=["time"]))     # Arrow prints the time of day too
    for fmt, read in (("parquet", pd.read_parquet), ("feather", pd.read_feather), ("csv", read_csv)):
        a = aio.write_table(table, str(tmp_path / f"t.{fmt}"), fmt)
        b = aio.write_output(frame, str(tmp_path / f"f.{fmt}"), fmt)
        pd.testing.assert_frame_equal(read(a), read(b), check_dtype=False)
