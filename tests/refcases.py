"""Shared test inputs: the reference test-suite's fixtures restated as plain arrays.

Every number here is copied as a NUMBER from the reference's tests (file:line cited); no
reference code is executed at test time.
"""
import numpy as np
import pandas as pd


def dataset_360_arrays():
    """aggfly/tests/test_aggregate.py:17-53 -- 4 x 12-hourly steps, 2x2 grid, lon 90/270."""
    np.random.seed(1216)
    x = np.linspace(0, 360, 3)
    longitude = (x[1:] + x[:-1]) / 2
    y = np.linspace(-90, 90, 3)
    latitude = (y[1:] + y[:-1]) / 2
    time = pd.date_range("2000-07-01", periods=4, freq="12h")
    arr = np.random.normal(20, 15, (len(time), len(latitude), len(longitude)))
    return arr, time, latitude, longitude


# aggfly/tests/test_aggregate.py:234-237 (weights sorted by cell_id, cell ids 0..3)
FIXTURE_WEIGHTS = np.array([0.18959496, 0.27482559, 0.09094742, 0.13207367])

# aggfly/tests/test_aggregate.py:275-280 -- rows = cells (lat, lon) row-major of the un-rescaled
# grid; columns bins_-99_20, bins_20_99, cooling_dday, tavg_1, tavg_2
GOLDEN_TIME_MATRIX = np.array([
    [0.0, 2.0, 44.945648, 62.472824, 1956.361671],
    [1.0, 1.0, 25.910298, 39.60287, 801.80304],
    [1.0, 1.0, 9.12584, 35.789426, 670.521066],
    [1.0, 1.0, 14.932308, 37.648473, 858.069229]])

# aggfly/tests/test_aggregate.py:311-313
GOLDEN_PANEL = np.array([[47.75461, 1245.594351]])


def golden_time_spec():
    """aggfly/tests/test_aggregate.py:255-270."""
    return dict(
        bins=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("aggregate", {"calc": "bins", "groupby": "month", "ddargs": [[-99, 20, 0], [20, 99, 0]]})],
        cooling_dday=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [20, 99, 0]}),
                      ("aggregate", {"calc": "sum", "groupby": "month"})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
              ("aggregate", {"calc": "sum", "groupby": "month"})],
    )


def golden_panel_spec():
    """aggfly/tests/test_aggregate.py:301-307."""
    return dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                      ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                      ("aggregate", {"calc": "sum", "groupby": "month"})])


def fixture_weights_frame():
    return pd.DataFrame({"cell_id": [0, 1, 2, 3], "index_right": [0, 0, 0, 0],
                         "weight": FIXTURE_WEIGHTS})


def wavg_loop_oracle(vals, time, grid_cell_ids, wdf, names):
    """Independent pure-loop weighted average (same role as the reference tests' helper,
    aggfly/tests/test_aggregate.py:578-601): vals[name] is (n_cells, n_time)."""
    pos = {int(c): i for i, c in enumerate(grid_cell_ids)}
    rows = []
    for r in np.sort(wdf["index_right"].unique()):
        sub = wdf[wdf["index_right"] == r]
        cidx = np.array([pos[int(c)] for c in sub["cell_id"]])
        wv = sub["weight"].to_numpy(dtype=float)
        for ti in range(len(time)):
            ok = np.ones(len(cidx), bool)
            for nm in names:
                ok &= ~np.isnan(vals[nm][cidx, ti])
            den = wv[ok].sum()
            if den == 0:
                continue
            row = {"region_id": r, "time": time[ti]}
            for nm in names:
                row[nm] = (wv[ok] * vals[nm][cidx, ti][ok]).sum() / den
            rows.append(row)
    return pd.DataFrame(rows)


SPATIAL_WDF = dict(cell_id=[0, 1, 2, 1, 2, 3], index_right=[0, 0, 0, 1, 1, 1],
                   weight=[0.5, 0.3, 0.2, 0.4, 0.4, 0.2])   # test_aggregate.py:625-629


def spatial_case_multiregion_nan():
    """aggfly/tests/test_aggregate.py:614-639."""
    time = pd.date_range("2000-07-01", periods=3, freq="D")
    vals = np.random.default_rng(7).normal(20, 5, (3, 2, 2))
    vals[1, 0, 0] = np.nan
    vals[2, 1, 1] = np.nan
    return vals, time, pd.DataFrame(SPATIAL_WDF)


def spatial_case_dropna_empty_group():
    """aggfly/tests/test_aggregate.py:642-664."""
    time = pd.date_range("2000-07-01", periods=2, freq="D")
    vals = np.random.default_rng(3).normal(20, 5, (2, 2, 2))
    vals[0, 0, 0] = vals[0, 0, 1] = vals[0, 1, 0] = np.nan
    return vals, time, pd.DataFrame(SPATIAL_WDF)
