"""Pins the CPU oracle (oracle/) against (a) the reference's own numba kernels, via the committed
fixtures in tests/golden/ref_kernels.npz, and (b) the reference test-suite's hard-coded numbers."""
import numpy as np
import pandas as pd
import pytest

from oracle import oracle as orc
from tests import refcases as rc


def _eq(a, b):
    """bit-for-bit, NaNs in the same places"""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.array_equal(a[m], b[m])


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("tag", ["clean", "nan"])
def test_stat_kernels_match_reference_numba(golden, dt, tag):
    cube, bounds = golden[f"cube_{dt}_{tag}"], golden["bounds"]
    for calc in orc.STAT_CODE:
        _eq(orc.block_stat(cube, bounds, calc), golden[f"stat_{calc}_{dt}_{tag}"])


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("tag", ["clean", "nan"])
def test_dd_bins_match_reference_numba(golden, dt, tag):
    cube, bounds, dda = golden[f"cube_{dt}_{tag}"], golden["bounds"], golden["ddargs"]
    _eq(orc.block_dd(cube, bounds, dda), golden[f"dd_{dt}_{tag}"])
    _eq(orc.block_bins(cube, bounds, dda), golden[f"bins_{dt}_{tag}"])


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("tag", ["clean", "nan"])
def test_sine_dd_matches_reference_numba(golden, dt, tag):
    # libm (gcc) vs numba's LLVM intrinsics: transcendental results may differ in the last ulp
    cube, bounds, dda = golden[f"cube_{dt}_{tag}"], golden["bounds"], golden["ddargs"]
    got, want = orc.block_sine_dd(cube, bounds, dda), golden[f"sine_dd_{dt}_{tag}"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    tol = 1e-5 if dt == "float32" else 1e-12
    assert np.allclose(got, want, rtol=tol, atol=tol, equal_nan=True)


@pytest.mark.parametrize("tag", ["full", "gap", "daily"])
def test_group_bounds_match_reference(golden, tag):
    t = pd.DatetimeIndex(golden[f"time_{tag}"].astype("datetime64[ns]"))
    for freq in ("1D", "ME", "YE", "W"):
        b, lab = orc.resample_groups(t, freq)
        assert np.array_equal(b, golden[f"bounds_{tag}_{freq}"])
        assert np.array_equal(lab.values.astype("datetime64[ns]").astype(np.int64), golden[f"labels_{tag}_{freq}"])
    lab1 = orc.resample_groups(t, "1D")[1]
    for f2 in ("ME", "YE", "W"):
        b2, lab2 = orc.resample_groups(lab1, f2)
        assert np.array_equal(b2, golden[f"bounds2_{tag}_{f2}"])
        assert np.array_equal(lab2.values.astype("datetime64[ns]").astype(np.int64), golden[f"labels2_{tag}_{f2}"])


def test_spatial_helpers_match_reference(golden):
    wdf = pd.DataFrame({"cell_id": golden["sp_cell_id"], "index_right": golden["sp_index_right"],
                        "weight": golden["sp_weight"]})
    ri, ci, wv, rids = orc.weight_triplets(wdf, np.arange(golden["sp_block"].shape[0]))
    assert np.array_equal(ri, golden["sp_region_idx"]) and np.array_equal(ci, golden["sp_cell_idx"])
    assert np.array_equal(wv, golden["sp_w_vals"]) and np.array_equal(rids, golden["sp_region_ids"])
    _eq(orc.scatter_block(golden["sp_block"], ri, ci, wv, len(rids)), golden["sp_scatter"])


# ---- the reference test-suite's hard-coded expectations ------------------------------------
def _ods():
    arr, time, lat, lon = rc.dataset_360_arrays()
    return orc.ODataset(arr, time, lat, lon, lon_is_360=True)


def test_reference_golden_time_matrix():
    out = orc.aggregate_time(_ods(), rc.golden_time_spec())
    cols = ["bins_-99_20", "bins_20_99", "cooling_dday", "tavg_1", "tavg_2"]
    assert list(out.keys()) == cols
    mat = np.stack([out[c][0].reshape(-1) for c in cols], axis=1)      # one month -> G == 1
    assert np.allclose(mat, rc.GOLDEN_TIME_MATRIX)
    assert list(out["tavg_1"][1]) == [pd.Timestamp("2000-07-31")]


def test_reference_golden_panel():
    w = orc.OWeights(rc.fixture_weights_frame(), np.arange(4),
                     pd.DataFrame({"geoid": ["region_1"]}), "geoid", zero_weight="nan")
    df = orc.aggregate_dataset(w, _ods(), **rc.golden_panel_spec())
    assert list(df.columns) == ["geoid", "time", "tavg_1", "tavg_2"]
    assert np.allclose(df[["tavg_1", "tavg_2"]].values, rc.GOLDEN_PANEL)


def test_reference_cftime_bounds():
    """aggfly/tests/test_aggregate.py:454-466."""
    t360 = orc.cal_range("360_day", 2000, 720)
    b_m, lab_m = orc.resample_groups(t360, "ME")
    assert set(np.diff(b_m).tolist()) == {30} and len(lab_m) == 24
    assert orc.resample_groups(t360, "YE")[0].tolist() == [0, 360, 720]
    b_nl, _ = orc.resample_groups(orc.cal_range("noleap", 2000, 365), "ME")
    assert np.diff(b_nl)[:3].tolist() == [31, 28, 31]


def test_reference_cftime_empty_bin():
    """aggfly/tests/test_aggregate.py:494-514: a missing month stays as a zero-width group -> NaN."""
    t = orc.cal_range("360_day", 2000, 90)
    keep = t.month != 2
    tg = orc.CalTime("360_day", t.year[keep], t.month[keep], t.day[keep])
    arr = np.random.default_rng(1).normal(15, 10, (len(tg), 2, 2))
    out = orc.aggregate_time(orc.ODataset(arr, tg, [-45.0, 45.0], [10.0, 100.0], False),
                             dict(v=[("aggregate", {"calc": "mean", "groupby": "month"})]))
    a = out["v"][0]
    assert a.shape[0] == 3 and np.all(np.isnan(a[1])) and np.all(np.isfinite(a[[0, 2]]))
    with pytest.raises(NotImplementedError, match="week"):
        orc.aggregate_time(orc.ODataset(arr, tg, [-45.0, 45.0], [10.0, 100.0], False),
                           dict(v=[("aggregate", {"calc": "mean", "groupby": "week"})]))


@pytest.mark.parametrize("case", [rc.spatial_case_multiregion_nan, rc.spatial_case_dropna_empty_group])
def test_reference_spatial_scenarios(case):
    vals, time, wdf = case()
    n_t = vals.shape[0]
    want = rc.wavg_loop_oracle({"v": vals.reshape(n_t, 4).T}, time.values, [0, 1, 2, 3], wdf, ["v"])
    w = orc.OWeights(wdf, np.arange(4), pd.DataFrame({"id": ["a", "b"]}), "id", zero_weight="area")
    got = orc.aggregate_space({"v": (vals, time)}, False, np.array([0.0, 1.0]), w)
    got = got.sort_values(["region_id", "time"]).reset_index(drop=True)
    want = want.sort_values(["region_id", "time"]).reset_index(drop=True)
    assert got.shape == want.shape
    assert (got[["region_id", "time"]].values == want[["region_id", "time"]].values).all()
    assert np.allclose(got["v"].values, want["v"].values)


def test_zero_weight_nan_policy_keeps_empty_region():
    """aggfly/tests/test_aggregate.py:1458-1468, 1507-1529 with a hand-built weights frame."""
    time = pd.date_range("2000-01-01", periods=2)
    arr = np.ones((2, 1, 4)); arr[1] = np.nan
    wdf = pd.DataFrame({"cell_id": [0, 1, 2, 3], "index_right": [0, 0, 1, 1], "weight": [0.5, 0.5, 0.0, 0.0]})
    shp = pd.DataFrame({"geoid": ["has_pop", "no_pop"]})
    spec = dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})])
    ds = orc.ODataset(arr, time, [0.5], [0.5, 1.5, 2.5, 3.5], False)
    df = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(4), shp, "geoid", "nan"), ds, **spec)
    assert len(df[df.geoid == "has_pop"]) == 1
    assert len(df[df.geoid == "no_pop"]) == 2 and df[df.geoid == "no_pop"].tavg.isna().all()
    df2 = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(4), shp, "geoid", "area"), ds, **spec)
    assert set(df2.geoid) == {"has_pop"}
