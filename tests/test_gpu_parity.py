"""Parity of the CUDA path (through the public API -> ctypes -> C-ABI -> kernels) against the
CPU oracle on the same seeded inputs.  Bar: bit-exact for group indices, bin counts and every
single-stripe result; rel 1e-12 where the stripe merge re-associates an fp64 sum (the
north-star tolerance for weighted float outputs is rel 1e-5)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

import aggfly_b200 as af
from aggfly_b200 import engine
from aggfly_b200.spec import ColSpec, LaneSpec, ProgramSpec, Stage
from oracle import oracle as orc
from tests import refcases as rc


@pytest.fixture(autouse=True, params=["tma", "ldg"])
def _variant(request):
    """Every test runs against both temporal-kernel variants: the TMA/shared-memory ring (default
    whenever the raster view is 16-byte aligned) and the direct-load kernel (any shape)."""
    import os
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    engine.OPTIONS["target_stripes"] = 0
    if request.param == "ldg":
        os.environ["AGF_DISABLE_TMA"] = "1"
    else:
        os.environ.pop("AGF_DISABLE_TMA", None)
        assert _lib_tma_eligible(60, 4) and _lib_tma_eligible(60, 8) and not _lib_tma_eligible(45, 4)
    yield
    os.environ.pop("AGF_DISABLE_TMA", None)
    engine.OPTIONS["target_stripes"] = 0


def _lib_tma_eligible(n_cells, esz):
    """The library's rule: a 16-byte aligned base (torch allocations are) and a row pitch that is a multiple of 16."""
    return (n_cells * esz) % 16 == 0


def _exact(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.array_equal(a[m], b[m]), np.abs(a[m] - b[m]).max()


def _close(a, b, rtol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


# ---- kernel level: arbitrary bounds, every calc, both dtypes, NaNs, empty group --------------
def _run_lanes(cube, bounds, lanes):
    import torch
    T, Y, X = cube.shape
    prog = ProgramSpec(input=None, in_dtype=cube.dtype, bounds1=np.asarray(bounds), bounds2=None,
                       lanes=[LaneSpec(c, dd) for c, dd in lanes], n_time=T)
    prog._source = None
    off = 3 if lanes[0][0] == "sine_dd" else 0
    if off:
        prog.lanes[:0] = [LaneSpec("_hidden_sum"), LaneSpec("_hidden_min"), LaneSpec("_hidden_max")]
    prog.cols = [ColSpec(off + i, None, 0.0, cube.dtype == np.float64, out_col=i) for i in range(len(lanes))]
    stage = Stage(programs=[prog], nodes=[None] * len(lanes), dtype=cube.dtype, labels=np.arange(len(bounds) - 1))
    res = engine.run_stage(stage, engine.to_device(cube).reshape(T, Y * X), Y * X)
    torch.cuda.synchronize()
    return np.moveaxis(res.X.cpu().numpy().reshape(len(bounds) - 1, Y, X, len(lanes)), -1, 1), res.V.cpu().numpy()


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("tag", ["clean", "nan"])
def test_level1_kernels_match_reference_numba_outputs(golden, dt, tag):
    cube, bounds, dda = golden[f"cube_{dt}_{tag}"], golden["bounds"], golden["ddargs"]
    stats = ["mean", "sum", "min", "max", "nanmean"]
    got, valid = _run_lanes(cube, bounds, [(c, None) for c in stats])
    for i, c in enumerate(stats):
        _exact(got[:, i], golden[f"stat_{c}_{dt}_{tag}"])
    assert np.array_equal(valid.reshape(got[:, 0].shape), ~np.isnan(got).any(axis=1))
    for calc in ("dd", "bins"):
        got, _ = _run_lanes(cube, bounds, [(calc, tuple(r)) for r in dda])
        _exact(np.moveaxis(got, 1, -1), golden[f"{calc}_{dt}_{tag}"])
    got, _ = _run_lanes(cube, bounds, [("sine_dd", tuple(r)) for r in dda])
    want = golden[f"sine_dd_{dt}_{tag}"]
    assert np.array_equal(np.isnan(np.moveaxis(got, 1, -1)), np.isnan(want))
    tol = 1e-5 if dt == "float32" else 1e-12                 # transcendental libm differences only
    assert np.allclose(np.moveaxis(got, 1, -1), want, rtol=tol, atol=tol, equal_nan=True)


# ---- reference test-suite goldens -------------------------------------------------------------
def _ref_dataset():
    arr, time, lat, lon = rc.dataset_360_arrays()
    return af.Dataset.from_arrays(arr, time, lat, lon, lon_is_360=True)


def _ref_weights(ds, zero_weight="nan"):
    w = af.weights_from_objects(ds, af.GeoRegions(pd.DataFrame({"geoid": ["region_1"]}), "geoid"), zero_weight=zero_weight)
    w.weights = rc.fixture_weights_frame()
    return w


def test_reference_golden_time_matrix():
    out = af.aggregate_time(dataset=_ref_dataset(), weights=None, **rc.golden_time_spec())
    cols = ["bins_-99_20", "bins_20_99", "cooling_dday", "tavg_1", "tavg_2"]
    assert list(out) == cols
    mat = np.stack([out[c].values.reshape(-1) for c in cols], axis=1)
    assert np.allclose(mat, rc.GOLDEN_TIME_MATRIX)
    assert list(out["tavg_1"].time) == [pd.Timestamp("2000-07-31")]


def test_reference_golden_panel():
    ds = _ref_dataset()
    df = af.aggregate_dataset(dataset=ds, weights=_ref_weights(ds), **rc.golden_panel_spec())
    assert list(df.columns) == ["geoid", "time", "tavg_1", "tavg_2"]
    assert np.allclose(df[["tavg_1", "tavg_2"]].values, rc.GOLDEN_PANEL)
    with pytest.warns(DeprecationWarning, match="no longer builds a Dask cluster"):
        df2 = af.aggregate_dataset(dataset=ds, weights=_ref_weights(ds), n_workers=50, processes=True,
                                   **rc.golden_panel_spec())
    assert "n_workers" not in df2.columns and np.array_equal(df2["tavg_2"].values, df["tavg_2"].values)
    with pytest.raises(ValueError, match="No dataset provided"):
        af.aggregate_dataset(weights=_ref_weights(ds), **rc.golden_panel_spec())


# ---- chains vs the oracle on seeded synthetic rasters --------------------------------------------
BINS13 = [[-20 + 5 * i, -15 + 5 * i, 0] for i in range(13)]
SPECS = {
    "c1_tavg_poly": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                               ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "c2_gdd": dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                        ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "c3_bins_and_poly": dict(
        temp_bins=[("aggregate", {"calc": "mean", "groupby": "date"}),
                   ("aggregate", {"calc": "bins", "groupby": "year", "ddargs": BINS13})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
              ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "c3b_daily": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13})],
                      tavg=[("aggregate", {"calc": "mean", "groupby": "date"})]),
    "monthly_mix": dict(
        tmax=[("aggregate", {"calc": "max", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})],
        tmin=[("aggregate", {"calc": "min", "groupby": "date"}), ("aggregate", {"calc": "min", "groupby": "month"})],
        hdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [-99, 18.3, 1]}),
             ("aggregate", {"calc": "sum", "groupby": "month"})],
        direct=[("aggregate", {"calc": "mean", "groupby": "month"})]),
    "weekly": dict(w=[("aggregate", {"calc": "sum", "groupby": "date"}), ("aggregate", {"calc": "max", "groupby": "week"})]),
    # daily minimum / maximum / mean of the hourly values (fixed lane layout KIND_MMS): two-level, per date, without a mean
    "tmin_tmax_tavg_monthly": dict(
        tmin=[("aggregate", {"calc": "min", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})],
        tmax=[("aggregate", {"calc": "max", "groupby": "date"}), ("aggregate", {"calc": "max", "groupby": "month"})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})]),
    "tmin_tmax_tavg_daily": dict(tmax=[("aggregate", {"calc": "max", "groupby": "date"})],
                                 tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                                 tmin=[("aggregate", {"calc": "min", "groupby": "date"})]),
    "tmin_tmax_yearly": dict(
        tmax=[("aggregate", {"calc": "max", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})],
        tmin=[("aggregate", {"calc": "min", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})]),
    # more degree-day thresholds than one four-lane program holds (several passes instead of the sixteen-lane general form)
    "dd6": dict(dd=[("aggregate", {"calc": "dd", "groupby": "date",
                                   "ddargs": [[0, 10, 0], [10, 20, 0], [20, 30, 0], [30, 99, 0], [-99, 0, 1], [10, 30, 0]]}),
                    ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "spline_and_pow": dict(
        s=[("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"transform": "spline"}),
           ("aggregate", {"calc": "sum", "groupby": "month"})],
        p=[("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"transform": "power", "exp": [[1, 3]]}),
           ("aggregate", {"calc": "sum", "groupby": "month"})],
        q=[("aggregate", {"calc": "mean", "groupby": "month"}), ("transform", {"transform": "power", "exp": np.arange(2, 4)})]),
    "dd_bins_of_daily_dd": dict(x=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [[10, 30, 0], [0.1, 17.3, 1]]}),
                                   ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "three_levels": dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "sum", "groupby": "month"}),
                            ("aggregate", {"calc": "max", "groupby": "year"})]),
    "nanmean_multipass": dict(a=[("aggregate", {"calc": "nanmean", "groupby": "date"}),
                                 ("aggregate", {"calc": "nanmean", "groupby": "month"})]),
    # the reference's example config (examples/era5_counties_area.yaml) + a heating degree-day lane
    "area_example": dict(
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})],
        gdd_10_30=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                   ("aggregate", {"calc": "sum", "groupby": "year"})],
        hdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [-99, 18.3, 1]}),
             ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # one mean lane + degree days feeding typed slots (bins of the daily mean, polynomial, dd sums)
    "mix_bins_poly_dd": dict(
        tb=[("aggregate", {"calc": "mean", "groupby": "date"}),
            ("aggregate", {"calc": "bins", "groupby": "month", "ddargs": BINS13})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 4)}),
              ("aggregate", {"calc": "sum", "groupby": "month"})],
        dd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [[10, 30, 0], [0.1, 17.3, 1]]}),
            ("aggregate", {"calc": "sum", "groupby": "month"})]),
    # daily panel: mean + degree days per date (single-level mixed layout)
    "daily_mean_dd": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                          gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]})],
                          tsum=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [-99, 18.3, 1]})]),
    "sine": dict(s=[("aggregate", {"calc": "sine_dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                    ("aggregate", {"calc": "sum", "groupby": "month"})],
                 h=[("aggregate", {"calc": "sine_dd", "groupby": "date", "ddargs": [[5, 18, 1], [12, 99, 0]]})]),
}
INEXACT = {"spline_and_pow": 1e-6, "sine": 1e-5, "mix_bins_poly_dd": 1e-14}   # pow: libm vs exact products     # powf / libm transcendental differences only


def _raster(dtype, nan, T=24 * 75 + 7, Y=5, X=12, seed=0):
    """Seeded test raster.  The default 5 x 12 = 60 cells make the row pitch a multiple of 16 bytes for both
    dtypes, so the "tma" variant of every test really runs the TMA kernels (a 45-cell raster silently fell back
    to the direct-load kernel)."""
    rng = np.random.default_rng(seed)
    t = pd.date_range("2001-11-20 03:00", periods=T, freq="h")
    hours = np.arange(T)
    base = 12 + 14 * np.sin(2 * np.pi * hours / (24 * 365))[:, None, None] + 5 * np.sin(2 * np.pi * hours / 24)[:, None, None]
    arr = (base + rng.normal(0, 6, (T, Y, X))).astype(dtype)
    if nan:
        arr[:, 0, 0] = np.nan
        arr[rng.random((T, Y, X)) < 0.001] = np.nan
        arr[24 * 30: 24 * 31, min(2, Y - 1), :] = np.nan
    lat = np.linspace(49.75, 24.0, Y)
    lon = np.linspace(235.0, 293.75, X)
    return arr, t, lat, lon


def _both_time(arr, t, lat, lon, spec):
    want = orc.aggregate_time(orc.ODataset(arr, t, lat, lon, True), spec)
    got = af.aggregate_time(dataset=af.Dataset.from_arrays(arr, t, lat, lon, True), weights=None, aggregator_dict=spec)
    assert list(got) == list(want)
    return got, want


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("nan", [False, True])
def test_chains_match_oracle_single_stripe(name, dtype, nan):
    engine.OPTIONS["target_stripes"] = 1
    arr, t, lat, lon = _raster(dtype, nan)
    got, want = _both_time(arr, t, lat, lon, SPECS[name])
    for k in want:
        w_arr, w_lab = want[k]
        assert got[k].values.dtype == w_arr.dtype, (k, got[k].values.dtype, w_arr.dtype)
        assert np.array_equal(pd.DatetimeIndex(got[k].time).values, pd.DatetimeIndex(w_lab).values)
        if name in INEXACT:
            _close(got[k].values, w_arr, INEXACT[name])
        else:
            _exact(got[k].values, w_arr)


@pytest.mark.parametrize("name", ["c1_tavg_poly", "c3_bins_and_poly", "monthly_mix", "weekly", "c3b_daily", "area_example",
                                  "mix_bins_poly_dd", "daily_mean_dd"])
@pytest.mark.parametrize("stripes", [0, 3, 17, 1000])
def test_chains_match_oracle_many_stripes(name, stripes):
    engine.OPTIONS["target_stripes"] = stripes
    arr, t, lat, lon = _raster("float32", True, seed=3)
    got, want = _both_time(arr, t, lat, lon, SPECS[name])
    for k in want:
        if "bins" in k or "tmin" in k or "tmax" in k or k == "w" or k.startswith("tb_"):
            _exact(got[k].values, want[k][0])              # counts / min / max do not depend on the split
        else:
            _close(got[k].values, want[k][0], 1e-12)


def test_ragged_time_axis_with_gaps_and_tiny_grid():
    arr, t, lat, lon = _raster("float32", True, T=24 * 20 + 5, Y=1, X=3, seed=9)
    keep = np.ones(len(t), bool); keep[24 * 3 + 7: 24 * 6 + 2] = False; keep[-3:] = False
    spec = dict(m=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "sum", "groupby": "month"})],
                b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [[0, 10, 0], [10, 99, 0]]}),
                   ("aggregate", {"calc": "sum", "groupby": "month"})])
    for stripes in (1, 5):
        engine.OPTIONS["target_stripes"] = stripes
        got, want = _both_time(arr[keep], t[keep], lat, lon, spec)
        for k in want:
            _close(got[k].values, want[k][0], 1e-12)
    d = [("aggregate", {"calc": "mean", "groupby": "date"})]
    got, want = _both_time(arr[keep], t[keep], lat, lon, dict(d=d))
    _exact(got["d"].values, want["d"][0])                   # empty days -> NaN rows kept
    assert np.isnan(want["d"][0][4]).all()


def test_noleap_calendar_chain():
    n = 365 * 3
    arr = np.random.default_rng(4).normal(15, 12, (n, 3, 4)).astype(np.float32)
    arr[100, 1, 1] = np.nan
    spec = dict(gdd_m=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                       ("aggregate", {"calc": "sum", "groupby": "month"})])
    want = orc.aggregate_time(orc.ODataset(arr, orc.cal_range("noleap", 2000, n), [0, 1, 2], [0, 1, 2, 3], False), spec)
    got = af.aggregate_time(dataset=af.Dataset.from_arrays(arr, af.CalendarIndex.range("noleap", 2000, n),
                                                          [0, 1, 2], [0, 1, 2, 3], False), weights=None, **spec)
    _close(got["gdd_m"].values, want["gdd_m"][0], 1e-12)
    assert len(got["gdd_m"].time) == 36 and got["gdd_m"].time.to_objects()[1].day == 28


# ---- end to end vs the oracle ---------------------------------------------------------------------
def _weights_case(lat, lon, rng, n_regions=7, zero_weight="nan", shuffle=True):
    ny, nx = len(lat), len(lon)
    rows = []
    for r in range(n_regions):
        cells = rng.choice(ny * nx + 4, size=rng.integers(1, 12), replace=False)   # a few ids off-grid
        for c in cells:
            rows.append((int(c), 10 + 3 * r, float(rng.random()) if r != 2 else 0.0))   # region row 2 has zero weight
    wdf = pd.DataFrame(rows, columns=["cell_id", "index_right", "weight"])
    if shuffle:
        wdf = wdf.sample(frac=1.0, random_state=1).reset_index(drop=True)
    shp = pd.DataFrame({"geoid": [f"r{r}" for r in range(n_regions + 1)]}, index=[10 + 3 * r for r in range(n_regions + 1)])
    return wdf, shp


@pytest.mark.parametrize("zero_weight", ["nan", "area", "drop"])
@pytest.mark.parametrize("lon_is_360", [True, False])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily", "monthly_mix"])
def test_aggregate_dataset_matches_oracle(name, lon_is_360, zero_weight):
    arr, t, lat, lon = _raster("float32", True, T=24 * 40, seed=11)
    if not lon_is_360:
        lon = lon - 360.0
    else:
        lon = np.concatenate([lon[4:], lon[:4] - 200.0])       # unsorted once relabelled: exercises the remap
    rng = np.random.default_rng(2)
    wdf, shp = _weights_case(lat, lon, rng)
    want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(arr.shape[1] * arr.shape[2]), shp, "geoid", zero_weight),
                                 orc.ODataset(arr, t, lat, lon, lon_is_360), aggregator_dict=SPECS[name])
    ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360)
    w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight=zero_weight)
    w.weights = wdf
    got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])
    assert list(got.columns) == list(want.columns)
    assert len(got) == len(want)
    assert (got["geoid"].values == want["geoid"].values).all()
    assert (got["time"].values == want["time"].values).all()
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    _close(got[vals].values, want[vals].values, 1e-11)


@pytest.mark.parametrize("case", [rc.spatial_case_multiregion_nan, rc.spatial_case_dropna_empty_group])
def test_reference_spatial_scenarios(case):
    vals, time, wdf = case()
    n_t = vals.shape[0]
    want = rc.wavg_loop_oracle({"v": vals.reshape(n_t, 4).T}, time.values, [0, 1, 2, 3], wdf, ["v"])
    ds = af.Dataset.from_arrays(vals, time, [0.0, 1.0], [0.0, 1.0], lon_is_360=False)
    w = af.GridWeights.from_frame(wdf, ds.grid, af.GeoRegions(pd.DataFrame({"id": ["a", "b"]}), "id"), zero_weight="area")
    got = af.SpatialAggregator([ds], w, names=["v"]).compute()
    assert got.shape == want.shape
    assert (got[["region_id", "time"]].values == want[["region_id", "time"]].values).all()
    assert np.allclose(got["v"].values, want["v"].values, rtol=1e-13)


def test_no_spec_aggregates_the_raw_series():
    arr, t, lat, lon = _raster("float32", True, T=30, seed=5)
    rng = np.random.default_rng(8)
    wdf, shp = _weights_case(lat, lon, rng)
    want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(arr.shape[1] * arr.shape[2]), shp, "geoid", "area"), orc.ODataset(arr, t, lat, lon, True))
    ds = af.Dataset.from_arrays(arr, t, lat, lon, True)
    w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="area"); w.weights = wdf
    got = af.aggregate_dataset(weights=w, dataset=ds)
    assert list(got.columns) == ["geoid", "time", "variable"] and len(got) == len(want)
    _close(got["variable"].values, want["variable"].values, 1e-12)


# ---- host-resident rasters: chunked feed overlapped with the kernels (stream.py) ---------------------
@pytest.mark.parametrize("feed", ["pinned", "pageable", "pageable_f64"])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily", "monthly_mix", "three_levels"])
def test_streamed_feed_is_bitwise_the_device_resident_result(name, feed):
    import torch
    from aggfly_b200 import stream
    dtype = "float64" if feed.endswith("f64") else "float32"
    arr, t, lat, lon = _raster(dtype, True, T=24 * 40 + 5, seed=13)
    rng = np.random.default_rng(3)
    wdf, shp = _weights_case(lat, lon, rng)
    want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(arr.shape[1] * arr.shape[2]), shp, "geoid", "nan"),
                                 orc.ODataset(arr, t, lat, lon, True), aggregator_dict=SPECS[name])

    def run(values):
        ds = af.Dataset.from_arrays(values, t, lat, lon, True)
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    engine.OPTIONS["target_stripes"] = 11
    old = dict(stream.OPTIONS)
    try:
        stream.OPTIONS.update(chunk_bytes=13 * arr[0].nbytes, staging_chunk_bytes=13 * arr[0].nbytes, staging_slots=3,
                              staging_threads=2)                                           # 13-row chunks
        resident = run(torch.from_numpy(arr).cuda())
        if feed == "pinned":
            host = torch.from_numpy(arr).pin_memory()
            assert host.is_pinned()
        else:
            host = arr
        stats_before = dict(stream.LAST_STATS)
        streamed = run(host)
        assert stream.LAST_STATS is not stats_before and stream.LAST_STATS["chunks"] == -(-arr.shape[0] // 13)
        assert stream.LAST_STATS["pinned"] == (feed == "pinned")
    finally:
        stream.OPTIONS.update(old)
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert list(streamed.columns) == list(want.columns) and len(streamed) == len(want)
    _exact(streamed[vals].values, resident[vals].values)          # same stripes, same merge order
    _close(streamed[vals].values, want[vals].values, 1e-11)


# ---- records longer than the device budget: a ring of device windows (stream._feed_ring) ---------------------
@pytest.mark.parametrize("feed", ["pinned", "pageable", "concat_pinned_parts"])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily", "monthly_mix", "three_levels"])
def test_ring_of_device_windows_is_bitwise_the_device_resident_result(name, feed):
    """The device holds two windows of about 100 rows at a time; the panel is bit for bit the one of the
    device-resident raster (same stripes, so the same merge order)."""
    import torch
    from aggfly_b200 import stream
    from aggfly_b200.dataset import TimeConcat
    arr, t, lat, lon = _raster("float32", True, T=24 * 40 + 5, seed=17)
    rng = np.random.default_rng(5)
    wdf, shp = _weights_case(lat, lon, rng)

    def run(values):
        ds = af.Dataset.from_arrays(values, t, lat, lon, True)
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    engine.OPTIONS["target_stripes"] = 11
    old = dict(stream.OPTIONS)
    row = arr[0].nbytes
    try:
        resident = run(torch.from_numpy(arr).cuda())
        stream.OPTIONS.update(chunk_bytes=13 * row, staging_chunk_bytes=13 * row, staging_slots=3, staging_threads=2,
                              device_raster_budget_bytes=1, ring_slot_bytes=100 * row, ring_slots=2)
        if feed == "pinned":
            host = torch.from_numpy(arr).pin_memory()
        elif feed == "pageable":
            host = arr
        else:               # three pinned "year buffers" concatenated lazily: their chunks are copied from in place
            cuts = [0, 24 * 13 + 7, 24 * 29, arr.shape[0]]
            keep = [torch.from_numpy(arr[a:b]).pin_memory() for a, b in zip(cuts[:-1], cuts[1:])]
            host = TimeConcat([k.numpy() for k in keep])
        streamed = run(host)
        st = stream.LAST_STATS
        assert st.get("ring") and st["ring_slots"] == 2
        if name.startswith("c3"):             # daily level-1 groups: many common stripe ends, windows of <= 100 rows
            assert st["windows"] >= 4 and st["ring_bytes"] < arr.nbytes / 3
        # (a month-level program next to a date-level one shares few stripe ends: the windows grow to what the cuts allow)
        if feed == "concat_pinned_parts":
            assert st["direct_chunks"] == st["chunks"]
        if feed == "pageable":
            assert st["direct_chunks"] == 0
    finally:
        stream.OPTIONS.update(old)
        stream.release_device_rasters()
    vals = [c for c in resident.columns if c not in ("geoid", "time")]
    assert list(streamed.columns) == list(resident.columns) and len(streamed) == len(resident)
    _exact(streamed[vals].values, resident[vals].values)


# ---- CF-packed integers in host memory: copied as stored, unpacked on the device (stream.feed_packed) -------------
@pytest.mark.parametrize("feed", ["pinned", "pageable"])
@pytest.mark.parametrize("stored,dtype", [("int16", "float32"), ("int16", "float64"), ("uint8", "float32"), ("int32", "float64")])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily"])
def test_packed_host_raster_is_unpacked_on_the_device(name, stored, dtype, feed):
    import torch
    from aggfly_b200 import stream
    from aggfly_b200.dataset import PackedRaster
    arr, t, lat, lon = _raster("float32", False, T=24 * 21 + 3, seed=23)
    info = np.iinfo(stored)
    scale, offset = (float(arr.max()) - float(arr.min())) / (info.max - info.min - 2), 3.25
    q = np.clip(np.rint((arr.astype(np.float64) - offset) / scale), info.min + 1, info.max).astype(stored)
    fill = float(info.min)
    q[5:9, 2, 3] = info.min                                        # _FillValue -> NaN
    rng = np.random.default_rng(9)
    wdf, shp = _weights_case(lat, lon, rng)

    def run(values):
        ds = af.Dataset.from_arrays(values, t, lat, lon, True)
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    decoded = (q.astype(np.float64) * scale + offset).astype(dtype)
    decoded[q == info.min] = np.nan
    host = torch.from_numpy(q).pin_memory() if feed == "pinned" else q
    packed = PackedRaster(host, scale, offset, fill, dtype)
    assert packed.shape == arr.shape and packed.dtype == np.dtype(dtype)
    assert np.array_equal(np.asarray(packed), decoded, equal_nan=True)
    assert np.array_equal(np.asarray(packed[24:48]), decoded[24:48], equal_nan=True)
    engine.OPTIONS["target_stripes"] = 7
    old = dict(stream.OPTIONS)
    try:
        stream.OPTIONS.update(chunk_bytes=17 * q[0].nbytes, staging_chunk_bytes=17 * q[0].nbytes, staging_slots=3, staging_threads=2)
        want = run(torch.from_numpy(decoded).cuda())
        got = run(packed)
        st = stream.LAST_STATS
        assert st.get("packed") and st["pinned"] == (feed == "pinned") and st["h2d_bytes"] == q.nbytes
        dev = engine.to_device(packed)                             # the device-resident route decodes the same way
        assert np.array_equal(dev.cpu().numpy(), decoded, equal_nan=True)
    finally:
        stream.OPTIONS.update(old)
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    _exact(got[vals].values, want[vals].values)


@pytest.mark.parametrize("source", ["one_packed_raster", "concat_of_packed_years", "pageable_packed"])
def test_packed_record_longer_than_the_device_goes_through_the_ring(source):
    """Packed integers + a record over the device budget: pinned pieces are copied as stored and unpacked straight into the
    ring windows; the panel is bit for bit that of the decoded raster held on the device."""
    import torch
    from aggfly_b200 import stream
    from aggfly_b200.dataset import PackedRaster, TimeConcat
    arr, t, lat, lon = _raster("float32", False, T=24 * 30 + 7, seed=29)
    scale, offset = 0.004, 11.5
    q = np.clip(np.rint((arr.astype(np.float64) - offset) / scale), -32000, 32000).astype(np.int16)
    q[40:44, 0, 1] = -32767
    decoded = (q.astype(np.float64) * scale + offset).astype(np.float32)
    decoded[q == -32767] = np.nan
    rng = np.random.default_rng(2)
    wdf, shp = _weights_case(lat, lon, rng)
    name = "c3_bins_and_poly"

    def run(values):
        ds = af.Dataset.from_arrays(values, t, lat, lon, True)
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    engine.OPTIONS["target_stripes"] = 9
    old = dict(stream.OPTIONS)
    row = decoded[0].nbytes
    try:
        want = run(torch.from_numpy(decoded).cuda())
        stream.OPTIONS.update(chunk_bytes=11 * row, staging_chunk_bytes=11 * row, staging_slots=3, staging_threads=2,
                              device_raster_budget_bytes=1, ring_slot_bytes=120 * row, ring_slots=2)
        if source == "one_packed_raster":
            values = PackedRaster(torch.from_numpy(q).pin_memory(), scale, offset, -32767.0)
        elif source == "pageable_packed":
            values = PackedRaster(q, scale, offset, -32767.0)
        else:
            cuts = [0, 24 * 11, 24 * 19 + 5, q.shape[0]]
            keep = [torch.from_numpy(q[a:b]).pin_memory() for a, b in zip(cuts[:-1], cuts[1:])]
            values = TimeConcat([PackedRaster(k, scale, offset, -32767.0) for k in keep])
        got = run(values)
        st = stream.LAST_STATS
        assert st.get("ring") and st["windows"] >= 3
        if source == "pageable_packed":
            assert st["unpacked_chunks"] == 0                          # decoded by the staging threads on the host
        else:
            assert st["unpacked_chunks"] == st["chunks"] and st["h2d_bytes"] == q.nbytes
    finally:
        stream.OPTIONS.update(old)
        stream.release_device_rasters()
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert len(got) == len(want) > 0
    _exact(got[vals].values, want[vals].values)


# ---- daily rasters: single-row inner groups collapse to one pass (spec.Planner._collapsed_lane) ----------
DAILY_SPECS = {
    "gdd_month": dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                           ("aggregate", {"calc": "sum", "groupby": "month"})]),
    "gdd_year_two": dict(x=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [[10, 30, 0], [0.1, 17.3, 1]]}),
                            ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "mixed_month": dict(
        gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}), ("aggregate", {"calc": "sum", "groupby": "month"})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})],
        hot=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [[25, 99, 0], [-99, 0, 0]]}),
             ("aggregate", {"calc": "sum", "groupby": "month"})],
        tx=[("aggregate", {"calc": "max", "groupby": "date"}), ("aggregate", {"calc": "max", "groupby": "month"})],
        nm=[("aggregate", {"calc": "nanmean", "groupby": "date"}), ("aggregate", {"calc": "nanmean", "groupby": "month"})]),
    "poly_month": dict(p=[("aggregate", {"calc": "mean", "groupby": "date"}),
                          ("transform", {"transform": "power", "exp": np.arange(1, 4)}),
                          ("aggregate", {"calc": "sum", "groupby": "month"})]),
    "daily_bins_of_daily": dict(b=[("aggregate", {"calc": "mean", "groupby": "date"}),
                                   ("aggregate", {"calc": "bins", "groupby": "month", "ddargs": BINS13})],
                                m=[("aggregate", {"calc": "mean", "groupby": "month"})]),
}


@pytest.mark.parametrize("name", list(DAILY_SPECS))
@pytest.mark.parametrize("calendar", ["standard", "noleap"])
@pytest.mark.parametrize("stripes", [1, 7])
def test_daily_raster_chains_match_oracle(name, calendar, stripes):
    engine.OPTIONS["target_stripes"] = stripes
    n = 365 * 2 + 40
    rng = np.random.default_rng(21)
    arr = (14 + 12 * np.sin(2 * np.pi * np.arange(n) / 365.0)[:, None, None] + rng.normal(0, 6, (n, 4, 7))).astype(np.float32)
    arr[rng.random(arr.shape) < 0.002] = np.nan
    arr[:, 0, 0] = np.nan
    lat, lon = np.linspace(40, 37, 4), np.linspace(250, 256, 7)
    if calendar == "standard":
        t = pd.date_range("1999-03-05", periods=n, freq="D")
        got, want = _both_time(arr, t, lat, lon, DAILY_SPECS[name])
    else:
        want = orc.aggregate_time(orc.ODataset(arr, orc.cal_range("noleap", 1999, n), lat, lon, True), DAILY_SPECS[name])
        got = af.aggregate_time(dataset=af.Dataset.from_arrays(arr, af.CalendarIndex.range("noleap", 1999, n), lat, lon, True),
                                weights=None, aggregator_dict=DAILY_SPECS[name])
        assert list(got) == list(want)
    for k in want:
        assert got[k].values.dtype == want[k][0].dtype, k
        if k == "p_3":
            _close(got[k].values, want[k][0], 1e-14)   # np.power(x, 3) is libm pow (< 1 ulp), ours is x*x*x rounded once
        elif stripes == 1 or "hot" in k or k in ("tx",) or k.startswith("b_"):
            _exact(got[k].values, want[k][0])
        else:
            _close(got[k].values, want[k][0], 1e-12)


def test_hourly_bins_and_mean_by_date_typed_lanes_many_shapes():
    """bins + mean per date on hourly data: the typed-lane kernels (int bin counters) with 3, 13 and
    27 bins, ragged first/last day, NaNs."""
    for nb in (3, 13, 27):
        bins = [[-30 + 3.0 * i, -27 + 3.0 * i + (0.5 if i % 2 else 0.0), 0] for i in range(nb)]
        spec = dict(hb=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": bins})],
                    tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                    tsum=[("aggregate", {"calc": "sum", "groupby": "date"})])
        for T in (24 * 12, 24 * 12 + 7):
            arr, t, lat, lon = _raster("float32", True, T=T, Y=3, X=172, seed=nb)            # 516 cells: two full CTAs + a partial one
            if T % 24 == 0:
                t = pd.date_range("2001-11-20 00:00", periods=T, freq="h")       # uniform 24-row groups
            got, want = _both_time(arr, t, lat, lon, spec)
            for k in want:
                _exact(got[k].values, want[k][0])


def test_sharded_call_world1_nccl_equals_plain_call():
    """aggregate_dataset_sharded through NCCL (world_size 1 on this box's GPU) == aggregate_dataset."""
    import socket
    import torch
    import torch.distributed as dist
    from aggfly_b200 import shard
    arr, t, lat, lon = _raster("float32", True, T=24 * 500, Y=4, X=6, seed=17)          # spans two year ends
    rng = np.random.default_rng(4)
    wdf, shp = _weights_case(lat, lon, rng)
    spec = SPECS["monthly_mix"]
    ds = af.Dataset.from_arrays(arr, t, lat, lon, True)
    w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
    w.weights = wdf
    plain = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        sharded = shard.aggregate_dataset_sharded(w, ds, spec, shard_by="year")
    finally:
        dist.destroy_process_group()
    assert list(sharded.columns) == list(plain.columns) and len(sharded) == len(plain)
    vals = [c for c in plain.columns if c not in ("geoid", "time")]
    assert (sharded["time"].values == plain["time"].values).all()
    _exact(sharded[vals].values, plain[vals].values)


@pytest.mark.parametrize("gs", ["1", "8", "16", "32"])
def test_every_spmm_variant_matches_oracle(gs, monkeypatch):
    """K2 picks lanes-per-(period, region) from the problem size; pin each variant in turn."""
    monkeypatch.setenv("AGF_SPMM_GS", gs)
    arr, t, lat, lon = _raster("float32", True, T=24 * 40, Y=6, X=12, seed=23)
    rng = np.random.default_rng(6)
    wdf, shp = _weights_case(lat, lon, rng, n_regions=9)
    big = pd.DataFrame({"cell_id": rng.permutation(72)[:50], "index_right": 10 + 3 * 8, "weight": rng.random(50)})
    wdf = pd.concat([wdf, big], ignore_index=True)                  # one region with more entries than a warp
    for name in ("c3b_daily", "c3_bins_and_poly", "monthly_mix", "c1_tavg_poly"):     # float2, double, float4, double2 loads
        want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(72), shp, "geoid", "nan"),
                                     orc.ODataset(arr, t, lat, lon, True), aggregator_dict=SPECS[name])
        ds = af.Dataset.from_arrays(arr, t, lat, lon, True)
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])
        vals = [c for c in want.columns if c not in ("geoid", "time")]
        assert len(got) == len(want) and (got["geoid"].values == want["geoid"].values).all()
        _close(got[vals].values, want[vals].values, 1e-11)


# ---- fused preprocess (preprocess.py): same bits as evaluating it with NumPy first ---------------------
@pytest.mark.parametrize("expr", ["kelvin_to_celsius", "(x - 32) * 5 / 9", "100 / x", "-x + 1"])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily", "monthly_mix"])
def test_fused_preprocess_equals_numpy_then_aggregate(name, dtype, expr):
    from aggfly_b200 import preprocess as pp
    engine.OPTIONS["target_stripes"] = 1
    arr, t, lat, lon = _raster(dtype, True, T=24 * 33 + 5, seed=29)
    raw = arr + np.asarray(273.15 if expr == "kelvin_to_celsius" else 40.0, dtype=arr.dtype)   # "Kelvin" / "Fahrenheit"-ish
    chain = pp.resolve(expr)
    pre = chain(raw)
    assert pre.dtype == raw.dtype
    want = orc.aggregate_time(orc.ODataset(pre, t, lat, lon, True), SPECS[name])
    got = af.aggregate_time(dataset=af.Dataset.from_arrays(raw, t, lat, lon, True, preprocess=expr), weights=None,
                            aggregator_dict=SPECS[name])
    assert list(got) == list(want)
    for k in want:
        _exact(got[k].values, want[k][0])


# ---- transforms the fused programs cannot carry: materialised by elementwise passes (never the CPU) ------
def test_materialised_transforms_and_interactions_match_oracle():
    arr, t, lat, lon = _raster("float32", True, T=24 * 45 + 3, Y=3, X=8, seed=31)
    rng = np.random.default_rng(12)
    n_days = len(pd.DatetimeIndex(t).normalize().unique())
    precip = rng.gamma(2.0, 1.5, (n_days, 3, 8)).astype(np.float32)              # a second daily variable
    precip64 = precip.astype(np.float64)
    hourly_w = rng.random((len(t), 3, 8)).astype(np.float32)
    spec = dict(
        poly=[("transform", {"transform": "power", "exp": np.arange(1, 3)}),      # powers of the HOURLY values
              ("aggregate", {"calc": "mean", "groupby": "month"})],
        txp=[("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"inter": precip}),
             ("aggregate", {"calc": "sum", "groupby": "month"})],
        txp64=[("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"inter": precip64}),
               ("aggregate", {"calc": "sum", "groupby": "month"})],
        hw=[("transform", {"inter": hourly_w}), ("aggregate", {"calc": "max", "groupby": "date"}),
            ("aggregate", {"calc": "mean", "groupby": "month"})],
        two=[("aggregate", {"calc": "mean", "groupby": "month"}), ("transform", {"transform": "power", "exp": [[2]]}),
             ("transform", {"transform": "spline"})])
    for stripes in (1, 4):
        engine.OPTIONS["target_stripes"] = stripes
        got, want = _both_time(arr, t, lat, lon, spec)
        for k in want:
            assert got[k].values.dtype == want[k][0].dtype, (k, got[k].values.dtype, want[k][0].dtype)
            _close(got[k].values, want[k][0], 1e-6 if got[k].values.dtype == np.float32 else 1e-12)
    # and end to end (the streamed host path runs the elementwise stages after the last chunk)
    wdf, shp = _weights_case(lat, lon, np.random.default_rng(1))
    want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(24), shp, "geoid", "nan"), orc.ODataset(arr, t, lat, lon, True),
                                 aggregator_dict=spec)
    ds = af.Dataset.from_arrays(arr, t, lat, lon, True)
    w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
    w.weights = wdf
    got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    _close(got[vals].values, want[vals].values, 1e-6)


def test_transform_of_the_raster_alone_and_shape_mismatch():
    arr, t, lat, lon = _raster("float32", True, T=30, Y=2, X=3, seed=33)
    got, want = _both_time(arr, t, lat, lon, dict(sq=[("transform", {"transform": "power", "exp": np.arange(2, 3)})]))
    assert got["sq_2"].values.dtype == np.float64
    _exact(got["sq_2"].values, want["sq_2"][0])
    with pytest.raises(AssertionError):
        af.aggregate_time(dataset=af.Dataset.from_arrays(arr, t, lat, lon, True), weights=None,
                          bad=[("transform", {"inter": np.ones((7, 2, 3), np.float32)})])


def test_dataset_transform_methods_match_numpy():
    arr, t, lat, lon = _raster("float32", True, T=50, Y=3, X=4, seed=41)
    ds = af.Dataset.from_arrays(arr, t, lat, lon, True, name="t2m")
    p2 = ds.power(np.int64(2))
    assert p2.values.dtype == np.float64 and p2.history == ["power2"]
    _exact(p2.values, np.power(arr, np.int64(2)))
    assert ds.power(2).values.dtype == np.float32                       # python int: stays float32
    other = np.random.default_rng(2).random(arr.shape).astype(np.float32)
    _exact(ds.interact(other).values, np.multiply(arr, other))
    s1, s2 = ds.spline()
    _exact(s1.values, arr)
    _exact(s2.values, (arr > 20) * (arr - 20))
    ds2 = ds.deepcopy()
    assert ds2.power(np.int64(2), update=True) is None and ds2.values.dtype == np.float64


@pytest.mark.parametrize("step_h", [3, 6])
@pytest.mark.parametrize("name", ["c1_tavg_poly", "c3_bins_and_poly", "c2_gdd", "area_example", "c3b_daily"])
def test_three_and_six_hourly_rasters_uniform_groups_of_8_and_4(name, step_h):
    n = (24 // step_h) * 70
    rng = np.random.default_rng(step_h)
    t = pd.date_range("2001-02-10", periods=n, freq=f"{step_h}h")
    arr = (11 + 9 * np.sin(np.arange(n) / 17.0)[:, None, None] + rng.normal(0, 6, (n, 4, 9))).astype(np.float32)
    arr[rng.random(arr.shape) < 0.002] = np.nan
    lat, lon = np.linspace(45, 42, 4), np.linspace(250, 258, 9)
    for stripes in (1, 5):
        engine.OPTIONS["target_stripes"] = stripes
        got, want = _both_time(arr, t, lat, lon, SPECS[name])
        for k in want:
            if stripes == 1 or "bins" in k:
                _exact(got[k].values, want[k][0])
            else:
                _close(got[k].values, want[k][0], 1e-12)


def test_preprocess_reaches_transforms_that_read_the_raster_directly():
    from aggfly_b200 import preprocess as pp
    arr, t, lat, lon = _raster("float32", True, T=24 * 20, Y=3, X=8, seed=43)
    raw = arr + np.float32(273.15)
    spec = dict(poly=[("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                      ("aggregate", {"calc": "mean", "groupby": "date"})],
                plain=[("aggregate", {"calc": "mean", "groupby": "date"})])
    want = orc.aggregate_time(orc.ODataset(pp.resolve("kelvin_to_celsius")(raw), t, lat, lon, True), spec)
    got = af.aggregate_time(dataset=af.Dataset.from_arrays(raw, t, lat, lon, True, preprocess="kelvin_to_celsius"),
                            weights=None, aggregator_dict=spec)
    for k in want:
        _exact(got[k].values, want[k][0])
    none = af.aggregate_dataset(weights=_ref_weights(_ref_dataset()), dataset=_ref_dataset())     # no spec: raw series
    assert list(none.columns) == ["geoid", "time", "variable"] and len(none) == 4
