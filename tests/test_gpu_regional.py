"""The one-kernel temporal + regional path (agf_temporal_regional_run, csrc/agf_regional.cuh) against the two-kernel path
(agf_temporal_run -> X -> agf_spmm_run) and against the CPU oracle.

Parity bar: identical rows, region ids and time labels; every value rel 1e-12 against the two-kernel path and rel 1e-11
against the oracle (north_star: 1e-5) -- the same terms w * x are added, in a different but FIXED association (two
interleaved phases per tile slot, then tiles in ascending order), so results are bit-identical from run to run and from
launch split to launch split."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

import aggfly_b200 as af
from aggfly_b200 import aggregate as agg_mod, engine, synthetic as syn
from oracle import oracle as orc

BINS13 = syn.BINS13
SPECS = {
    "daily_bins_mean": syn.SPECS["daily_bins_mean"],                                              # 13 bin lanes + mean (typed, LPS 8)
    "tavg": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})]),                     # one lane
    "tsum": dict(tsum=[("aggregate", {"calc": "sum", "groupby": "date"})]),
    "gdd": dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]})]),   # degree days per date
    "tavg_gdd": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],                  # mixed mean + dd lanes
                     gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]})]),
    "tavg_3dd": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                     dd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [[0, 10, 0], [10, 20, 0], [20, 99, 1]]})]),
    "bins6_mean_sum": dict(b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13[3:9]})],
                           tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                           tsum=[("aggregate", {"calc": "sum", "groupby": "date"})]),
    "bins14_two_sums": dict(b=[("aggregate", {"calc": "bins", "groupby": "date",
                                               "ddargs": [[-99, -20, 0]] + BINS13})],
                            tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                            tsum=[("aggregate", {"calc": "sum", "groupby": "date"})]),
    "tmin_tmax_tavg": dict(tmax=[("aggregate", {"calc": "max", "groupby": "date"})],             # fixed min / max / mean layout
                           tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                           tmin=[("aggregate", {"calc": "min", "groupby": "date"})]),
    "tavg_poly": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),                   # columns transformed from one lane
                            ("transform", {"transform": "power", "exp": np.arange(1, 4)})]),
}


@pytest.fixture(autouse=True)
def _gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    saved = dict(engine.OPTIONS)
    yield
    engine.OPTIONS.clear()
    engine.OPTIONS.update(saved)


def _case(n_lat, n_lon, days, seed, ocean=0.2, regions=(9, 5), zero_weight="nan"):
    rng = np.random.default_rng(seed)
    T = 24 * days
    t = pd.date_range("2001-03-01", periods=T, freq="h")
    lat = 49.75 - 0.25 * np.arange(n_lat)
    lon = 235.0 + 0.25 * np.arange(n_lon)
    arr = (12 + 14 * np.sin(np.arange(T) / 24.0 * 0.7)[:, None, None] + 5 * np.sin(2 * np.pi * (np.arange(T) % 24) / 24)[:, None, None]
           + rng.normal(0, 4, (T, n_lat, n_lon))).astype(np.float32)
    if ocean:
        tiles = rng.random(((n_lat + 3) // 4, (n_lon + 3) // 4)) < ocean
        mask = np.repeat(np.repeat(tiles, 4, 0), 4, 1)[:n_lat, :n_lon]
        arr[:, mask] = np.nan
    arr[30:33, n_lat // 2, n_lon // 2] = np.nan                    # a cell that is invalid on ONE day only
    gd = syn.GridDef(lat, lon, True, (regions[0], regions[1], -125.125, -125.125 + 0.25 * n_lon, 49.875 - 0.25 * n_lat, 49.875))
    ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360=True)
    w = af.weights_from_objects(ds, syn.tessellation(gd), zero_weight=zero_weight)
    w.calculate_weights()
    return arr, t, lat, lon, ds, w


def _frames(ds, w, spec, device=True):
    import torch
    d = ds
    if device:
        d = af.Dataset.from_arrays(torch.from_numpy(np.asarray(ds.values)).cuda(), ds.time, ds.latitude, ds.longitude,
                                   lon_is_360=ds.lon_is_360)
    engine.OPTIONS["regional"] = False
    two = af.aggregate_dataset(weights=w, dataset=d, aggregator_dict=spec)
    assert "spmm (issue)" in agg_mod.LAST_TRACE["phases_ms"] or "spmm + d2h" in agg_mod.LAST_TRACE["phases_ms"]
    engine.OPTIONS["regional"] = True
    one = af.aggregate_dataset(weights=w, dataset=d, aggregator_dict=spec)
    assert any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"]), agg_mod.LAST_TRACE
    return one, two


def _same_frame(a, b, rtol):
    assert list(a.columns) == list(b.columns) and len(a) == len(b) > 0
    rid = a.columns[0]
    assert np.array_equal(np.asarray(a[rid]), np.asarray(b[rid]))
    assert np.array_equal(np.asarray(a["time"]).astype("datetime64[ns]"), np.asarray(b["time"]).astype("datetime64[ns]"))
    cols = [c for c in a.columns if c not in (rid, "time")]
    x, y = a[cols].to_numpy(float), b[cols].to_numpy(float)
    assert np.array_equal(np.isnan(x), np.isnan(y))
    ok = ~np.isnan(y)
    # relative to the value, with a floor for averages that cancel to nearly zero (terms of magnitude ~30)
    err = np.abs(x[ok] - y[ok]) / np.maximum(np.abs(y[ok]), 1e-2)
    assert float(err.max(initial=0.0)) <= rtol, float(err.max())


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("shape", [(40, 64), (21, 100), (13, 36)])
def test_one_kernel_path_matches_two_kernel_path_and_oracle(name, shape):
    spec = SPECS[name]
    arr, t, lat, lon, ds, w = _case(shape[0], shape[1], days=9, seed=shape[0] + len(name))
    one, two = _frames(ds, w, spec)
    _same_frame(one, two, 1e-12)
    rid = w.georegions.regionid
    want = orc.aggregate_dataset(orc.OWeights(w.weights, w.grid.cell_id, w.georegions.shp, rid, w.zero_weight),
                                 orc.ODataset(arr, t, lat, lon, True), aggregator_dict=spec)
    _same_frame(one.reset_index(drop=True), want.reset_index(drop=True), 1e-11)


def _adversarial_bits(rng, shape):
    """float32 bit patterns within +-70000 ulps of the 14 edges of BINS13 (0.0 is one of them: +-0, denormals of both
    signs): every value sits next to a threshold, on either side of the bfloat16 grid points around it."""
    edges = np.array([b[0] for b in BINS13] + [BINS13[-1][1]], dtype=np.float32)
    e = edges[rng.integers(0, len(edges), shape)]
    k = rng.integers(-70000, 70001, shape)
    k[rng.random(shape) < 0.3] //= 4096                      # a third of them within a few ulps
    bits = e.view(np.uint32).astype(np.int64)
    mag, sign = bits & 0x7FFFFFFF, bits >> 31
    zero = mag == 0
    sign = np.where(zero, rng.integers(0, 2, shape), sign)   # around 0.0 the offset is a magnitude, the sign is drawn
    mag = np.where(zero, np.abs(k), np.maximum(mag + k, 0))
    return ((sign << 31) | mag).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("mode", ["near_edges", "with_equal_values", "nan_payloads"])
def test_packed_edge_counting_is_exact_next_to_the_thresholds(mode):
    """The bfloat16-packed edge compares of the daily-bins kernel (rg_count_above_packed) against the two-kernel path and
    the oracle on values chosen to break a truncating compare: a miscounted value moves a bin's average by >= 1e-3."""
    rng = np.random.default_rng(77)
    n_lat, n_lon, days = 24, 64, 6
    arr, t, lat, lon, ds, w = _case(n_lat, n_lon, days=days, seed=5, ocean=0.0)
    vals = _adversarial_bits(rng, arr.shape)
    low = vals.view(np.uint32) & 0xFFFF
    keep = np.zeros(vals.shape, dtype=bool)
    if mode == "with_equal_values":
        keep[:, :8, :] = True                                 # values equal to an edge (slow path) in the first tile row only
    vals = np.where((low == 0) & ~keep, np.float32(7.3), vals)   # elsewhere nothing the equality screen would catch
    if mode == "nan_payloads":
        for (y, x, pat) in [(3, 5, 0x7F800001), (3, 6, 0xFF800001), (9, 40, 0x7FFFFFFF), (10, 41, 0x7FC00000), (11, 1, 0xFFC00000)]:
            vals[:, y, x] = np.array([pat], dtype=np.uint32).view(np.float32)[0]     # all-NaN cells (fast path)
        vals[24:30, 15, 20] = np.nan                                                  # a partly-NaN period (slow path)
    ds = af.Dataset.from_arrays(vals, t, lat, lon, lon_is_360=True)
    spec = SPECS["daily_bins_mean"]
    one, two = _frames(ds, w, spec)
    _same_frame(one, two, 1e-12)
    rid = w.georegions.regionid
    want = orc.aggregate_dataset(orc.OWeights(w.weights, w.grid.cell_id, w.georegions.shp, rid, w.zero_weight),
                                 orc.ODataset(vals, t, lat, lon, True), aggregator_dict=spec)
    _same_frame(one.reset_index(drop=True), want.reset_index(drop=True), 1e-11)


@pytest.mark.parametrize("groupby", ["date", "month"])
def test_edge_form_of_the_two_kernel_path_is_exact_next_to_the_thresholds(groupby):
    """The same adversarial values through the temporal kernels of the two-kernel path (l1_acc_group's edge form: the
    per-date kernel for ``date``, the ragged-group kernel with its tile / eight-row batches and per-value tails for
    ``month``) against the oracle."""
    rng = np.random.default_rng(78)
    n_lat, n_lon, days = 16, 64, 40                                # 1 March .. 9 April: two ragged month groups
    arr, t, lat, lon, ds, w = _case(n_lat, n_lon, days=days, seed=6, ocean=0.0)
    vals = _adversarial_bits(rng, arr.shape)
    low = vals.view(np.uint32) & 0xFFFF
    keep = np.zeros(vals.shape, dtype=bool)
    keep[:, :4, :] = True                                          # values equal to an edge in the first rows only
    vals = np.where((low == 0) & ~keep, np.float32(7.3), vals)
    for (y, x, pat) in [(8, 5, 0x7F800001), (8, 6, 0xFF800001), (9, 40, 0x7FFFFFFF), (10, 41, 0x7FC00000)]:
        vals[:, y, x] = np.array([pat], dtype=np.uint32).view(np.float32)[0]          # all-NaN cells
    vals[24:30, 12, 20] = np.nan                                                       # partly-NaN batches
    vals[100, 13, 21] = np.nan
    ds = af.Dataset.from_arrays(vals, t, lat, lon, lon_is_360=True)
    spec = dict(hbins=[("aggregate", {"calc": "bins", "groupby": groupby, "ddargs": BINS13})],
                tavg=[("aggregate", {"calc": "mean", "groupby": groupby})])
    engine.OPTIONS["regional"] = False
    import torch
    d = af.Dataset.from_arrays(torch.from_numpy(vals).cuda(), t, lat, lon, lon_is_360=True)
    got = af.aggregate_dataset(weights=w, dataset=d, aggregator_dict=spec)
    rid = w.georegions.regionid
    want = orc.aggregate_dataset(orc.OWeights(w.weights, w.grid.cell_id, w.georegions.shp, rid, w.zero_weight),
                                 orc.ODataset(vals, t, lat, lon, True), aggregator_dict=spec)
    _same_frame(got.reset_index(drop=True), want.reset_index(drop=True), 1e-11)


@pytest.mark.parametrize("zero_weight", ["nan", "area"])
def test_host_fed_raster_and_zero_weight_rows(zero_weight):
    """The streamed feed (pageable NumPy raster -> staging ring -> launches per period range) and the row-drop rules."""
    from aggfly_b200 import stream
    arr, t, lat, lon, ds, w = _case(40, 64, days=40, seed=5, ocean=0.45, regions=(6, 4), zero_weight=zero_weight)
    arr[:, :11, :12] = np.nan                     # every cell of the north-west region is ocean: its rows are dropped
    saved = dict(stream.OPTIONS)
    stream.OPTIONS.update(staging_chunk_bytes=1 << 20, chunk_bytes=1 << 20)        # many chunks -> several launches
    try:
        one, two = _frames(ds, w, SPECS["daily_bins_mean"], device=False)
    finally:
        stream.OPTIONS.clear()
        stream.OPTIONS.update(saved)
    _same_frame(one, two, 1e-12)
    assert len(one) < 24 * 40                                                       # some (region, day) rows were dropped


@pytest.mark.parametrize("name", ["daily_bins_mean", "tavg_gdd"])
def test_tiles_with_hundreds_of_tiny_regions_read_their_tables_from_global_memory(name):
    """2 560 one-cell regions on a 40 x 64 grid: 256 slots per tile, more than the kernel's shared-memory tables hold."""
    arr, t, lat, lon, ds, w = _case(40, 64, days=6, seed=21, regions=(64, 40))
    one, two = _frames(ds, w, SPECS[name])
    _same_frame(one, two, 1e-12)
    assert one["geoid"].nunique() > 1500


def test_period_ranges_launched_separately_do_not_change_the_result():
    """A streamed feed launches the kernel per period range as rows land: same bits as one launch."""
    import torch
    from aggfly_b200.aggregate import _device_csr, _plan
    arr, t, lat, lon, ds, w = _case(48, 128, days=41, seed=11, regions=(14, 9))
    dsd = af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, lon_is_360=True)
    csr = _device_csr(w, dsd)
    names, stage = _plan(dsd, SPECS["daily_bins_mean"])
    flat = dsd.values.reshape(len(t), -1)
    rr = engine.RegionalRunner(stage, csr, len(lat), len(lon), want_den=True)
    whole = rr.run(flat)
    torch.cuda.synchronize()
    p0, d0 = whole.panel.clone(), whole.den.clone()
    st = torch.cuda.current_stream()
    rr.begin_streamed(st)
    rr.panel.fill_(-7.0)
    for g0, g1 in ((0, 5), (5, 6), (6, 30), (30, 41)):
        rr._launch(flat, g0, g1, st)
    torch.cuda.synchronize()
    assert torch.equal(rr.panel.view(torch.int64), p0.view(torch.int64)) and torch.equal(rr.den, d0)
    rr.close()


def test_runner_denominators_and_repeatability():
    import torch
    from aggfly_b200.aggregate import _device_csr, _plan
    arr, t, lat, lon, ds, w = _case(48, 128, days=12, seed=3, regions=(14, 9))
    spec = SPECS["daily_bins_mean"]
    dsd = af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, lon_is_360=True)
    csr = _device_csr(w, dsd)
    names, stage = _plan(dsd, spec)
    flat = dsd.values.reshape(len(t), -1)
    rr = engine.RegionalRunner(stage, csr, len(lat), len(lon), want_den=True)
    assert rr.supported and rr.info.lanes_per_slot == 8
    res = rr.run(flat)
    torch.cuda.synchronize()
    p0, d0 = res.panel.clone(), res.den.clone()
    two = engine.StageRunner(stage, len(lat) * len(lon))
    r2 = two.run(flat)
    p2, d2 = engine.run_spmm(csr, r2, want_den=True)
    torch.cuda.synchronize()
    assert torch.equal(torch.isnan(p0), torch.isnan(p2))
    ok = ~torch.isnan(p2)
    assert float(((p0[ok] - p2[ok]).abs() / p2[ok].abs().clamp_min(1e-2)).max()) <= 1e-12
    assert float(((d0 - d2).abs() / d2.abs().clamp_min(1e-300))[d2 > 0].max()) <= 1e-12 and torch.equal(d0 == 0, d2 == 0)
    for _ in range(5):
        res = rr.run(flat)
        torch.cuda.synchronize()
        assert torch.equal(res.panel.view(torch.int64), p0.view(torch.int64)) and torch.equal(res.den, d0)
    rr.close(); two.close()


def test_unsupported_programs_fall_back_to_the_two_kernel_path():
    """Two-level chains, float64 rasters, ragged periods: no regional instantiation -> same answer through K1 + K2."""
    import torch
    arr, t, lat, lon, ds, w = _case(16, 64, days=40, seed=8)
    engine.OPTIONS["regional"] = True
    spec = dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})])
    df = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    assert len(df) > 0 and not any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"])
    ds64 = af.Dataset.from_arrays(arr.astype(np.float64), t, lat, lon, lon_is_360=True)
    df = af.aggregate_dataset(weights=w, dataset=ds64, aggregator_dict=SPECS["tavg"])
    assert len(df) > 0 and not any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"])
    # nanmean lanes have no regional instantiation: the library says so, the host falls back
    spec = dict(tnm=[("aggregate", {"calc": "nanmean", "groupby": "date"})])
    df = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    assert len(df) > 0 and not any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"])
    # ... a daily minimum alone has one (the fixed min / max / mean layout)
    spec = dict(tmin=[("aggregate", {"calc": "min", "groupby": "date"})])
    df = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    assert len(df) > 0 and any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"])


def test_full_size_daily_panel_is_repeatable_and_matches_the_two_kernel_path():
    """C3b at its BASELINE size: 721 x 1440 x 8760 -> 45 000 regions x 365 days x 14 columns."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a GPU that holds the 36.4 GB raster and the 21 GB of columns of the two-kernel path")
    from aggfly_b200.aggregate import _device_csr, _plan
    wl = syn.make_workload("c3b_global_daily")
    raster = wl.raster(torch.device("cuda", 0), seed=1218)
    ds = wl.dataset(raster)
    w = wl.weights(ds)
    csr = _device_csr(w, ds)
    names, stage = _plan(ds, wl.spec)
    flat = raster.reshape(wl.n_time, wl.n_cells)
    rr = engine.RegionalRunner(stage, csr, 721, 1440)
    assert rr.supported
    first = None
    for it in range(10):
        res = rr.run(flat)
        torch.cuda.synchronize()
        if first is None:
            first = res.panel.clone()
        else:
            assert torch.equal(res.panel.view(torch.int64), first.view(torch.int64)), f"launch {it} differs"
    rr.close()
    two = engine.StageRunner(stage, wl.n_cells)
    p2 = engine.run_spmm(csr, two.run(flat))
    torch.cuda.synchronize()
    two.close()
    assert torch.equal(torch.isnan(first), torch.isnan(p2))
    ok = ~torch.isnan(p2)
    assert float(((first[ok] - p2[ok]).abs() / p2[ok].abs().clamp_min(1e-2)).max()) <= 1e-12


@pytest.mark.parametrize("feed", ["pinned", "pageable"])
def test_ring_of_device_windows_feeds_the_one_kernel_path_bit_identically(feed):
    """A host record longer than the device budget: the one-kernel path scans it window by window (row0 of
    agf_temporal_regional_run); same association per period, so the panel is bit for bit the resident one."""
    import torch
    from aggfly_b200 import stream
    spec = SPECS["daily_bins_mean"]
    arr, t, lat, lon, ds, w = _case(21, 100, days=50, seed=77)
    engine.OPTIONS["regional"] = True
    dres = af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, lon_is_360=True)
    resident = af.aggregate_dataset(weights=w, dataset=dres, aggregator_dict=spec)
    old = dict(stream.OPTIONS)
    row = arr[0].nbytes
    try:
        stream.OPTIONS.update(chunk_bytes=31 * row, staging_chunk_bytes=31 * row, staging_slots=3, staging_threads=2,
                              device_raster_budget_bytes=1, ring_slot_bytes=24 * 17 * row, ring_slots=2)
        host = torch.from_numpy(arr).pin_memory() if feed == "pinned" else arr
        dh = af.Dataset.from_arrays(host, t, lat, lon, lon_is_360=True)
        streamed = af.aggregate_dataset(weights=w, dataset=dh, aggregator_dict=spec)
        assert any(k.startswith("temporal + regional") for k in agg_mod.LAST_TRACE["phases_ms"]), agg_mod.LAST_TRACE
        st = stream.LAST_STATS
        assert st.get("ring") and st["windows"] >= 3 and st["ring_bytes"] < arr.nbytes
    finally:
        stream.OPTIONS.update(old)
        stream.release_device_rasters()
    _same_frame(streamed, resident, 0.0)
