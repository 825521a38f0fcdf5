"""Host logic without a GPU: DSL -> symbolic series -> programs, key/dtype rules, error
behaviour (reference: aggfly/aggregate/aggregate.py:36-162, 285-303; temporal.py:57-163)."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

from aggfly_b200 import _lib
from aggfly_b200.spec import Graph, Planner, TemporalAggregator, build_desc, compile_spec
from tests import refcases as rc

T_HOURLY = pd.date_range("2001-01-01", periods=24 * 90, freq="h")


def _plan(spec, dtype=np.float32, t=T_HOURLY):
    g = Graph(dtype, t)
    outs = compile_spec(g, spec)
    return g, outs, Planner(g).plan(list(outs.values()))


def test_keys_follow_reference_naming():
    g, outs, _ = _plan(rc.golden_time_spec())
    assert list(outs) == ["bins_-99_20", "bins_20_99", "cooling_dday", "tavg_1", "tavg_2"]
    _, outs, _ = _plan(dict(t=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("transform", {"transform": "spline"}),
                               ("aggregate", {"calc": "sum", "groupby": "year"})]))
    assert list(outs) == ["t_spline1", "t_spline2"]
    _, outs, _ = _plan(dict(b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [[0.5, 10, 0], [10, 20.25, 0]]})]))
    assert list(outs) == ["b_0.5_10", "b_10_20.25"]


def test_power_dtype_promotion_follows_numpy():
    g, outs, _ = _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("transform", {"transform": "power", "exp": np.arange(1, 3)})]))
    assert all(n.dtype == np.float64 for n in outs.values())            # numpy ints promote f32 -> f64
    g, outs, _ = _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("transform", {"transform": "power", "exp": [[1, 2]]})]))
    assert all(n.dtype == np.float32 for n in outs.values())            # python ints do not
    g, outs, _ = _plan(dict(a=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                               ("aggregate", {"calc": "sum", "groupby": "year"})]))
    assert outs["a"].dtype == np.float32                                # stays f32 end to end


def test_shared_prefix_is_one_lane_and_one_program():
    g, outs, stage = _plan(dict(
        temp_bins=[("aggregate", {"calc": "mean", "groupby": "date"}),
                   ("aggregate", {"calc": "bins", "groupby": "year",
                                  "ddargs": [[-20 + 5 * i, -15 + 5 * i, 0] for i in range(13)]})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
              ("aggregate", {"calc": "sum", "groupby": "year"})]))
    assert len(stage.programs) == 1
    p = stage.programs[0]
    assert len(p.lanes) == 1 and len(p.slots) == 15 and len(p.cols) == 15 and p.two_level
    assert stage.dtype == np.float64
    assert [c.out_col for c in p.cols] == list(range(15))
    desc, keep = build_desc(p, stage.dtype)
    n_str, n_rec, kl, ks, kd = (C.c_int32() for _ in range(5))
    _lib.check(_lib.lib().agf_program_plan(C.byref(desc), 1038240, 0, 148, None, 0, C.byref(n_str), C.byref(n_rec),
                                           C.byref(kl), C.byref(ks), C.byref(kd)))
    assert (kl.value, ks.value, kd.value) == (1, 20, 0)       # 16 bin counters + 4 power sums (typed slots)


def test_hourly_bins_plus_mean_is_single_level_diag():
    g, outs, stage = _plan(dict(
        hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [[i, i + 5, 0] for i in range(0, 65, 5)]})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"})]))
    assert len(stage.programs) == 1 and not stage.programs[0].two_level
    assert len(stage.programs[0].lanes) == 14 and stage.dtype == np.float32


def test_mixed_depth_outputs_share_one_stage():
    g, outs, stage = _plan(dict(
        a=[("aggregate", {"calc": "mean", "groupby": "month"})],
        b=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "max", "groupby": "month"})]))
    assert len(stage.programs) == 2
    assert sorted(c.out_col for p in stage.programs for c in p.cols) == [0, 1]


def test_three_levels_materialise_the_inner_series():
    g, outs, stage = _plan(dict(
        a=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "sum", "groupby": "month"}),
           ("aggregate", {"calc": "max", "groupby": "year"})]))
    assert len(stage.inputs) == 1 and stage.inputs[0].programs[0].two_level
    assert not stage.programs[0].two_level and stage.programs[0].input is not g.raw


def test_unfusable_level2_calc_goes_multi_pass():
    g, outs, stage = _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}),
                                   ("aggregate", {"calc": "nanmean", "groupby": "month"})]))
    assert len(stage.inputs) == 1 and not stage.programs[0].two_level


def test_many_lanes_split_into_several_programs():
    dd = [[i, i + 1, 0] for i in range(40)]
    g, outs, stage = _plan(dict(b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": dd}),
                                   ("aggregate", {"calc": "sum", "groupby": "year"})]))
    # a sum of bin counts over non-empty inner groups is the bin count over the composed groups: single-level programs
    # with typed bin lanes, at most 28 of them each
    assert len(outs) == 40 and [len(p.lanes) for p in stage.programs] == [28, 12]
    assert all(not p.two_level and p.input is g.raw for p in stage.programs)
    assert sorted(c.out_col for p in stage.programs for c in p.cols) == list(range(40))
    # ... but not when an inner group is empty (its bins are NaN and poison the outer sum): the two-level form stays
    g2, outs2, stage2 = _plan(dict(b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": dd}),
                                      ("aggregate", {"calc": "sum", "groupby": "year"})]), t=T_HOURLY.delete(slice(48, 72)))       # a whole day missing: an empty date group
    assert len(outs2) == 40 and len(stage2.programs) == 3 and all(p.two_level for p in stage2.programs)   # 16 + 16 + 8


def test_degree_day_lanes_stay_within_the_four_lane_kernels():
    """Six dd thresholds: two two-level programs of <= 4 lanes (the specialised kernels), not one sixteen-lane diagonal
    program on the general ragged-group kernel; a mean lane shares a program with dd lanes up to four in total."""
    dd = [[0, 10, 0], [10, 20, 0], [20, 30, 0], [30, 99, 0], [-99, 0, 1], [10, 30, 0]]
    g, outs, stage = _plan(dict(dd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": dd}),
                                    ("aggregate", {"calc": "sum", "groupby": "year"})],
                                tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                                      ("aggregate", {"calc": "mean", "groupby": "year"})]))
    assert len(outs) == 7 and all(p.two_level for p in stage.programs)
    assert [len(p.lanes) for p in stage.programs] == [4, 3]
    assert sorted(c.out_col for p in stage.programs for c in p.cols) == list(range(7))
    # bins that feed a MEAN keep the diagonal form (the mean of daily counts is not a count)
    g, outs, stage = _plan(dict(b=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": dd}),
                                   ("aggregate", {"calc": "mean", "groupby": "year"})]))
    assert len(stage.programs) == 1 and stage.programs[0].two_level and len(stage.programs[0].lanes) == 6


def test_errors_match_reference():
    with pytest.raises(ValueError, match="multiple ddargs"):
        _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}),
                      ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                      ("aggregate", {"calc": "bins", "groupby": "month", "ddargs": [[0, 1, 0], [1, 2, 0]]})]))
    with pytest.raises(ValueError, match="No valid transform"):
        _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"transform": "log"})]))
    with pytest.raises(KeyError):
        _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "decade"})]))
    with pytest.raises(ValueError, match="unsupported calc"):
        _plan(dict(a=[("aggregate", {"calc": "median", "groupby": "date"})]))
    with pytest.raises(NotImplementedError, match="week"):
        from aggfly_b200.timeaxis import CalendarIndex
        g = Graph(np.float64, CalendarIndex.range("360_day", 2000, 60))
        outs = compile_spec(g, dict(v=[("aggregate", {"calc": "mean", "groupby": "week"})]))
        Planner(g).plan(list(outs.values()))


def test_temporal_aggregator_attributes():
    a = TemporalAggregator("dd", "date", ddargs=[10, 30, 0])
    assert (a.calc, a.groupby, a.multi_dd, a.kwargs) == ("dd", "1D", False, {"ddargs": [10, 30, 0]})
    assert TemporalAggregator("bins", "month", ddargs=[[0, 1, 0], [1, 2, 0]]).multi_dd


def test_single_row_inner_groups_collapse_to_one_level():
    """Daily data grouped by date: the inner step is per-value, so dd/date -> sum/month is ONE
    single-level program over the composed (month) bounds -- no partial records, no finalize."""
    from aggfly_b200.timeaxis import CalendarIndex
    t = CalendarIndex.range("noleap", 1950, 365 * 3)
    spec = dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                     ("aggregate", {"calc": "sum", "groupby": "month"})],
                tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "month"})],
                hot=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [30, 99, 0]}),
                     ("aggregate", {"calc": "sum", "groupby": "month"})])
    g, outs, stage = _plan(spec, t=t)
    assert len(stage.programs) == 1 and not stage.programs[0].two_level and not stage.inputs
    p = stage.programs[0]
    assert [(l.calc, l.dd) for l in p.lanes] == [("dd_r", (10.0, 30.0, 0.0)), ("mean", None), ("bins", (30.0, 99.0, 0.0))]
    assert len(p.bounds1) == 37 and list(p.bounds1[:4]) == [0, 31, 59, 90] and p.bounds1[-1] == 365 * 3
    # hourly data: groups of 24 rows do not collapse -- except the sum of bin counts, which is the bin count over the
    # month whatever the inner group length (typed bin lanes in a pass of their own)
    _, _, stage = _plan(spec)
    assert sorted(q.two_level for q in stage.programs) == [False, True]
    single = [q for q in stage.programs if not q.two_level][0]
    assert [(l.calc, l.dd) for l in single.lanes] == [("bins", (30.0, 99.0, 0.0))] and len(single.bounds1) == 4
    # a transform between the steps keeps the two-level form (the power applies to the daily value)
    _, _, stage = _plan(dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}),
                                ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                                ("aggregate", {"calc": "sum", "groupby": "month"})]), t=t)
    assert all(q.two_level for q in stage.programs)
    # dd -> mean is not a plain sum of rounded terms: stays two-level
    _, _, stage = _plan(dict(a=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                                ("aggregate", {"calc": "mean", "groupby": "month"})]), t=t)
    assert all(q.two_level for q in stage.programs)


def test_csr_lowering_is_cached_on_disk_by_content(tmp_path):
    from aggfly_b200 import weights as W
    wdf = pd.DataFrame({"cell_id": [3, 1, 0, 2, 9], "index_right": [7, 7, 2, 2, 2], "weight": [0.5, 0.25, 1.0, 2.0, 3.0]})
    lon_order = np.array([1, 0])
    a = W.lower_to_csr_cached(wdf, np.arange(4), 2, 2, lon_order, str(tmp_path))
    files = list(tmp_path.rglob("*.npz"))
    assert len(files) == 1 and files[0].parent.name.startswith("mod-") and files[0].parent.parent.name == "DeviceCSR"
    b = W.lower_to_csr_cached(wdf, np.arange(4), 2, 2, lon_order, str(tmp_path))       # served from disk
    ref = W.lower_to_csr(wdf, np.arange(4), 2, 2, lon_order)
    for x in (a, b):
        assert np.array_equal(x.row_ptr, ref.row_ptr) and np.array_equal(x.cell_idx, ref.cell_idx)
        assert np.array_equal(x.w, ref.w) and np.array_equal(x.region_ids, ref.region_ids) and x.n_cells == 4
    assert list(ref.region_ids) == [2, 7] and list(ref.row_ptr) == [0, 2, 4]           # cell 9 is off-grid: dropped
    assert list(ref.cell_idx) == [1, 3, 2, 0]                                          # columns swapped by lon_order
    wdf2 = wdf.assign(weight=wdf["weight"] * 2)
    W.lower_to_csr_cached(wdf2, np.arange(4), 2, 2, lon_order, str(tmp_path))
    assert len(list(tmp_path.rglob("*.npz"))) == 2                                     # different content, new key


def test_time_sel_selects_like_pandas_partial_string_indexing():
    import aggfly_b200 as af
    from aggfly_b200.timeaxis import CalendarIndex
    t = pd.date_range("2000-11-01", "2002-02-28 23:00", freq="h")
    arr = np.arange(len(t), dtype=np.float32)[:, None, None] * np.ones((1, 2, 2), np.float32)
    ra = af.RasterArray(arr, ("time", "latitude", "longitude"), {"time": t, "latitude": [1.0, 0.0], "longitude": [0.0, 1.0]})
    ds = af.Dataset(ra, time_sel="2001")
    assert len(ds.time) == 8760 and ds.time[0] == pd.Timestamp("2001-01-01") and ds.values.shape[0] == 8760
    assert float(ds.values[0, 0, 0]) == float(t.get_loc("2001-01-01 00:00"))
    assert len(af.Dataset(ra, time_sel="2001-06").time) == 720
    assert len(af.Dataset(ra, time_sel=slice("2001-01-01", "2001-01-02")).time) == 48          # both ends inclusive
    from aggfly_b200.dataset import time_selection
    assert time_selection(CalendarIndex.range("noleap", 2000, 365 * 3), "2001") == (365, 730)


def test_dask_client_helpers_are_warn_once_noops():
    """aggfly/__init__.py:1-11 exports start_dask_client / shutdown_dask_client / is_distributed /
    distributed_client (aggregate_utils.py:9-102); scripts that call them must keep running."""
    import warnings
    import aggfly_b200 as af
    from aggfly_b200 import aggregate as agg
    agg._SHIM_WARNED.clear()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert af.start_dask_client(n_workers=4, threads_per_worker=1, processes=False) is None
        assert af.start_dask_client() is None                                  # second call: silent
        assert af.is_distributed() is False and af.distributed_client() is None
        assert af.shutdown_dask_client() == {"n_workers": 2, "threads_per_worker": 2, "cap_numba_threads": 1}
        assert af.shutdown_dask_client() is None
    assert sorted(str(w.message).split("(")[0] for w in rec) == [
        "aggfly_b200.distributed_client", "aggfly_b200.is_distributed", "aggfly_b200.shutdown_dask_client",
        "aggfly_b200.start_dask_client"]


def test_specs_with_opaque_values_are_never_plan_cached():
    """A plan captures an ``inter`` array by value; keying the cache on id() would hand back a stale plan."""
    from aggfly_b200.aggregate import _spec_fingerprint
    small = {"a": [("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"transform": "power", "exp": np.arange(1, 3)})]}
    assert _spec_fingerprint(small) is not None and _spec_fingerprint(small) == _spec_fingerprint(
        {"a": [("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"transform": "power", "exp": np.arange(1, 3)})]})
    big = {"a": [("aggregate", {"calc": "mean", "groupby": "date"}), ("transform", {"inter": np.zeros((3, 5, 5))})]}
    assert _spec_fingerprint(big) is None
    assert _spec_fingerprint({"a": [("transform", {"inter": object()})]}) is None


def test_csr_disk_cache_round_trips_string_region_ids(tmp_path):
    """np.savez pickles object arrays and np.load refuses them: string ids are stored as unicode and come back as objects."""
    import pandas as pd
    from aggfly_b200.weights import lower_to_csr_cached
    wdf = pd.DataFrame({"cell_id": [0, 1, 2, 3], "index_right": ["b", "b", "a", "c"], "weight": [0.5, 0.5, 1.0, 1.0]})
    first = lower_to_csr_cached(wdf, np.arange(6), 2, 3, None, str(tmp_path))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")                         # an unreadable entry would warn
        again = lower_to_csr_cached(wdf, np.arange(6), 2, 3, None, str(tmp_path))
    assert list(again.region_ids) == list(first.region_ids) == ["a", "b", "c"]
    assert np.array_equal(again.row_ptr, first.row_ptr) and np.array_equal(again.cell_idx, first.cell_idx)
