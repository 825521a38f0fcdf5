"""Time sharding (aggfly_b200/shard.py): plan, gather and panel assembly on CPU with the gloo
backend, world_size 2 -- the host logic of the multi-GPU path.  The per-rank compute (CUDA in the
product) is stood in for by the oracle here, which is what lets the assembled panel be compared
with one oracle call over the whole record."""
import os
import socket

import numpy as np
import pandas as pd
import pytest

from aggfly_b200 import shard
from aggfly_b200.timeaxis import CalendarIndex


def test_balanced_ranges_and_plan_cut_at_year_boundaries():
    assert shard.balanced_ranges(5, 2) == [(0, 3), (3, 5)]
    assert shard.balanced_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    t = pd.date_range("2001-01-01", "2004-12-31 23:00", freq="h")              # 2004 is a leap year
    shards = shard.plan_time_shards(t, 2)
    assert shards == [(0, 2 * 8760), (2 * 8760, len(t))]
    shards = shard.plan_time_shards(t, 4)
    assert [b - a for a, b in shards] == [8760, 8760, 8760, 8784]
    shards = shard.plan_time_shards(t, 8)                                       # more ranks than years
    assert sum(b - a for a, b in shards) == len(t) and sum(b > a for a, b in shards) == 4
    assert all(t[a].dayofyear == 1 and t[a].hour == 0 for a, b in shards if b > a)
    # partial first year, month shards, noleap calendar
    t2 = pd.date_range("2001-11-15", "2002-03-10", freq="D")
    for a, b in shard.plan_time_shards(t2, 3, "month"):
        assert a == 0 or t2[a].day == 1
    tn = CalendarIndex.range("noleap", 1950, 365 * 6)
    assert shard.plan_time_shards(tn, 3) == [(0, 730), (730, 1460), (1460, 2190)]


def test_check_shardable_rejects_weeks_and_coarser_groups():
    ok = dict(a=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "sum", "groupby": "year"})])
    shard.check_shardable(ok, "year")
    with pytest.raises(ValueError, match="week"):
        shard.check_shardable(dict(a=[("aggregate", {"calc": "sum", "groupby": "week"})]), "year")
    with pytest.raises(ValueError, match="coarser"):
        shard.check_shardable(ok, "month")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aggfly_b200 as af
        from aggfly_b200 import aggregate as agg, engine
        from oracle import oracle as orc

        rng = np.random.default_rng(5)                                 # same data on every rank
        t = pd.date_range("2001-06-01", "2003-09-30 23:00", freq="6h")
        arr = (15 + rng.normal(0, 8, (len(t), 3, 4))).astype(np.float32)
        arr[5:40, 1, 1] = np.nan
        lat, lon = np.array([40.0, 39.0, 38.0]), np.array([250.0, 251.0, 252.0, 253.0])
        wdf = pd.DataFrame({"cell_id": [0, 1, 5, 6, 7, 11, 2], "index_right": [4, 4, 4, 9, 9, 9, 2],
                            "weight": [0.2, 0.3, 0.5, 1.0, 2.0, 0.5, 0.0]})
        shp = pd.DataFrame({"geoid": ["z", "a", "b"]}, index=[2, 4, 9])
        spec = dict(hot=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [20, 99, 0]}),
                         ("aggregate", {"calc": "sum", "groupby": "month"})],
                    tavg=[("aggregate", {"calc": "mean", "groupby": "month"})])
        ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360=True)
        w = af.GridWeights.from_frame(wdf, af.Grid(af.lon_to_180(lon), lat), af.GeoRegions(shp, "geoid"), zero_weight="nan")

        # stand-ins for the CUDA stages: oracle temporal + oracle scatter, same tensor contract
        class _Res:
            pass

        def fake_temporal(sub, spec_):
            out = orc.aggregate_time(orc.ODataset(np.asarray(sub.values), sub.time, sub.latitude, sub.longitude, True), spec_)
            res = _Res()
            res.X = np.stack([out[k][0].reshape(out[k][0].shape[0], -1) for k in out], axis=2)      # [G, cells, NC]
            res.labels = list(out.values())[0][1]
            return list(out), res, None

        class _Csr:
            host = af.lower_to_csr(wdf, w.grid.cell_id, 3, 4, ds.lon_sort_order())

        def fake_spmm(csr, res):
            h = csr.host
            G, _, NC = res.X.shape
            valid = ~np.isnan(res.X).any(axis=2)                                                    # [G, cells]
            panel = np.full((h.n_regions, G, NC), np.nan)
            for r in range(h.n_regions):
                e = slice(h.row_ptr[r], h.row_ptr[r + 1])
                cells, ww = h.cell_idx[e], h.w[e]
                for g in range(G):
                    m = valid[g, cells]
                    den = (ww * m).sum()
                    if den != 0:
                        panel[r, g] = (np.where(m[:, None], res.X[g][cells, :], 0.0) * ww[:, None]).sum(axis=0) / den
            return torch.from_numpy(panel)

        agg._temporal_device = fake_temporal
        agg._device_csr = lambda weights, dataset: _Csr
        engine.run_spmm = fake_spmm
        torch.cuda.current_device = lambda: 0
        real_empty = torch.empty
        torch.empty = lambda *a, **k: real_empty(*a, **{**k, "device": "cpu"}) if "device" in k else real_empty(*a, **k)

        got = shard.aggregate_dataset_sharded(w, ds, spec, shard_by="year")
        want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(12), shp, "geoid", "nan"),
                                     orc.ODataset(arr, t, lat, lon, True), aggregator_dict=spec)
        assert list(got.columns) == list(want.columns), (list(got.columns), list(want.columns))
        assert len(got) == len(want) and (got["geoid"].values == want["geoid"].values).all()
        assert (got["time"].values == want["time"].values).all()
        for c in ("hot", "tavg"):
            assert np.allclose(got[c].values, want[c].values, rtol=1e-12, equal_nan=True), c
        # the raw collective: ragged period counts per rank
        local = torch.full((2, 3 if rank == 0 else 1, 2), float(rank + 1), dtype=torch.float64)
        full = shard.gather_panels(local, [3, 1])
        assert full.shape == (2, 4, 2) and full[:, :3].eq(1).all() and full[:, 3:].eq(2).all()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_panel_equals_single_call_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
