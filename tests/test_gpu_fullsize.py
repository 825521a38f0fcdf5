"""BASELINE.json's full-size configuration (C3: global 0.25deg hourly year, 721 x 1440 x 8760 f32 =
36.4 GB, 45 000 regions) checked through size-independent properties -- the oracle cannot run at
this size in seconds:

* a plain PyTorch float64 restatement of the chain on a random SAMPLE of cells (bit-exact bins,
  rel 1e-12 sums);
* conservation: the 13 bins plus the two open-ended bins partition the valid days of every cell;
* invariance to the time-stripe split (bins bit-exact, sums to re-association error);
* the regional average of every column lies within the range of the region's valid cells, and a
  region whose cells are all invalid is NaN;
* time sharding: the year run as two half-year shards gives the same daily panel rows.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BINS15 = [[-999.0, -20.0, 0]] + [[-20 + 5 * i, -15 + 5 * i, 0] for i in range(13)] + [[45.0, 999.0, 0]]


@pytest.fixture(scope="module")
def c3():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a GPU that holds the 36.4 GB raster")
    from aggfly_b200 import synthetic as syn
    wl = syn.make_workload("c3_global_bins")
    dev = torch.device("cuda", 0)
    raster = wl.raster(dev, seed=1218)
    ds = wl.dataset(raster)
    w = wl.weights(ds)
    yield wl, raster, ds, w
    del raster
    torch.cuda.empty_cache()


def _run(wl, raster, ds, spec, stripes=0):
    import torch
    from aggfly_b200 import engine
    from aggfly_b200.aggregate import _plan
    names, stage = _plan(ds, spec)
    runner = engine.StageRunner(stage, wl.n_cells, raster.device, target_stripes=stripes)
    res = runner.run(raster.reshape(wl.n_time, wl.n_cells))
    torch.cuda.synchronize()
    return names, res, runner


def _spec():
    return dict(
        temp_bins=[("aggregate", {"calc": "mean", "groupby": "date"}),
                   ("aggregate", {"calc": "bins", "groupby": "year", "ddargs": BINS15})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
              ("aggregate", {"calc": "sum", "groupby": "year"})])


def test_sampled_cells_match_a_float64_torch_restatement_and_bins_partition_the_year(c3):
    import torch
    wl, raster, ds, w = c3
    names, res, runner = _run(wl, raster, ds, _spec())
    X = res.X[0]                                                     # [cells, 17]
    assert names[:15] == [f"temp_bins_{a}_{b}" for a, b, _ in BINS15] and names[15:] == ["tavg_1", "tavg_2"]
    # ---- sample: daily mean = float32(sum of 24 float64 values / 24), like the reference kernel
    g = torch.Generator(device="cpu").manual_seed(5)
    idx = torch.randperm(wl.n_cells, generator=g)[:20000].to(raster.device)
    sub = raster.reshape(wl.n_time, wl.n_cells)[:, idx].double().reshape(365, 24, -1)
    s = sub[:, 0]
    for h in range(1, 24):
        s = s + sub[:, h]                                            # time-ordered fp64 accumulation
    daily = (s / 24.0).float()                                       # stored in the raster dtype
    d64 = daily.double()
    got = X[idx]
    for j, (lo, hi, _) in enumerate(BINS15):
        want = ((d64 > lo) & (d64 < hi)).sum(0).double()
        want = torch.where(torch.isnan(d64).any(0), want, want)     # NaN days simply do not count
        assert torch.equal(got[:, j], want), names[j]
    t1, t2 = d64.clone(), d64 * d64
    a1, a2 = t1[0], t2[0]
    for d in range(1, 365):
        a1, a2 = a1 + t1[d], a2 + t2[d]
    for col, want in ((15, a1), (16, a2)):
        ok = ~torch.isnan(want)
        assert torch.equal(torch.isnan(got[:, col]), ~ok)
        diff = (got[ok, col] - want[ok]).abs().max().item()
        assert torch.equal(got[ok, col], want[ok]), (names[col], diff)   # single stripe per cell: same order, same bits
    # ---- every cell: the 15 bins partition its valid days (a daily mean exactly on an edge is in no bin)
    total = X[:, :15].sum(1)
    valid_cell = ~torch.isnan(X[:, 15])
    assert bool((total[valid_cell] <= 365).all())
    assert float((total[valid_cell] == 365).double().mean()) > 0.9999
    assert bool((total[~valid_cell] == 0).all())                     # all-NaN (ocean) cells count nothing
    assert torch.equal(res.V[0].bool(), valid_cell)                  # shared validity mask
    runner.close()


def test_stripe_split_does_not_change_the_result(c3):
    import torch
    wl, raster, ds, w = c3
    _, one, r1 = _run(wl, raster, ds, _spec(), stripes=1)
    X1 = one.X.clone()
    _, many, r2 = _run(wl, raster, ds, _spec(), stripes=7)
    assert r2.programs[0].info.n_stripes == 7
    assert torch.equal(X1[..., :15], many.X[..., :15])               # counts
    a, b = X1[..., 15:], many.X[..., 15:]
    ok = ~torch.isnan(a)
    assert torch.equal(ok, ~torch.isnan(b))
    assert float(((a[ok] - b[ok]).abs() / a[ok].abs().clamp_min(1e-300)).max()) < 1e-12
    r1.close(); r2.close()


def test_regional_averages_lie_within_their_cells_range(c3):
    import torch
    from aggfly_b200 import engine
    from aggfly_b200.aggregate import _device_csr
    wl, raster, ds, w = c3
    names, res, runner = _run(wl, raster, ds, _spec())
    csr = _device_csr(w, ds)
    panel, den = engine.run_spmm(csr, res, want_den=True)
    torch.cuda.synchronize()
    R, nnz = csr.host.n_regions, csr.host.nnz
    rows = torch.repeat_interleave(torch.arange(R, device=raster.device), (csr.row_ptr[1:] - csr.row_ptr[:-1]).long())
    cells = csr.cell_idx.long()
    valid = res.V[0].bool()[cells] & (csr.w > 0)
    for col in (3, 9, 15, 16):
        x = res.X[0][cells, col]
        lo = torch.full((R,), float("inf"), dtype=torch.float64, device=raster.device)
        hi = torch.full((R,), float("-inf"), dtype=torch.float64, device=raster.device)
        lo.scatter_reduce_(0, rows[valid], x[valid], "amin")
        hi.scatter_reduce_(0, rows[valid], x[valid], "amax")
        p = panel[:, 0, col]
        has = den[:, 0] != 0
        assert torch.equal(torch.isnan(p), ~has)
        tol = 1e-9 * hi[has].abs().clamp_min(1.0)
        assert bool((p[has] >= lo[has] - tol).all()) and bool((p[has] <= hi[has] + tol).all()), names[col]
    # denominators: sum of the weights of the valid cells
    want_den = torch.zeros(R, dtype=torch.float64, device=raster.device).index_add_(0, rows, csr.w * res.V[0][cells].double())
    assert float(((den[:, 0] - want_den).abs() / want_den.abs().clamp_min(1e-300))[want_den > 0].max()) < 1e-12
    assert int(has.sum()) > 30000 and int((~has).sum()) > 0        # land and all-ocean regions both occur
    runner.close()


def test_daily_panel_of_two_half_year_shards_equals_the_whole_year(c3):
    """Time sharding at full size: rows of a period depend only on that period's hours."""
    import torch
    from aggfly_b200 import engine, shard
    from aggfly_b200.aggregate import _device_csr
    wl, raster, ds, w = c3
    spec = dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})],
                hot=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [25, 99, 0]})])
    csr = _device_csr(w, ds)
    _, whole, r0 = _run(wl, raster, ds, spec)
    p_whole = engine.run_spmm(csr, whole)
    (a0, a1), (b0, b1) = shard.plan_time_shards(ds.time, 2, "month")
    assert a0 == 0 and a1 == b0 and b1 == wl.n_time and a1 % 24 == 0
    parts = []
    for lo, hi in ((a0, a1), (b0, b1)):
        sub = ds.isel_time(lo, hi)
        from aggfly_b200.aggregate import _plan
        names, stage = _plan(sub, spec)
        rr = engine.StageRunner(stage, wl.n_cells, raster.device)
        res = rr.run(raster[lo:hi].reshape(hi - lo, wl.n_cells))
        parts.append(engine.run_spmm(csr, res))
        torch.cuda.synchronize()
        rr.close()
    both = torch.cat(parts, dim=1)
    assert both.shape == p_whole.shape == (csr.host.n_regions, 365, 2)
    assert torch.equal(torch.isnan(both), torch.isnan(p_whole))
    ok = ~torch.isnan(both)
    assert torch.equal(both[ok], p_whole[ok])
    r0.close()
