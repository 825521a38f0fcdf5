"""Synthetic ERA5-/CMIP6-shaped workloads for benchmarks, smoke and tests.

There is no network and no real data in the build environment, so every benchmark config is
generated: an hourly (or daily) temperature raster with a latitudinal gradient, a seasonal and a
diurnal cycle plus noise (so every bin / degree-day threshold is exercised), an optional all-NaN
"ocean" mask, and a rectangular tessellation of pseudo-regions whose exact area x cos(lat)
(x secondary raster) weights come from ``GridWeights.calculate_weights``.

The raster is produced on the device with torch (data generation is not part of the product
path) and is reproducible from ``(seed, year)``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import pandas as pd

from .dataset import Dataset
from .timeaxis import CalendarIndex
from .weights import GeoRegions, GridWeights, weights_from_objects

BINS13 = [[-20 + 5 * i, -15 + 5 * i, 0] for i in range(13)]

SPECS: Dict[str, dict] = {
    # configs[0]: tavg date-mean -> power 1..2 -> year sum
    "tavg_poly": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                            ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                            ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # examples/era5_counties_area.yaml: daily mean -> annual mean, plus growing degree days -> annual sum
    "area_example": dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("aggregate", {"calc": "mean", "groupby": "year"})],
                         gdd_10_30=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                                    ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # configs[1]: growing degree-days [10, 30, 0] date -> year sum
    "gdd": dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                     ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # configs[2]: temperature bins + daily-mean polynomial, one call, one pass
    "bins_poly": dict(
        temp_bins=[("aggregate", {"calc": "mean", "groupby": "date"}),
                   ("aggregate", {"calc": "bins", "groupby": "year", "ddargs": BINS13})],
        tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
              ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
              ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # C3b: daily panel (hourly bins per date + daily mean); write traffic is not negligible
    "daily_bins_mean": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13})],
                            tavg=[("aggregate", {"calc": "mean", "groupby": "date"})]),
    # bins of the HOURLY values by year (SURVEY a-1: "[bins(multi) on hourly]"): one ragged 8760-row group per cell
    "hourly_bins_year": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "year", "ddargs": BINS13})],
                             tavg=[("aggregate", {"calc": "mean", "groupby": "year"})]),
    # the reference tests' own chain shape (tests/test_aggregate.py:275-280): hourly bins per date summed over the year,
    # next to the polynomial of the daily mean
    "bins_date_year_poly": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13}),
                                       ("aggregate", {"calc": "sum", "groupby": "year"})],
                                tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                                      ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                                      ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "bins_date_year": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13}),
                                  ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # probes of other chain shapes at the global size (tools/gpu_r2_probe.sh)
    "probe_minmaxmean": dict(tmin=[("aggregate", {"calc": "min", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})],
                             tmax=[("aggregate", {"calc": "max", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})],
                             tavg=[("aggregate", {"calc": "mean", "groupby": "date"}), ("aggregate", {"calc": "mean", "groupby": "year"})]),
    "probe_dd6": dict(dd=[("aggregate", {"calc": "dd", "groupby": "date",
                                         "ddargs": [[0, 10, 0], [10, 20, 0], [20, 30, 0], [30, 99, 0], [-99, 0, 1], [10, 30, 0]]}),
                          ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "probe_bins_mean": dict(hbins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": BINS13}),
                                   ("aggregate", {"calc": "mean", "groupby": "year"})]),
    "probe_golden": dict(bins=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [[-99, 20, 0], [20, 99, 0]]}),
                               ("aggregate", {"calc": "sum", "groupby": "year"})],
                         cooling_dday=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [20, 99, 0]}),
                                       ("aggregate", {"calc": "sum", "groupby": "year"})],
                         tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                               ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                               ("aggregate", {"calc": "sum", "groupby": "year"})]),
    "probe_sine": dict(sdd=[("aggregate", {"calc": "sine_dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                            ("aggregate", {"calc": "sum", "groupby": "year"})]),
    # configs[4]: degree-days by month (daily input)
    "gdd_month": dict(gdd=[("aggregate", {"calc": "dd", "groupby": "date", "ddargs": [10, 30, 0]}),
                           ("aggregate", {"calc": "sum", "groupby": "month"})]),
}


@dataclass
class GridDef:
    latitude: np.ndarray
    longitude: np.ndarray
    lon_is_360: bool
    regions: tuple        # (n_rx, n_ry, lon_min, lon_max, lat_min, lat_max) in -180..180 coordinates


def conus_grid() -> GridDef:
    lat = 49.75 - 0.25 * np.arange(104)                 # descending, like ERA5 files
    lon = 235.0 + 0.25 * np.arange(236)                 # 0-360 convention
    return GridDef(lat, lon, True, (62, 50, -125.125, -66.125, 23.875, 49.875))


def global_grid() -> GridDef:
    lat = 90.0 - 0.25 * np.arange(721)
    lon = 0.25 * np.arange(1440)
    return GridDef(lat, lon, True, (300, 150, -180.0, 180.0, -60.0, 75.0))


def cmip_grid() -> GridDef:
    lat = -89.5 + np.arange(180)
    lon = 0.5 + np.arange(360)
    return GridDef(lat, lon, True, (90, 60, -180.0, 180.0, -60.0, 75.0))


def small_grid(n_lat=40, n_lon=64) -> GridDef:
    lat = 49.75 - 0.25 * np.arange(n_lat)
    lon = 235.0 + 0.25 * np.arange(n_lon)
    return GridDef(lat, lon, True, (8, 5, -125.125, -125.125 + 0.25 * n_lon, 49.875 - 0.25 * n_lat, 49.875))


GRIDS = {"conus": conus_grid, "global": global_grid, "cmip": cmip_grid, "small": small_grid}


def tessellation(gd: GridDef, regionid: str = "geoid") -> GeoRegions:
    n_rx, n_ry, lon0, lon1, lat0, lat1 = gd.regions
    xe = np.linspace(lon0, lon1, n_rx + 1)
    ye = np.linspace(lat0, lat1, n_ry + 1)
    jj, ii = np.meshgrid(np.arange(n_ry), np.arange(n_rx), indexing="ij")
    ids = [f"R{j:03d}_{i:03d}" for j, i in zip(jj.ravel(), ii.ravel())]
    return GeoRegions.from_rectangles(ids, xe[ii.ravel()], xe[ii.ravel() + 1], ye[jj.ravel()], ye[jj.ravel() + 1],
                                      regionid=regionid)


def synth_raster(gd: GridDef, n_time: int, seed: int, device, hourly: bool = True, ocean_frac: float = 0.0,
                 dtype="float32", chunk_days: int = 8, noise_sigma: float = 3.0):
    """values[T, lat, lon] on ``device`` (torch tensor):
    27 cos(lat) - 6 + 9 sign(lat) sin(2 pi (doy - 110) / 365) + 4 sin(2 pi (hour - 9) / 24) + 3 N(0, 1) [deg C].
    ``noise_sigma`` (default 3: every hour and cell independent -- much rougher than a reanalysis field) is only varied
    by the sensitivity runs of tools/regional_bench.py."""
    import torch
    tdt = torch.float32 if dtype == "float32" else torch.float64
    ny, nx = len(gd.latitude), len(gd.longitude)
    out = torch.empty((n_time, ny, nx), dtype=tdt, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    phi = torch.deg2rad(torch.as_tensor(gd.latitude, dtype=torch.float32, device=device))
    base = (27.0 * torch.cos(phi) - 6.0)[None, :, None]
    hemi = torch.sign(phi)[None, :, None]
    per_day = 24 if hourly else 1
    step = chunk_days * per_day
    for t0 in range(0, n_time, step):
        t1 = min(n_time, t0 + step)
        t = torch.arange(t0, t1, device=device, dtype=torch.float32)
        doy = torch.floor(t / per_day) % 365
        season = 9.0 * torch.sin(2 * np.pi * (doy - 110.0) / 365.0)[:, None, None]
        v = base + hemi * season
        if hourly:
            v = v + 4.0 * torch.sin(2 * np.pi * ((t % 24) - 9.0) / 24.0)[:, None, None]
        noise = torch.randn((t1 - t0, ny, nx), generator=gen, device=device, dtype=torch.float32)
        out[t0:t1] = (v + float(noise_sigma) * noise).to(tdt)
    if ocean_frac > 0:
        g2 = torch.Generator(device=device)
        g2.manual_seed(int(seed) + 7919)
        # blocky mask (8 x 8 cell tiles) so "ocean" cells are spatially coherent
        tiles = torch.rand(((ny + 7) // 8, (nx + 7) // 8), generator=g2, device=device) < ocean_frac
        mask = tiles.repeat_interleave(8, 0).repeat_interleave(8, 1)[:ny, :nx]
        out[:, mask] = float("nan")
    return out


@dataclass
class Workload:
    name: str
    grid: GridDef
    spec_name: str
    n_time: int
    time: object
    hourly: bool = True
    secondary: bool = False
    ocean_frac: float = 0.25
    description: str = ""

    @property
    def spec(self):
        return SPECS[self.spec_name]

    @property
    def n_cells(self) -> int:
        return len(self.grid.latitude) * len(self.grid.longitude)

    @property
    def cell_steps(self) -> int:
        """cell-hours (cell-days for daily inputs) one pass aggregates"""
        return self.n_time * self.n_cells

    def raster(self, device, seed: int, noise_sigma: float = 3.0):
        return synth_raster(self.grid, self.n_time, seed, device, hourly=self.hourly, ocean_frac=self.ocean_frac,
                            noise_sigma=noise_sigma)

    def dataset(self, values) -> Dataset:
        return Dataset.from_arrays(values, self.time, self.grid.latitude, self.grid.longitude,
                                   lon_is_360=self.grid.lon_is_360, name=self.name)

    def weights(self, dataset: Dataset) -> GridWeights:
        regions = tessellation(self.grid)
        sec = None
        if self.secondary:          # log-normal "population" on the weight grid
            rng = np.random.default_rng(1217)
            sec = rng.lognormal(0.0, 1.5, size=self.n_cells)
        w = weights_from_objects(dataset, regions, sec)
        w.calculate_weights()
        return w


def _hourly_year(year=2001):
    return pd.date_range(f"{year}-01-01", periods=8760, freq="h")


def make_workload(name: str) -> Workload:
    if name == "c1_conus_tavg":
        return Workload(name, conus_grid(), "tavg_poly", 8760, _hourly_year(), ocean_frac=0.0,
                        description="CONUS 0.25deg hourly 104x236x8760, 3100 pseudo-counties, area weights, "
                                    "tavg date-mean -> power 1..2 -> year sum")
    if name == "c2_conus_gdd":
        return Workload(name, conus_grid(), "gdd", 8760, _hourly_year(), secondary=True, ocean_frac=0.0,
                        description="CONUS grid, population-weighted, dd[10,30,0] date -> year sum")
    if name == "c3c_global_area_example":
        return Workload(name, global_grid(), "area_example", 8760, _hourly_year(),
                        description="global 0.25deg hourly year, the reference's example spec: daily mean -> annual mean "
                                    "+ dd[10,30,0] -> annual sum (mixed mean / degree-day lanes, one pass)")
    if name == "c3_global_bins":
        return Workload(name, global_grid(), "bins_poly", 8760, _hourly_year(),
                        description="global 0.25deg hourly 721x1440x8760 (36.4 GB f32), 45000 pseudo admin-2 "
                                    "regions, daily-mean -> 13 yearly bins + tavg power 1..2 year sum, one pass")
    if name == "c3b_global_daily":
        return Workload(name, global_grid(), "daily_bins_mean", 8760, _hourly_year(),
                        description="global 0.25deg hourly year, daily panel: 13 hourly bins per date + daily mean")
    if name == "c3d_global_hourly_bins":
        return Workload(name, global_grid(), "hourly_bins_year", 8760, _hourly_year(),
                        description="global 0.25deg hourly year: 13 bins of the hourly values per year + annual mean")
    if name == "c3e_global_bins_date_year":
        return Workload(name, global_grid(), "bins_date_year_poly", 8760, _hourly_year(),
                        description="global 0.25deg hourly year: hourly bins per date -> year sum, + daily mean -> power 1..2 -> year sum")
    if name == "c3f_global_bins_date_year_only":
        return Workload(name, global_grid(), "bins_date_year", 8760, _hourly_year(),
                        description="global 0.25deg hourly year: hourly bins per date -> year sum")
    if name.startswith("probe_") and name in SPECS:
        return Workload(name, global_grid(), name, 8760, _hourly_year(), description="global hourly year, chain-shape probe " + name)
    if name == "c5_cmip_gdd":
        n = 365 * 150
        return Workload(name, cmip_grid(), "gdd_month", n, CalendarIndex.range("noleap", 1950, n), hourly=False,
                        secondary=True, description="CMIP6-shaped 1deg daily noleap 180x360x54750, cropland-weighted "
                                                    "dd[10,30,0] by month")
    if name == "small":
        return Workload(name, small_grid(), "bins_poly", 24 * 60, pd.date_range("2001-01-01", periods=24 * 60, freq="h"),
                        description="tiny smoke workload")
    raise ValueError(f"unknown workload {name!r}")
