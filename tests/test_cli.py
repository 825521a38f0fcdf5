"""``aggfly run config.yaml`` on this engine: config schema / validation (reference:
aggfly/cli/config.py:214-386, tests aggfly/tests/test_cli.py), the host-only commands, and -- on a
GPU -- ``run`` against the hand-written API calls it stands for (test_cli.py:426-458)."""
import os

import numpy as np
import pandas as pd
import pytest
import yaml
from click.testing import CliRunner

import aggfly_b200 as af
from aggfly_b200 import cli, io, runconfig
from tests.test_weights import _write_dbf, _write_shp

SPEC = {"tavg": [["aggregate", {"calc": "mean", "groupby": "date"}],
                 ["transform", {"transform": "power", "exp": [1, 2]}],
                 ["aggregate", {"calc": "sum", "groupby": "year"}]],
        "temp_bins": [["aggregate", {"calc": "mean", "groupby": "date"}],
                      ["aggregate", {"calc": "bins", "groupby": "year", "ddargs": [[0, 5, 0], [5, 10, 0], [10, 15, 0]]}]]}


def _project(tmp_path, years=(2001, 2002), fmt="parquet", **dataset_extra):
    lat, lon = 40.0 - 0.5 * np.arange(8), 250.0 + 0.5 * np.arange(10)               # 0-360 longitudes
    rng = np.random.default_rng(0)
    for y in years:
        t = pd.date_range(f"{y}-01-01", f"{y}-03-31 23:00", freq="h")
        vals = (281.0 + rng.normal(0, 6, (len(t), 8, 10))).astype(np.float32)      # Kelvin
        np.savez(tmp_path / f"t2m_{y}.npz", t2m=vals, time=t.values, latitude=lat, longitude=lon)
    sq = lambda x0, y0, s: np.array([[x0, y0], [x0, y0 + s], [x0 + s, y0 + s], [x0 + s, y0], [x0, y0]], dtype=float)   # noqa: E731
    _write_shp(tmp_path / "regions.shp", [[sq(-109.7, 37.1, 1.6)], [sq(-108.1, 36.4, 2.2)], [sq(-60.0, 0.0, 1.0)]])
    _write_dbf(tmp_path / "regions.dbf", ["08001", "08003", "far"])
    np.savez(tmp_path / "pop.npz", values=rng.random((16, 20)) + 0.1, latitude=40.125 - 0.25 * np.arange(16),
             longitude=-110.125 + 0.25 * np.arange(20))
    cfg = {"regions": {"path": str(tmp_path / "regions.shp"), "regionid": "GEOID"},
           "dataset": {"path": str(tmp_path / "t2m_{year}.npz"), "var": "t2m", "preprocess": "kelvin_to_celsius",
                       "lon_is_360": True, **dataset_extra},
           "weights": {"project_dir": str(tmp_path / "proj"), "secondary": {"type": "pop", "path": str(tmp_path / "pop.npz")}},
           "aggregate": {"engine": "auto", "variables": SPEC},
           "years": f"{years[0]}:{years[-1]}",
           "execution": {"backend": "threads"},
           "output": {"path": str(tmp_path / f"out.{fmt}")}}
    path = tmp_path / "config.yaml"
    path.write_text(yaml.safe_dump(cfg))
    return str(path), cfg


def test_config_roundtrip_and_aggregator_dict(tmp_path):
    path, raw = _project(tmp_path)
    cfg = runconfig.load_config(path)
    assert cfg.templated and cfg.years == [2001, 2002] and len(cfg.resolved_paths()) == 2
    assert cfg.output_format == "parquet" and cfg.engine == "auto" and cfg.zero_weight == "nan"
    d = cfg.to_aggregator_dict()
    assert isinstance(d["tavg"][1][1]["exp"], np.ndarray) and d["tavg"][0] == ("aggregate", {"calc": "mean", "groupby": "date"})
    assert runconfig.parse_years("1980:1983", []) == [1980, 1981, 1982, 1983] and runconfig.parse_years(7, []) == [7]


def test_every_problem_is_reported_at_once():
    bad = {"regions": {"path": "r.shp"}, "dataset": {"path": "d_{year}.nc", "preprocess": "x-1", "preprocess_from": "f.py"},
           "weights": {"zero_weight": "bogus", "secondary": {"type": "gold"}},
           "aggregate": {"engine": "fortran", "variables": {
               "a": [["aggregate", {"calc": "median", "groupby": "decade"}], ["transform", {}], ["smooth", {}]],
               "b": [["transform", {"exp": [1, 2]}], ["aggregate", {"calc": "bins", "groupby": "year", "ddargs": [[0, 1, 0], [1, 2, 0]]}]]}},
           "execution": {"backend": "mpi"}, "output": {"path": "out.xlsx"}}
    with pytest.raises(runconfig.ConfigError) as e:
        runconfig.parse_config(bad)
    msgs = "\n".join(e.value.errors)
    for needle in ("regions.regionid is required", "dataset.var is required", "at most one of 'preprocess'",
                   "preprocess_from must be", "zero_weight 'bogus'", "secondary.type 'gold'", "secondary.path is required",
                   "engine 'fortran'", "calc 'median'", "groupby 'decade'", "transform step needs one of",
                   "unknown step type 'smooth'", "cannot combine a multi-'ddargs'", "backend 'mpi'", "output.format 'xlsx'",
                   "contains '{year}' but no 'years'"):
        assert needle in msgs, needle
    with pytest.raises(runconfig.ConfigError, match="non-empty YAML mapping"):
        runconfig.parse_config(None)


def test_validate_and_weights_commands_need_no_gpu(tmp_path):
    path, _ = _project(tmp_path)
    r = CliRunner().invoke(cli.cli, ["validate", path])
    assert r.exit_code == 0 and "Config OK: 2 variable(s), 2 dataset path(s)" in r.output
    r = CliRunner().invoke(cli.cli, ["weights", path])
    assert r.exit_code == 0, r.output
    assert "2 regions" in r.output and os.path.isdir(tmp_path / "proj" / "tmp" / "DeviceCSR")
    broken = tmp_path / "broken.yaml"
    broken.write_text(yaml.safe_dump({"regions": {}}))
    r = CliRunner().invoke(cli.cli, ["validate", str(broken)])
    assert r.exit_code == 1 and "regions.path is required" in r.output
    cfg = yaml.safe_load(open(path))
    cfg["dataset"]["preprocess"] = "import os"
    broken.write_text(yaml.safe_dump(cfg))
    r = CliRunner().invoke(cli.cli, ["validate", str(broken)])
    assert r.exit_code != 0 and "preprocess" in r.output


def test_readers_npz_netcdf3_geojson_and_clip(tmp_path):
    path, _ = _project(tmp_path)
    ds = io.dataset_from_path(str(tmp_path / "t2m_2001.npz"), var="t2m", preprocess="kelvin_to_celsius")
    assert ds.shape == (24 * 90, 8, 10) and ds.dtype == np.float32 and len(ds.pre_ops) == 1
    assert ds.time[0] == pd.Timestamp("2001-01-01") and ds.lon_is_360
    # the same raster as a NetCDF-3 file (int16-packed) through scipy
    from scipy.io import netcdf_file
    f = netcdf_file(str(tmp_path / "t2m.nc"), "w")
    for name, n in (("time", 48), ("latitude", 8), ("longitude", 10)):
        f.createDimension(name, n)
    tv = f.createVariable("time", "i4", ("time",)); tv[:] = np.arange(48); tv.units = "hours since 1900-01-01 00:00:00.0"
    la = f.createVariable("latitude", "f4", ("latitude",)); la[:] = ds.latitude
    lo = f.createVariable("longitude", "f4", ("longitude",)); lo[:] = ds.longitude
    v = f.createVariable("t2m", "i2", ("time", "latitude", "longitude"))
    v.scale_factor, v.add_offset = 0.01, 280.0
    packed = np.round((np.asarray(ds.values[:48], dtype=np.float64) - 280.0) / 0.01).astype(np.int16)
    v[:] = packed
    f.close()
    nc = io.dataset_from_path(str(tmp_path / "t2m.nc"), var="t2m")
    assert nc.dtype == np.float64 and nc.time[1] == pd.Timestamp("1900-01-01 01:00")
    assert np.allclose(nc.values, packed * 0.01 + 280.0) and np.allclose(nc.latitude, ds.latitude)
    # clipping to the regions' extent keeps a block with one cell of margin
    gr = io.georegions_from_path(str(tmp_path / "regions.shp"), "GEOID", ["08001", "08003"])
    small = io.clip_to_extent(ds, *cli.region_extent(gr))
    assert small.shape[1] < 8 and small.shape[2] <= 10 and small.values.base is not None
    assert small.latitude.max() >= 38.7 and small.latitude.min() <= 36.5 and small.shape[1] == 6
    # geojson regions
    import json
    gj = {"type": "FeatureCollection", "features": [
        {"type": "Feature", "properties": {"id": "a"},
         "geometry": {"type": "Polygon", "coordinates": [[[0, 0], [2, 0], [2, 2], [0, 2], [0, 0]], [[0.5, 0.5], [1.5, 0.5], [1.5, 1.5], [0.5, 1.5], [0.5, 0.5]]]}},
        {"type": "Feature", "properties": {"id": "b"},
         "geometry": {"type": "MultiPolygon", "coordinates": [[[[2, 2], [3, 2], [3, 3], [2, 3], [2, 2]]], [[[3, 3], [4, 3], [4, 4], [3, 4], [3, 3]]]]}}]}
    (tmp_path / "r.geojson").write_text(json.dumps(gj))
    gr = io.georegions_from_path(str(tmp_path / "r.geojson"), "id")
    lat = lon = np.arange(0.5, 4.0)
    w = af.weights_from_objects(af.Dataset.from_arrays(np.zeros((1, 4, 4)), pd.date_range("2000-01-01", periods=1), lat, lon, False),
                                gr, cosine_area=False)
    w.calculate_weights()
    assert np.isclose(w.weights.groupby("id").area_weight.sum()["a"], 3.0)
    assert np.isclose(w.weights.groupby("id").area_weight.sum()["b"], 2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["parquet", "feather", "csv"])
def test_run_matches_the_hand_written_api(tmp_path, fmt):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    path, raw = _project(tmp_path, fmt=fmt)
    r = CliRunner().invoke(cli.cli, ["run", path, "-v"])
    assert r.exit_code == 0, r.output
    out = tmp_path / f"out.{fmt}"
    got = {"parquet": pd.read_parquet, "feather": pd.read_feather, "csv": lambda p: pd.read_csv(p, dtype={"GEOID": str})}[fmt](out)
    # the script a user would write by hand
    regions = io.georegions_from_path(raw["regions"]["path"], "GEOID")
    frames, w = [], None
    for y in (2001, 2002):
        ds = io.dataset_from_path(raw["dataset"]["path"].format(year=y), var="t2m", preprocess="kelvin_to_celsius", lon_is_360=True)
        if w is None:
            w = af.weights_from_objects(ds, regions, secondary_weights=io.secondary_weights_from_path(raw["weights"]["secondary"]["path"]))
            w.calculate_weights()
        cfg = runconfig.load_config(path)
        frames.append(af.aggregate_dataset(dataset=ds, weights=w, aggregator_dict=cfg.to_aggregator_dict()))
    want = pd.concat(frames, ignore_index=True)
    assert list(got.columns) == list(want.columns) and len(got) == len(want) == 4     # 2 regions x 2 years ("far" has no cells)
    assert list(got.GEOID) == list(want.GEOID)
    vals = [c for c in want.columns if c not in ("GEOID", "time")]
    # clipping (on in the CLI run, off in the hand-written one) never changes results (test_cli.py:461-476)
    assert np.allclose(got[vals].values.astype(float), want[vals].values, rtol=1e-12)


@pytest.mark.gpu
def test_run_from_zarr_stores_matches_the_npz_run(tmp_path):
    """The reference's main workflow: ``aggfly run`` over one zarr store per year (one in the converter's
    time-contiguous Blosc layout, one time-major zstd v3), clipped to the regions -- same panel as from the arrays."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from aggfly_b200 import zarrio
    path, raw = _project(tmp_path)
    r = CliRunner().invoke(cli.cli, ["run", path])
    assert r.exit_code == 0, r.output
    want = pd.read_parquet(tmp_path / "out.parquet")
    layouts = {2001: dict(dims=("latitude", "longitude", "time"), chunks={"latitude": 3, "longitude": 4}, zarr_format=2, compressor="blosc"),
               2002: dict(dims=("time", "latitude", "longitude"), chunks={"time": 24 * 7}, zarr_format=3, compressor="zstd")}
    for y, lay in layouts.items():
        z = np.load(tmp_path / f"t2m_{y}.npz")
        zarrio.write_dataset(str(tmp_path / f"t2m_{y}.zarr"), z["t2m"], pd.DatetimeIndex(z["time"]), z["latitude"], z["longitude"],
                             var="t2m", **lay)
    raw["dataset"]["path"] = str(tmp_path / "t2m_{year}.zarr")
    raw["output"]["path"] = str(tmp_path / "out_zarr.parquet")
    (tmp_path / "config_zarr.yaml").write_text(yaml.safe_dump(raw))
    r = CliRunner().invoke(cli.cli, ["run", str(tmp_path / "config_zarr.yaml")])
    assert r.exit_code == 0, r.output
    got = pd.read_parquet(tmp_path / "out_zarr.parquet")
    assert list(got.columns) == list(want.columns) and list(got.GEOID) == list(want.GEOID) and list(got.time) == list(want.time)
    vals = [c for c in want.columns if c not in ("GEOID", "time")]
    assert np.allclose(got[vals].values.astype(float), want[vals].values.astype(float), rtol=1e-12, atol=0, equal_nan=True)


# ---- `aggfly info` (reference: aggfly/cli/info.py, tests aggfly/tests/test_cli.py:77-115) -------------------------------
def _info_store(tmp_path, name, calendar="standard", lon_360=True, var="t2m", units="K", fmt=3):
    from aggfly_b200 import zarrio
    lon = (np.arange(0.0, 360.0, 90.0) if lon_360 else np.arange(-180.0, 180.0, 90.0))
    lat = np.array([10.0, 0.0, -10.0])
    t = pd.date_range("2001-01-01", periods=48, freq="h") if calendar == "standard" else af.CalendarIndex.range(calendar, 2001, 48, "D")
    return zarrio.write_dataset(str(tmp_path / name), np.zeros((48, 3, 4), np.float32), t, lat, lon, var=var, chunks={"time": 24},
                                zarr_format=fmt, compressor="zstd", attrs={"units": units} if units else None)


def test_info_datetime64_era5_like(tmp_path):
    r = CliRunner().invoke(cli.cli, ["info", _info_store(tmp_path, "era5.zarr")])
    assert r.exit_code == 0, r.output
    for needle in ("data variables : t2m", "dims   : time=48, latitude=3, longitude=4", "chunks : time=24, latitude=3, longitude=4",
                   "xycoords   : [longitude, latitude]", "lon_is_360: true", "timecoord  : time", "units  : K", "time steps : 48",
                   "time span  : 2001-01-01 00:00:00 .. 2001-01-02 23:00:00"):
        assert needle in r.output, needle
    assert "cftime" not in r.output                                    # a standard calendar is not flagged


def test_info_360day_cmip6_like_and_errors(tmp_path):
    path = _info_store(tmp_path, "cmip6.zarr", calendar="360_day", lon_360=False, var="tas", units="", fmt=2)
    r = CliRunner().invoke(cli.cli, ["info", path, "--var", "tas"])
    assert r.exit_code == 0, r.output
    assert "calendar   : 360_day" in r.output and "cftime / non-standard" in r.output and "lon_is_360: false" in r.output
    r = CliRunner().invoke(cli.cli, ["info", path, "--var", "nope"])
    assert r.exit_code != 0 and "not found" in r.output
    r = CliRunner().invoke(cli.cli, ["info", path, "--storage-options", "{not json}"])
    assert r.exit_code != 0 and "not valid JSON" in r.output
    r = CliRunner().invoke(cli.cli, ["info", str(tmp_path / "missing.tif")])
    assert r.exit_code != 0 and "Could not open" in r.output


def test_info_npz(tmp_path):
    path, _ = _project(tmp_path, years=(2001,))
    r = CliRunner().invoke(cli.cli, ["info", str(tmp_path / "t2m_2001.npz")])
    assert r.exit_code == 0, r.output
    assert "t2m:" in r.output and "lon_is_360: true" in r.output and "time steps : 2160" in r.output


def test_regions_command(tmp_path):
    """`aggfly regions` (aggfly/cli/main.py:50-81, aggfly/regions/georegions.py:326-428) from the file headers."""
    _project(tmp_path, years=(2001,))
    r = CliRunner().invoke(cli.cli, ["regions", str(tmp_path / "regions.shp"), "--uniqueness"])
    assert r.exit_code == 0, r.output
    for needle in ("geometry   : Polygon  features=3", "bounds     : lon -109.7000 .. -59.0000 | lat 0.0000 .. 38.7000",
                   "fields     : 1", "GEOID", "first 3 row(s)", "08003", "regionid candidates", "crs        : NONE"):
        assert needle in r.output, needle
    d = cli.describe_regions(str(tmp_path / "regions.shp"), rows=0)
    assert d["features"] == 3 and d["head"] is None and d["unique_columns"] is None
    os.remove(tmp_path / "regions.dbf")
    r = CliRunner().invoke(cli.cli, ["regions", str(tmp_path / "regions.shp")])
    assert r.exit_code == 0 and "no column to use as regionid" in r.output
    r = CliRunner().invoke(cli.cli, ["regions", str(tmp_path / "nope.gpkg")])
    assert r.exit_code != 0 and "ValueError" in r.output


def test_validate_reports_missing_paths(tmp_path):
    path, raw = _project(tmp_path, years=(2001,))
    r = CliRunner().invoke(cli.cli, ["validate", path, "--strict"])
    assert r.exit_code == 0 and "Config OK" in r.output and "Warnings" not in r.output
    raw["years"] = "2001:2002"                                        # 2002 was never written
    (tmp_path / "c2.yaml").write_text(yaml.safe_dump(raw))
    r = CliRunner().invoke(cli.cli, ["validate", str(tmp_path / "c2.yaml")])
    assert r.exit_code == 0 and "Warnings:" in r.output and "t2m_2002.npz does not exist" in r.output and "Config OK" in r.output
    r = CliRunner().invoke(cli.cli, ["validate", str(tmp_path / "c2.yaml"), "--strict"])
    assert r.exit_code == 1 and "Config OK" not in r.output
