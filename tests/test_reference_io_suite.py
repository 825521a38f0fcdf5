"""The reference's own tests of the data formats either side of the hot path, restated against this package's native
readers (no xarray / zarr-python / geopandas here, so the fixtures are written with ``zarrio`` / scipy / the shapefile
helpers of tests/test_weights.py).  Each test cites the reference test it mirrors (aggfly/tests/test_aggregate.py)."""
import os

import numpy as np
import pandas as pd
import pytest

import aggfly_b200 as af
from aggfly_b200 import zarrio
from aggfly_b200.io import _auto_chunks
from aggfly_b200.zarrio import looks_like_zarr
from tests import refcases as rc
from tests.test_weights import _write_shp


def test_auto_chunks_policy():
    """:667-679."""
    c = _auto_chunks({"latitude": 721, "longitude": 1440, "time": 8784}, 4, 256)       # short series: full-time chunk
    assert c["time"] == -1 and c["latitude"] == c["longitude"] and c["latitude"] >= 32
    c = _auto_chunks({"latitude": 721, "longitude": 1440, "time": 350640}, 4, 256)     # long hourly series: time split
    assert 0 < c["time"] < 350640 and c["time"] * c["latitude"] * c["longitude"] * 4 <= 256 * 1024 * 1024
    c = _auto_chunks({"latitude": 2, "longitude": 2, "time": 4}, 8, 256)               # tiny grid: tile capped by extent
    assert c["latitude"] <= 2 and c["longitude"] <= 2


def test_dataset_to_zarr_roundtrip(tmp_path):
    """:693-709 on the reference's ``dataset_360`` fixture (:17-53)."""
    arr, t, lat, lon = rc.dataset_360_arrays()
    ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360=True, name="t2m")
    store = str(tmp_path / "roundtrip.zarr")
    out = af.dataset_to_zarr(ds, store, overwrite=True)
    assert os.path.isdir(store) and isinstance(out, af.Dataset)
    a = out.values.array
    assert a.chunks[a.dims.index("time")] == len(t)                                    # time is one chunk
    assert a.dims == ("latitude", "longitude", "time") and out.values.single_time_chunk
    assert np.allclose(np.transpose(arr, (1, 2, 0)), np.transpose(np.asarray(out.values), (1, 2, 0)))
    assert out.lon_is_360 == ds.lon_is_360 and (out.time == t).all()
    with pytest.raises(FileExistsError):                                               # refuses to clobber
        af.dataset_to_zarr(ds, store, overwrite=False)


def test_dataset_to_zarr_compresses(tmp_path):
    """:712-720."""
    t = pd.date_range("2000-01-01", periods=100, freq="D")
    ds = af.Dataset.from_arrays(np.zeros((100, 50, 50), np.float32), t, np.arange(50.0), np.arange(50.0), lon_is_360=False)
    store = str(tmp_path / "compress.zarr")
    assert af.dataset_to_zarr(ds, store, return_dataset=False, overwrite=True) is None
    on_disk = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(store) for f in fs)
    assert on_disk < 100 * 50 * 50 * 4


def test_zarr_from_path_agnostic(tmp_path):
    """:723-747 -- a source with ERA5-style ``valid_time / lat / lon`` names converts through the same normalisation
    parameters (the source here is NetCDF-3, written with scipy)."""
    from scipy.io import netcdf_file
    t = pd.date_range("2016-06-01", periods=48, freq="h")
    vals = np.arange(48 * 20 * 30, dtype=np.float32).reshape(48, 20, 30)
    nc = str(tmp_path / "src.nc")
    with netcdf_file(nc, "w") as f:
        f.createDimension("valid_time", 48), f.createDimension("lat", 20), f.createDimension("lon", 30)
        tv = f.createVariable("valid_time", "i4", ("valid_time",))
        tv[:] = np.arange(48)
        tv.units = "hours since 2016-06-01 00:00:00"
        f.createVariable("lat", "f8", ("lat",))[:] = np.linspace(10, 20, 20)
        f.createVariable("lon", "f8", ("lon",))[:] = np.linspace(-100, -80, 30)
        f.createVariable("t2m", "f4", ("valid_time", "lat", "lon"))[:] = vals
    out = af.zarr_from_path(nc, var="t2m", store=str(tmp_path / "out.zarr"), xycoords=("lon", "lat"), timecoord="valid_time",
                            lon_is_360=False)
    assert isinstance(out, af.Dataset) and set(out.values.array.dims) >= {"latitude", "longitude", "time"}
    assert out.values.single_time_chunk                                                # time contiguous
    assert np.allclose(np.asarray(out.values), vals) and (out.time == t).all()
    assert np.allclose(out.latitude, np.linspace(10, 20, 20)) and np.allclose(out.longitude, np.linspace(-100, -80, 30))


def test_cftime_roundtrip_dataset_from_path(tmp_path):
    """:549-562 -- a noleap store keeps its calendar through dataset_from_path (the monthly aggregation of that test runs
    on the GPU in tests/test_gpu_parity.py; here: the calendar and the 12 month groups)."""
    t = af.CalendarIndex.range("noleap", 2001, 365, "D")
    store = zarrio.write_dataset(str(tmp_path / "cmip.zarr"), np.random.default_rng(0).random((365, 2, 2)), t, [-45.0, 45.0],
                                 [10.0, 100.0], var="tas", zarr_format=3, compressor="zstd")
    ds = af.dataset_from_path(store, var="tas", lon_is_360=False, xycoords=("longitude", "latitude"), timecoord="time")
    assert isinstance(ds.time, af.CalendarIndex) and ds.time.calendar == "noleap" and ds.dtype == np.float64
    bounds, labels = af.group_bounds(ds.time, af.translate_groupby("month"))
    assert len(bounds) - 1 == 12 and np.diff(bounds).tolist() == [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def test_looks_like_zarr_by_name_without_touching_the_filesystem():
    """:1055-1062."""
    assert looks_like_zarr("/nonexistent/x.zarr") and looks_like_zarr("s3://bucket/store.zarr/")
    for p in ["/nonexistent/f.nc", "/nonexistent/F.NC4", "/nonexistent/x.tif", "/nonexistent/y.grib2", "/nonexistent/z.hdf5"]:
        assert not looks_like_zarr(p), p


def test_zarr_store_without_suffix_or_with_explicit_engine(tmp_path):
    """:1065-1101 -- a store directory not named *.zarr is recognised by its metadata; ``engine="zarr"`` is accepted."""
    t = pd.date_range("2000-01-01", periods=3)
    for fmt, name in ((2, "storedir"), (3, "storedir3")):
        store = zarrio.write_dataset(str(tmp_path / name), np.ones((3, 4, 4)), t, np.arange(4.0) + 0.5, np.arange(4.0) + 0.5,
                                     var="t2m", zarr_format=fmt, compressor="zstd")
        assert looks_like_zarr(store) and not looks_like_zarr(str(tmp_path / "does_not_exist"))
        assert af.dataset_from_path(store, var="t2m").shape[0] == 3
        assert af.dataset_from_path(store, var="t2m", engine="zarr").shape[0] == 3


def _write_dbf2(path, columns: dict, flen: int = 8):
    """dBASE III table of character fields (one per key of ``columns``)."""
    import struct
    names = list(columns)
    n = len(columns[names[0]])
    hdr = struct.pack("<4BIHH20x", 3, 24, 1, 1, n, 32 + 32 * len(names) + 1, 1 + flen * len(names))
    fields = b"".join(nm.encode().ljust(11, b"\x00") + b"C" + b"\x00" * 4 + bytes([flen, 0]) + b"\x00" * 14 for nm in names)
    body = b"".join(b" " + b"".join(str(columns[nm][i]).encode().ljust(flen) for nm in names) for i in range(n))
    open(path, "wb").write(hdr + fields + b"\x0d" + body + b"\x1a")


def _regions_file(tmp_path, names=("a", "b", "c")):
    sq = lambda i: np.array([[i, 0], [i, 1], [i + 1, 1], [i + 1, 0], [i, 0]], dtype=float)   # noqa: E731
    _write_shp(tmp_path / "r.shp", [[sq(i)] for i in range(3)])
    _write_dbf2(tmp_path / "r.dbf", {"fips": ["01", "02", "03"], "name": list(names)})
    return str(tmp_path / "r.shp")


def test_shapefile_info_reports_fields_bounds_and_unique_columns(tmp_path, capsys):
    """:1168-1183."""
    info = af.shapefile_info(_regions_file(tmp_path), n=2, uniqueness=True)
    assert info["features"] == 3 and set(info["fields"]) == {"fips", "name"} and len(info["total_bounds"]) == 4
    assert tuple(info["total_bounds"]) == (0.0, 0.0, 3.0, 1.0)
    assert info["head"] is not None and len(info["head"]) == 2 and "geometry" not in info["head"].columns
    assert set(info["unique_columns"]) == {"fips", "name"}
    assert "regionid candidates" in capsys.readouterr().out


def test_shapefile_info_flags_a_non_unique_column(tmp_path):
    """:1186-1194."""
    info = af.shapefile_info(_regions_file(tmp_path, names=("a", "b", "a")), n=0, uniqueness=True)
    assert info["unique_columns"] == ["fips"] and info["head"] is None


def test_shapefile_info_handles_a_file_with_no_attributes(tmp_path):
    """:1197-1212 (GeoJSON without properties)."""
    import json
    path = tmp_path / "geom_only.geojson"
    path.write_text(json.dumps({"type": "FeatureCollection", "features": [
        {"type": "Feature", "properties": {}, "geometry": {"type": "Polygon", "coordinates": [[[0, 0], [1, 0], [1, 1], [0, 1], [0, 0]]]}}]}))
    info = af.shapefile_info(str(path))
    assert info["fields"] == [] and info["head"] is None and info["features"] == 1


# ---- secondary rasters (:1222-1322): GeoTIFF, zarr with variable + coordinate selection, NetCDF ----------------------------
def _cropland_zarr(tmp_path):
    """:1236-1253 -- a `layer` variable indexed by a `crop` coordinate (zarr v2, fixed-width unicode labels like xarray writes)."""
    root = str(tmp_path / "cropland.zarr")
    y = x = np.arange(0, 4.0) + 0.5
    import json
    os.makedirs(root)
    json.dump({"zarr_format": 2}, open(root + "/.zgroup", "w"))
    zarrio.write_array(root + "/layer", np.stack([np.ones((4, 4)), np.full((4, 4), 7.0)]), [1, 4, 4], ["crop", "y", "x"])
    zarrio.write_array(root + "/crop", np.array(["corn", "soybeans"], dtype="<U8"), [-1], ["crop"], compressor=None, fill_value=0)
    json.dump(dict(json.load(open(root + "/crop/.zarray")), fill_value=""), open(root + "/crop/.zarray", "w"))
    zarrio.write_array(root + "/y", y, [-1], ["y"])
    zarrio.write_array(root + "/x", x, [-1], ["x"])
    return root


def _pop_tif(tmp_path):
    """:1222-1233."""
    from PIL import Image, TiffImagePlugin
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550], ifd.tagtype[33550] = (1.0, 1.0, 0.0), 12
    ifd[33922], ifd.tagtype[33922] = (0.0, 0.0, 0.0, 0.0, 4.0, 0.0), 12
    path = str(tmp_path / "pop.tif")
    Image.fromarray(np.arange(16, dtype=np.float32).reshape(4, 4)).save(path, tiffinfo=ifd)
    return path


def test_pop_and_crop_wrappers_equal_the_generic_loader(tmp_path):
    """:1256-1300."""
    tif = _pop_tif(tmp_path)
    pop, gen = af.pop_weights_from_path(tif), af.secondary_weights_from_path(tif, wtype="pop")
    assert pop.wtype == gen.wtype == "pop" and np.array_equal(pop.values, gen.values)
    assert np.array_equal(pop.latitude, [3.5, 2.5, 1.5, 0.5]) and np.array_equal(pop.longitude, [0.5, 1.5, 2.5, 3.5])
    store = _cropland_zarr(tmp_path)
    crop = af.crop_weights_from_path(store, crop="soybeans", feed="rainfed", crs="WGS84")
    gen = af.secondary_weights_from_path(store, var="layer", sel={"crop": "soybeans"}, wtype="soybeans", cache_identifier="rainfed",
                                         crs="WGS84")
    assert np.array_equal(crop.values, gen.values) and (crop.wtype, crop.cache_identifier) == (gen.wtype, gen.cache_identifier)
    assert float(crop.values.max()) == 7.0                                       # soybeans == 7.0, corn == 1.0
    assert float(af.crop_weights_from_path(store, crop="corn", crs="WGS84").values.max()) == 1.0
    rain, irr = (af.crop_weights_from_path(store, crop="corn", feed=f, crs="WGS84") for f in ("rainfed", "irrigated"))
    assert (rain.feed, irr.feed) == ("rainfed", "irrigated") and rain.cache_identifier != irr.cache_identifier
    with pytest.raises(KeyError, match="wheat"):
        af.crop_weights_from_path(store, crop="wheat")


def test_open_raster_supports_tif_zarr_and_netcdf(tmp_path):
    """:1304-1322."""
    from scipy.io import netcdf_file
    assert af.secondary_weights_from_path(_pop_tif(tmp_path)).values.shape == (4, 4)
    z = af.secondary_weights_from_path(_cropland_zarr(tmp_path), var="layer", sel={"crop": "corn"})
    assert float(z.values.max()) == 1.0 and z.values.shape == (4, 4)
    nc = str(tmp_path / "x.nc")
    with netcdf_file(nc, "w") as f:
        f.createDimension("y", 2), f.createDimension("x", 2)
        f.createVariable("y", "f8", ("y",))[:] = [0.5, 1.5]
        f.createVariable("x", "f8", ("x",))[:] = [0.5, 1.5]
        f.createVariable("layer", "f8", ("y", "x"))[:] = np.ones((2, 2))
    got = af.secondary_weights_from_path(nc, var="layer")
    assert np.array_equal(got.values, np.ones((2, 2))) and np.array_equal(got.latitude, [0.5, 1.5])
    with pytest.raises(NotImplementedError, match="Unsupported raster format"):
        af.secondary_weights_from_path(str(tmp_path / "x.bogus"))


def test_cropland_weights_flow_into_the_weight_frame(tmp_path):
    """A cropland layer from a zarr store weights the cells of a region like an in-memory raster does."""
    store = _cropland_zarr(tmp_path)
    lat = lon = np.arange(0, 4.0) + 0.5
    ds = af.Dataset.from_arrays(np.ones((3, 4, 4)), pd.date_range("2000-01-01", periods=3), lat, lon, lon_is_360=False)
    regions = af.GeoRegions.from_rectangles(["r1"], lon_min=[0.0], lon_max=[4.0], lat_min=[0.0], lat_max=[4.0])
    a = af.weights_from_objects(ds, regions, secondary_weights=af.crop_weights_from_path(store, crop="soybeans"))
    b = af.weights_from_objects(ds, regions, secondary_weights=af.SecondaryWeights(np.full((4, 4), 7.0), lat, lon))
    a.calculate_weights(), b.calculate_weights()
    assert len(a.weights) == 16 and np.allclose(a.weights.sort_values("cell_id").weight.values, b.weights.sort_values("cell_id").weight.values)
