"""Weights builder without GEOS (aggfly_b200/geometry.py + csrc/agf_geom.cu): the reference's known
answers (aggfly/tests/test_aggregate.py:191-237) and its property tests (:838-870, :1348-1529)
restated without geopandas/shapely/xarray.  Host-only: no GPU needed."""
import struct
import warnings

import numpy as np
import pandas as pd
import pytest

import aggfly_b200 as af
from aggfly_b200 import geometry as geo
from aggfly_b200.weights import SecondaryWeights
from tests import refcases as rc


# ---- an independent (pure Python) clip, used as the checker for the library's C++ ------------------
def _py_clip_area(ring, x0, x1, y0, y1):
    def pass_(P, inside, cut):
        out = []
        for i in range(len(P)):
            a, b = P[i - 1], P[i]
            ia, ib = inside(a), inside(b)
            if ia != ib:
                out.append(cut(a, b))
            if ib:
                out.append(b)
        return out
    P = [tuple(p) for p in ring]
    for inside, cut in [
            (lambda p: p[0] >= x0, lambda a, b: (x0, a[1] + (b[1] - a[1]) * (x0 - a[0]) / (b[0] - a[0]))),
            (lambda p: p[0] <= x1, lambda a, b: (x1, a[1] + (b[1] - a[1]) * (x1 - a[0]) / (b[0] - a[0]))),
            (lambda p: p[1] >= y0, lambda a, b: (a[0] + (b[0] - a[0]) * (y0 - a[1]) / (b[1] - a[1]), y0)),
            (lambda p: p[1] <= y1, lambda a, b: (a[0] + (b[0] - a[0]) * (y1 - a[1]) / (b[1] - a[1]), y1))]:
        if not P:
            return 0.0
        P = pass_(P, inside, cut)
    return geo.ring_signed_area(np.array(P)) if len(P) >= 3 else 0.0


def _reference_fixture_weights():
    """aggfly/tests/test_aggregate.py:67-177: hull of 20 seeded points, 4x4 random secondary raster,
    the 2x2 0-360 dataset."""
    arr, t, lat, lon = rc.dataset_360_arrays()
    ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360=True)
    np.random.seed(1216)
    px, py = np.random.uniform(-180, 180, 20), np.random.uniform(-90, 90, 20)
    regions = af.GeoRegions.from_polygons(["region_1"], [geo.convex_hull(np.c_[px, py])])
    np.random.seed(1216)
    x, y = np.linspace(-180, 180, 5), np.linspace(-90, 90, 5)
    sec = SecondaryWeights(np.random.rand(1, 4, 4), (y[1:] + y[:-1]) / 2, (x[1:] + x[:-1]) / 2)
    return ds, regions, sec


def test_reference_known_answers_area_raster_and_final_weights():
    ds, regions, sec = _reference_fixture_weights()
    w = af.weights_from_objects(ds, regions, sec)
    assert w.cosine_area is False and w.zero_weight == "nan"
    w.calculate_weights()
    wdf = w.weights.sort_values("cell_id")
    assert list(wdf.cell_id) == [0, 1, 2, 3] and set(wdf.geoid) == {"region_1"}
    assert np.allclose(wdf.area_weight, [0.68526356, 0.82993589, 0.39051704, 0.82911388])       # :223-226
    assert np.allclose(wdf.raster_weight, [0.67392287, 0.80659155, 0.56727215, 0.38801016])     # :229-232
    assert np.allclose(wdf.weight, rc.FIXTURE_WEIGHTS)                                            # :234-237
    assert list(w.grid.longitude) == [-90.0, 90.0]                                                # relabelled + sorted


def test_nonsquare_grid_circle_matches_true_rectangle_overlap_times_cos_lat():
    """aggfly/tests/test_aggregate.py:838-870 with the expected values from an independent clip."""
    dlon, dlat = 1.25, 1.0
    lon = np.arange(-10, 10, dlon) + dlon / 2
    lat = np.arange(-8, 8, dlat) + dlat / 2
    ang = np.linspace(0, 2 * np.pi, 64, endpoint=False)
    circle = np.c_[5 * np.cos(ang), 5 * np.sin(ang)]
    ds = af.Dataset.from_arrays(np.zeros((1, len(lat), len(lon)), np.float32), pd.date_range("2000-01-01", periods=1),
                                lat, lon, lon_is_360=False)
    w = af.weights_from_objects(ds, af.GeoRegions.from_polygons(["r1"], [circle]))
    assert w.cosine_area is True
    w.calculate_weights()
    wdf = w.weights
    assert ((wdf.area_weight > 1e-9) & (wdf.area_weight < 0.99)).sum() > 5
    want = [abs(_py_clip_area(circle, r.longitude - dlon / 2, r.longitude + dlon / 2, r.latitude - dlat / 2,
                              r.latitude + dlat / 2)) / (dlon * dlat) * np.cos(np.radians(r.latitude))
            for r in wdf.itertuples()]
    assert np.allclose(wdf.area_weight.values, want, rtol=1e-12, atol=0)
    # cells wholly inside are exactly cos(lat); total area is conserved
    inner = wdf[(np.abs(wdf.longitude) < 2) & (np.abs(wdf.latitude) < 2)]
    assert np.array_equal(inner.area_weight.values, np.cos(np.radians(inner.latitude.values)))
    raw = wdf.area_weight / np.cos(np.radians(wdf.latitude))
    assert np.isclose(raw.sum() * dlon * dlat, abs(geo.ring_signed_area(circle)), rtol=1e-12)


def test_holes_multipart_and_descending_latitudes():
    lon = np.arange(-3.5, 4.0, 1.0)
    lat = np.arange(3.5, -4.0, -1.0)                                  # descending, like ERA5
    shell = np.array([[-3.2, -3.2], [3.2, -3.2], [3.2, 3.2], [-3.2, 3.2]])
    hole = np.array([[-1.5, -1.5], [1.5, -1.5], [1.5, 1.5], [-1.5, 1.5]])
    island = np.array([[-0.25, -0.25], [0.25, -0.25], [0.25, 0.25], [-0.25, 0.25]]) + 0.5
    rings = geo.orient_polygon(shell, [hole]) + geo.orient_polygon(island)
    region, cell, frac = geo.cell_overlaps([rings, geo.orient_polygon(island + 10)], lon, lat, 1.0, 1.0)
    assert set(region) == {0}                                         # the second region lies off the grid
    area = dict(zip(cell, frac))
    cid = lambda la, lo: int(np.where(lat == la)[0][0]) * len(lon) + int(np.where(lon == lo)[0][0])   # noqa: E731
    assert area[cid(2.5, 2.5)] == 1.0                                 # interior cell: exactly 1
    assert np.isclose(area[cid(3.5, 0.5)], 0.2) and np.isclose(area[cid(3.5, 3.5)], 0.04)
    assert np.isclose(area[cid(1.5, 1.5)], 0.75)                      # the hole's corner takes a quarter
    assert np.isclose(area[cid(0.5, 0.5)], 0.25)                      # inside the hole: only the island
    assert cid(-0.5, -0.5) not in area                                # inside the hole, no island: absent
    assert np.isclose(sum(frac), 6.4 ** 2 - 9 + 0.25)
    assert list(cell) == sorted(cell)                                 # cell_id ascending within a region


def test_rectangles_and_their_polygon_form_agree():
    lon, lat = 235.0 + 0.25 * np.arange(40), 49.75 - 0.25 * np.arange(30)
    ds = af.Dataset.from_arrays(np.zeros((1, 30, 40), np.float32), pd.date_range("2000-01-01", periods=1), lat, lon, True)
    x0, x1 = np.array([-124.9, -122.0, -119.3]), np.array([-122.0, -119.3, -116.0])
    y0, y1 = np.array([43.1, 44.0, 42.6]), np.array([48.2, 49.9, 47.0])
    rect = af.weights_from_objects(ds, af.GeoRegions.from_rectangles(list("abc"), x0, x1, y0, y1))
    rect.calculate_weights()
    polys = [np.array([[a, c], [b, c], [b, d], [a, d]]) for a, b, c, d in zip(x0, x1, y0, y1)]
    poly = af.weights_from_objects(ds, af.GeoRegions.from_polygons(list("abc"), polys))
    poly.calculate_weights()
    a = rect.weights.sort_values(["index_right", "cell_id"]).reset_index(drop=True)
    b = poly.weights.sort_values(["index_right", "cell_id"]).reset_index(drop=True)
    assert list(a.cell_id) == list(b.cell_id) and list(a.geoid) == list(b.geoid)
    assert np.allclose(a.weight, b.weight, rtol=1e-10, atol=1e-15)


# ---- secondary raster: missing values and the zero_weight policies (reference :1348-1529) --------------
def _box_dataset():
    lat = lon = np.arange(0, 4.0) + 0.5
    ds = af.Dataset.from_arrays(np.ones((3, 4, 4)), pd.date_range("2000-01-01", periods=3), lat, lon, lon_is_360=False)
    box = lambda a, b, c, d: np.array([[a, c], [b, c], [b, d], [a, d]], dtype=float)   # noqa: E731
    return ds, lat, lon, box


def _partial_pop(covered_rows, lat, lon):
    vals = np.ones((1, 4, 4))
    vals[0, covered_rows:, :] = np.nan
    return SecondaryWeights(vals, lat, lon)


def test_missing_raster_values_become_zero_weight_not_nan():
    ds, lat, lon, box = _box_dataset()
    regions = af.GeoRegions.from_polygons(["r1"], [box(0, 4, 0, 4)])
    w = af.weights_from_objects(ds, regions, secondary_weights=_partial_pop(2, lat, lon))
    with pytest.warns(UserWarning, match="no secondary raster value"):
        w.calculate_weights()
    assert not w.weights.weight.isna().any() and (w.weights.weight >= 0).all()
    assert (w.weights.weight > 0).sum() == 8 and np.isclose(w.weights.weight.sum(), 1.0)
    w0 = af.weights_from_objects(ds, regions, secondary_weights=_partial_pop(0, lat, lon))
    with pytest.warns(UserWarning, match="no secondary raster value"):
        w0.calculate_weights()
    assert (w0.weights.weight == 0).all()
    w_area = af.weights_from_objects(ds, regions, secondary_weights=_partial_pop(0, lat, lon), zero_weight="area")
    with pytest.warns(UserWarning):
        w_area.calculate_weights()
    assert np.allclose(w_area.weights.weight, w_area.weights.area_weight)
    clean = af.weights_from_objects(ds, regions, secondary_weights=_partial_pop(4, lat, lon))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        clean.calculate_weights()
    assert (clean.weights.weight > 0).all()


def test_zero_weight_policies_on_a_region_without_population():
    ds, lat, lon, box = _box_dataset()
    regions = af.GeoRegions.from_polygons(["has_pop", "no_pop"], [box(0, 2, 0, 4), box(2, 4, 0, 4)])
    vals = np.ones((4, 4))
    vals[:, 2:] = 0.0
    sw = SecondaryWeights(vals, lat, lon)
    w = af.weights_from_objects(ds, regions, secondary_weights=sw)
    w.calculate_weights()
    assert (w.weights.loc[w.weights.geoid == "no_pop", "weight"] == 0).all()
    assert (w.weights.loc[w.weights.geoid == "has_pop", "weight"] > 0).all()
    w = af.weights_from_objects(ds, regions, secondary_weights=sw, zero_weight="area")
    with pytest.warns(UserWarning, match="fall back to AREA weights"):
        w.calculate_weights()
    assert (w.weights.loc[w.weights.geoid == "no_pop", "weight"] > 0).all()
    w = af.weights_from_objects(ds, regions, secondary_weights=sw, zero_weight="drop")
    with pytest.warns(UserWarning, match="DROPPED"):
        w.calculate_weights()
    assert set(w.weights.geoid) == {"has_pop"}
    with pytest.raises(ValueError, match="zero_weight must be one of"):
        af.weights_from_objects(ds, regions, secondary_weights=sw, zero_weight="bogus")
    with pytest.warns(DeprecationWarning, match="default_to_area_weights is deprecated"):
        assert af.weights_from_objects(ds, regions, secondary_weights=sw, default_to_area_weights=True).zero_weight == "area"


def test_rescale_is_block_mean_when_aligned_and_area_weighted_otherwise():
    src = np.arange(16.0).reshape(4, 4)
    fine = np.arange(0.5, 4.0)
    out = geo.rescale_raster_to_grid(src, fine, fine, np.array([1.0, 3.0]), np.array([1.0, 3.0]), 2.0, 2.0)
    assert np.allclose(out, src.reshape(2, 2, 2, 2).mean(axis=(1, 3)))
    src[0, 0] = np.nan                                                 # nodata pixels are left out of the mean
    out = geo.rescale_raster_to_grid(src, fine, fine, np.array([1.0, 3.0]), np.array([1.0, 3.0]), 2.0, 2.0)
    assert np.isclose(out[0, 0], (1 + 4 + 5) / 3)
    out = geo.rescale_raster_to_grid(np.array([[1.0, 3.0]]), [0.5], [0.5, 1.5], np.array([0.5]), np.array([0.75]), 1.0, 1.5)
    assert np.isclose(out[0, 0], (1.0 * 1.0 + 3.0 * 0.5) / 1.5)       # 1.0 of the first pixel, 0.5 of the second
    assert np.isnan(geo.rescale_raster_to_grid(src, fine, fine, np.array([10.0]), np.array([10.0]), 1.0, 1.0)[0, 0])


# ---- shapefile reader -----------------------------------------------------------------------------------
def _write_shp(path, polygons):
    recs = []
    for i, rings in enumerate(polygons):
        pts = np.concatenate(rings)
        parts = np.cumsum([0] + [len(r) for r in rings[:-1]])
        body = struct.pack("<i4d2i", 5, pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max(), len(rings), len(pts))
        body += struct.pack(f"<{len(rings)}i", *parts) + pts.astype("<f8").tobytes()
        recs.append(struct.pack(">2i", i + 1, len(body) // 2) + body)
    payload = b"".join(recs)
    hdr = struct.pack(">i5ii", 9994, 0, 0, 0, 0, 0, (100 + len(payload)) // 2) + struct.pack("<2i8d", 1000, 5, *([0.0] * 8))
    open(path, "wb").write(hdr + payload)


def _write_dbf(path, names):
    flen = 8
    hdr = struct.pack("<4BIHH20x", 3, 24, 1, 1, len(names), 32 + 32 + 1, 1 + flen)
    field = b"GEOID".ljust(11, b"\x00") + b"C" + b"\x00" * 4 + bytes([flen, 0]) + b"\x00" * 14
    body = b"".join(b" " + n.encode().ljust(flen) for n in names)
    open(path, "wb").write(hdr + field + b"\x0d" + body + b"\x1a")


def test_shapefile_roundtrip_with_attributes(tmp_path):
    sq = np.array([[0, 0], [0, 2], [2, 2], [2, 0], [0, 0]], dtype=float)             # clockwise + closed, as in the format
    hole = np.array([[0.5, 0.5], [1.5, 0.5], [1.5, 1.5], [0.5, 1.5], [0.5, 0.5]], dtype=float)
    _write_shp(tmp_path / "r.shp", [[sq, hole], [sq + 2]])
    _write_dbf(tmp_path / "r.dbf", ["06001", "06003"])
    regions = af.GeoRegions.from_shapefile(str(tmp_path / "r.shp"), regionid="GEOID")
    assert list(regions.shp.GEOID) == ["06001", "06003"] and len(regions.shp.rings[0]) == 2
    lat = lon = np.arange(0.5, 4.0)
    ds = af.Dataset.from_arrays(np.zeros((1, 4, 4)), pd.date_range("2000-01-01", periods=1), lat, lon, lon_is_360=False)
    w = af.weights_from_objects(ds, regions, cosine_area=False)
    w.calculate_weights()
    got = w.weights.groupby("GEOID").area_weight.sum()
    assert np.isclose(got["06001"], 3.0) and np.isclose(got["06003"], 4.0)
    assert (w.weights[w.weights.GEOID == "06001"].area_weight == 0.75).all()


def test_real_shapefile_if_available():
    """The reference ships one real polygon file (benchmarks/data/usa_simple_noHI.shp, no .dbf).  It is
    only present in the build container; elsewhere this test is skipped."""
    import os
    import time
    path = "/root/reference/benchmarks/data/usa_simple_noHI.shp"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    regions = af.GeoRegions.from_shapefile(path)
    assert len(regions.shp) >= 1 and all(len(r) >= 1 for r in regions.shp.rings)
    lat, lon = 49.875 - 0.25 * np.arange(104), 235.125 + 0.25 * np.arange(236)
    ds = af.Dataset.from_arrays(np.zeros((1, 104, 236), np.float32), pd.date_range("2000-01-01", periods=1), lat, lon, True)
    w = af.weights_from_objects(ds, regions, cosine_area=False)
    t0 = time.perf_counter()
    w.calculate_weights()
    assert time.perf_counter() - t0 < 30
    wdf = w.weights
    assert (wdf.area_weight > 0).all() and (wdf.area_weight <= 1.0).all() and (wdf.area_weight == 1.0).any()
    # area conservation, per region: sum of covered fractions x cell area == |polygon area| inside the grid box
    for ridx, rings in zip(regions.shp.index, regions.shp.rings):
        inside = sum(_py_clip_area(r, -125.0, -66.0, 24.0, 50.0) for r in rings)
        got = wdf.loc[wdf.index_right == ridx, "area_weight"].sum() * 0.0625
        assert np.isclose(got, abs(inside), rtol=1e-9), (ridx, got, inside)


def test_geotiff_secondary_raster_equals_the_npz_one(tmp_path):
    """Secondary rasters usually come as GeoTIFF (aggfly/weights/secondary_weights.py:201-245; the reference's
    example config points at a LandScan .tif): tiepoint / pixel-scale georeferencing, GDAL_NODATA, LZW / deflate."""
    from PIL import Image, TiffImagePlugin
    from aggfly_b200 import io
    from aggfly_b200.dataset import Grid
    rng = np.random.default_rng(4)
    vals = rng.random((16, 20)).astype(np.float32) * 100
    vals[3, 5] = -9999.0
    lat, lon = 40.125 - 0.25 * np.arange(16), -110.125 + 0.25 * np.arange(20)
    np.savez(tmp_path / "pop.npz", values=vals, latitude=lat, longitude=lon)
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550], ifd.tagtype[33550] = (0.25, 0.25, 0.0), 12                                  # ModelPixelScale
    ifd[33922], ifd.tagtype[33922] = (0.0, 0.0, 0.0, -110.25, 40.25, 0.0), 12                # ModelTiepoint: outer corner
    ifd[42113], ifd.tagtype[42113] = "-9999", 2                                             # GDAL_NODATA
    grid = Grid(-110.0 + 0.5 * np.arange(9), 40.0 - 0.5 * np.arange(7))
    want = io.secondary_weights_from_path(str(tmp_path / "pop.npz"), nodata=-9999.0)
    for comp in ("tiff_lzw", "tiff_adobe_deflate", None):
        path = str(tmp_path / f"pop_{comp}.tif")
        Image.fromarray(vals).save(path, tiffinfo=ifd, compression=comp)
        got = io.secondary_weights_from_path(path)
        assert got.nodata == -9999.0 and np.array_equal(got.values, vals.astype(float))
        assert np.allclose(got.latitude, lat) and np.allclose(got.longitude, lon)
        assert np.array_equal(got.on_grid(grid), want.on_grid(grid), equal_nan=True)
    Image.fromarray(vals).save(str(tmp_path / "bare.tif"))
    with pytest.raises(ValueError, match="georeferencing"):
        io.secondary_weights_from_path(str(tmp_path / "bare.tif"))
    ifd[33922] = (0.0, 0.0, 0.0, 500000.0, 4400000.0, 0.0)                                   # UTM-like coordinates
    Image.fromarray(vals).save(str(tmp_path / "utm.tif"), tiffinfo=ifd)
    with pytest.raises(NotImplementedError, match="projected"):
        io.secondary_weights_from_path(str(tmp_path / "utm.tif"))
