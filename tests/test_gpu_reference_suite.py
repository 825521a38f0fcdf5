"""End-to-end scenarios of the reference's own test-suite (aggfly/tests/test_aggregate.py) restated
without xarray / geopandas and run through the CUDA engine: zero_weight policies (:1458-1529), missing
secondary-raster values (:1360-1418), cosine-area vs population weighting (:936-991), the golden panel
from weights computed HERE (:283-313 with :191-237)."""
import warnings

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

import aggfly_b200 as af
from aggfly_b200 import geometry as geo
from aggfly_b200.weights import SecondaryWeights
from tests import refcases as rc

TAVG = dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"})])


@pytest.fixture(autouse=True)
def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _box(a, b, c, d):
    return np.array([[a, c], [b, c], [b, d], [a, d]], dtype=float)


def _two_regions_one_empty(arr=None):
    lat = lon = np.arange(0, 4.0) + 0.5
    arr = np.ones((2, 4, 4)) if arr is None else arr
    ds = af.Dataset.from_arrays(arr, pd.date_range("2000-01-01", periods=2), lat, lon, lon_is_360=False)
    gr = af.GeoRegions.from_polygons(["has_pop", "no_pop"], [_box(0, 2, 0, 4), _box(2, 4, 0, 4)])
    vals = np.ones((4, 4))
    vals[:, 2:] = 0.0
    return ds, gr, SecondaryWeights(vals, lat, lon)


def test_golden_panel_from_weights_built_here():
    """The reference's end-to-end numbers with nothing hand-fed: hull region, secondary raster, 0-360 grid."""
    arr, t, lat, lon = rc.dataset_360_arrays()
    ds = af.Dataset.from_arrays(arr, t, lat, lon, lon_is_360=True)
    np.random.seed(1216)
    px, py = np.random.uniform(-180, 180, 20), np.random.uniform(-90, 90, 20)
    regions = af.GeoRegions.from_polygons(["region_1"], [geo.convex_hull(np.c_[px, py])])
    np.random.seed(1216)
    x, y = np.linspace(-180, 180, 5), np.linspace(-90, 90, 5)
    sec = SecondaryWeights(np.random.rand(1, 4, 4), (y[1:] + y[:-1]) / 2, (x[1:] + x[:-1]) / 2)
    w = af.weights_from_objects(ds, regions, sec)
    w.calculate_weights()
    df = af.aggregate_dataset(dataset=ds, weights=w, **rc.golden_panel_spec())
    assert list(df.columns) == ["geoid", "time", "tavg_1", "tavg_2"]
    assert np.allclose(df[["tavg_1", "tavg_2"]].values, rc.GOLDEN_PANEL)             # :311-313


def test_zero_weight_policies_end_to_end():
    ds, gr, sw = _two_regions_one_empty()
    w = af.weights_from_objects(ds, gr, secondary_weights=sw)
    w.calculate_weights()
    df = af.aggregate_dataset(dataset=ds, weights=w, **TAVG)
    assert set(df.geoid) == {"has_pop", "no_pop"}                                      # default "nan": kept, reported NaN
    assert df.loc[df.geoid == "no_pop", "tavg"].isna().all() and df.loc[df.geoid == "has_pop", "tavg"].notna().all()
    w = af.weights_from_objects(ds, gr, secondary_weights=sw, zero_weight="area")
    with pytest.warns(UserWarning, match="fall back to AREA weights"):
        w.calculate_weights()
    df = af.aggregate_dataset(dataset=ds, weights=w, **TAVG)
    assert set(df.geoid) == {"has_pop", "no_pop"} and df.tavg.notna().all()
    w = af.weights_from_objects(ds, gr, secondary_weights=sw, zero_weight="drop")
    with pytest.warns(UserWarning, match="DROPPED"):
        w.calculate_weights()
    assert set(af.aggregate_dataset(dataset=ds, weights=w, **TAVG).geoid) == {"has_pop"}


def test_nan_policy_still_drops_rows_with_missing_climate_data():
    arr = np.ones((2, 4, 4))
    arr[1] = np.nan
    ds, gr, sw = _two_regions_one_empty(arr)
    w = af.weights_from_objects(ds, gr, secondary_weights=sw)
    w.calculate_weights()
    df = af.aggregate_dataset(dataset=ds, weights=w, **TAVG)
    assert len(df[df.geoid == "has_pop"]) == 1                                          # its all-NaN day is dropped
    empty = df[df.geoid == "no_pop"]
    assert len(empty) == 2 and empty.tavg.isna().all()                                 # the empty region keeps both days


def test_missing_raster_values_do_not_drop_the_region():
    lat = lon = np.arange(0, 4.0) + 0.5
    ds = af.Dataset.from_arrays(np.ones((3, 4, 4)), pd.date_range("2000-01-01", periods=3), lat, lon, lon_is_360=False)
    vals = np.ones((1, 4, 4))
    vals[0, 2:, :] = np.nan
    w = af.weights_from_objects(ds, af.GeoRegions.from_polygons(["r1"], [_box(0, 4, 0, 4)]),
                                secondary_weights=SecondaryWeights(vals, lat, lon))
    with pytest.warns(UserWarning, match="no secondary raster value"):
        w.calculate_weights()
    df = af.aggregate_dataset(dataset=ds, weights=w, **TAVG)
    assert len(df) == 3 and df.tavg.notna().all()


def _lat_span():
    lon, lat = np.arange(0, 3, 1.0) + 0.5, np.arange(0, 30, 1.0) + 0.5
    arr = np.broadcast_to(lat[None, :, None], (2, len(lat), len(lon))).copy()           # temperature == latitude
    ds = af.Dataset.from_arrays(arr, pd.date_range("2000-01-01", periods=2, freq="D"), lat, lon, lon_is_360=False)
    return ds, af.GeoRegions.from_polygons(["r1"], [_box(0, 3, 0, 30)]), lat


def _tavg(ds, w):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w.calculate_weights()
    return float(af.aggregate_dataset(dataset=ds, weights=w, **TAVG).tavg.iloc[0])


def test_cosine_area_and_population_weighting_relations():
    ds, gr, lat = _lat_span()
    lon_src, lat_src = np.arange(0, 3, 0.25) + 0.125, np.arange(0, 30, 0.25) + 0.125
    ones = SecondaryWeights(np.ones((len(lat_src), len(lon_src))), lat_src, lon_src)
    assert af.weights_from_objects(ds, gr).cosine_area is True
    assert af.weights_from_objects(ds, gr, secondary_weights=ones).cosine_area is False
    assert af.weights_from_objects(ds, gr, secondary_weights=ones, cosine_area=True).cosine_area is True
    # population uniform per unit of physical area (pixel counts ~ cos(lat)) == plain area weighting   (:952-975)
    uniform = SecondaryWeights(np.broadcast_to(np.cos(np.radians(lat_src))[:, None], (len(lat_src), len(lon_src))).copy(),
                               lat_src, lon_src)
    a = _tavg(ds, af.weights_from_objects(ds, gr, secondary_weights=uniform))
    b = _tavg(ds, af.weights_from_objects(ds, gr))
    assert np.isclose(a, b, atol=1e-4)
    assert b < float(lat.mean())                                                      # cos(lat) favours low latitudes
    # equal population in every cell == the unweighted mean of the cells                               (:978-991)
    assert np.isclose(_tavg(ds, af.weights_from_objects(ds, gr, secondary_weights=ones)), float(lat.mean()), atol=1e-4)
