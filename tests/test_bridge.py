"""Adapters from reference-shaped objects (INTEGRATION.md route A) -- host logic only, no GPU."""
from types import SimpleNamespace

import numpy as np
import pandas as pd

from aggfly_b200 import bridge
from aggfly_b200.timeaxis import CalendarIndex


class _DA:
    """Just enough of xarray.DataArray: dims, transpose(...).values, get_index."""

    def __init__(self, values, dims, coords):
        self._v, self.dims, self._c = values, tuple(dims), coords

    def transpose(self, *dims):
        return _DA(np.transpose(self._v, [self.dims.index(d) for d in dims]), dims, self._c)

    @property
    def values(self):
        return self._v

    def get_index(self, d):
        return self._c[d]


def test_dataset_adapter_moves_time_first_and_keeps_axes():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(2, 3, 5)).astype(np.float32)                 # (lat, lon, time) like clean_dims
    t = pd.date_range("2000-01-01", periods=5, freq="h")
    ref = SimpleNamespace(da=_DA(v, ("latitude", "longitude", "time"),
                                 {"time": t, "latitude": np.array([1.0, 0.0]), "longitude": np.array([10.0, 190.0, 350.0])}),
                          lon_is_360=True, name="t2m")
    ds = bridge.dataset_from_aggfly(ref)
    assert ds.shape == (5, 2, 3) and ds.dtype == np.float32 and ds.lon_is_360
    assert np.array_equal(ds.values, np.transpose(v, (2, 0, 1)))
    assert list(ds.time) == list(t) and list(ds.latitude) == [1.0, 0.0]
    assert bridge.dataset_from_aggfly(ds) is ds


def test_dataset_adapter_reads_cftime_like_index_fields():
    n = 4
    idx = SimpleNamespace(calendar="noleap", year=np.full(n, 2001), month=np.full(n, 2), day=np.arange(26, 30) % 28 + 1,
                          hour=np.zeros(n, int))
    t = bridge._time_index(idx)
    assert isinstance(t, CalendarIndex) and len(t) == n and t.calendar == "noleap"


def test_weights_adapter_carries_frame_ids_policy_and_cell_numbering():
    wdf = pd.DataFrame({"cell_id": [5, 7, 9], "index_right": [3, 3, 8], "weight": [0.5, 0.5, 1.0]})
    shp = pd.DataFrame({"GEOID": ["a", "b"], "geometry": [None, None]}, index=[3, 8])
    ref = SimpleNamespace(weights=wdf, zero_weight="drop",
                          grid=SimpleNamespace(longitude=np.array([-170.0, -10.0, 10.0]), latitude=np.array([1.0, 0.0]),
                                               cell_id=np.array([4, 5, 6, 7, 8, 9]), lon_is_360=False),
                          georegions=SimpleNamespace(shp=shp, regionid="GEOID"))
    w = bridge.weights_from_aggfly(ref)
    assert w.zero_weight == "drop" and w.weights is wdf
    assert list(w.grid.cell_id) == [4, 5, 6, 7, 8, 9]
    assert list(w.georegions.shp.index) == [3, 8] and list(w.georegions.shp.columns) == ["GEOID"]
    assert bridge.weights_from_aggfly(w) is w
