"""Light data model for the hot path: ``Dataset`` (raster + axes) and ``Grid`` (cell ids).

The reference wraps an ``xarray.DataArray`` (aggfly/dataset/dataset.py:21-118) and only uses it,
on this path, for: the values, the time index, the lat/lon vectors, ``lon_is_360`` /
``rescale_longitude`` (:419-440) and ``grid.cell_id`` (aggfly/dataset/grid.py:74-80).  xarray is
not a dependency here: a ``Dataset`` holds a time-major ``values[time, lat, lon]`` array (numpy,
pinned/pageable torch CPU tensor, or torch CUDA tensor) plus plain axes.  If xarray is
importable, a ``DataArray`` is accepted and unwrapped.
"""
from __future__ import annotations

from copy import copy
from typing import Any, Optional, Sequence

import numpy as np
import pandas as pd

from .preprocess import FusedPreprocess, constants_for, resolve
from .timeaxis import CalendarIndex


def lon_to_180(longitude):
    """aggfly/dataset/grid_utils.py:16-31."""
    return (np.asarray(longitude, dtype=float) + 180) % 360 - 180


def lon_to_360(longitude):
    """aggfly/dataset/grid_utils.py:34-49."""
    lon = np.asarray(longitude, dtype=float)
    return (lon < 0) * (lon + 360) + (lon >= 0) * lon


class Grid:
    """Cell numbering of a lat x lon grid (aggfly/dataset/grid.py:56-87, 137-147):
    ``cell_id = lat_index * n_lon + lon_index`` over the grid's own (lat, lon) order."""

    def __init__(self, longitude, latitude, name=None, lon_is_360=False):
        self.longitude = np.asarray(longitude, dtype=float)
        self.latitude = np.asarray(latitude, dtype=float)
        self.name = name
        self.lon_is_360 = lon_is_360
        self.index = np.arange(len(self.latitude) * len(self.longitude)).reshape(
            len(self.latitude), len(self.longitude))
        self.cell_id = self.index.flatten()
        self.resolution_lon, self.resolution_lat = self.get_resolution()

    def get_resolution(self):
        res_lon = abs(np.diff(self.longitude).mean()) if len(self.longitude) > 1 else 0.0
        res_lat = abs(np.diff(self.latitude).mean()) if len(self.latitude) > 1 else 0.0
        if res_lon == 0.0:
            res_lon = res_lat
        if res_lat == 0.0:
            res_lat = res_lon
        return res_lon, res_lat

    @property
    def resolution(self):
        return max(self.resolution_lon, self.resolution_lat)


class RasterArray:
    """Minimal stand-in for the ``xarray.DataArray`` the reference's ``Dataset`` wraps."""

    def __init__(self, data, dims: Sequence[str], coords: dict):
        self.data, self.dims, self.coords = data, tuple(dims), dict(coords)


class TimeConcat:
    """Lazy concatenation along time of rasters ``[T_i, lat, lon]`` that live in different files (NumPy arrays,
    memory maps, ``zarrio.ChunkedRaster``): what ``xr.open_mfdataset`` gives the reference for a list / glob of
    paths (aggfly/dataset/dataset.py:686-695).  Slicing with step-1 slices stays lazy (a slice inside one part IS
    that part's own slice); NumPy sees the data through ``__array__``, so the host feed's worker threads read /
    decode only the rows of the chunk they stage."""

    lazy_rows = True
    ndim = 3

    def __init__(self, parts):
        parts = [p for p in parts if p.shape[0] > 0] or list(parts[:1])
        if not parts:
            raise ValueError("nothing to concatenate")
        tail = {tuple(p.shape[1:]) for p in parts}
        if len(tail) != 1:
            raise ValueError(f"parts have different grids: {sorted(tail)}")
        dts = {np.dtype(p.dtype) for p in parts}
        self.dtype = np.dtype(np.float64) if (len(dts) > 1 or next(iter(dts)) not in (np.float32, np.float64)) else next(iter(dts))
        self.parts = list(parts)
        self.starts = np.concatenate([[0], np.cumsum([p.shape[0] for p in parts])]).astype(np.int64)

    @property
    def shape(self):
        return (int(self.starts[-1]),) + tuple(self.parts[0].shape[1:])

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) <= 3 and all(isinstance(k, slice) and k.step in (None, 1) for k in key):
            a, b, _ = key[0].indices(self.shape[0])
            b = max(a, b)
            rest = tuple(key[1:])
            out = []
            for p, s0, s1 in zip(self.parts, self.starts[:-1], self.starts[1:]):
                lo, hi = max(a, int(s0)), min(b, int(s1))
                if hi > lo:
                    out.append(p[(slice(lo - int(s0), hi - int(s0)),) + rest])
            if len(out) == 1 and np.dtype(out[0].dtype) == self.dtype:
                return out[0]
            if not out:
                out = [self.parts[0][(slice(0, 0),) + rest]]
            return TimeConcat(out)
        return np.asarray(self)[key]

    def __array__(self, dtype=None, copy=None):
        out = np.concatenate([np.asarray(p, dtype=self.dtype) for p in self.parts], axis=0)
        return out if dtype is None else out.astype(dtype, copy=False)

    def __repr__(self):
        return f"<TimeConcat {self.shape} {self.dtype} of {len(self.parts)} parts>"


class PackedRaster:
    """CF-packed integers ``stored[T, lat, lon]`` (int16 / int32 / uint8 ...) in HOST memory with their
    ``scale_factor`` / ``add_offset`` / ``_FillValue`` -- what an ERA5 NetCDF variable is on disk, and what
    ``xr.open_dataset(..., mask_and_scale=True)`` decodes lazily for the reference (aggfly/dataset/dataset.py:700-707).
    Here the stored integers cross PCIe as they are (2 bytes per value instead of 4) and are decoded ON THE DEVICE by
    ``agf_tile_place_run`` (``stored * scale_factor + add_offset`` in double, then rounded to ``dtype``; fill -> NaN)
    while the next chunk is copied (stream.feed_packed).  ``stored`` may be a NumPy array or a (pinned) CPU torch
    tensor; slicing along time stays lazy."""

    is_packed_raster = True
    ndim = 3

    def __init__(self, stored, scale_factor: float = 1.0, add_offset: float = 0.0, fill_value=None, dtype=np.float32):
        self.stored = stored
        self.scale, self.offset = float(scale_factor), float(add_offset)
        self.fill = None if fill_value is None else float(fill_value)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.float32, np.float64):
            raise TypeError("decoded dtype must be float32 or float64")
        if len(stored.shape) != 3:
            raise ValueError("stored values must be [time, lat, lon]")

    @property
    def shape(self):
        return tuple(int(v) for v in self.stored.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def __len__(self):
        return self.shape[0]

    def stored_numpy(self) -> np.ndarray:
        return self.stored.numpy() if _is_torch(self.stored) else np.asarray(self.stored)

    def __getitem__(self, key):
        k0 = key[0] if isinstance(key, tuple) else key
        rest = key[1:] if isinstance(key, tuple) else ()
        if isinstance(k0, slice) and k0.step in (None, 1) and all(isinstance(k, slice) and k == slice(None) for k in rest):
            return PackedRaster(self.stored[k0], self.scale, self.offset, self.fill, self.dtype)
        return np.asarray(self)[key]

    def __array__(self, dtype=None, copy=None):
        """Host materialisation (tests, small cases): the arithmetic agf_tile_place_run does on the device."""
        raw = self.stored_numpy()
        out = (raw.astype(np.float64) * self.scale + self.offset).astype(self.dtype)
        if self.fill is not None:
            out[raw.astype(np.float64) == self.fill] = np.nan
        return out if dtype is None else out.astype(dtype, copy=False)

    def __repr__(self):
        return f"<PackedRaster {self.shape} {self.stored.dtype} -> {self.dtype} scale={self.scale} offset={self.offset}>"


def time_selection(time, time_sel):
    """Row range [a, b) of ``.sel(time=time_sel)`` on a sorted axis: a partial date string ("2001",
    "2001-06", "2001-06-15"), a timestamp, or a slice of those (both ends inclusive, like pandas/xarray)."""
    if isinstance(time, CalendarIndex):
        key = str(time_sel.start if isinstance(time_sel, slice) else time_sel)
        if isinstance(time_sel, slice) or len(key) != 4:
            raise NotImplementedError("time_sel on a non-standard calendar supports a single year ('YYYY')")
        idx = np.nonzero(np.asarray(time.year) == int(key))[0]
        return (int(idx[0]), int(idx[-1]) + 1) if len(idx) else (0, 0)
    t = pd.DatetimeIndex(time)
    if isinstance(time_sel, slice):
        sl = t.slice_indexer(time_sel.start, time_sel.stop)
    else:
        loc = t.get_loc(time_sel if not isinstance(time_sel, (int, np.integer)) else str(time_sel))
        sl = loc if isinstance(loc, slice) else slice(int(loc), int(loc) + 1) if np.ndim(loc) == 0 else None
        if sl is None:                                            # boolean mask (non-unique axis)
            idx = np.nonzero(loc)[0]
            sl = slice(int(idx[0]), int(idx[-1]) + 1)
    return int(sl.start or 0), int(len(t) if sl.stop is None else sl.stop)


def _fusable(preprocess):
    """FusedPreprocess for names / expressions / FusedPreprocess objects, None for other callables."""
    if isinstance(preprocess, (str, FusedPreprocess)):
        return resolve(preprocess) or FusedPreprocess([])
    return None


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Dataset:
    """Raster + axes.  ``Dataset(da, xycoords, timecoord, lon_is_360=True, preprocess=None, ...)``
    keeps the reference's constructor (aggfly/dataset/dataset.py:47-118) for a ``RasterArray`` /
    ``xarray.DataArray``; ``Dataset.from_arrays`` is the direct route."""

    def __init__(self, da, xycoords=("longitude", "latitude"), timecoord="time", time_sel=None,
                 lon_is_360=True, preprocess=None, georegions=None, time_fix=False, name=None):
        if not isinstance(da, RasterArray):
            da = _from_xarray(da)
        xdim, ydim = xycoords
        dims = list(da.dims)
        for need in (xdim, ydim, timecoord):
            if need not in dims:
                raise ValueError(f"dimension {need!r} not found in {dims}")
        values = da.data
        order = [dims.index(timecoord), dims.index(ydim), dims.index(xdim)]
        if order != [0, 1, 2]:
            values = values.permute(*order).contiguous() if _is_torch(values) else np.transpose(values, order)
        fused = None
        if preprocess is not None:
            fused = _fusable(preprocess)
            if fused is None:
                values = preprocess(values)              # arbitrary callable: applied to the array now
        self._init(values, da.coords[timecoord], da.coords[ydim], da.coords[xdim], lon_is_360, name)
        if time_sel is not None:                                  # da.sortby("time").sel(time=time_sel), dataset.py:90-91
            a, b = time_selection(self.time, time_sel)
            self.values, self.time = self.values[a:b], self.time[a:b]
        self.georegions = georegions
        if fused is not None:
            self.pre_ops = constants_for(fused.ops, self.dtype)

    def _init(self, values, time, latitude, longitude, lon_is_360, name):
        if not isinstance(time, CalendarIndex):
            time = pd.DatetimeIndex(time)
        if not time.is_monotonic_increasing:                          # da.sortby("time"), dataset.py:87
            if isinstance(time, CalendarIndex):
                order = np.argsort(time.ordinal_hours(), kind="stable")
            else:
                order = np.argsort(time.values, kind="stable")
            time = time[order]
            values = values[order] if not _is_torch(values) else values[list(order)]
        if (not _is_torch(values) and not getattr(values, "is_chunked_raster", False) and not getattr(values, "lazy_rows", False)
                and not getattr(values, "is_packed_raster", False)):
            values = np.asarray(values)
            if values.dtype not in (np.float32, np.float64):
                values = values.astype(np.float64)
        self.values = values
        self.time = time
        self.latitude = np.asarray(latitude, dtype=float)
        self.longitude = np.asarray(longitude, dtype=float)
        self.lon_is_360 = bool(lon_is_360)
        self.name = name
        self.history = []
        self.georegions = None
        self.pre_ops = []             # fused preprocess chain (preprocess.py), applied by the kernels
        if tuple(self.values.shape) != (len(self.time), len(self.latitude), len(self.longitude)):
            raise ValueError(f"values shape {tuple(self.values.shape)} does not match axes "
                             f"({len(self.time)}, {len(self.latitude)}, {len(self.longitude)})")
        self.grid = Grid(self.longitude, self.latitude, name, self.lon_is_360)

    @classmethod
    def from_arrays(cls, values, time, latitude, longitude, lon_is_360=True, name=None, preprocess=None) -> "Dataset":
        """values[time, lat, lon] (numpy or torch, float32/float64), time = DatetimeIndex | CalendarIndex.
        ``preprocess``: builtin name (``"kelvin_to_celsius"``), arithmetic expression in ``x`` or a
        ``FusedPreprocess`` -- fused into the temporal kernel, the stored raster stays untouched."""
        self = cls.__new__(cls)
        self._init(values, time, latitude, longitude, lon_is_360, name)
        fused = None if preprocess is None else _fusable(preprocess)
        if preprocess is not None and fused is None:
            raise TypeError("from_arrays(preprocess=...) takes a builtin name, an expression in x or a "
                            "FusedPreprocess; apply other callables to the array yourself")
        if fused is not None:
            self.pre_ops = constants_for(fused.ops, self.dtype)
        return self

    # -- reference API surface used on the hot path ----------------------------------------------
    @property
    def da(self) -> RasterArray:
        return RasterArray(self.values, ("time", "latitude", "longitude"),
                           {"time": self.time, "latitude": self.latitude, "longitude": self.longitude})

    @property
    def dtype(self) -> np.dtype:
        v = self.values
        if _is_torch(v):
            import torch
            return np.dtype(np.float32 if v.dtype == torch.float32 else np.float64)
        return v.dtype

    @property
    def shape(self):
        return tuple(self.values.shape)

    def deepcopy(self) -> "Dataset":
        new = copy(self)                      # the raster itself is immutable on this path
        new.history = list(self.history)
        return new

    def isel_time(self, row0: int, row1: int) -> "Dataset":
        """Rows [row0, row1) of the time axis (a view of the raster, no copy)."""
        new = copy(self)
        new.values = self.values[row0:row1]
        new.time = self.time[row0:row1]
        new.history = list(self.history)
        return new

    # -- elementwise transforms (aggfly/dataset/dataset.py:442-518), run by the library's kernel ---------
    def _transformed(self, params: dict):
        from .aggregate import aggregate_time
        out = aggregate_time(self, None, {self.name or "x": [("transform", params)]})
        return list(out.values())

    def power(self, exp, update: bool = False) -> Optional["Dataset"]:
        """``np.power(values, exp)`` with NumPy's promotion (a numpy-integer exponent makes float32 data
        float64, a Python int does not)."""
        new = self._transformed({"transform": "power", "exp": [[exp]]})[0]
        new.history = list(self.history) + [f"power{exp}"]
        return self._maybe_update(new, update)

    def interact(self, inter, update: bool = False) -> Optional["Dataset"]:
        """Multiply by another Dataset / array of the same shape."""
        new = self._transformed({"inter": inter})[0]
        new.history = list(self.history) + ["interacted"]
        return self._maybe_update(new, update)

    def spline(self):
        """[x, (x > 20) * (x - 20)]  (aggfly/dataset/dataset.py:475-481)."""
        return self._transformed({"transform": "spline"})

    def _maybe_update(self, new: "Dataset", update: bool):
        if update:
            self.values, self.history = new.values, new.history
            return None
        return new

    def lon_sort_order(self) -> np.ndarray:
        """Column order the reference's ``rescale_longitude`` puts a 0-360 raster in
        (relabel to -180..180, then ``sortby('longitude')``; dataset.py:419-440,
        grid_utils.py:52-73).  Identity for a -180..180 dataset."""
        if not self.lon_is_360:
            return np.arange(len(self.longitude))
        return np.argsort(lon_to_180(self.longitude), kind="stable")

    def rescale_longitude(self) -> None:
        """Relabel + sort longitudes (a host-side, metadata + column permutation operation)."""
        if self.lon_is_360:
            order = np.argsort(lon_to_180(self.longitude), kind="stable")
            self.longitude = lon_to_180(self.longitude)[order]
            self.lon_is_360 = False
        else:
            order = np.argsort(lon_to_360(self.longitude), kind="stable")
            self.longitude = lon_to_360(self.longitude)[order]
            self.lon_is_360 = True
        self.values = self.values[..., list(order)] if _is_torch(self.values) else self.values[..., order]
        self.grid = Grid(self.longitude, self.latitude, self.name, self.lon_is_360)

    def __repr__(self):
        return (f"<aggfly_b200.Dataset {self.name or ''} time={len(self.time)} lat={len(self.latitude)} "
                f"lon={len(self.longitude)} dtype={self.dtype} lon_is_360={self.lon_is_360}>")


def _from_xarray(da: Any) -> RasterArray:
    try:
        import xarray as xr                                    # optional
    except Exception as exc:                                   # pragma: no cover
        raise TypeError("Dataset expects a RasterArray (xarray is not installed)") from exc
    if not isinstance(da, xr.DataArray):
        raise TypeError(f"Dataset expects a RasterArray or xarray.DataArray, got {type(da)}")
    coords = {}
    for d in da.dims:
        idx = da.get_index(d)
        if type(idx).__name__ == "CFTimeIndex":
            idx = CalendarIndex(idx.calendar, idx.year, idx.month, idx.day, idx.hour)
        coords[d] = idx
    return RasterArray(da.values, da.dims, coords)
