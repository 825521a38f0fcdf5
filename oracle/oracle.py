"""CPU oracle for aggfly's ``aggregate_dataset`` path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``aggfly_b200``) never does; it fails loudly when
its CUDA library is missing instead of falling back to anything here.

Parity status: PINNED -- see ``tests/test_oracle.py``: the arithmetic (``oracle/agf_oracle.c``) is
checked against outputs of the reference's own numba kernels (``tests/golden/ref_kernels.npz``,
made by ``tests/golden/make_golden.py``) and the chain / spatial / panel logic below against the
hard-coded expectations of the reference's test-suite (``aggfly/tests/test_aggregate.py:275-280,
311-313, 454-466, 494-514, 614-664``).

What is restated (reference file:line):

* ``resample_groups``        <- aggfly/aggregate/nb_kernels.py:80-115 (pandas branch verbatim in
  spirit: ``Series.resample(freq).count()``; the cftime branch is restated with integer
  calendar arithmetic because neither xarray nor cftime exist in this image)
* ``numba_resample``         <- aggfly/aggregate/nb_kernels.py:271-305 (dtype rule :260,:266)
* ``TemporalStep``           <- aggfly/aggregate/temporal.py:57-163, 221-263, 441-456
* ``aggregate_time``         <- aggfly/aggregate/aggregate.py:36-78, 101-162, 285-303
* ``aggregate_space``        <- aggfly/aggregate/spatial.py:57-199 (lon rescale: dataset/dataset.py:419-440,
  dataset/grid_utils.py:16-73; cell ids: dataset/grid.py:74-80,137-147)
* ``aggregate_dataset``      <- aggfly/aggregate/aggregate.py:210-282
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libagf_oracle.so")
_lib = None

STAT_CODE = {"mean": 0, "sum": 1, "min": 2, "max": 3, "nanmean": 4}   # nb_kernels.py:33
FREQ = {"date": "1D", "month": "ME", "year": "YE", "week": "W"}       # temporal.py:456


def build(force: bool = False) -> str:
    """Compile oracle/agf_oracle.c -> oracle/libagf_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "agf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       env={k: v for k, v in os.environ.items() if k != "CC"})
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        for suf in ("f32", "f64"):
            getattr(L, f"orc_block_stat_{suf}").argtypes = [p, i64, i64, p, i64, ctypes.c_int, p]
            for k in ("dd", "bins", "sine_dd"):
                getattr(L, f"orc_block_{k}_{suf}").argtypes = [p, i64, i64, p, i64, p, i64, p]
        L.orc_scatter_block.argtypes = [p, i64, p, p, p, i64, i64, p]
        L.orc_max_threads.restype = ctypes.c_int
        L.orc_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype) -> str:
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle kernels take float32/float64 cubes, got {dtype}")


# --------------------------------------------------------------------------------------------
# kernels (thin wrappers over agf_oracle.c)
# --------------------------------------------------------------------------------------------
def block_stat(cube: np.ndarray, bounds: np.ndarray, calc: str) -> np.ndarray:
    cube = np.ascontiguousarray(cube)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    T, NY, NX = cube.shape
    G = len(bounds) - 1
    out = np.empty((G, NY, NX), cube.dtype)          # nb_kernels.py:260 -- input dtype preserved
    getattr(lib(), f"orc_block_stat_{_suffix(cube.dtype)}")(
        _ptr(cube), NY, NX, _ptr(bounds), G, STAT_CODE[calc], _ptr(out))
    return out


def _block_dd_like(kind: str, cube, bounds, ddargs) -> np.ndarray:
    cube = np.ascontiguousarray(cube)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    dda = np.ascontiguousarray(np.atleast_2d(np.asarray(ddargs, dtype=np.float64)))  # :295
    T, NY, NX = cube.shape
    G, D = len(bounds) - 1, dda.shape[0]
    out = np.empty((G, NY, NX, D), cube.dtype)       # nb_kernels.py:266
    getattr(lib(), f"orc_block_{kind}_{_suffix(cube.dtype)}")(
        _ptr(cube), NY, NX, _ptr(bounds), G, _ptr(dda), D, _ptr(out))
    return out


def block_dd(cube, bounds, ddargs):
    return _block_dd_like("dd", cube, bounds, ddargs)


def block_bins(cube, bounds, ddargs):
    return _block_dd_like("bins", cube, bounds, ddargs)


def block_sine_dd(cube, bounds, ddargs):
    return _block_dd_like("sine_dd", cube, bounds, ddargs)


def scatter_block(block, region_idx, cell_idx, w_vals, n_regions) -> np.ndarray:
    """spatial.py:181-186.  ``block`` may be f32; the product w*block is formed in fp64."""
    block = np.ascontiguousarray(block, dtype=np.float64)
    region_idx = np.ascontiguousarray(region_idx, dtype=np.int64)
    cell_idx = np.ascontiguousarray(cell_idx, dtype=np.int64)
    w_vals = np.ascontiguousarray(w_vals, dtype=np.float64)
    out = np.empty((n_regions, block.shape[1]), np.float64)
    lib().orc_scatter_block(_ptr(block), block.shape[1], _ptr(region_idx), _ptr(cell_idx),
                            _ptr(w_vals), len(w_vals), n_regions, _ptr(out))
    return out


# --------------------------------------------------------------------------------------------
# time axes and group bounds
# --------------------------------------------------------------------------------------------
@dataclass
class CalTime:
    """A non-standard-calendar time axis (stand-in for xarray.CFTimeIndex): one entry per step."""
    calendar: str                      # "noleap" | "360_day"
    year: np.ndarray
    month: np.ndarray
    day: np.ndarray
    hour: np.ndarray = None

    def __post_init__(self):
        self.year = np.asarray(self.year, dtype=np.int64)
        self.month = np.asarray(self.month, dtype=np.int64)
        self.day = np.asarray(self.day, dtype=np.int64)
        self.hour = (np.zeros_like(self.year) if self.hour is None
                     else np.asarray(self.hour, dtype=np.int64))

    def __len__(self):
        return len(self.year)

    def keys(self) -> np.ndarray:
        return ((self.year * 13 + self.month) * 32 + self.day) * 24 + self.hour


_NOLEAP_MDAYS = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])


def _cal_mdays(calendar: str) -> np.ndarray:
    if calendar == "noleap":
        return _NOLEAP_MDAYS
    if calendar == "360_day":
        return np.full(12, 30)
    raise ValueError(f"unsupported calendar {calendar!r}")


def cal_range(calendar: str, start_year: int, ndays: int) -> CalTime:
    """Daily axis starting on Jan 1 of start_year (what the reference tests build)."""
    md = _cal_mdays(calendar)
    y, m, d = [], [], []
    yy, mm, dd = start_year, 1, 1
    for _ in range(ndays):
        y.append(yy); m.append(mm); d.append(dd)
        dd += 1
        if dd > md[mm - 1]:
            dd, mm = 1, mm + 1
            if mm > 12:
                mm, yy = 1, yy + 1
    return CalTime(calendar, y, m, d)


def _cal_groups(t: CalTime, freq: str):
    """cftime branch of resample_groups (nb_kernels.py:100-110): bins are calendar days /
    months / years from the first to the last stamp, empty interior bins kept (count 0)."""
    md = _cal_mdays(t.calendar)
    ylen = int(md.sum())
    cum = np.concatenate([[0], np.cumsum(md)])
    if freq == "1D":
        ids = t.year * ylen + cum[t.month - 1] + (t.day - 1)
    elif freq == "ME":
        ids = t.year * 12 + (t.month - 1)
    elif freq == "YE":
        ids = t.year.copy()
    else:
        raise NotImplementedError("groupby='week' is not supported on non-standard CF calendars")
    if np.any(np.diff(t.keys()) < 0):
        raise ValueError("numba engine requires a monotonic-increasing time index")
    first, last = int(ids[0]), int(ids[-1])
    counts = np.bincount(ids - first, minlength=last - first + 1)
    bounds = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    bins = np.arange(first, last + 1)
    if freq == "1D":
        yy, doy = bins // ylen, bins % ylen
        mm = np.searchsorted(cum, doy, side="right")
        labels = CalTime(t.calendar, yy, mm, doy - cum[mm - 1] + 1)
    elif freq == "ME":                                    # label = last day of the month
        yy, mm = bins // 12, bins % 12 + 1
        labels = CalTime(t.calendar, yy, mm, md[mm - 1])
    else:                                                 # label = last day of the year
        labels = CalTime(t.calendar, bins, np.full_like(bins, 12), np.full_like(bins, md[11]))
    return bounds, labels


def resample_groups(tindex, freq: str):
    """(bounds int64[G+1], labels) -- nb_kernels.py:80-115."""
    if isinstance(tindex, CalTime):
        return _cal_groups(tindex, freq)
    tindex = pd.DatetimeIndex(tindex)
    if not tindex.is_monotonic_increasing:
        raise ValueError("numba engine requires a monotonic-increasing time index "
                         "(xarray's resample path enforces the same).")
    counts = pd.Series(1, index=tindex).resample(freq).count()
    bounds = np.concatenate([[0], np.cumsum(counts.values)]).astype(np.int64)
    return bounds, pd.DatetimeIndex(counts.index)


# --------------------------------------------------------------------------------------------
# data containers (plain arrays instead of xarray objects)
# --------------------------------------------------------------------------------------------
@dataclass
class ODataset:
    """values[time, lat, lon] + axes; the oracle's stand-in for aggfly.Dataset."""
    values: np.ndarray
    time: Any                               # pd.DatetimeIndex | CalTime
    latitude: np.ndarray
    longitude: np.ndarray
    lon_is_360: bool = True

    def copy(self):
        return ODataset(self.values, self.time, self.latitude, self.longitude, self.lon_is_360)


@dataclass
class OWeights:
    """Stand-in for GridWeights as consumed by the hot path (spatial.py:62-69, aggregate.py:276-280)."""
    weights: pd.DataFrame                   # columns cell_id, index_right, weight
    cell_id: np.ndarray                     # weights.grid.cell_id
    shp: pd.DataFrame                       # weights.georegions.shp (index = shapefile row index)
    regionid: str
    zero_weight: str = "area"               # spatial.py:69 default when attr missing


# --------------------------------------------------------------------------------------------
# temporal chain
# --------------------------------------------------------------------------------------------
class TemporalStep:
    """temporal.py:19-263 restricted to the numba engine (the numeric contract)."""

    def __init__(self, calc, groupby, ddargs=None, pre_compute=False, engine="auto"):
        self.calc = calc
        self.freq = FREQ[groupby]
        self.ddargs = ddargs
        self.multi_dd = ddargs is not None and np.array(ddargs).ndim > 1   # temporal.py:150-163
        if calc not in ("mean", "nanmean", "sum", "min", "max", "dd", "bins", "sine_dd"):
            raise ValueError(f"unsupported calc {calc!r}")

    def execute(self, arr: np.ndarray, time):
        if self.freq == "W" and isinstance(time, CalTime):                   # temporal.py:221-227
            raise NotImplementedError("groupby='week' is not supported on non-standard CF calendars")
        bounds, labels = resample_groups(time, self.freq)
        if self.calc in STAT_CODE:
            return block_stat(arr, bounds, self.calc), labels
        fn = {"dd": block_dd, "bins": block_bins, "sine_dd": block_sine_dd}[self.calc]
        out = fn(arr, bounds, self.ddargs)                                   # [G, Y, X, D]
        if not self.multi_dd:
            return out[..., 0], labels                                       # nb_kernels.py:303-304
        if out.shape[-1] == 1:
            raise ValueError("2-D ddargs with a single row is not supported (reference bug, "
                             "temporal.py:250-253)")
        return [np.ascontiguousarray(out[..., d]) for d in range(out.shape[-1])], labels


def _transform(arr: np.ndarray, key: str, params: dict):
    """aggregate.py:36-78 + dataset.py:442-481, 527-543."""
    if "exp" in params:
        exp = params["exp"]
        if not isinstance(exp, list):
            exp = [exp]
        exps = exp[0]
        return [np.power(arr, e) for e in exps], [f"{key}_{e}" for e in exps]
    if "inter" in params:
        other = params["inter"]
        other = other.values if isinstance(other, ODataset) else np.asarray(other)
        assert arr.shape == other.shape
        return [np.multiply(arr, other)], [key]
    if "spline" in params.get("transform", ""):
        return [arr, (arr > 20) * (arr - 20)], [f"{key}_spline1", f"{key}_spline2"]
    raise ValueError("No valid transform argument provided.")


def aggregate_time(dataset: ODataset, aggregator_dict: Dict[str, list]) -> Dict[str, Tuple[np.ndarray, Any]]:
    """aggregate.py:101-162 -> {output name: (values[G, lat, lon], labels)}."""
    out: Dict[str, Tuple[np.ndarray, Any]] = {}
    for key, steps in aggregator_dict.items():
        keys = [key]
        data = [(dataset.values, dataset.time)]
        for kind, params in steps:
            if kind == "aggregate":
                step = params if isinstance(params, TemporalStep) else TemporalStep(**params)
                res = [step.execute(a, t) for a, t in data]
                if step.multi_dd:
                    if len(res) > 1:
                        raise ValueError("Cannot aggregate multiple datasets with multiple ddargs, "
                                         "e.g., multiple polynomials for multiple bins")
                    arrs, labels = res[0]
                    data = [(a, labels) for a in arrs]
                    keys = [f"{key}_{x[0]}_{x[1]}" for x in step.ddargs]      # aggregate.py:299
                else:
                    data = res
            elif kind == "transform":
                nd, nk = [], []
                for (a, t), k in zip(data, keys):
                    a2, k2 = _transform(a, k, dict(params))
                    nd.extend((x, t) for x in a2)
                    nk.extend(k2)
                data, keys = nd, nk
        out.update(dict(zip(keys, data)))
    return out


# --------------------------------------------------------------------------------------------
# spatial
# --------------------------------------------------------------------------------------------
def lon_to_180(lon):
    return (np.asarray(lon, dtype=float) + 180) % 360 - 180          # grid_utils.py:16-31


def rescale_longitude(values: np.ndarray, longitude: np.ndarray):
    """dataset.py:419-440 for a 0-360 dataset: relabel, then sortby('longitude') (stable)."""
    l180 = lon_to_180(longitude)
    order = np.argsort(l180, kind="stable")
    return values[..., order], l180[order]


def weight_triplets(wdf: pd.DataFrame, cell_ids: np.ndarray):
    """spatial.py:157-178."""
    cellpos = {int(c): i for i, c in enumerate(cell_ids)}
    region_ids = np.sort(wdf["index_right"].unique())
    regionpos = {r: i for i, r in enumerate(region_ids)}
    rows = wdf["index_right"].map(regionpos).to_numpy()
    cols = wdf["cell_id"].map(cellpos).to_numpy()
    keep = ~pd.isna(cols)
    return (rows[keep].astype(np.intp), cols[keep].astype(np.intp),
            wdf["weight"].to_numpy(dtype=float)[keep], region_ids)


def aggregate_space(dataset_dict: Dict[str, Tuple[np.ndarray, Any]], lon_is_360: bool,
                    longitude: np.ndarray, weights: OWeights) -> pd.DataFrame:
    """spatial.py:71-154 on plain arrays.  Every entry is (values[G, lat, lon], labels)."""
    names = list(dataset_dict.keys())
    label_sets = [dataset_dict[n][1] for n in names]
    first = label_sets[0]
    for lab in label_sets[1:]:
        same = (len(lab) == len(first)) and (
            np.array_equal(lab.keys(), first.keys()) if isinstance(first, CalTime)
            else bool((pd.DatetimeIndex(lab) == pd.DatetimeIndex(first)).all()))
        if not same:
            raise ValueError("all outputs of one call must share one output time axis")
    time = first
    n_time = len(time)
    arrs = {}
    for nm in names:
        v = dataset_dict[nm][0]
        if lon_is_360:
            v, _ = rescale_longitude(v, longitude)                    # spatial.py:60
        arrs[nm] = v.reshape(n_time, -1).T                            # (cell, time), cell = lat-major
    cell_ids = np.asarray(weights.cell_id)
    region_idx, cell_idx, w_vals, region_ids = weight_triplets(weights.weights, cell_ids)
    n_regions = len(region_ids)
    valid = None
    for nm in names:
        v = ~np.isnan(arrs[nm])
        valid = v if valid is None else (valid & v)
    den = scatter_block(valid.astype(float), region_idx, cell_idx, w_vals, n_regions)
    res = {}
    for nm in names:
        num = scatter_block(np.where(valid, arrs[nm], 0.0), region_idx, cell_idx, w_vals, n_regions)
        with np.errstate(invalid="ignore", divide="ignore"):
            res[nm] = np.divide(num, den, out=np.full_like(den, np.nan), where=den != 0)
    if isinstance(time, CalTime):
        tvals = np.array([f"{y:04d}-{m:02d}-{d:02d}" for y, m, d in zip(time.year, time.month, time.day)],
                         dtype=object)
    else:
        tvals = pd.DatetimeIndex(time).values
    out = pd.DataFrame({"region_id": np.repeat(region_ids, n_time), "time": np.tile(tvals, n_regions)})
    for nm in names:
        out[nm] = res[nm].reshape(-1)
    if weights.zero_weight == "nan":                                  # spatial.py:144-151
        wsum = weights.weights.groupby("index_right")["weight"].sum()
        zero_regions = set(wsum.index[~(wsum > 0)])
        keep = out["region_id"].isin(zero_regions) | out[names].notna().all(axis=1)
        out = out.loc[keep].reset_index(drop=True)
    else:
        out = out.dropna(subset=names).reset_index(drop=True)
    return out


def aggregate_dataset(weights: OWeights, dataset: ODataset = None,
                      aggregator_dict: Dict[str, list] = None, **kwargs) -> pd.DataFrame:
    """aggregate.py:210-282."""
    if dataset is None:
        raise ValueError("No dataset provided.")
    stale = {k: kwargs.pop(k) for k in ("n_workers", "threads_per_worker", "processes",
                                        "memory_limit", "cluster_args") if k in kwargs}
    if stale:
        warnings.warn("aggregate_dataset no longer builds a Dask cluster", DeprecationWarning, stacklevel=2)
    if aggregator_dict is None and kwargs:
        aggregator_dict = kwargs
    if aggregator_dict is not None:
        dd = aggregate_time(dataset, aggregator_dict)
    else:
        dd = {"variable": (dataset.values, dataset.time)}
    df = aggregate_space(dd, dataset.lon_is_360, dataset.longitude, weights)
    df = weights.shp[[weights.regionid]].merge(df, left_index=True, right_on="region_id").drop(
        columns="region_id")
    return df
