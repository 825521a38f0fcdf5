"""Readers / writers around the hot path: rasters, region files, secondary rasters, panels.

The reference opens everything through xarray / geopandas / rioxarray (aggfly/dataset/dataset.py:636-740,
aggfly/regions/georegions.py:220-323, aggfly/weights/secondary_weights.py:201-245).  None of those
exist in this image, so the formats that can be read with numpy / scipy / the standard library are
read natively, and xarray is used only if it happens to be importable:

=====================  =========================================================================
``.npz``               arrays ``values`` (or the ``var`` name) ``[time, lat, lon]``, ``time``
                       (datetime64 or ISO strings), ``latitude``, ``longitude``  -- also the format of
                       secondary rasters (``values[lat, lon]``, ``latitude``, ``longitude``)
``.npy``               the raster alone, memory-mapped; axes from ``<path>.axes.npz``
``.nc`` (NetCDF-3)     ``scipy.io.netcdf_file`` (memory-mapped); CF ``units = "<unit> since <date>"``
                       time axes, ``scale_factor`` / ``add_offset`` unpacked like xarray (float64)
zarr directory store   v2 / v3, any axis order and chunking (aggfly_b200.zarrio): opened lazily, chunks
                       are decoded by host threads and placed on the device (``stream.feed_chunked``)
list / glob of paths   one dataset concatenated along time (``dataset.TimeConcat``, lazy), parts in any of the formats above
anything else          ``xarray.open_dataset`` when xarray is installed
``.tif`` / ``.tiff``   secondary rasters: north-up single-band GeoTIFF in geographic coordinates (Pillow)
``.shp``               polygons (+ ``.dbf`` attributes), aggfly_b200.geometry
``.geojson`` / ``.json``  FeatureCollection of Polygon / MultiPolygon features
=====================  =========================================================================
"""
from __future__ import annotations

import json
import os
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from .dataset import Dataset
from .geometry import orient_polygon
from .weights import GeoRegions, SecondaryWeights
from . import hdf5io as _hdf5io
from .zarrio import looks_like_zarr, open_raster, write_dataset


# ---------------------------------------------------------------------------------------------
# rasters
# ---------------------------------------------------------------------------------------------
def _cf_time(values: np.ndarray, units: str, calendar: str = "standard"):
    unit, _, origin = units.partition(" since ")
    unit = unit.strip().lower().rstrip("s")
    step = {"second": "s", "minute": "m", "hour": "h", "day": "D"}.get(unit)
    if step is None:
        raise ValueError(f"unsupported CF time unit {units!r}")
    if calendar not in ("standard", "gregorian", "proleptic_gregorian"):
        raise NotImplementedError(f"calendar {calendar!r} in a NetCDF-3 file: build a CalendarIndex yourself "
                                  "(Dataset.from_arrays)")
    vals = np.asarray(values, dtype=np.float64)
    origin = pd.Timestamp(origin.strip())
    if origin.tzinfo is not None:
        origin = origin.tz_convert(None)
    ns = {"s": 1e9, "m": 60e9, "h": 3600e9, "D": 86400e9}[step]
    return pd.DatetimeIndex(origin.value + np.round(vals * ns).astype(np.int64))


def dataset_from_path(path, var: Optional[str] = None, xycoords: Sequence[str] = ("longitude", "latitude"),
                      timecoord: str = "time", lon_is_360: bool = True, preprocess=None, time_sel=None,
                      name: Optional[str] = None, **kwargs) -> Dataset:
    """``af.dataset_from_path`` (aggfly/dataset/dataset.py:636-740) for the formats listed in the module
    docstring.  ``preprocess``: builtin name / expression in ``x`` (fused on the device) or a callable."""
    xdim, ydim = xycoords
    if isinstance(path, (list, tuple)) or (isinstance(path, str) and any(c in path for c in "*?[")):
        return _dataset_from_paths(path, var, xycoords, timecoord, lon_is_360, preprocess, time_sel, name, **kwargs)
    ext = os.path.splitext(path.rstrip("/"))[1].lower()
    keepalive = None
    if ext == ".npz":
        z = np.load(path, allow_pickle=False)
        key = var if (var and var in z.files) else "values"
        values, time, lat, lon = z[key], z[timecoord], z[ydim], z[xdim]
        time = pd.DatetimeIndex(time.astype("datetime64[ns]") if time.dtype.kind == "M" else pd.to_datetime(time.astype(str)))
    elif ext == ".npy":
        values = np.load(path, mmap_mode="r")
        z = np.load(path + ".axes.npz", allow_pickle=False)
        time, lat, lon = pd.DatetimeIndex(z[timecoord].astype("datetime64[ns]")), z[ydim], z[xdim]
    elif ext in (".nc", ".nc3", ".cdf") and _is_netcdf3(path):
        from scipy.io import netcdf_file
        f = netcdf_file(path, "r", mmap=True, maskandscale=False)
        if var not in f.variables:
            raise KeyError(f"{path}: variable {var!r} not found (have {sorted(f.variables)})")
        v = f.variables[var]
        dims = list(v.dimensions)
        order = [dims.index(timecoord), dims.index(ydim), dims.index(xdim)]
        values = np.transpose(v.data, order) if order != [0, 1, 2] else v.data
        scale, offset = getattr(v, "scale_factor", None), getattr(v, "add_offset", None)
        if scale is not None or offset is not None:                 # CF packing -> float64, like xarray
            values = np.asarray(values, dtype=np.float64) * (1.0 if scale is None else float(scale)) \
                + (0.0 if offset is None else float(offset))
        elif values.dtype.byteorder == ">":
            values = values.astype(values.dtype.newbyteorder("="))   # NetCDF-3 is big-endian on disk
        tv = f.variables[timecoord]
        units = tv.units.decode() if isinstance(tv.units, bytes) else tv.units
        cal = getattr(tv, "calendar", b"standard")
        time = _cf_time(tv.data, units, cal.decode() if isinstance(cal, bytes) else cal)
        lat, lon = np.array(f.variables[ydim].data, dtype=float), np.array(f.variables[xdim].data, dtype=float)
        keepalive = f                                                # the raster is a view of the mapped file
    elif kwargs.get("engine") in (None, "zarr") and looks_like_zarr(path):      # dataset.py:697-701
        values, time, lat, lon = open_raster(path, var, xycoords, timecoord)
    elif kwargs.get("engine") in (None, "netcdf4", "h5netcdf") and os.path.isfile(path) and _hdf5io.looks_like_hdf5(path):
        # NetCDF-4 = HDF5: chunks are read, inflated and un-shuffled by host threads into pinned slots and placed /
        # unpacked on the device (hdf5io.py); no xarray / netCDF4 / h5py needed
        values, time, lat, lon = _hdf5io.open_raster(path, var, xycoords, timecoord)
    else:
        try:
            import xarray as xr                                      # optional dependency
        except Exception as exc:
            raise ImportError(f"{path}: reading this format needs xarray (not installed); natively supported: "
                              ".npz, .npy (+ .axes.npz), NetCDF-3 and NetCDF-4 .nc, zarr directory stores") from exc
        kwargs.pop("chunks", None)
        dsx = xr.open_dataset(path, **kwargs)
        return Dataset(dsx[var], xycoords=xycoords, timecoord=timecoord, time_sel=time_sel, lon_is_360=lon_is_360,
                       preprocess=preprocess, name=name)
    ds = Dataset.from_arrays(values, time, lat, lon, lon_is_360=lon_is_360, name=name or var,
                             preprocess=preprocess if isinstance(preprocess, str) else None)
    ds._keepalive = keepalive
    if preprocess is not None and not isinstance(preprocess, str):
        ds.values = np.asarray(preprocess(np.asarray(ds.values)))
    if time_sel is not None:
        ds = select_time(ds, time_sel)
    return ds


def _dataset_from_paths(paths, var, xycoords, timecoord, lon_is_360, preprocess, time_sel, name, **kwargs) -> Dataset:
    """Several files of one variable (a list, or a glob pattern) as ONE dataset concatenated along time -- the
    reference's ``xr.open_mfdataset`` branch (aggfly/dataset/dataset.py:686-695).  Every file is opened lazily
    on its own; the parts are ordered by their first time stamp, must share the grid and must not overlap in time."""
    import glob
    from .dataset import TimeConcat
    from .timeaxis import CalendarIndex
    if isinstance(paths, str):
        pattern, paths = paths, sorted(glob.glob(paths))
        if not paths:
            raise FileNotFoundError(f"no file matches {pattern!r}")
    if not paths:
        raise ValueError("empty list of paths")
    fused = preprocess if isinstance(preprocess, str) else None
    parts = [dataset_from_path(p, var=var, xycoords=xycoords, timecoord=timecoord, lon_is_360=lon_is_360,
                               preprocess=preprocess, name=name, **kwargs) for p in paths]
    cal = {type(d.time).__name__ + (d.time.calendar if isinstance(d.time, CalendarIndex) else "") for d in parts}
    if len(cal) != 1:
        raise ValueError(f"the files do not share one calendar: {sorted(cal)}")
    key = (lambda d: int(d.time.ordinal_hours()[0])) if isinstance(parts[0].time, CalendarIndex) else (lambda d: d.time[0].value)
    parts = sorted([d for d in parts if len(d.time)], key=key) or parts[:1]
    first = parts[0]
    for d in parts[1:]:
        if not (np.array_equal(d.latitude, first.latitude) and np.array_equal(d.longitude, first.longitude)):
            raise ValueError("the files are on different grids")
    if isinstance(first.time, CalendarIndex):
        time = CalendarIndex(first.time.calendar, *[np.concatenate([getattr(d.time, f) for d in parts])
                                                     for f in ("year", "month", "day", "hour")])
    else:
        time = pd.DatetimeIndex(np.concatenate([d.time.values for d in parts]))
    if not time.is_monotonic_increasing:
        raise ValueError("the files overlap in time")
    values = TimeConcat([d.values for d in parts]) if len(parts) > 1 else first.values
    ds = Dataset.from_arrays(values, time, first.latitude, first.longitude, lon_is_360=lon_is_360, name=name or var,
                             preprocess=fused)
    ds._keepalive = [getattr(d, "_keepalive", None) for d in parts]
    if time_sel is not None:
        ds = select_time(ds, time_sel)
    return ds


def _auto_chunks(sizes: dict, itemsize: int, target_mb: float) -> dict:
    """Chunking policy of the reference's converter (aggfly/dataset/zarr_convert.py:31-47): keep the time
    axis whole when a square spatial tile of at least 32 cells fits the byte budget (tile capped at 256),
    otherwise 128-cell tiles and as many time steps as the budget allows."""
    T, Y, X = sizes["time"], sizes["latitude"], sizes["longitude"]
    budget = max(1, int(target_mb * 1024 * 1024 / itemsize))
    side = int((budget / T) ** 0.5)
    if side >= 32:
        side = int(min(side, 256, Y, X))
        return {"time": -1, "latitude": side, "longitude": side}
    side = int(min(128, Y, X))
    return {"time": int(min(max(1, budget // (side * side)), T)), "latitude": side, "longitude": side}


def dataset_to_zarr(dataset: Dataset, store: str, chunking="auto", target_mb: float = 256, overwrite: bool = False,
                    return_dataset: bool = True, zarr_format: int = 3, compressor: Optional[str] = "zstd"):
    """``af.dataset_to_zarr`` (aggfly/dataset/zarr_convert.py:50-121): write the raster as a time-contiguous
    store with dims ``(latitude, longitude, time)`` and reopen it lazily."""
    import shutil
    values = np.asarray(dataset.values.cpu() if hasattr(dataset.values, "cpu") else dataset.values)
    name = dataset.name or "variable"
    sizes = {"time": values.shape[0], "latitude": values.shape[1], "longitude": values.shape[2]}
    if isinstance(chunking, str) and chunking == "auto":
        chunks = _auto_chunks(sizes, values.dtype.itemsize, target_mb)
    elif isinstance(chunking, dict):
        chunks = dict(chunking)
    else:
        raise ValueError("chunking must be 'auto' or a dict of chunk sizes")
    if os.path.exists(store):
        if not overwrite:
            raise FileExistsError(f"{store} exists; pass overwrite=True to replace it")
        shutil.rmtree(store)
    write_dataset(store, values, dataset.time, dataset.latitude, dataset.longitude, var=name,
                  dims=("latitude", "longitude", "time"), chunks=chunks, zarr_format=zarr_format, compressor=compressor)
    if not return_dataset:
        return None
    new = dataset_from_path(store, var=name, lon_is_360=dataset.lon_is_360)
    new.pre_ops = list(getattr(dataset, "pre_ops", []))
    return new


def zarr_from_path(path: str, var: str, store: str, *, xycoords=("longitude", "latitude"), timecoord: str = "time",
                   lon_is_360: bool = True, preprocess=None, chunking="auto", target_mb: float = 256,
                   overwrite: bool = False, **kwargs):
    """Load any readable source and convert it in one call (aggfly/dataset/zarr_convert.py:124-155)."""
    ds = dataset_from_path(path, var=var, xycoords=xycoords, timecoord=timecoord, lon_is_360=lon_is_360,
                           preprocess=preprocess, **kwargs)
    return dataset_to_zarr(ds, store, chunking=chunking, target_mb=target_mb, overwrite=overwrite)


def _is_netcdf3(path: str) -> bool:
    with open(path, "rb") as f:
        return f.read(3) == b"CDF"


def select_time(ds: Dataset, time_sel) -> Dataset:
    """``da.sel(time=time_sel)``: a partial date string ("2001", "2001-06"), a timestamp or a slice."""
    from .dataset import time_selection
    a, b = time_selection(ds.time, time_sel)
    return ds.isel_time(a, b)


def clip_to_extent(ds: Dataset, lon_min: float, lon_max: float, lat_min: float, lat_max: float) -> Dataset:
    """Keep the block of cells that can touch the box (one cell of margin), like the reference's
    ``clip_to_regions`` (aggfly/dataset/dataset.py:225-312): fewer cells to move and scan.  The box is
    in -180..180 longitudes; a 0-360 dataset whose kept columns would wrap is left unclipped in
    longitude."""
    from copy import copy
    from .dataset import Grid, lon_to_180
    lat, lon = ds.latitude, ds.longitude
    lon180 = lon_to_180(lon) if ds.lon_is_360 else lon
    dlat, dlon = ds.grid.resolution_lat, ds.grid.resolution_lon
    iy = np.nonzero((lat >= lat_min - 1.5 * dlat) & (lat <= lat_max + 1.5 * dlat))[0]
    ix = np.nonzero((lon180 >= lon_min - 1.5 * dlon) & (lon180 <= lon_max + 1.5 * dlon))[0]
    if len(iy) == 0 or len(ix) == 0:
        raise ValueError("the regions do not overlap the dataset's grid")
    y0, y1 = int(iy[0]), int(iy[-1]) + 1
    x0, x1 = (int(ix[0]), int(ix[-1]) + 1) if len(ix) == ix[-1] - ix[0] + 1 else (0, len(lon))
    new = copy(ds)
    new.values = ds.values[:, y0:y1, x0:x1]
    new.latitude, new.longitude = lat[y0:y1], lon[x0:x1]
    new.grid = Grid(new.longitude, new.latitude, ds.name, ds.lon_is_360)
    new.history = list(ds.history) + ["clipped"]
    return new


# ---------------------------------------------------------------------------------------------
# regions / secondary rasters
# ---------------------------------------------------------------------------------------------
def georegions_from_path(path: str, regionid: Optional[str] = None, region_list=None) -> GeoRegions:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".shp":
        return GeoRegions.from_shapefile(path, regionid, region_list)
    if ext in (".geojson", ".json"):
        fc = json.load(open(path))
        feats = fc["features"] if fc.get("type") == "FeatureCollection" else [fc]
        polys, props = [], []
        for ft in feats:
            g = ft["geometry"]
            parts = [g["coordinates"]] if g["type"] == "Polygon" else g["coordinates"] if g["type"] == "MultiPolygon" else None
            if parts is None:
                raise ValueError(f"{path}: geometry type {g['type']!r} is not a polygon")
            rings = []
            for part in parts:
                rings += orient_polygon(np.asarray(part[0], float)[:, :2], [np.asarray(h, float)[:, :2] for h in part[1:]])
            polys.append(rings)
            props.append(ft.get("properties") or {})
        attrs = pd.DataFrame(props)
        if regionid is None:
            regionid = "region_id"
            attrs[regionid] = np.arange(len(polys))
        gr = GeoRegions.from_polygons(attrs[regionid], polys, regionid, attrs)
        if region_list is not None:
            gr = GeoRegions(gr.shp[gr.shp[regionid].isin(region_list)], regionid)
        return gr
    raise ImportError(f"{path}: only .shp and .geojson region files are read natively (no geopandas here)")


def describe_regions(path: str, rows: int = 5, uniqueness: bool = False) -> dict:
    """Fields, feature count, bounds and the first rows of a shapefile (from the .shp / .shx / .dbf HEADERS, no
    geometry is read) or a GeoJSON file; ``uniqueness`` also lists the columns that could serve as ``regionid``."""
    import struct
    from .geometry import read_dbf
    ext = os.path.splitext(path)[1].lower()
    out = {"path": path, "fields": [], "dtypes": [], "head": None, "unique_columns": None, "crs": None}
    if ext == ".shp":
        with open(path, "rb") as f:
            head = f.read(100)
        if len(head) < 100 or struct.unpack(">i", head[:4])[0] != 9994:
            raise ValueError(f"{path}: not an ESRI shapefile")
        stype = struct.unpack("<i", head[32:36])[0]
        out["geometry_type"] = {0: "Null", 1: "Point", 3: "LineString", 5: "Polygon", 15: "PolygonZ", 25: "PolygonM"}.get(stype, f"type {stype}")
        out["total_bounds"] = struct.unpack("<4d", head[36:68])
        base = os.path.splitext(path)[0]
        if os.path.exists(base + ".shx"):
            out["features"] = (os.path.getsize(base + ".shx") - 100) // 8
        if out["total_bounds"][0] >= out["total_bounds"][2] or "features" not in out:
            # a writer that left the header box empty / no index file: walk the record headers (each polygon record
            # carries its own box; vertices are skipped)
            boxes, n = [], 0
            with open(path, "rb") as f:
                f.seek(100)
                while True:
                    rh = f.read(8)
                    if len(rh) < 8:
                        break
                    clen = struct.unpack(">ii", rh)[1]
                    rec = f.read(36)
                    if len(rec) >= 36 and struct.unpack("<i", rec[:4])[0] in (3, 5, 8, 13, 15, 18, 23, 25, 28):
                        boxes.append(struct.unpack("<4d", rec[4:36]))
                    f.seek(2 * clen - len(rec), 1)
                    n += 1
            out["features"] = n
            if boxes:
                b = np.asarray(boxes)
                out["total_bounds"] = (b[:, 0].min(), b[:, 1].min(), b[:, 2].max(), b[:, 3].max())
        if os.path.exists(base + ".prj"):
            out["crs"] = open(base + ".prj").read().strip()[:60]
        attrs = read_dbf(base + ".dbf") if os.path.exists(base + ".dbf") else None
        out["driver"] = "ESRI Shapefile"
    elif ext in (".geojson", ".json"):
        gr = georegions_from_path(path)
        attrs = gr.shp.drop(columns=[c for c in ("rings", "region_id") if c in gr.shp.columns])
        pts = np.concatenate([r for rings in gr.shp["rings"] for r in rings])
        out.update(geometry_type="Polygon", total_bounds=(pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max()),
                   features=len(gr.shp), driver="GeoJSON", crs="OGC:CRS84 (GeoJSON default)")
        attrs = attrs if len(attrs.columns) else None
    else:
        raise ValueError(f"{path}: only .shp and .geojson region files are read natively")
    if attrs is not None:
        out.setdefault("features", len(attrs))
        out["fields"], out["dtypes"] = list(attrs.columns), [str(attrs[c].dtype) for c in attrs.columns]
        out["head"] = attrs.head(rows) if rows else None
        if uniqueness:
            out["unique_columns"] = [c for c in attrs.columns if attrs[c].notna().all() and not attrs[c].duplicated().any()]
    return out


def print_regions_info(d: dict, echo=print) -> None:
    echo(f"{d['path']}")
    echo(f"  driver     : {d.get('driver')}")
    echo(f"  geometry   : {d.get('geometry_type')}  features={d.get('features')}")
    echo(f"  crs        : {d['crs']}" if d["crs"] else "  crs        : NONE (no .prj); coordinates are taken as longitude / latitude")
    xmin, ymin, xmax, ymax = d["total_bounds"]
    echo(f"  bounds     : lon {xmin:.4f} .. {xmax:.4f} | lat {ymin:.4f} .. {ymax:.4f}")
    if xmin >= 0 and xmax > 180:
        echo("               longitudes run 0\u2013360, not -180\u2013180")
    if not d["fields"]:
        echo("  fields     : none \u2014 this file has no attribute table, so there is")
        echo("               no column to use as regionid")
        return
    echo(f"  fields     : {len(d['fields'])}")
    for f, t in zip(d["fields"], d["dtypes"]):
        echo(f"      {f:<24} {t}")
    if d["head"] is not None:
        echo(f"  first {len(d['head'])} row(s) (geometry omitted):")
        for line in d["head"].to_string().splitlines():
            echo(f"      {line}")
    if d["unique_columns"] is not None:
        if d["unique_columns"]:
            echo(f"  unique across all {d['features']} features (regionid candidates):")
            echo(f"      {', '.join(d['unique_columns'])}")
        else:
            echo("  no column is unique across all features \u2014 none can serve as a")
            echo("  regionid on its own")


def shapefile_info(path: str, n: int = 5, uniqueness: bool = False) -> dict:
    """``af.shapefile_info`` (aggfly/regions/georegions.py:326-428): print a summary of a regions file (fields, feature
    count, bounds, the first ``n`` rows, optionally the columns that are unique across all features) and return it."""
    d = describe_regions(path, n, uniqueness)
    print_regions_info(d)
    return d


def _read_geotiff(path: str):
    """(values[lat, lon], latitude, longitude, nodata) of a north-up, single-band GeoTIFF in geographic coordinates
    (LandScan / GPW / cropland rasters: aggfly/weights/secondary_weights.py:201-245 opens them with rioxarray).
    Pixels are decoded by Pillow's libtiff (strips / tiles, LZW / deflate); the georeferencing comes from the
    ModelPixelScale + ModelTiepoint tags (or ModelTransformation without rotation), nodata from GDAL_NODATA."""
    try:
        from PIL import Image
    except Exception as exc:                                         # pragma: no cover
        raise ImportError(f"{path}: reading GeoTIFF needs Pillow") from exc
    Image.MAX_IMAGE_PIXELS = None                                    # global rasters are far above Pillow's bomb guard
    with Image.open(path) as img:
        tags = img.tag_v2
        scale, tie, xform = tags.get(33550), tags.get(33922), tags.get(34264)
        nodata = tags.get(42113)
        if getattr(img, "n_frames", 1) != 1 or len(img.getbands()) != 1:
            raise ValueError(f"{path}: expected one band, found {len(img.getbands())} band(s) / {getattr(img, 'n_frames', 1)} page(s)")
        values = np.asarray(img)
    ny, nx = values.shape
    if scale is not None and tie is not None:
        sx, sy = float(scale[0]), float(scale[1])
        i0, j0, x0, y0 = float(tie[0]), float(tie[1]), float(tie[3]), float(tie[4])
        lon = x0 + (np.arange(nx) + 0.5 - i0) * sx                    # tiepoint = outer corner of pixel (i0, j0)
        lat = y0 - (np.arange(ny) + 0.5 - j0) * sy
    elif xform is not None:
        m = [float(v) for v in xform]
        if m[1] != 0.0 or m[4] != 0.0:
            raise NotImplementedError(f"{path}: rotated / sheared GeoTIFF")
        lon = m[3] + (np.arange(nx) + 0.5) * m[0]
        lat = m[7] + (np.arange(ny) + 0.5) * m[5]
    else:
        raise ValueError(f"{path}: no georeferencing tags (ModelPixelScale / ModelTiepoint)")
    if isinstance(nodata, (bytes, str)):
        txt = (nodata.decode() if isinstance(nodata, bytes) else nodata).strip("\x00 ").strip()
        nodata = float(txt) if txt else None
    if abs(lat).max() > 90.0001 or abs(lon).max() > 360.0001:
        raise NotImplementedError(f"{path}: projected coordinates (only geographic lat / lon rasters are averaged onto the grid)")
    return values, lat, lon, nodata


_Y_NAMES, _X_NAMES = ("y", "latitude", "lat"), ("x", "longitude", "lon")


def _select_2d(values, dims, coords, sel, path):
    """``da.sel(**sel)`` on the non-spatial axes, then the single remaining band: -> (values[y, x], y name, x name)."""
    dims = list(dims)
    for dim, want in (sel or {}).items():
        if dim not in dims:
            raise KeyError(f"{path}: no dimension {dim!r} to select on (have {dims})")
        labels = [v.decode() if isinstance(v, bytes) else str(v) for v in np.asarray(coords[dim]).tolist()] \
            if np.asarray(coords[dim]).dtype.kind in "SUO" else np.asarray(coords[dim]).tolist()
        key = str(want) if labels and isinstance(labels[0], str) else want
        if key not in labels:
            raise KeyError(f"{path}: {dim}={want!r} not found (have {labels})")
        values = np.take(values, labels.index(key), axis=dims.index(dim))
        dims.remove(dim)
    ydim = next((d for d in dims if d.lower() in _Y_NAMES), None)
    xdim = next((d for d in dims if d.lower() in _X_NAMES), None)
    if ydim is None or xdim is None:
        raise ValueError(f"{path}: no y / x (latitude / longitude) dimensions among {dims}")
    for d in [d for d in dims if d not in (ydim, xdim)]:
        if values.shape[dims.index(d)] != 1:
            raise ValueError(f"{path}: dimension {d!r} has {values.shape[dims.index(d)]} entries; pass sel={{{d!r}: ...}}")
        values = np.take(values, 0, axis=dims.index(d))
        dims.remove(d)
    if dims.index(ydim) > dims.index(xdim):
        values = np.asarray(values).T
    return np.asarray(values), ydim, xdim


def secondary_weights_from_path(path: str, var: Optional[str] = None, sel: Optional[dict] = None, nodata: Optional[float] = None,
                                wtype: str = "secondary", cache_identifier: Optional[str] = None, crs=None, name: Optional[str] = None,
                                **kwargs) -> SecondaryWeights:
    """``af.secondary_weights_from_path`` (aggfly/weights/secondary_weights.py:112-245): a raster on a regular
    latitude / longitude grid from ``.npz`` (values, latitude, longitude), a geographic GeoTIFF, a zarr store or a
    NetCDF-3 file (``var`` + ``sel={"crop": "corn"}`` coordinate selection like the reference's ``open_raster``).
    ``wtype`` / ``cache_identifier`` label the weights (cache discrimination); ``crs`` is accepted for signature
    compatibility -- only geographic rasters are supported, nothing is reprojected."""
    ext = os.path.splitext(path.rstrip("/"))[1].lower()
    label = name or os.path.basename(path.rstrip("/"))
    if ext in (".tif", ".tiff"):
        values, lat, lon, file_nodata = _read_geotiff(path)
        nodata = file_nodata if nodata is None else nodata
    elif ext == ".npz":
        z = np.load(path, allow_pickle=False)
        values, lat, lon = z["values" if var is None or var not in z.files else var], z["latitude"], z["longitude"]
    elif looks_like_zarr(path):
        from .zarrio import ZarrGroup
        g = ZarrGroup(path)
        if var is None:
            cands = [n for n in g.names() if g[n].dims and g[n].ndim >= 2 and n not in g[n].dims]
            if len(cands) != 1:
                raise KeyError(f"{path}: pass var= (arrays: {cands})")
            var = cands[0]
        a = g[var]
        fv, scale, offset = a.cf_packing()
        raw = a.read()
        values = raw.astype(np.float64) * scale + offset if (scale != 1.0 or offset != 0.0) else raw.astype(np.float64)
        if fv is not None:
            values[raw.astype(np.float64) == fv] = np.nan
        coords = {d: g[d].read() for d in a.dims if d in g}
        values, ydim, xdim = _select_2d(values, a.dims, coords, sel, path)
        lat, lon = np.asarray(coords[ydim], dtype=float), np.asarray(coords[xdim], dtype=float)
    elif ext in (".nc", ".nc3", ".cdf") and _is_netcdf3(path):
        from scipy.io import netcdf_file
        with netcdf_file(path, "r", mmap=False, maskandscale=True) as f:
            if var is None:
                cands = [n for n, v in f.variables.items() if len(v.dimensions) >= 2]
                if len(cands) != 1:
                    raise KeyError(f"{path}: pass var= (variables: {cands})")
                var = cands[0]
            v = f.variables[var]
            coords = {d: np.array(f.variables[d].data) for d in v.dimensions if d in f.variables}
            values = np.ma.filled(np.ma.asarray(v[:], dtype=np.float64), np.nan)
            values, ydim, xdim = _select_2d(values, v.dimensions, coords, sel, path)
            lat, lon = np.asarray(coords[ydim], dtype=float), np.asarray(coords[xdim], dtype=float)
    else:
        raise NotImplementedError(f"Unsupported raster format: {path} (.npz, GeoTIFF, zarr and NetCDF-3 are read natively)")
    out = SecondaryWeights(values, lat, lon, nodata=nodata, name=label)
    out.wtype, out.cache_identifier, out.path = wtype, cache_identifier, path
    return out


def pop_weights_from_path(path: str, **kwargs) -> SecondaryWeights:
    """``af.pop_weights_from_path`` (aggfly/weights/pop_weights.py): a population raster, ``wtype="pop"``."""
    kwargs.setdefault("wtype", "pop")
    return secondary_weights_from_path(path, **kwargs)


def crop_weights_from_path(path: str, crop: str = "corn", feed: Optional[str] = "total", var: str = "layer", **kwargs) -> SecondaryWeights:
    """``af.crop_weights_from_path`` (aggfly/weights/crop_weights.py): the ``crop`` layer of a cropland store; the feed
    regime only discriminates caches."""
    out = secondary_weights_from_path(path, var=var, sel={"crop": crop}, wtype=crop, cache_identifier=feed, **kwargs)
    out.crop, out.feed = crop, feed
    return out


# ---------------------------------------------------------------------------------------------
# panels
# ---------------------------------------------------------------------------------------------
def write_table(table, path: str, fmt: Optional[str] = None) -> str:
    """``write_output`` for a ``pyarrow.Table`` (``aggregate_dataset_table``): parquet / feather / csv written by Arrow itself."""
    import pyarrow as pa
    fmt = fmt or {"pq": "parquet"}.get(os.path.splitext(path)[1].lstrip(".").lower(),
                                       os.path.splitext(path)[1].lstrip(".").lower())
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if fmt == "parquet":
        import pyarrow.parquet as pq
        pq.write_table(table, path)
    elif fmt == "feather":
        import pyarrow.feather as pf
        pf.write_feather(table, path)
    elif fmt == "csv":
        import pyarrow.csv as pc
        pc.write_csv(table, path)
    else:
        raise ValueError(f"output format {fmt!r} not in ['csv', 'feather', 'parquet']")
    return path


def write_output(df: pd.DataFrame, path: str, fmt: Optional[str] = None) -> str:
    """aggfly/cli/pipeline.py:159-172: parquet / feather / csv by ``fmt`` or the extension."""
    fmt = fmt or {"pq": "parquet"}.get(os.path.splitext(path)[1].lstrip(".").lower(),
                                       os.path.splitext(path)[1].lstrip(".").lower())
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    out = df.copy()
    if len(out) and not isinstance(out["time"].iloc[0], (pd.Timestamp, np.datetime64)):
        out["time"] = out["time"].map(lambda t: t.isoformat() if hasattr(t, "isoformat") else str(t))   # cftime-like labels
    if fmt == "parquet":
        out.to_parquet(path, index=False)
    elif fmt == "feather":
        out.reset_index(drop=True).to_feather(path)
    elif fmt == "csv":
        out.to_csv(path, index=False)
    else:
        raise ValueError(f"output format {fmt!r} not in ['csv', 'feather', 'parquet']")
    return path
