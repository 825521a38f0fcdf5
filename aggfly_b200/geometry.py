"""Polygon regions without geopandas/shapely: ring storage, a minimal ESRI shapefile reader, and
the call into the library's exact polygon/cell overlap (``agf_overlap_*``, csrc/agf_geom.cu).

The reference builds its weights from a GeoDataFrame with GEOS (aggfly/regions/georegions.py,
aggfly/weights/grid_weights.py:238-421); this image has neither, so ``GeoRegions`` can also carry
plain ring arrays and ``GridWeights.calculate_weights`` clips them itself.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import _lib


def ring_signed_area(ring: np.ndarray) -> float:
    x, y = ring[:, 0], ring[:, 1]
    return 0.5 * float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def orient_polygon(shell: np.ndarray, holes: Sequence[np.ndarray] = ()) -> List[np.ndarray]:
    """[shell, *holes] with the shell counter-clockwise and every hole clockwise."""
    shell = np.asarray(shell, dtype=float)
    rings = [shell if ring_signed_area(shell) >= 0 else shell[::-1]]
    for h in holes:
        h = np.asarray(h, dtype=float)
        rings.append(h if ring_signed_area(h) <= 0 else h[::-1])
    return rings


def convex_hull(points: np.ndarray) -> np.ndarray:
    """Andrew's monotone chain; counter-clockwise ring without the closing vertex."""
    pts = sorted(map(tuple, np.asarray(points, dtype=float)))

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def cell_overlaps(region_rings: Sequence[Sequence[np.ndarray]], lon: np.ndarray, lat: np.ndarray,
                  dlon: float, dlat: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(region position, cell_id, covered fraction of the cell) for every overlapping pair;
    ``region_rings[r]`` = the rings of region r (holes oriented opposite to shells)."""
    L = _lib.lib()
    n_regions = len(region_rings)
    region_ptr = np.zeros(n_regions + 1, dtype=np.int64)
    ring_sizes, chunks = [], []
    for r, rings in enumerate(region_rings):
        region_ptr[r + 1] = region_ptr[r] + len(rings)
        for ring in rings:
            ring = np.ascontiguousarray(ring, dtype=np.float64)
            if ring.ndim != 2 or ring.shape[1] != 2:
                raise ValueError(f"region {r}: a ring must be an (n, 2) array of lon/lat vertices")
            ring_sizes.append(len(ring))
            chunks.append(ring)
    ring_ptr = np.zeros(len(ring_sizes) + 1, dtype=np.int64)
    np.cumsum(ring_sizes, out=ring_ptr[1:])
    xy = np.ascontiguousarray(np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 2)), dtype=np.float64)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    dp, i64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    h, n = C.c_void_p(), C.c_int64()
    _lib.check(L.agf_overlap_create(C.byref(h), n_regions, region_ptr.ctypes.data_as(i64p), ring_ptr.ctypes.data_as(i64p),
                                    xy.ctypes.data_as(dp), len(lon), lon.ctypes.data_as(dp), float(dlon),
                                    len(lat), lat.ctypes.data_as(dp), float(dlat), C.byref(n)))
    try:
        region = np.empty(n.value, dtype=np.int32)
        cell = np.empty(n.value, dtype=np.int64)
        frac = np.empty(n.value, dtype=np.float64)
        if n.value:
            _lib.check(L.agf_overlap_fetch(h, region.ctypes.data_as(i32p), cell.ctypes.data_as(i64p), frac.ctypes.data_as(dp)))
    finally:
        L.agf_overlap_destroy(h)
    return region, cell, frac


# ---------------------------------------------------------------------------------------------
# ESRI shapefile (.shp polygons + optional .dbf attributes)
# ---------------------------------------------------------------------------------------------
def read_shp_polygons(path: str) -> List[List[np.ndarray]]:
    """Rings of every record of a Polygon / PolygonZ / PolygonM shapefile (shape types 5, 15, 25).
    Outer rings are clockwise and holes counter-clockwise in the format, which is all the overlap
    code needs (opposite orientations subtract).  Null shapes give an empty ring list."""
    data = open(path, "rb").read()
    if len(data) < 100 or struct.unpack(">i", data[:4])[0] != 9994:
        raise ValueError(f"{path}: not an ESRI shapefile")
    out, pos = [], 100
    while pos + 8 <= len(data):
        _, clen = struct.unpack(">ii", data[pos:pos + 8])
        rec = data[pos + 8: pos + 8 + 2 * clen]
        pos += 8 + 2 * clen
        stype = struct.unpack("<i", rec[:4])[0]
        if stype == 0:
            out.append([])
            continue
        if stype not in (5, 15, 25):
            raise ValueError(f"{path}: shape type {stype} is not a polygon")
        n_parts, n_points = struct.unpack("<ii", rec[36:44])
        parts = np.frombuffer(rec, dtype="<i4", count=n_parts, offset=44)
        pts = np.frombuffer(rec, dtype="<f8", count=2 * n_points, offset=44 + 4 * n_parts).reshape(n_points, 2)
        ends = list(parts[1:]) + [n_points]
        out.append([np.array(pts[a:b]) for a, b in zip(parts, ends)])
    return out


def read_dbf(path: str) -> pd.DataFrame:
    """Attribute table of a shapefile (dBASE III: C / N / F / L / D fields as strings or numbers)."""
    data = open(path, "rb").read()
    n_rec, hdr_len, rec_len = struct.unpack("<IHH", data[4:12])
    fields, pos = [], 32
    while data[pos] != 0x0D:
        name = data[pos:pos + 11].split(b"\x00")[0].decode("latin-1")
        fields.append((name, chr(data[pos + 11]), data[pos + 16], data[pos + 17]))
        pos += 32
    cols = {f[0]: [] for f in fields}
    for i in range(n_rec):
        rec = data[hdr_len + i * rec_len: hdr_len + (i + 1) * rec_len]
        off = 1
        for name, ftype, flen, fdec in fields:
            raw = rec[off:off + flen].decode("latin-1").strip()
            off += flen
            if ftype in ("N", "F") and raw not in ("", "*" * flen):
                cols[name].append(float(raw) if (fdec or "." in raw or "e" in raw.lower()) else int(raw))
            elif ftype in ("N", "F"):
                cols[name].append(np.nan)
            else:
                cols[name].append(raw)
    return pd.DataFrame(cols)


def rescale_raster_to_grid(values: np.ndarray, src_lat: np.ndarray, src_lon: np.ndarray,
                           grid_lat: np.ndarray, grid_lon: np.ndarray, dlat: float, dlon: float,
                           nodata: Optional[float] = None) -> np.ndarray:
    """Secondary raster -> climate grid: the area-weighted mean of the source pixels under each grid
    cell, over the pixels that have data (``rio.reproject_match(..., Resampling.average)`` in
    aggfly/weights/secondary_weights.py:40-109; for the usual aligned integer-factor case this is
    the plain block mean GDAL computes).  Cells with no valid pixel under them come out NaN."""
    v = np.asarray(values, dtype=np.float64)
    src_lat, src_lon = np.asarray(src_lat, float), np.asarray(src_lon, float)
    if v.shape != (len(src_lat), len(src_lon)):
        raise ValueError(f"raster shape {v.shape} does not match its axes ({len(src_lat)}, {len(src_lon)})")
    ok = np.isfinite(v)
    if nodata is not None:
        ok &= v != nodata

    def overlap(src, dst, d_dst):
        d_src = abs(np.diff(src).mean()) if len(src) > 1 else d_dst
        lo = np.maximum(dst[:, None] - d_dst / 2, src[None, :] - d_src / 2)
        hi = np.minimum(dst[:, None] + d_dst / 2, src[None, :] + d_src / 2)
        return np.clip(hi - lo, 0.0, None)                      # [n_dst, n_src] overlap lengths

    wy = overlap(src_lat, np.asarray(grid_lat, float), dlat)
    wx = overlap(src_lon, np.asarray(grid_lon, float), dlon)
    num = wy @ np.where(ok, v, 0.0) @ wx.T
    den = wy @ ok.astype(np.float64) @ wx.T
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(den > 0, num / den, np.nan)
