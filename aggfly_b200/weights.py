"""Weights side of the hot path: the ``weights`` frame contract and its lowering to a
device-resident CSR.

The reference's ``GridWeights`` (aggfly/weights/grid_weights.py) computes polygon / cell overlap
with geopandas+shapely (out of scope here) and hands the aggregation a
DataFrame with columns ``cell_id``, ``index_right``, ``weight`` (:379-521), a ``grid`` whose
``cell_id`` numbers the -180..180-sorted grid (:614-648) and a ``zero_weight`` policy (:125).  This
module keeps exactly that contract:

* :class:`GeoRegions` / :class:`GridWeights` -- light containers with the attributes the hot path
  reads (``weights.weights``, ``weights.grid.cell_id``, ``weights.zero_weight``,
  ``weights.georegions.shp[[regionid]]``);
* :func:`weights_from_objects` -- same name/signature as the reference; ``calculate_weights()``
  returns what ``get_area_weights`` / ``get_weighted_area_weights`` return: exact cell / region
  overlap fractions (x cos(lat), x a secondary raster normalised per region) for polygon regions
  (rings clipped against every candidate cell by the library, csrc/agf_geom.cu -- no GEOS) and,
  through a closed form, for axis-aligned rectangles (synthetic tessellations);
* :func:`lower_to_csr` -- ``_weight_triplets`` (aggfly/aggregate/spatial.py:157-178) + the cell
  re-ordering of ``rescale_longitude`` folded into ``cell_idx`` (raster memory order), so the
  raster itself is never permuted on the device.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Optional

import numpy as np
import pandas as pd

from .dataset import Dataset, Grid, lon_to_180

ZERO_WEIGHT_POLICIES = ("nan", "area", "drop")


class GeoRegions:
    """Region table: ``shp`` (one row per region, index = the ``index_right`` the weights frame
    refers to) and the name of the id column (aggfly/regions/georegions.py:22-217)."""

    def __init__(self, shp: pd.DataFrame, regionid: str = "geoid", region_list=None):
        if regionid not in shp.columns:
            raise ValueError(f"regionid column {regionid!r} not in {list(shp.columns)}")
        if region_list is not None:
            shp = shp[shp[regionid].isin(region_list)]
        self.shp = shp
        self.regionid = regionid

    @property
    def regions(self):
        return self.shp[self.regionid]

    @classmethod
    def from_polygons(cls, ids, polygons, regionid="geoid", attrs: Optional[pd.DataFrame] = None) -> "GeoRegions":
        """Regions given as rings in lon/lat degrees (-180..180).  ``polygons[i]`` is one ``(n, 2)``
        array (a simple polygon), or a list of rings ``[shell, hole, ...]`` / several parts; holes must be
        oriented opposite to their shell (``geometry.orient_polygon`` does that)."""
        rings = []
        for poly in polygons:
            if isinstance(poly, np.ndarray) and poly.ndim == 2:
                poly = [poly]
            rings.append([np.asarray(r, dtype=float) for r in poly])
        shp = pd.DataFrame({regionid: list(ids)}) if attrs is None else attrs.reset_index(drop=True).copy()
        if regionid not in shp.columns:
            shp[regionid] = list(ids)
        shp["rings"] = rings
        return cls(shp, regionid)

    @classmethod
    def from_shapefile(cls, path: str, regionid: Optional[str] = None, region_list=None) -> "GeoRegions":
        """``georegions_from_path`` (aggfly/regions/georegions.py:220-323) for an ESRI shapefile, without
        geopandas: polygons from the ``.shp``, attributes from the ``.dbf`` next to it (if any)."""
        import os
        from .geometry import read_dbf, read_shp_polygons
        rings = read_shp_polygons(path)
        dbf = os.path.splitext(path)[0] + ".dbf"
        attrs = read_dbf(dbf) if os.path.exists(dbf) else pd.DataFrame(index=range(len(rings)))
        if len(attrs) != len(rings):
            raise ValueError(f"{path}: {len(rings)} shapes but {len(attrs)} attribute records")
        if regionid is None:
            regionid = "region_id" if "region_id" not in attrs.columns else "region_id_"
            attrs[regionid] = np.arange(len(rings))
        attrs["rings"] = rings
        return cls(attrs, regionid, region_list)

    @classmethod
    def from_rectangles(cls, ids, lon_min, lon_max, lat_min, lat_max, regionid="geoid") -> "GeoRegions":
        shp = pd.DataFrame({regionid: list(ids), "lon_min": np.asarray(lon_min, float),
                            "lon_max": np.asarray(lon_max, float), "lat_min": np.asarray(lat_min, float),
                            "lat_max": np.asarray(lat_max, float)})
        return cls(shp, regionid)


class GridWeights:
    """Container with the reference's attribute names (aggfly/weights/grid_weights.py:103-138)."""

    def __init__(self, grid: Grid, georegions: GeoRegions, raster_weights=None, chunks=30,
                 project_dir: Optional[str] = None, simplify=None, zero_weight: str = "nan",
                 default_to_area_weights: Optional[bool] = None, cosine_area: Optional[bool] = None,
                 verbose: bool = True):
        assert not grid.lon_is_360                                        # grid_weights.py:104
        self.grid = grid
        self.georegions = georegions
        self.raster_weights = raster_weights
        self.project_dir = project_dir
        if default_to_area_weights is not None:                           # :112-119
            warnings.warn('default_to_area_weights is deprecated; use zero_weight="area" (True) or '
                          'zero_weight="drop" (False).', DeprecationWarning, stacklevel=2)
            zero_weight = "area" if default_to_area_weights else "drop"
        if zero_weight not in ZERO_WEIGHT_POLICIES:
            raise ValueError(f"zero_weight must be one of {sorted(ZERO_WEIGHT_POLICIES)}, got {zero_weight!r}")
        self.zero_weight = zero_weight
        if cosine_area is None:
            cosine_area = raster_weights is None                          # :133-135
        self.cosine_area = cosine_area
        self.weights: Optional[pd.DataFrame] = None
        self._csr_cache = {}

    @classmethod
    def from_frame(cls, weights: pd.DataFrame, grid: Grid, georegions: GeoRegions,
                   zero_weight: str = "nan") -> "GridWeights":
        """Wrap a precomputed weights frame (e.g. the reference's cached ``.feather``)."""
        for col in ("cell_id", "index_right", "weight"):
            if col not in weights.columns:
                raise ValueError(f"weights frame lacks column {col!r}")
        self = cls(grid, georegions, zero_weight=zero_weight)
        self.weights = weights
        return self

    def calculate_weights(self) -> None:
        """Exact area (x cos(lat), x secondary raster) weights for rectangular regions."""
        shp = self.georegions.shp
        need = {"lon_min", "lon_max", "lat_min", "lat_max"}
        lon, lat = self.grid.longitude, self.grid.latitude
        dlon, dlat = self.grid.resolution_lon, self.grid.resolution_lat
        if "rings" in shp.columns:
            self.weights = self._finish(self._polygon_area_weights(shp, lon, lat, dlon, dlat))
            return
        if not need.issubset(shp.columns):
            raise NotImplementedError(
                "calculate_weights() needs polygon rings (GeoRegions.from_polygons / from_shapefile) or "
                "rectangles (GeoRegions.from_rectangles); a frame computed elsewhere can be wrapped with "
                "GridWeights.from_frame")
        rows = []
        for ridx, r in zip(shp.index, shp.itertuples(index=False)):
            ox = np.clip(np.minimum(lon + dlon / 2, r.lon_max) - np.maximum(lon - dlon / 2, r.lon_min), 0, None)
            oy = np.clip(np.minimum(lat + dlat / 2, r.lat_max) - np.maximum(lat - dlat / 2, r.lat_min), 0, None)
            jy, jx = np.nonzero(oy)[0], np.nonzero(ox)[0]
            if len(jy) == 0 or len(jx) == 0:
                continue
            frac = np.outer(oy[jy], ox[jx]) / (dlon * dlat)               # cell area fraction inside
            yy, xx = np.meshgrid(jy, jx, indexing="ij")
            aw = frac * (np.cos(np.deg2rad(lat[yy])) if self.cosine_area else 1.0)
            rows.append(pd.DataFrame({"cell_id": (yy * len(lon) + xx).ravel(), "index_right": ridx,
                                      "area_weight": aw.ravel(), "longitude": lon[xx].ravel(),
                                      "latitude": lat[yy].ravel()}))
        self.weights = self._finish(pd.concat(rows, ignore_index=True))

    def _polygon_area_weights(self, shp, lon, lat, dlon, dlat) -> pd.DataFrame:
        """get_area_weights (aggfly/weights/grid_weights.py:379-421) for polygon regions: interior
        cells 1, border cells their covered fraction, x cos(latitude) when ``cosine_area``."""
        from .geometry import cell_overlaps
        region, cell, frac = cell_overlaps(list(shp["rings"]), lon, lat, dlon, dlat)
        yy, xx = cell // len(lon), cell % len(lon)
        aw = frac * (np.cos(np.radians(lat[yy])) if self.cosine_area else 1.0)
        return pd.DataFrame({"cell_id": np.asarray(self.grid.cell_id)[cell], "index_right": np.asarray(shp.index)[region],
                             "area_weight": aw, "longitude": lon[xx], "latitude": lat[yy]})

    def _finish(self, w: pd.DataFrame) -> pd.DataFrame:
        """Secondary-raster weighting, zero-weight policy, region ids (grid_weights.py:423-521, 194-196)."""
        shp = self.georegions.shp
        if self.raster_weights is None:
            w["weight"] = w["area_weight"]
        else:
            # aggfly/weights/grid_weights.py:447-489: missing / non-finite raster values count as zero;
            # weight = area_weight * raster_weight / (the region's summed raster_weight)
            raster = self.raster_weights
            if isinstance(raster, SecondaryWeights):
                raster = raster.on_grid(self.grid)
            # raster is laid out like the weights grid: position = lat_index * n_lon + lon_index
            pos = np.searchsorted(np.asarray(self.grid.cell_id), w["cell_id"].to_numpy()) \
                if not np.array_equal(self.grid.cell_id, np.arange(len(self.grid.cell_id))) else w["cell_id"].to_numpy()
            rw = np.asarray(raster, dtype=float).reshape(-1)[pos]
            n_missing = int((~np.isfinite(rw)).sum())
            if n_missing:
                warnings.warn(f"{n_missing} of {len(w)} cell-region pairs had no secondary raster value (outside its "
                              "extent, or entirely nodata) and were given zero weight. A region with no valid cells "
                              "at all falls back to whatever the zero_weight policy specifies.", stacklevel=2)
            w["raster_weight"] = np.where(np.isfinite(rw), rw, 0.0)
            tot = w["raster_weight"].groupby(w["index_right"]).transform("sum")
            w["total_weight"] = tot
            with np.errstate(invalid="ignore", divide="ignore"):
                w["weight"] = np.where(tot > 0, w["area_weight"] * (w["raster_weight"] / tot), 0.0)
            empty = ~(tot > 0)
            w["zero_weight"] = empty
            if empty.any():
                if self.zero_weight == "area":
                    warnings.warn("regions with no secondary weight fall back to AREA weights", UserWarning)
                    w.loc[empty, "weight"] = w.loc[empty, "area_weight"]
                elif self.zero_weight == "drop":
                    warnings.warn("regions with no secondary weight are DROPPED", UserWarning)
                    w = w.loc[~empty].reset_index(drop=True)
        self._csr_cache.clear()
        return shp[[self.georegions.regionid]].merge(w, right_on="index_right", left_index=True)


class SecondaryWeights:
    """A secondary raster (population, cropland ...) on its own regular lat/lon grid
    (aggfly/weights/secondary_weights.py:13-109 without rioxarray): ``on_grid`` averages it onto the
    climate grid like ``rescale_raster_to_grid`` (``Resampling.average``, nodata excluded)."""

    def __init__(self, values, latitude, longitude, nodata: Optional[float] = None, name: Optional[str] = None):
        self.values = np.asarray(values, dtype=float)
        if self.values.ndim == 3 and self.values.shape[0] == 1:
            self.values = self.values[0]                        # a single band
        self.latitude = np.asarray(latitude, dtype=float)
        self.longitude = np.asarray(longitude, dtype=float)
        self.nodata, self.name = nodata, name

    def on_grid(self, grid: Grid) -> np.ndarray:
        from .geometry import rescale_raster_to_grid
        return rescale_raster_to_grid(self.values, self.latitude, self.longitude, grid.latitude, grid.longitude,
                                      grid.resolution_lat, grid.resolution_lon, self.nodata)


def weights_from_objects(clim: Dataset, georegions: GeoRegions, secondary_weights=None,
                         project_dir: Optional[str] = None, **kwargs) -> GridWeights:
    """aggfly/weights/grid_weights.py:614-648: the weight grid is the dataset's grid after the
    0-360 -> -180..180 relabel + sort."""
    if clim.lon_is_360:
        order = clim.lon_sort_order()
        grid = Grid(lon_to_180(clim.longitude)[order], clim.latitude, clim.name, False)
    else:
        grid = Grid(clim.longitude, clim.latitude, clim.name, False)
    return GridWeights(grid, georegions, secondary_weights, project_dir=project_dir, **kwargs)


# ---------------------------------------------------------------------------------------------
# CSR lowering
# ---------------------------------------------------------------------------------------------
@dataclass
class HostCSR:
    row_ptr: np.ndarray      # int32[R+1]
    cell_idx: np.ndarray     # int32[nnz], raster memory order
    w: np.ndarray            # float64[nnz]
    region_ids: np.ndarray   # sorted unique index_right (row r <-> region_ids[r])
    n_cells: int

    @property
    def n_regions(self) -> int:
        return len(self.region_ids)

    @property
    def nnz(self) -> int:
        return len(self.w)


def lower_to_csr(wdf: pd.DataFrame, grid_cell_id: np.ndarray, n_lat: int, n_lon: int,
                 lon_order: Optional[np.ndarray] = None) -> HostCSR:
    """weights frame -> CSR over (region row, raster cell).

    * rows: ``region_ids = sort(unique(index_right))``                       spatial.py:165-166
    * cols: position of ``cell_id`` in ``grid.cell_id``; absent cells dropped  :168-172
    * entries keep the frame's order inside each region (fp64 sum order of ``np.add.at``)
    * ``lon_order`` (argsort of the relabelled longitudes) maps a position in the sorted grid
      back to the raster's memory column: cell (i, j_sorted) -> i * n_lon + lon_order[j_sorted]
    """
    grid_cell_id = np.asarray(grid_cell_id)
    n_cells = n_lat * n_lon
    if len(grid_cell_id) != n_cells:
        raise ValueError(f"weights grid has {len(grid_cell_id)} cells, the raster has {n_cells}")
    region_ids = np.sort(pd.unique(wdf["index_right"]))
    rows = np.searchsorted(region_ids, wdf["index_right"].to_numpy())
    cid = wdf["cell_id"].to_numpy()
    if np.array_equal(grid_cell_id, np.arange(n_cells)):
        ok = (cid >= 0) & (cid < n_cells)
        pos = np.where(ok, cid, 0).astype(np.int64)
    else:
        sorter = np.argsort(grid_cell_id, kind="stable")
        loc = np.searchsorted(grid_cell_id, cid, sorter=sorter)
        loc = np.clip(loc, 0, n_cells - 1)
        pos = sorter[loc]
        ok = grid_cell_id[pos] == cid
    rows, pos = rows[ok], pos[ok]
    w = wdf["weight"].to_numpy(dtype=np.float64)[ok]
    if lon_order is not None:
        lon_order = np.asarray(lon_order, dtype=np.int64)
        pos = (pos // n_lon) * n_lon + lon_order[pos % n_lon]
    order = np.argsort(rows, kind="stable")                      # group by region, keep frame order
    rows, pos, w = rows[order], pos[order], w[order]
    row_ptr = np.zeros(len(region_ids) + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=len(region_ids)), out=row_ptr[1:])
    if len(pos) > np.iinfo(np.int32).max or n_lat * n_lon > np.iinfo(np.int32).max:
        raise ValueError(f"{len(pos)} weight entries / {n_lat * n_lon} cells exceed the 32-bit indices of the device CSR")
    return HostCSR(row_ptr.astype(np.int32), pos.astype(np.int32), np.ascontiguousarray(w),
                   region_ids, n_cells)


# ---------------------------------------------------------------------------------------------
# on-disk cache of the lowered CSR (next to the reference's feather cache, same keying idea)
# ---------------------------------------------------------------------------------------------
def csr_cache_key(wdf: pd.DataFrame, grid_cell_id: np.ndarray, n_lat: int, n_lon: int,
                  lon_order: Optional[np.ndarray]) -> str:
    """sha256 over everything the lowering depends on, truncated like the reference's cache ids
    (aggfly/cache/project_cache.py:207-226 keeps 15 hex digits)."""
    import hashlib
    h = hashlib.sha256()
    for col in ("cell_id", "index_right", "weight"):
        h.update(np.ascontiguousarray(wdf[col].to_numpy()).tobytes())
    h.update(np.ascontiguousarray(np.asarray(grid_cell_id)).tobytes())
    h.update(np.asarray([n_lat, n_lon], dtype=np.int64).tobytes())
    h.update(b"none" if lon_order is None else np.ascontiguousarray(np.asarray(lon_order, dtype=np.int64)).tobytes())
    return h.hexdigest()[:15]


def lower_to_csr_cached(wdf: pd.DataFrame, grid_cell_id: np.ndarray, n_lat: int, n_lon: int,
                        lon_order: Optional[np.ndarray] = None, project_dir: Optional[str] = None) -> HostCSR:
    """``lower_to_csr`` memoised under ``{project_dir}/tmp/DeviceCSR/mod-{sha}/{sha}.npz`` (the layout of
    the reference's ProjectCache, aggfly/cache/project_cache.py:26-57): the yearly loop of a pipeline and
    the ranks of a sharded run lower the weights once."""
    import os
    if not project_dir:
        return lower_to_csr(wdf, grid_cell_id, n_lat, n_lon, lon_order)
    sha = csr_cache_key(wdf, grid_cell_id, n_lat, n_lon, lon_order)
    d = os.path.join(project_dir, "tmp", "DeviceCSR", f"mod-{sha}")
    path = os.path.join(d, f"{sha}.npz")
    if os.path.exists(path):
        try:
            with np.load(path) as z:
                rid = z["region_ids"]
                if "region_ids_object" in z.files and bool(z["region_ids_object"]):
                    rid = rid.astype(object)                       # string ids come back as the object array they were
                return HostCSR(z["row_ptr"], z["cell_idx"], z["w"], rid, int(z["n_cells"]))
        except Exception as exc:                                   # unreadable cache entry: say so and rebuild it
            import warnings
            warnings.warn(f"DeviceCSR cache entry {path} is unreadable ({type(exc).__name__}: {exc}); rebuilding it")
    csr = lower_to_csr(wdf, grid_cell_id, n_lat, n_lon, lon_order)
    os.makedirs(d, exist_ok=True)
    tmp = f"{path}.{os.getpid()}.tmp.npz"
    rid = np.asarray(csr.region_ids)
    as_object = rid.dtype == object
    if as_object:                                                  # np.savez would pickle an object array (unloadable by default)
        if not all(isinstance(v, str) for v in rid):
            return csr                                             # mixed / exotic ids: not cached
        rid = rid.astype(str)
    np.savez(tmp, row_ptr=csr.row_ptr, cell_idx=csr.cell_idx, w=csr.w, region_ids=rid, region_ids_object=np.bool_(as_object),
             n_cells=np.int64(csr.n_cells))
    os.replace(tmp, path)                                          # atomic: ranks may race
    return csr
