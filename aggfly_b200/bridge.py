"""Adapters from the reference's objects to this engine (route A).

``aggfly.aggregate_dataset(..., engine="cuda")`` would call ``aggregate_dataset_from_aggfly`` with
the reference's own ``GridWeights`` / ``Dataset`` instances.  Everything is duck-typed on the
attributes the reference's spatial/temporal code reads (aggfly/aggregate/aggregate.py:276-280,
aggfly/aggregate/spatial.py:57-69, 86-103), so neither aggfly nor xarray is imported here.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd

from .dataset import Dataset, Grid
from .timeaxis import CalendarIndex
from .weights import GeoRegions, GridWeights


def _time_index(idx):
    """DatetimeIndex stays; a cftime index (no cftime here: read its fields) -> CalendarIndex."""
    if isinstance(idx, CalendarIndex):
        return idx
    if type(idx).__name__ == "CFTimeIndex" or hasattr(idx, "calendar"):
        hour = getattr(idx, "hour", None)
        return CalendarIndex(idx.calendar, np.asarray(idx.year), np.asarray(idx.month), np.asarray(idx.day),
                             None if hour is None else np.asarray(hour))
    return pd.DatetimeIndex(idx)


def dataset_from_aggfly(ds) -> Dataset:
    """The reference's ``Dataset`` (an ``xarray.DataArray`` wrapper with dims normalised to
    ``latitude, longitude, time``; aggfly/dataset/grid_utils.py:299-324) or anything shaped like it.

    The raster is transposed to the time-major layout the kernels scan (a no-op view for files
    stored time-major, which is how ERA5/CMIP6 netCDF and the reference's converted zarr stores are
    laid out on disk) and materialised (``np.asarray`` computes a dask-backed array)."""
    if isinstance(ds, Dataset):
        return ds
    da = ds.da
    dims = list(da.dims)
    tdim = "time" if "time" in dims else dims[-1]
    ydim = "latitude" if "latitude" in dims else dims[0]
    xdim = "longitude" if "longitude" in dims else dims[1]
    if hasattr(da, "transpose"):
        values = np.asarray(da.transpose(tdim, ydim, xdim).values)
        coord = lambda d: da.get_index(d) if hasattr(da, "get_index") else da.coords[d]     # noqa: E731
    else:                                                   # RasterArray-like: data + dims + coords
        values = np.transpose(np.asarray(da.data), [dims.index(tdim), dims.index(ydim), dims.index(xdim)])
        coord = lambda d: da.coords[d]                                                      # noqa: E731
    return Dataset.from_arrays(values, _time_index(coord(tdim)), np.asarray(coord(ydim), dtype=float),
                               np.asarray(coord(xdim), dtype=float), lon_is_360=bool(getattr(ds, "lon_is_360", True)),
                               name=getattr(ds, "name", None))


def weights_from_aggfly(weights) -> GridWeights:
    """The reference's ``GridWeights``: only the frame, the grid's ``cell_id`` vector, the region
    frame and the ``zero_weight`` policy cross over (aggfly/aggregate/spatial.py:57-69)."""
    if isinstance(weights, GridWeights):
        return weights
    g = weights.grid
    grid = Grid(np.asarray(g.longitude, dtype=float), np.asarray(g.latitude, dtype=float),
                getattr(g, "name", None), bool(getattr(g, "lon_is_360", False)))
    shp = weights.georegions.shp
    regionid = weights.georegions.regionid
    frame = pd.DataFrame({regionid: np.asarray(shp[regionid])}, index=shp.index)
    out = GridWeights.from_frame(weights.weights, grid, GeoRegions(frame, regionid),
                                 zero_weight=getattr(weights, "zero_weight", "area"))
    ref_ids = np.asarray(g.cell_id).reshape(-1)
    if len(ref_ids) != len(grid.cell_id):
        raise ValueError(f"weights.grid.cell_id has {len(ref_ids)} entries for a {len(grid.latitude)} x "
                         f"{len(grid.longitude)} grid")
    grid.cell_id = ref_ids                                   # whatever numbering the frame refers to
    return out


def aggregate_dataset_from_aggfly(weights, dataset, aggregator_dict: Optional[dict] = None, **kwargs) -> pd.DataFrame:
    from .aggregate import aggregate_dataset
    return aggregate_dataset(weights_from_aggfly(weights), dataset_from_aggfly(dataset),
                             aggregator_dict=aggregator_dict, engine="cuda", **kwargs)
