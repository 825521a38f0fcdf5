"""aggfly_b200 -- a B200-native (sm_100a) engine for aggfly's ``aggregate_dataset`` hot path.

Drop-in surface (same names as ``import aggfly as af``): ``aggregate_dataset``,
``aggregate_time``, ``aggregate_space``, ``TemporalAggregator``, ``SpatialAggregator``, ``Dataset``,
``Grid``, ``GridWeights``, ``GeoRegions``, ``weights_from_objects``.  Everything numeric runs in
hand-written CUDA kernels behind the C-ABI in ``include/aggfly_b200.h``; there is no CPU fallback.
"""
from .aggregate import (ALLOWED_ENGINE, SpatialAggregator, aggregate_dataset, aggregate_dataset_table, aggregate_space,
                        aggregate_time, distributed_client, is_distributed, resolve_engine, shutdown_dask_client,
                        start_dask_client)
from .dataset import Dataset, Grid, PackedRaster, RasterArray, TimeConcat, lon_to_180, lon_to_360
from .io import (crop_weights_from_path, dataset_from_path, dataset_to_zarr, georegions_from_path, pop_weights_from_path,
                 secondary_weights_from_path, shapefile_info, write_output, write_table, zarr_from_path)
from .shard import aggregate_dataset_sharded
from .spec import TemporalAggregator
from .timeaxis import CalendarIndex, CFDate, group_bounds, translate_groupby
from .weights import GeoRegions, GridWeights, SecondaryWeights, lower_to_csr, weights_from_objects

__version__ = "0.1.0"

__all__ = [
    "aggregate_dataset", "aggregate_time", "aggregate_space", "TemporalAggregator", "SpatialAggregator",
    "resolve_engine", "ALLOWED_ENGINE", "Dataset", "Grid", "RasterArray", "GridWeights", "GeoRegions",
    "weights_from_objects", "SecondaryWeights", "lower_to_csr", "CalendarIndex", "CFDate", "group_bounds", "translate_groupby",
    "lon_to_180", "lon_to_360", "dataset_from_path", "georegions_from_path", "secondary_weights_from_path",
    "write_output", "aggregate_dataset_sharded", "dataset_to_zarr", "zarr_from_path", "shapefile_info", "pop_weights_from_path", "crop_weights_from_path",
    "start_dask_client", "shutdown_dask_client", "is_distributed", "distributed_client", "PackedRaster", "TimeConcat", "aggregate_dataset_table", "write_table",
]
