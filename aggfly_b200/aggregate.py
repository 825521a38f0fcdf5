"""The reference's public aggregation API on the CUDA engine.

Same names, signatures, return types and error behaviour as ``aggfly/aggregate/aggregate.py``
(``aggregate_dataset`` :210-282, ``aggregate_time`` :101-162, ``aggregate_space`` :165-198) and
``aggfly/aggregate/spatial.py`` (``SpatialAggregator``); the only API addition is the engine value
``"cuda"`` (what ``"auto"`` resolves to here).  ``"dask"`` / ``"numba"`` are accepted so existing
scripts keep running, and are served by the same CUDA path -- there is no CPU engine in this
package.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Union

import numpy as np
import pandas as pd

from . import _lib as _lib_mod
from . import engine as _engine
from . import stream as _stream
from .dataset import Dataset
from .spec import Graph, Planner, TemporalAggregator, compile_spec
from .timeaxis import label_values, labels_equal
from .weights import GridWeights, lower_to_csr_cached

ALLOWED_ENGINE = ("auto", "cuda", "dask", "numba")          # reference: cli/config.py:27 + "cuda"

_DEPRECATED_CLUSTER_KWARGS = ("n_workers", "threads_per_worker", "processes", "memory_limit", "cluster_args")


def resolve_engine(engine: str, da=None, calc: str = None) -> str:
    """aggfly/aggregate/nb_kernels.py:59-74 -- every known engine resolves to "cuda" here."""
    if engine not in ALLOWED_ENGINE:
        raise ValueError(f"engine must be 'cuda', 'dask', 'numba', or 'auto', got {engine!r}")
    return "cuda"


# ---------------------------------------------------------------------------------------------
# dask client helpers of the reference (aggfly/aggregate/aggregate_utils.py:9-102, exported at
# aggfly/__init__.py:1-11).  There is no dask cluster behind this engine: scripts that bracket their calls with
# start_dask_client() / shutdown_dask_client() keep running, each helper says so ONCE and does nothing.
# ---------------------------------------------------------------------------------------------
_SHIM_WARNED = set()
_SHIM_ARGS = None          # what start_dask_client was called with (shutdown_dask_client returns it, like :89-102)


def _shim_warn(name: str) -> None:
    if name not in _SHIM_WARNED:
        _SHIM_WARNED.add(name)
        warnings.warn(f"aggfly_b200.{name}() is a no-op: the CUDA engine of the current device replaces the dask cluster "
                      "(multi-GPU runs are launched with torchrun, see aggregate_dataset_sharded)", UserWarning, stacklevel=3)


def is_distributed() -> bool:
    """aggregate_utils.py:9-23 -- True when a dask client is running; never the case here."""
    _shim_warn("is_distributed")
    return False


def distributed_client():
    """aggregate_utils.py:25-35 -- the global dask client; always None here."""
    _shim_warn("distributed_client")
    return None


def start_dask_client(n_workers: int = 2, threads_per_worker: int = 2, cap_numba_threads: int = 1, **kwargs):
    """aggregate_utils.py:38-86 -- accepted for drop-in compatibility, starts nothing, returns None."""
    global _SHIM_ARGS
    _shim_warn("start_dask_client")
    _SHIM_ARGS = {"n_workers": n_workers, "threads_per_worker": threads_per_worker,
                  "cap_numba_threads": cap_numba_threads, **kwargs}
    return None


def shutdown_dask_client():
    """aggregate_utils.py:89-102 -- returns the arguments of the matching start_dask_client() call, or None."""
    global _SHIM_ARGS
    _shim_warn("shutdown_dask_client")
    args, _SHIM_ARGS = _SHIM_ARGS, None
    return args


# ---------------------------------------------------------------------------------------------
# temporal
# ---------------------------------------------------------------------------------------------
class _DeviceRaster:
    """The call's raster on the device, flattened to [T, cells]."""

    def __init__(self, dataset: Dataset):
        self.tensor = _engine.to_device(dataset.values)
        T = self.tensor.shape[0]
        self.n_lat, self.n_lon = len(dataset.latitude), len(dataset.longitude)
        self.n_cells = self.n_lat * self.n_lon
        self.flat = self.tensor.reshape(T, self.n_cells)


def _compile(dataset: Dataset, aggregator_dict: Optional[Dict[str, list]]):
    graph = Graph(dataset.dtype, dataset.time, getattr(dataset, "pre_ops", None))
    if aggregator_dict is None:
        outputs = {"variable": graph.raw}                               # aggregate.py:269-270
    else:
        outputs = compile_spec(graph, aggregator_dict)
    return graph, outputs


def _spec_fingerprint(aggregator_dict) -> Optional[tuple]:
    """Hashable identity of a spec by CONTENT (None: not cacheable).  A spec that carries an opaque object -- an
    ``inter`` Dataset / tensor, a large array -- is never cached: its identity (``id``) says nothing about its
    contents, which the caller may have reassigned or mutated since the plan captured them."""
    def fp(v):
        if isinstance(v, np.ndarray):
            if v.size > 64:
                raise TypeError("large array: not cacheable")
            return ("arr", v.dtype.str, v.shape, v.tobytes())
        if isinstance(v, (list, tuple)):
            return tuple(fp(x) for x in v)
        if isinstance(v, dict):
            return tuple(sorted((k, fp(x)) for k, x in v.items()))
        if isinstance(v, TemporalAggregator):
            return ("agg", v.calc, v.groupby, fp(v.ddargs))
        if isinstance(v, (str, int, float, bool, type(None), np.integer, np.floating)):
            return (type(v).__name__, v)
        raise TypeError("opaque value: not cacheable")
    try:
        return None if aggregator_dict is None else tuple((k, fp(v)) for k, v in aggregator_dict.items())
    except Exception:
        return None


_PLAN_CACHE: Dict[tuple, tuple] = {}          # a yearly loop plans the same spec on same-shaped axes again and again


def _plan(dataset: Dataset, aggregator_dict: Optional[Dict[str, list]]):
    """One stage for the whole call; every output must end on the same time axis (they are merged
    into one panel -- the reference would union the axes and NaN-fill, spatial.py:90-92).  Plans are
    memoised on (spec, time axis, dtype, preprocess): planning costs 3-4 ms, as much as scanning a
    global year on the device."""
    t = dataset.time
    tkey = (id(t), len(t))
    key = (_spec_fingerprint(aggregator_dict), tkey, str(dataset.dtype), tuple(getattr(dataset, "pre_ops", [])))
    hit = _PLAN_CACHE.get(key) if key[0] is not None else None
    if hit is not None and hit[0] is t:                     # the id is only trusted while the object is alive
        return hit[1], hit[2]
    names, stage = _plan_uncached(dataset, aggregator_dict)
    if key[0] is not None:
        if len(_PLAN_CACHE) >= 16:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = (t, names, stage)
    return names, stage


def _plan_uncached(dataset: Dataset, aggregator_dict: Optional[Dict[str, list]]):
    graph, outputs = _compile(dataset, aggregator_dict)
    names = list(outputs.keys())
    nodes = [outputs[n] for n in names]
    first = graph.labels(nodes[0])
    for n, node in zip(names, nodes):
        if not labels_equal(graph.labels(node), first):
            raise ValueError(f"output {n!r} ends on a different time axis than {names[0]!r}; all outputs "
                             "of one call must share one output time axis")
    return names, Planner(graph).plan(nodes)


def _is_device_tensor(values) -> bool:
    return type(values).__module__.startswith("torch") and values.is_cuda


def _temporal_device(dataset: Dataset, aggregator_dict, target_stripes: int = 0):
    """Temporal chains of one call -> X / V on the device.  A raster already on the device is
    scanned in place; a host raster is streamed (chunked copies overlapped with the kernels)."""
    names, stage = _plan(dataset, aggregator_dict)
    if _is_device_tensor(dataset.values):
        raster = _DeviceRaster(dataset)
        res = _engine.run_stage(stage, raster.flat, raster.n_cells, target_stripes=target_stripes)
        return names, res, raster
    import torch
    n_cells = len(dataset.latitude) * len(dataset.longitude)
    whole_time_chunks = getattr(dataset.values, "single_time_chunk", False)
    if not (target_stripes or _engine.OPTIONS["target_stripes"]) and not whole_time_chunks:
        # stripes of about four copy chunks: the kernels of a stripe start when its rows have landed.  (A store
        # whose chunks span the whole time axis delivers every row at once: planned like a resident raster.)
        nbytes = int(np.prod(dataset.shape)) * dataset.dtype.itemsize
        target_stripes = -int(min(64, nbytes // (4 * _stream.OPTIONS["chunk_bytes"])))
        if nbytes > _stream.OPTIONS["device_raster_budget_bytes"]:
            # a record longer than the device budget goes through a ring of device windows: one stripe per window
            target_stripes = -int(-(-nbytes // _stream.OPTIONS["ring_slot_bytes"]))
    import time
    t0 = time.perf_counter()
    runner = _engine.StageRunner(stage, n_cells, target_stripes=target_stripes)
    t1 = time.perf_counter()
    try:
        res, raster = _stream.feed_and_run(runner, dataset.values, n_cells)
        t2 = time.perf_counter()
    finally:
        torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        runner.close()
    _stream.check_device_decompress()
    global LAST_FEED_TRACE
    LAST_FEED_TRACE = {"runner_ms": (t1 - t0) * 1e3, "issue_ms": (t2 - t1) * 1e3, "drain_ms": (t3 - t2) * 1e3,
                       "close_ms": (time.perf_counter() - t3) * 1e3}
    return names, res, raster


LAST_FEED_TRACE: dict = {}


def _temporal_regional(dataset: Dataset, aggregator_dict, csr):
    """Many-period panels: temporal scan + regional average in one kernel (engine.RegionalRunner), no per-cell X.
    Returns (names, PanelResult) or None when the call is not a candidate / the library has no instantiation for it --
    the caller then runs the two-kernel path."""
    names, stage = _plan(dataset, aggregator_dict)
    n_lat, n_lon = len(dataset.latitude), len(dataset.longitude)
    if not _engine.regional_candidate(stage, n_lon):
        return None
    import torch
    runner = _engine.RegionalRunner(stage, csr, n_lat, n_lon)
    try:
        if not runner.supported:
            return None
        # scratch rows of the regions that straddle tiles + the panel must fit next to the raster (a very long daily
        # record): otherwise the two-kernel path decides (it streams X per call and fails loudly if that does not fit either)
        need = int(runner.info.workspace_bytes) + csr.host.n_regions * len(stage.labels) * len(names) * 8
        if need > 0.9 * torch.cuda.mem_get_info()[0]:
            return None
        try:
            if _is_device_tensor(dataset.values):
                raster = _DeviceRaster(dataset)
                res = runner.run(raster.flat)
            else:
                res, raster = _stream.feed_and_run(runner, dataset.values, n_lat * n_lon)
        except _lib_mod.AgfUnsupported:                     # e.g. a raster view that is not 16-byte aligned
            return None
        finally:
            torch.cuda.current_stream().synchronize()
    finally:
        runner.close()
    _stream.check_device_decompress()
    return names, res


def aggregate_time(dataset: Dataset, weights: GridWeights = None,
                   aggregator_dict: Dict[str, Union[list, TemporalAggregator]] = None,
                   engine: str = "auto", **kwargs) -> Dict[str, Dataset]:
    """Temporal chains only: {output name: Dataset with values[G, lat, lon]} (aggregate.py:101-162).
    Outputs may end on different time axes here (one stage per axis)."""
    resolve_engine(engine)
    if aggregator_dict is None:
        if not kwargs:
            raise ValueError("No arguments provided.")
        aggregator_dict = kwargs
    graph, outputs = _compile(dataset, aggregator_dict)
    groups: List[List[str]] = []
    for name, node in outputs.items():
        for grp in groups:
            if labels_equal(graph.labels(outputs[grp[0]]), graph.labels(node)):
                grp.append(name)
                break
        else:
            groups.append([name])
    raster = _DeviceRaster(dataset)
    planner = Planner(graph)
    done: Dict[str, Dataset] = {}
    for grp in groups:
        res = _engine.run_stage(planner.plan([outputs[n] for n in grp]), raster.flat, raster.n_cells)
        X = res.X.cpu().numpy()                                         # [G, cells, n_cols]
        for c, name in enumerate(grp):
            vals = X[:, :, c].astype(res.nodes[c].dtype, copy=False)
            vals = np.ascontiguousarray(vals).reshape(len(res.labels), raster.n_lat, raster.n_lon)
            done[name] = Dataset.from_arrays(vals, res.labels, dataset.latitude, dataset.longitude,
                                             lon_is_360=dataset.lon_is_360, name=name)
    return {name: done[name] for name in outputs}


def _execute_single_step(agg: TemporalAggregator, dataset: Dataset):
    """TemporalAggregator.execute (temporal.py:165-263): a Dataset, or a list for multi-ddargs."""
    out = aggregate_time(dataset, None, {"_step": [("aggregate", agg)]})
    vals = list(out.values())
    return vals if agg.multi_dd else vals[0]


# ---------------------------------------------------------------------------------------------
# spatial
# ---------------------------------------------------------------------------------------------
def _device_csr(weights: GridWeights, dataset: Dataset) -> _engine.DeviceCSR:
    import torch
    lon_order = dataset.lon_sort_order() if dataset.lon_is_360 else None
    key = (torch.cuda.current_device(), len(dataset.latitude), len(dataset.longitude),
           None if lon_order is None else lon_order.tobytes(), id(weights.weights))
    cache = getattr(weights, "_csr_cache", None)
    if cache is None:
        cache = {}
        try:
            weights._csr_cache = cache
        except Exception:
            pass
    if key not in cache:
        host = lower_to_csr_cached(weights.weights, weights.grid.cell_id, len(dataset.latitude),
                                   len(dataset.longitude), lon_order, getattr(weights, "project_dir", None))
        cache[key] = _engine.DeviceCSR(host)
    return cache[key]


def _zero_weight_regions(weights) -> np.ndarray:
    """Regions whose weights sum to <= 0 (spatial.py:144-151), computed once per weights frame: the
    groupby over ~1 M rows would otherwise be the largest host cost of a yearly call."""
    frame = weights.weights
    hit = getattr(weights, "_zero_regions", None)
    if hit is None or hit[0] is not frame:
        wsum = frame.groupby("index_right")["weight"].sum()
        hit = (frame, np.asarray(wsum.index[~(wsum > 0)]))
        try:
            weights._zero_regions = hit
        except Exception:
            pass
    return hit[1]


def _assemble_panel(panel: np.ndarray, names: List[str], labels, region_ids: np.ndarray,
                    weights: GridWeights) -> pd.DataFrame:
    """spatial.py:136-153: long frame (regions outer, periods inner), then the row-drop rules.  The kept rows
    are selected first and the frame is built once from them (a daily panel has 16 M candidate rows)."""
    R, G, NC = panel.shape
    flat = panel.reshape(R * G, NC)
    keep = ~np.isnan(flat).any(axis=1)
    if getattr(weights, "zero_weight", "area") == "nan":
        zero = _zero_weight_regions(weights)
        if len(zero):
            keep |= np.repeat(np.isin(region_ids, zero), G)
    idx = np.flatnonzero(keep)
    data = {"region_id": np.asarray(region_ids)[idx // G], "time": label_values(labels)[idx % G]}
    sel = flat[idx]
    for c, nm in enumerate(names):
        data[nm] = sel[:, c]
    return pd.DataFrame(data)


_D2H_SLOTS: dict = {}        # (dtype, elements) -> pinned staging slots of _to_host, kept across calls


def _to_host(t) -> np.ndarray:
    """Device tensor -> fresh NumPy array.  A daily panel is 1.7 GB: ``tensor.cpu()`` spent 0.8 s on it (2 GB/s -- the
    page faults of the fresh pageable destination, taken one at a time behind every staged copy).  Here 64 MB pieces go to
    a small ring of cached pinned slots at PCIe speed and worker threads copy them into the result, so the first-touch
    page faults of the destination are taken by several cores while the next pieces are already in flight."""
    import torch
    nbytes = t.numel() * t.element_size()
    if t.device.type != "cuda" or nbytes < (64 << 20):
        return t.cpu().numpy()
    from concurrent.futures import ThreadPoolExecutor
    flat = t.contiguous().view(-1)
    n = flat.shape[0]
    piece = (64 << 20) // t.element_size()
    key = (str(t.dtype), piece)
    if key not in _D2H_SLOTS:
        _D2H_SLOTS.clear()
        _D2H_SLOTS[key] = [torch.empty(piece, dtype=t.dtype, pin_memory=True) for _ in range(4)]
    slots = _D2H_SLOTS[key]
    out = np.empty(n, dtype=slots[0].numpy().dtype)
    stream = torch.cuda.current_stream(t.device)

    def land(slot, ev, a, m):
        ev.synchronize()
        np.copyto(out[a:a + m], slots[slot][:m].numpy())

    futs = [None] * len(slots)
    with ThreadPoolExecutor(max_workers=len(slots)) as pool:
        for i, a in enumerate(range(0, n, piece)):
            k = i % len(slots)
            if futs[k] is not None:
                futs[k].result()                                   # the slot's previous piece has left it
            m = min(piece, n - a)
            slots[k][:m].copy_(flat[a:a + m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            futs[k] = pool.submit(land, k, ev, a, m)
        for f in futs:
            if f is not None:
                f.result()
    return out.reshape(tuple(t.shape))


def _panel_frame(panel, names: List[str], labels, region_ids: np.ndarray, weights: GridWeights) -> pd.DataFrame:
    """The frame ``aggregate_dataset`` returns, from the columns of ``_panel_columns`` (no copies)."""
    got = _panel_columns(panel, names, labels, region_ids, weights)
    if isinstance(got, pd.DataFrame):
        return got
    data, index = got
    return pd.DataFrame(data, index=index, copy=False)


def _panel_table(panel, names: List[str], labels, region_ids: np.ndarray, weights: GridWeights):
    """The same rows as a ``pyarrow.Table`` built straight from the gathered columns (no pandas frame in between): what the
    panel writer takes (aggfly/cli/pipeline.py:150, 159-172 concatenates and writes pandas frames).  Labels of a
    non-standard calendar are written as ISO strings, like ``io.write_output`` does."""
    import pyarrow as pa
    got = _panel_columns(panel, names, labels, region_ids, weights)
    if isinstance(got, pd.DataFrame):                                  # literal route (duplicated region index, no device room)
        df = got.copy()
        if len(df) and not isinstance(df["time"].iloc[0], (pd.Timestamp, np.datetime64)):
            df["time"] = df["time"].map(lambda t: t.isoformat() if hasattr(t, "isoformat") else str(t))
        return pa.Table.from_pandas(df, preserve_index=False)
    data, _index = got
    cols = {}
    for k, v in data.items():
        if k == "time" and getattr(v, "dtype", None) == object:
            v = np.array([t.isoformat() if hasattr(t, "isoformat") else str(t) for t in v], dtype=object)
        cols[k] = pa.array(v)
    return pa.table(cols)


def _panel_columns(panel, names: List[str], labels, region_ids: np.ndarray, weights: GridWeights):
    """Device panel ``[R, G, n_cols]`` (torch, float64) -> ({column name: array}, index) of the long frame of
    spatial.py:136-153 with its row-drop rules, already joined with the region ids (aggregate.py:276-280).

    Same rows, order, columns and index as ``_assemble_panel`` + ``shp[[rid]].merge(...)``, but the row selection, the
    gather of the kept rows and the transposition to column-major happen on the device the panel is on; the host
    only receives contiguous columns (a daily panel has 16 M candidate rows x 14 columns: the NumPy / pandas route
    cost 2.8 s per call, far more than the copy and the kernels together)."""
    import torch
    shp = weights.georegions.shp
    rid = weights.georegions.regionid
    if shp.index.has_duplicates:                      # a join that repeats blocks: keep the literal route
        df = _assemble_panel(panel.cpu().numpy(), names, labels, region_ids, weights)
        return shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
    R, G, NC = (int(v) for v in panel.shape)
    dev = panel.device
    if dev.type == "cuda" and torch.cuda.mem_get_info(dev)[0] < 3 * panel.numel() * panel.element_size():
        # no room for the gathered + transposed copies next to the panel: the literal route on the host
        df = _assemble_panel(panel.cpu().numpy(), names, labels, region_ids, weights)
        return shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
    flat = panel.reshape(R * G, NC)
    keep = ~torch.isnan(flat).any(dim=1)
    if getattr(weights, "zero_weight", "area") == "nan":
        zero = _zero_weight_regions(weights)
        if len(zero):
            zmask = torch.from_numpy(np.isin(np.asarray(region_ids), zero)).to(dev)
            keep |= zmask[:, None].expand(R, G).reshape(-1)
    # the inner join with the region frame: regions in shp order, only those present on both sides
    pos = pd.Index(np.asarray(region_ids)).get_indexer(shp.index)          # shp row -> region row (-1: no cells anywhere)
    order = pos[pos >= 0]
    shp_row_of_region = np.full(R, -1, dtype=np.int64)
    shp_row_of_region[order] = np.flatnonzero(pos >= 0)
    region_col = shp[rid].array                          # taken from, so the column keeps the region frame's dtype
    identity = len(order) == R and bool((order == np.arange(R)).all())
    tvals = label_values(labels)
    if identity and bool(keep.all()):
        cols = _to_host(flat.t().contiguous())
        reg_col, time_col, index = region_col.take(np.repeat(shp_row_of_region, G)), np.tile(tvals, R), pd.RangeIndex(R * G)
    else:
        if identity:
            idx = torch.nonzero(keep).squeeze(1)
            index = None
        else:
            order_t = torch.from_numpy(order.astype(np.int64)).to(dev)
            cand = (order_t[:, None] * G + torch.arange(G, device=dev)[None, :]).reshape(-1)      # rows in output order
            idx = cand[keep[cand]]
            index = (torch.cumsum(keep.to(torch.int64), 0) - 1)[idx].cpu().numpy()     # the row's number in the un-joined frame
        cols = _to_host(flat.index_select(0, idx).t().contiguous())
        # region / period of every kept row, split on the device (two int32 columns instead of an int64 one + host div / mod)
        r_h = _to_host(torch.div(idx, G, rounding_mode="floor").to(torch.int32))
        g_h = _to_host((idx % G).to(torch.int32))
        reg_col, time_col = region_col.take(shp_row_of_region[r_h]), tvals[g_h]
        index = pd.RangeIndex(len(r_h)) if index is None else pd.Index(index)
    data = {rid: reg_col, "time": time_col}
    for c, nm in enumerate(names):
        data[nm] = cols[c]
    return data, index


class SpatialAggregator:
    """aggfly/aggregate/spatial.py:37-154 on the CUDA engine: takes temporally-reduced Datasets
    (one per output name, dims time x lat x lon, sharing one time axis)."""

    def __init__(self, dataset: Union[list, Dataset], weights: GridWeights, names: Union[str, List[str]] = "climate"):
        self.dataset = dataset if isinstance(dataset, list) else [dataset]
        self.weights_obj = weights
        self.grid = weights.grid
        self.weights = weights.weights
        self.names = [names] if isinstance(names, str) else list(names)
        self.zero_weight = getattr(weights, "zero_weight", "area")

    def compute(self, npartitions: int = None) -> pd.DataFrame:
        import torch
        d0 = self.dataset[0]
        for d in self.dataset[1:]:
            if not labels_equal(d.time, d0.time):
                raise ValueError("all outputs of one call must share one output time axis")
        n_cells = len(d0.latitude) * len(d0.longitude)
        G = len(d0.time)
        dt = np.result_type(*[d.dtype for d in self.dataset])
        X = torch.stack([_engine.to_device(d.values).reshape(G, n_cells).to(_engine._tdtype(dt))
                         for d in self.dataset], dim=2).contiguous()         # [G, cells, names]
        # validity mask = AND over names of ~isnan  (spatial.py:114-119), by the library's kernel
        V = torch.empty((G, n_cells), dtype=torch.uint8, device=X.device)
        _engine.valid_mask(X, dt, V)
        res = _engine.StageResult(X, V, dt, d0.time, None)
        csr = _device_csr(self.weights_obj, d0)
        panel = _engine.run_spmm(csr, res).cpu().numpy()
        return _assemble_panel(panel, self.names, d0.time, csr.host.region_ids, self.weights_obj)


def aggregate_space(dataset_dict: Dict[str, Dataset], weights: GridWeights, npartitions=None, **kwargs) -> pd.DataFrame:
    """aggregate.py:165-198."""
    return SpatialAggregator(list(dataset_dict.values()), weights, names=list(dataset_dict.keys())).compute()


# ---------------------------------------------------------------------------------------------
# the hot path
# ---------------------------------------------------------------------------------------------
def aggregate_dataset(weights: GridWeights, dataset: Dataset = None,
                      aggregator_dict: Dict[str, Union[list, TemporalAggregator]] = None,
                      dataset_dict=None, engine: str = "auto", **kwargs) -> pd.DataFrame:
    """Gridded raster -> region x period panel (aggregate.py:210-282).

    One pass: fused temporal kernel(s) over the raster, CSR weighted average onto regions, panel
    assembly.  Output frame: ``[regionid, "time", *names]``, regions in shapefile-row order, periods
    ascending, rows with NaN dropped (except zero-weight regions under ``zero_weight="nan"``).
    """
    if dataset is None:
        raise ValueError("No dataset provided.")
    resolve_engine(engine)
    stale = {k: kwargs.pop(k) for k in _DEPRECATED_CLUSTER_KWARGS if k in kwargs}
    if stale:
        warnings.warn(
            f"aggregate_dataset no longer builds a Dask cluster; {sorted(stale)} is/are ignored. "
            "aggfly_b200 runs on the CUDA engine of the current device.",
            DeprecationWarning, stacklevel=2)
    as_table = bool(kwargs.pop("_as_arrow_table", False))
    if aggregator_dict is None and kwargs:
        aggregator_dict = kwargs
    if aggregator_dict is None and dataset_dict is not None:
        df = aggregate_space(dataset_dict, weights)
        rid = weights.georegions.regionid
        return weights.georegions.shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
    tr = _Trace()
    csr = _device_csr(weights, dataset)               # cached on the weights object after the first call
    tr.mark("csr")
    fused = _temporal_regional(dataset, aggregator_dict, csr)
    if fused is not None:
        names, pres = fused
        tr.mark("temporal + regional, one kernel (+ host feed)")
        if as_table:
            df = _panel_table(pres.panel, names, pres.labels, csr.host.region_ids, weights)
        elif _engine.OPTIONS.get("device_panel_frame", False):
            df = _panel_frame(pres.panel, names, pres.labels, csr.host.region_ids, weights)
        else:
            df = _assemble_panel(pres.panel.cpu().numpy(), names, pres.labels, csr.host.region_ids, weights)
            rid = weights.georegions.regionid
            df = weights.georegions.shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
        tr.mark("panel frame")
        tr.done()
        return df
    names, res, raster = _temporal_device(dataset, aggregator_dict)
    tr.mark("temporal (+ host feed)")
    if as_table:
        panel = _engine.run_spmm(csr, res)
        tr.mark("spmm (issue)")
        df = _panel_table(panel, names, res.labels, csr.host.region_ids, weights)
        tr.mark("panel table (row selection on the device, d2h, arrow columns)")
    elif _engine.OPTIONS.get("device_panel_frame", False):
        panel = _engine.run_spmm(csr, res)
        tr.mark("spmm (issue)")
        df = _panel_frame(panel, names, res.labels, csr.host.region_ids, weights)
        tr.mark("panel frame (row selection on the device, d2h, region join)")
    else:
        panel = _engine.run_spmm(csr, res).cpu().numpy()
        tr.mark("spmm + d2h")
        df = _assemble_panel(panel, names, res.labels, csr.host.region_ids, weights)
        tr.mark("assemble")
        rid = weights.georegions.regionid
        df = weights.georegions.shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
        tr.mark("merge")
    tr.done()
    return df


def aggregate_dataset_table(weights: GridWeights, dataset: Dataset = None, aggregator_dict=None, engine: str = "auto", **kwargs):
    """``aggregate_dataset`` whose result is a ``pyarrow.Table`` with the same rows and columns, built straight from the
    panel's gathered columns -- what ``aggfly run`` concatenates and hands to the parquet / feather / csv writer
    (``io.write_table``) without a pandas frame in between (aggfly/cli/pipeline.py:150, 159-172)."""
    if aggregator_dict is None and not kwargs:
        raise ValueError("aggregate_dataset_table needs an aggregator_dict")
    return aggregate_dataset(weights, dataset, aggregator_dict, engine=engine, _as_arrow_table=True, **kwargs)


class _Trace:
    """Wall-clock phases of one call; printed when AGF_TRACE is set, always kept in LAST_TRACE."""

    def __init__(self):
        import time
        self._t = time.perf_counter
        self.t0 = self.last = self._t()
        self.phases = []

    def mark(self, name):
        now = self._t()
        self.phases.append((name, (now - self.last) * 1e3))
        self.last = now

    def done(self):
        import os
        global LAST_TRACE
        LAST_TRACE = {"total_ms": (self.last - self.t0) * 1e3, "phases_ms": dict(self.phases)}
        if os.environ.get("AGF_TRACE"):
            print("[aggfly_b200] " + ", ".join(f"{k} {v:.1f} ms" for k, v in self.phases), flush=True)


LAST_TRACE: dict = {}
