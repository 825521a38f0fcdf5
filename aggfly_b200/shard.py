"""Time sharding across the GPUs of one box: one process per GPU, replicated CSR, one panel gather.

The reference processes years serially and concatenates the yearly frames
(aggfly/cli/pipeline.py:138-150); inside one call the spatial step is one task per time chunk
(aggfly/aggregate/spatial.py:189-199).  Panel rows of different periods never interact, so the
time axis is cut at *outer-period* boundaries (years by default), every rank aggregates its
periods with its own replica of the weights, and the only exchange is one all-gather of the small
``float64[R, G_local, n_cols]`` panels (NCCL over NVLink/NVSwitch for CUDA tensors; gloo for the
CPU tests of this host logic).  There is no collective on the data path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from .timeaxis import CalendarIndex, group_bounds

SHARD_FREQ = {"year": "YE", "month": "ME", "date": "1D"}


def balanced_ranges(n_units: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [begin, end) unit ranges, earlier ranks take the remainder; ranks beyond
    ``n_units`` get empty ranges."""
    base, extra = divmod(n_units, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def plan_time_shards(time, world: int, shard_by: str = "year") -> List[Tuple[int, int]]:
    """Row ranges [row0, row1) per rank, cut at ``shard_by`` period boundaries of ``time`` and
    balanced by ROW count (a leading partial year counts for what it holds)."""
    if shard_by not in SHARD_FREQ:
        raise ValueError(f"shard_by must be one of {sorted(SHARD_FREQ)}, got {shard_by!r}")
    bounds, _ = group_bounds(time, SHARD_FREQ[shard_by])
    bounds = np.asarray(bounds, dtype=np.int64)
    cuts = np.unique(bounds[1:])                             # candidate end rows (ascending)
    total = int(bounds[-1])
    out, row0 = [], int(bounds[0])
    for r in range(world):
        if r == world - 1:
            row1 = total
        else:
            want = int(bounds[0]) + (total - int(bounds[0])) * (r + 1) / world
            j = int(np.argmin(np.abs(cuts - want)))
            row1 = max(row0, int(cuts[j]))
        out.append((row0, row1))
        row0 = row1
    return out


def check_shardable(aggregator_dict: Optional[dict], shard_by: str) -> None:
    """Every group of every aggregate step must lie inside one shard: periods nest as
    date < month < year; weeks straddle all of them."""
    order = {"date": 0, "month": 1, "year": 2}
    if aggregator_dict is None:
        return
    for name, steps in aggregator_dict.items():
        for kind, params in steps:
            if kind != "aggregate":
                continue
            gb = params.groupby if hasattr(params, "groupby") and not isinstance(params, dict) else params.get("groupby")
            gb = {"1D": "date", "ME": "month", "YE": "year", "W": "week"}.get(gb, gb)
            if gb == "week":
                raise ValueError(f"output {name!r} groups by week, which straddles {shard_by} boundaries; "
                                 "time sharding needs date / month / year steps")
            if order[gb] > order[shard_by]:
                raise ValueError(f"output {name!r} groups by {gb}, coarser than the shard period {shard_by!r}")


def gather_panels(local, g_sizes: Sequence[int], group=None):
    """All-gather per-rank panels ``[R, G_r, NC]`` (torch tensors, same R / NC / dtype / device on every
    rank; ``g_sizes[r]`` = G of rank r, known to everybody from the shard plan) into one
    ``[R, sum(G_r), NC]`` tensor, periods in rank order.  One collective, padded to max(G_r)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(g_sizes) == world
    R, _, NC = local.shape
    gmax = max(1, max(g_sizes))
    send = local
    if local.shape[1] != gmax:
        send = torch.zeros((R, gmax, NC), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    recv = torch.empty(world * R * gmax * NC, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous().view(-1), group=group)     # flat: same on nccl and gloo
    recv = recv.view(world, R, gmax, NC)
    return torch.cat([recv[r, :, : g_sizes[r]] for r in range(world)], dim=1)


def concat_labels(parts):
    parts = [p for p in parts if p is not None and len(p)]
    if isinstance(parts[0], CalendarIndex):
        return CalendarIndex(parts[0].calendar, np.concatenate([p.year for p in parts]),
                             np.concatenate([p.month for p in parts]), np.concatenate([p.day for p in parts]),
                             np.concatenate([p.hour for p in parts]))
    return pd.DatetimeIndex(np.concatenate([pd.DatetimeIndex(p).values for p in parts]))


def aggregate_dataset_sharded(weights, dataset, aggregator_dict=None, shard_by: str = "year", group=None,
                              **kwargs) -> pd.DataFrame:
    """``aggregate_dataset`` with the time axis sharded over the ranks of ``group`` (default: the
    world).  Every rank passes the same ``dataset`` description (its ``values`` may be any array-like
    that supports row slicing -- only the rank's own rows are touched) and receives the full panel,
    identical to a single call over the whole record."""
    import torch
    import torch.distributed as dist
    from . import aggregate as _agg
    from . import engine as _engine

    if aggregator_dict is None and kwargs:
        aggregator_dict, kwargs = kwargs, {}
    check_shardable(aggregator_dict, shard_by)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shards = plan_time_shards(dataset.time, world, shard_by)
    subs = [dataset.isel_time(r0, r1) if r1 > r0 else None for r0, r1 in shards]
    # every rank plans every shard's label axis on the host (cheap), so the gather needs no size exchange
    plans = [None if s is None else _agg._plan(s, aggregator_dict) for s in subs]
    labels = [None if p is None else p[1].labels for p in plans]
    g_sizes = [0 if l is None else len(l) for l in labels]
    names = next(p[0] for p in plans if p is not None)
    csr = _agg._device_csr(weights, dataset)
    R, NC = csr.host.n_regions, len(names)
    if subs[rank] is not None:
        _, res, _raster = _agg._temporal_device(subs[rank], aggregator_dict)
        local = _engine.run_spmm(csr, res)
    else:
        local = torch.empty((R, 0, NC), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    panel = gather_panels(local, g_sizes, group).cpu().numpy()
    df = _agg._assemble_panel(panel, names, concat_labels(labels), csr.host.region_ids, weights)
    rid = weights.georegions.regionid
    return weights.georegions.shp[[rid]].merge(df, left_index=True, right_on="region_id").drop(columns="region_id")
