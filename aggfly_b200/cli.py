"""``aggfly run config.yaml`` on the CUDA engine (``python -m aggfly_b200 run ...``).

Same commands and flags as the reference's click group for the parts that touch the hot path
(aggfly/cli/main.py:16-255: ``validate``, ``weights``, ``run``) and the same orchestration order as
``run_pipeline`` (aggfly/cli/pipeline.py:124-172): regions -> sample layer -> weights (cached) ->
``aggregate_dataset`` for every ``{year}`` path -> concat -> write.  What changes underneath:

* the yearly loop is time-sharded over the ranks of a ``torch.distributed`` job when there is one
  (``torchrun --nproc-per-node 8 -m aggfly_b200 run ...``): rank r aggregates years r, r+W, ... and
  rank 0 concatenates the gathered frames -- the reference runs the years serially;
* ``dataset.preprocess`` names / expressions are fused into the temporal kernel;
* ``execution.backend`` / ``n_workers`` are accepted and ignored (there is no dask here).
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Callable, List, Optional

import click
import pandas as pd

from . import io as _io
from . import preprocess as _pp
from . import runconfig as _cfg


def _log_fn(verbose: bool) -> Callable[[str], None]:
    return (lambda m: click.echo(m, err=True)) if verbose else (lambda m: None)


def resolve_preprocess(config: _cfg.RunConfig):
    """Builtin name / expression -> fused chain; ``preprocess_from: file.py:func`` -> that callable
    (trusted user code, applied to the array on the host, as in aggfly/cli/preprocess.py:116-175)."""
    if config.preprocess_from:
        path, _, func = str(config.preprocess_from).rpartition(":")
        if not os.path.exists(path):
            raise _pp.PreprocessError(f"preprocess_from: file not found: {path}")
        spec = importlib.util.spec_from_file_location("_aggfly_user_preprocess", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if not hasattr(mod, func):
            raise _pp.PreprocessError(f"preprocess_from: {path} has no function {func!r}")
        return getattr(mod, func)
    if config.preprocess is None:
        return None
    _pp.resolve(str(config.preprocess))                      # validate now, fail before any data is read
    return str(config.preprocess)


def load_dataset(config: _cfg.RunConfig, path: str, georegions=None):
    ds = _io.dataset_from_path(path, var=config.var, xycoords=config.xycoords, timecoord=config.timecoord,
                               time_sel=config.time_sel, lon_is_360=config.lon_is_360,
                               preprocess=resolve_preprocess(config), name=config.var)
    if georegions is not None and config.clip_to_regions:
        box = region_extent(georegions)
        if box is not None:
            ds = _io.clip_to_extent(ds, *box)
    return ds


def region_extent(georegions):
    import numpy as np
    shp = georegions.shp
    if "rings" in shp.columns:
        pts = [r for rings in shp["rings"] for r in rings if len(r)]
        if not pts:
            return None
        allp = np.concatenate(pts)
        return float(allp[:, 0].min()), float(allp[:, 0].max()), float(allp[:, 1].min()), float(allp[:, 1].max())
    if {"lon_min", "lon_max", "lat_min", "lat_max"}.issubset(shp.columns):
        return float(shp.lon_min.min()), float(shp.lon_max.max()), float(shp.lat_min.min()), float(shp.lat_max.max())
    return None


def compute_weights(config: _cfg.RunConfig, log=lambda m: None):
    from .weights import weights_from_objects
    log(f"Loading regions: {config.regions_path}")
    georegions = _io.georegions_from_path(config.regions_path, config.regionid, config.region_list)
    path0 = config.resolved_paths()[0]
    log(f"Building weights from sample layer: {path0}")
    sample = load_dataset(config, path0, georegions)
    secondary = None
    if config.secondary is not None:                       # aggfly/cli/pipeline.py:60-73
        sc = config.secondary
        if sc.type == "pop":
            secondary = _io.pop_weights_from_path(sc.path)
        elif sc.type == "crop":
            secondary = _io.crop_weights_from_path(sc.path, crop=sc.crop or "corn", feed=sc.feed)
        else:
            secondary = _io.secondary_weights_from_path(sc.path)
    w = weights_from_objects(sample, georegions, secondary_weights=secondary, project_dir=config.project_dir,
                             zero_weight=config.zero_weight)
    w.calculate_weights()
    return w, georegions, sample


def run_pipeline(config: _cfg.RunConfig, log=lambda m: None, as_table: bool = False):
    """The panel (on rank 0; None on the other ranks of a distributed job): a DataFrame, or with ``as_table`` a
    ``pyarrow.Table`` whose yearly parts were built straight from the device panels' columns (what ``aggfly run`` writes)."""
    from .aggregate import aggregate_dataset, aggregate_dataset_table
    if as_table:
        aggregate_dataset = aggregate_dataset_table                      # noqa: F811 -- same signature, Arrow result
    rank, world = _dist_info()
    weights, georegions, sample = compute_weights(config, log)
    paths = config.resolved_paths()
    spec = config.to_aggregator_dict()
    frames: List[pd.DataFrame] = []
    for i, path in enumerate(paths):
        if i % world != rank:
            continue
        log(f"[rank {rank}] Aggregating [{i + 1}/{len(paths)}]: {path}")
        ds = sample if i == 0 else load_dataset(config, path, georegions)
        frames.append((i, aggregate_dataset(dataset=ds, weights=weights, aggregator_dict=spec, engine=config.engine)))
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, frames)               # small frames; the rasters never leave their rank
        frames = [f for part in gathered for f in part]
        if rank != 0:
            return None
    frames = [f for _, f in sorted(frames, key=lambda t: t[0])]
    if as_table:
        import pyarrow as pa
        return pa.concat_tables(frames) if len(frames) > 1 else frames[0]
    return pd.concat(frames, ignore_index=True) if len(frames) > 1 else frames[0]


def _dist_info():
    try:
        import torch
        import torch.distributed as dist
    except Exception:
        return 0, 1
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        if not dist.is_initialized():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            from .stream import bind_host_to_device
            bind_host_to_device(local)                          # staging buffers on the GPU's NUMA node
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------------------------
# click group
# ---------------------------------------------------------------------------------------------
@click.group()
def cli():
    """aggfly on the B200 CUDA engine."""


def _load(config_path: str) -> _cfg.RunConfig:
    try:
        return _cfg.load_config(config_path)
    except _cfg.ConfigError as e:
        click.echo("Config is invalid:", err=True)
        for msg in e.errors:
            click.echo(f"  - {msg}", err=True)
        raise SystemExit(1)


def check_paths(cfg: _cfg.RunConfig) -> List[str]:
    """Local input paths that do not exist (aggfly/cli/main.py:118-127 reports them as warnings); URLs are skipped."""
    missing = []
    paths = [("regions.path", cfg.regions_path)] + [("dataset.path", p) for p in cfg.resolved_paths()]
    if cfg.secondary is not None:
        paths.append(("weights.secondary.path", cfg.secondary.path))
    for label, p in paths:
        if p and "://" not in str(p) and not os.path.exists(str(p)):
            missing.append(f"{label}: {p} does not exist")
    return missing


@cli.command()
@click.argument("config", type=click.Path())
@click.option("--strict", is_flag=True, help="Treat unresolved input paths as errors (exit nonzero), not warnings.")
def validate(config, strict):
    """Check a config file (schema, step lists, preprocess) without reading any data; local input paths are
    checked for existence (warnings, or errors with --strict)."""
    cfg = _load(config)
    try:
        resolve_preprocess(cfg)
    except _pp.PreprocessError as e:
        raise click.ClickException(f"preprocess: {e}")
    n = len(cfg.resolved_paths())
    problems = check_paths(cfg)
    if problems:
        click.echo("Errors:" if strict else "Warnings:", err=strict)
        for msg in problems:
            click.echo(f"  - {msg}", err=strict)
        if strict:
            raise SystemExit(1)
    click.echo(f"Config OK: {len(cfg.variables)} variable(s), {n} dataset path(s), engine={cfg.engine}, "
               f"output={cfg.output_path} ({cfg.output_format}).")


@cli.command()
@click.argument("config", type=click.Path())
@click.option("--project-dir", default=None, help="Override weights.project_dir (weight cache).")
@click.option("-v", "--verbose", is_flag=True)
def weights(config, project_dir, verbose):
    """Build the weights (and the lowered CSR cache) for a config."""
    cfg = _load(config)
    if project_dir is not None:
        cfg.project_dir = project_dir
    w, _, sample = compute_weights(cfg, _log_fn(verbose))
    from .weights import lower_to_csr_cached
    csr = lower_to_csr_cached(w.weights, w.grid.cell_id, len(sample.latitude), len(sample.longitude),
                              sample.lon_sort_order() if sample.lon_is_360 else None, cfg.project_dir)
    click.echo(f"Weights: {len(w.weights)} cell-region pairs, {csr.n_regions} regions, {csr.nnz} CSR entries.")
    click.echo(f"Cached under: {cfg.project_dir}" if cfg.project_dir else
               "No weights.project_dir set — the lowered CSR was computed but not cached.")


@cli.command()
@click.argument("config", type=click.Path())
@click.option("-o", "--output", default=None, help="Override output.path from the config.")
@click.option("--engine", type=click.Choice(sorted(_cfg.ALLOWED_ENGINE)), default=None,
              help="Override the temporal engine (every value runs the CUDA engine here).")
@click.option("--years", default=None, help="Override years for a {year}-templated dataset path (e.g. 1980:1990).")
@click.option("--project-dir", default=None, help="Override weights.project_dir (weight cache).")
@click.option("--backend", type=click.Choice(sorted(_cfg.ALLOWED_BACKEND)), default=None,
              help="Accepted for compatibility; there is no dask backend on this engine.")
@click.option("--n-workers", type=int, default=None, help="Accepted for compatibility; ignored.")
@click.option("-v", "--verbose", is_flag=True, help="Print per-step progress.")
def run(config, output, engine, years, project_dir, backend, n_workers, verbose):
    """Run the full aggregation pipeline from a config file."""
    cfg = _load(config)
    if output is not None:
        cfg.output_path = output
        ext = output.rsplit(".", 1)[-1].lower() if "." in output else ""
        cfg.output_format = {"pq": "parquet"}.get(ext, ext) or cfg.output_format
    if engine is not None:
        cfg.engine = engine
    if project_dir is not None:
        cfg.project_dir = project_dir
    if backend is not None:
        cfg.backend = backend
    if years is not None:
        errs: List[str] = []
        cfg.years = _cfg.parse_years(years, errs)
        if errs:
            raise click.ClickException("; ".join(errs))
    try:
        resolve_preprocess(cfg)
    except _pp.PreprocessError as e:
        raise click.ClickException(f"preprocess: {e}")
    try:
        table = run_pipeline(cfg, log=_log_fn(verbose), as_table=True)
    except Exception as e:
        if verbose:
            raise
        raise click.ClickException(f"{type(e).__name__}: {e}")
    if table is not None:
        _io.write_table(table, cfg.output_path, cfg.output_format)        # Arrow columns -> file, no pandas frame
        click.echo(f"Wrote {table.num_rows} rows to {cfg.output_path} ({cfg.output_format}).")


# ---------------------------------------------------------------------------------------------
# info: what a config author needs to know about a raster (aggfly/cli/info.py), metadata only
# ---------------------------------------------------------------------------------------------
_COORD_ALIASES = {"lon": ("longitude", "lon", "x", "nav_lon"), "lat": ("latitude", "lat", "y", "nav_lat"),
                  "time": ("time", "valid_time", "t")}


def _alias(names, kind):
    lowered = {str(n).lower(): n for n in names}
    return next((lowered[c] for c in _COORD_ALIASES[kind] if c in lowered), None)


def describe_raster(path: str) -> dict:
    """Variables (dims, shape, chunks, units) and coordinate facts of a raster file, reading coordinate
    arrays only: zarr directory stores, .npz, NetCDF-3."""
    import numpy as np
    from . import zarrio
    out = {"variables": {}, "coords": {}}
    if zarrio.looks_like_zarr(path):
        g = zarrio.ZarrGroup(path)
        arrays = {}
        for n in g.names():
            try:
                arrays[n] = g[n]
            except NotImplementedError:                      # auxiliary arrays of types the reader does not decode (strings ...)
                continue
        dims_all = {d for a in arrays.values() for d in (a.dims or ())}
        for n, a in arrays.items():
            if n in dims_all and a.ndim == 1:
                out["coords"][n] = a
            else:
                out["variables"][n] = {"dims": a.dims or tuple(f"dim_{i}" for i in range(a.ndim)), "shape": a.shape,
                                       "chunks": a.chunks, "units": a.attrs.get("units")}
        tname = _alias(out["coords"], "time")
        if tname is not None:
            t = zarrio.decode_time(out["coords"][tname])
            out["time"] = (tname, str(out["coords"][tname].attrs.get("calendar", "standard")).lower(), t)
        out["coords"] = {n: (None if n == tname else np.asarray(a.read(), dtype=float)) for n, a in out["coords"].items()}
        return out
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        z = np.load(path, allow_pickle=False)
        one_d = {n: z[n] for n in z.files if z[n].ndim == 1}
        for n in z.files:
            if z[n].ndim > 1:
                out["variables"][n] = {"dims": ("time", "latitude", "longitude")[-z[n].ndim:], "shape": z[n].shape, "chunks": None,
                                       "units": None}
        tname = _alias(one_d, "time")
        if tname is not None:
            tv = one_d.pop(tname)
            t = pd.DatetimeIndex(tv.astype("datetime64[ns]") if tv.dtype.kind == "M" else pd.to_datetime(tv.astype(str)))
            out["time"] = (tname, "standard", t)
            out["coords"][tname] = None
        out["coords"].update({n: np.asarray(v, dtype=float) for n, v in one_d.items()})
        return out
    if ext in (".nc", ".nc3", ".cdf") and _io._is_netcdf3(path):
        from scipy.io import netcdf_file
        with netcdf_file(path, "r", mmap=False, maskandscale=False) as f:
            dec = lambda v: v.decode() if isinstance(v, bytes) else v                    # noqa: E731
            for n, v in f.variables.items():
                if len(v.dimensions) == 1 and v.dimensions[0] == n:
                    out["coords"][n] = np.array(v.data, dtype=float)
                else:
                    out["variables"][n] = {"dims": tuple(v.dimensions), "shape": tuple(v.shape), "chunks": None,
                                           "units": dec(getattr(v, "units", None))}
            tname = _alias(out["coords"], "time")
            if tname is not None:
                tv = f.variables[tname]
                cal = str(dec(getattr(tv, "calendar", "standard"))).lower()
                t = _io._cf_time(out["coords"][tname], dec(tv.units), cal) if cal in ("standard", "gregorian", "proleptic_gregorian") else None
                out["time"] = (tname, cal, t)
                out["coords"][tname] = None
        return out
    raise click.ClickException(f"Could not open {path!r}: natively readable are zarr directory stores, .npz and NetCDF-3 files")


@cli.command()
@click.argument("path", type=str)
@click.option("--var", default=None, help="Report only this data variable.")
@click.option("--storage-options", default=None, help="JSON dict for a remote storage backend (validated; local paths only here).")
def info(path, var, storage_options):
    """Inspect a raster dataset (dims, calendar, lon convention, time span): the values a config needs
    for ``xycoords``, ``timecoord``, ``lon_is_360`` and ``preprocess``."""
    import json
    import numpy as np
    from .timeaxis import CalendarIndex
    if storage_options is not None:
        try:
            json.loads(storage_options)
        except json.JSONDecodeError as e:
            raise click.ClickException(f"--storage-options is not valid JSON: {e}")
    try:
        d = describe_raster(path)
    except click.ClickException:
        raise
    except Exception as e:
        raise click.ClickException(f"Could not open {path!r}: {e}")
    names = list(d["variables"])
    if var is not None and var not in names:
        raise click.ClickException(f"Variable {var!r} not found. Available: {', '.join(names) or '(none)'}")
    click.echo(f"Dataset: {path}")
    click.echo(f"  data variables : {', '.join(names) or '(none)'}")
    for name in ([var] if var else names):
        v = d["variables"][name]
        click.echo(f"  {name}:")
        click.echo("    dims   : " + ", ".join(f"{dn}={n}" for dn, n in zip(v["dims"], v["shape"])))
        if v["chunks"] is not None:
            click.echo("    chunks : " + ", ".join(f"{dn}={c}" for dn, c in zip(v["dims"], v["chunks"])))
        if v["units"]:
            click.echo(f"    units  : {v['units']}")
    lon_name, lat_name = _alias(d["coords"], "lon"), _alias(d["coords"], "lat")
    click.echo("  config hints:")
    if lon_name and lat_name:
        click.echo(f"    xycoords   : [{lon_name}, {lat_name}]")
    if lon_name:
        lo, hi = float(np.nanmin(d["coords"][lon_name])), float(np.nanmax(d["coords"][lon_name]))
        click.echo(f"    lon range  : {lo:.4g} .. {hi:.4g}  \u2192 lon_is_360: {str(hi > 180.0).lower()}")
    if "time" in d:
        tname, calendar, t = d["time"]
        nonstd = isinstance(t, CalendarIndex) or calendar not in ("standard", "gregorian", "proleptic_gregorian")
        click.echo(f"    timecoord  : {tname}")
        click.echo(f"    calendar   : {calendar}{'  (cftime / non-standard)' if nonstd else ''}")
        if t is not None:
            click.echo(f"    time steps : {len(t)}")
            if len(t):
                first, last = (t.to_objects()[[0, -1]] if isinstance(t, CalendarIndex) else (t[0], t[-1]))
                click.echo(f"    time span  : {first} .. {last}")


# ---------------------------------------------------------------------------------------------
# regions: what a config author needs to know about a regions file (aggfly/regions/georegions.py:326-428)
# ---------------------------------------------------------------------------------------------
def describe_regions(path: str, rows: int = 5, uniqueness: bool = False) -> dict:
    return _io.describe_regions(path, rows, uniqueness)


@cli.command()
@click.argument("path", type=click.Path())
@click.option("-n", "--rows", default=5, show_default=True, help="Rows to preview. 0 skips the preview.")
@click.option("--uniqueness/--no-uniqueness", default=False, help="Also report which columns are unique across every feature.")
@click.option("-v", "--verbose", is_flag=True, help="Show the full traceback on error.")
def regions(path, rows, uniqueness, verbose):
    """Inspect a regions file to work out its region id column (fields, feature count, bounds, a few rows)."""
    try:
        d = describe_regions(path, rows, uniqueness)
    except Exception as e:
        if verbose:
            raise
        raise click.ClickException(f"{type(e).__name__}: {e}")
    _io.print_regions_info(d, echo=click.echo)


def main(argv=None):
    cli.main(args=argv, prog_name="aggfly")


if __name__ == "__main__":
    main(sys.argv[1:])
