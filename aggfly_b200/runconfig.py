"""``aggfly run`` configuration: the reference's YAML schema (aggfly/cli/config.py:52-113, 214-386),
validated the same way -- every problem is collected and reported at once -- with one more accepted
``aggregate.engine`` value, ``"cuda"`` (which ``"auto"`` resolves to in this package).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import yaml

ALLOWED_CALCS = {"mean", "nanmean", "sum", "min", "max", "dd", "bins", "sine_dd"}
DD_CALCS = {"dd", "bins", "sine_dd"}
ALLOWED_GROUPBY = {"date", "month", "year", "week"}
ALLOWED_ENGINE = {"auto", "cuda", "dask", "numba"}
ALLOWED_BACKEND = {"threads", "processes", "none"}
ALLOWED_FORMAT = {"parquet", "feather", "csv"}
ALLOWED_SECONDARY = {"pop", "crop", "generic"}
ALLOWED_ZERO_WEIGHT = {"nan", "area", "drop"}
STEP_TYPES = {"aggregate", "transform"}


class ConfigError(Exception):
    """All the problems of a config, one per line."""

    def __init__(self, errors):
        self.errors = list(errors)
        super().__init__("\n".join(f"- {e}" for e in self.errors))


@dataclass
class SecondaryWeightsConfig:
    type: str
    path: str
    crop: Optional[str] = None
    feed: Optional[str] = None


@dataclass
class RunConfig:
    regions_path: str
    regionid: str
    region_list: Optional[List[str]]
    dataset_path: str
    var: str
    preprocess: Optional[str]
    preprocess_from: Optional[str]
    lon_is_360: bool
    timecoord: str
    xycoords: Tuple[str, str]
    time_sel: Optional[str]
    clip_to_regions: bool
    project_dir: Optional[str]
    secondary: Optional[SecondaryWeightsConfig]
    zero_weight: str
    engine: str
    variables: Dict[str, List]
    years: Optional[List[int]]
    backend: str
    output_path: str
    output_format: str
    extra: Dict[str, object] = field(default_factory=dict)   # reader options the CUDA path does not need

    @property
    def templated(self) -> bool:
        return "{year}" in self.dataset_path

    def resolved_paths(self) -> List[str]:
        if not self.templated:
            return [self.dataset_path]
        return [self.dataset_path.format(year=y) for y in (self.years or [])]

    def to_aggregator_dict(self) -> Dict[str, List]:
        """``variables`` as ``aggregate_dataset`` wants them: steps as tuples, every ``exp`` a NumPy array
        (so that the library's ``exp[0]`` indexing sees the list of exponents; aggfly/cli/config.py:98-113)."""
        out = {}
        for name, steps in self.variables.items():
            fixed = []
            for kind, params in steps:
                params = dict(params)
                if kind == "transform" and "exp" in params:
                    params["exp"] = np.array(params["exp"])
                fixed.append((kind, params))
            out[name] = fixed
        return out


def parse_years(spec, errors: List[str]) -> Optional[List[int]]:
    """``"1980:1990"`` (inclusive) | list | int | None."""
    if spec is None:
        return None
    if isinstance(spec, bool):
        errors.append("years: must be a range 'start:end', a list, or an int")
    elif isinstance(spec, int):
        return [spec]
    elif isinstance(spec, list):
        try:
            return [int(y) for y in spec]
        except (TypeError, ValueError):
            errors.append(f"years: list must contain integers, got {spec!r}")
    elif isinstance(spec, str):
        try:
            if ":" in spec:
                lo, hi = spec.split(":")
                return list(range(int(lo), int(hi) + 1))
            return [int(spec)]
        except ValueError:
            errors.append(f"years: could not parse {spec!r} (use 'start:end' or an int)")
    else:
        errors.append(f"years: unsupported type {type(spec).__name__}")
    return None


def _check_steps(name: str, steps, errors: List[str]) -> None:
    if not isinstance(steps, list) or not steps:
        errors.append(f"aggregate.variables.{name}: must be a non-empty list of steps")
        return
    n_out = 1
    for i, step in enumerate(steps):
        loc = f"aggregate.variables.{name}[{i}]"
        if not (isinstance(step, (list, tuple)) and len(step) == 2):
            errors.append(f"{loc}: each step must be [step_type, params]")
            continue
        kind, params = step
        if kind not in STEP_TYPES:
            errors.append(f"{loc}: unknown step type {kind!r} (expected one of {sorted(STEP_TYPES)})")
            continue
        if not isinstance(params, dict):
            errors.append(f"{loc}: params must be a mapping")
            continue
        if kind == "aggregate":
            calc, groupby = params.get("calc"), params.get("groupby")
            if calc not in ALLOWED_CALCS:
                errors.append(f"{loc}: calc {calc!r} not in {sorted(ALLOWED_CALCS)}")
            if groupby not in ALLOWED_GROUPBY:
                errors.append(f"{loc}: groupby {groupby!r} not in {sorted(ALLOWED_GROUPBY)}")
            if calc in DD_CALCS:
                dd = params.get("ddargs")
                if not isinstance(dd, list) or not dd:
                    errors.append(f"{loc}: calc {calc!r} requires a non-empty 'ddargs' list")
                elif isinstance(dd[0], list) and n_out > 1:
                    errors.append(f"aggregate.variables.{name}: cannot combine a multi-'ddargs' (bins) step with a "
                                  "multi-output transform (e.g. multiple exponents) — the library rejects this at runtime")
        else:
            is_spline = params.get("transform") == "spline" or "spline" in params
            if not ("exp" in params or "inter" in params or is_spline):
                errors.append(f"{loc}: transform step needs one of 'exp' (power), 'inter', or transform: spline")
            if "exp" in params:
                if not isinstance(params["exp"], (list, int)):
                    errors.append(f"{loc}: 'exp' must be an int or a list of ints")
                else:
                    n_out = len(params["exp"]) if isinstance(params["exp"], list) else 1
            if is_spline:
                n_out = 2


def parse_config(raw) -> RunConfig:
    if not isinstance(raw, dict) or not raw:
        raise ConfigError(["config must be a non-empty YAML mapping"])
    errors: List[str] = []

    def section(key):
        val = raw.get(key)
        if val is None:
            return {}
        if not isinstance(val, dict):
            errors.append(f"{key}: must be a mapping")
            return {}
        return val

    regions, dataset, weights = section("regions"), section("dataset"), section("weights")
    aggregate, execution, output = section("aggregate"), section("execution"), section("output")

    for key, sec, label in (("path", regions, "regions.path"), ("regionid", regions, "regions.regionid"),
                            ("path", dataset, "dataset.path"), ("var", dataset, "dataset.var")):
        if not sec.get(key):
            errors.append(f"{label} is required")
    preprocess, preprocess_from = dataset.get("preprocess"), dataset.get("preprocess_from")
    if preprocess is not None and preprocess_from is not None:
        errors.append("dataset: set at most one of 'preprocess' and 'preprocess_from'")
    if preprocess_from is not None and ":" not in str(preprocess_from):
        errors.append("dataset.preprocess_from must be 'path/to/file.py:function'")
    xycoords = dataset.get("xycoords", ["longitude", "latitude"])
    if not (isinstance(xycoords, list) and len(xycoords) == 2):
        errors.append("dataset.xycoords must be a 2-item list [lon_name, lat_name]")
        xycoords = ["longitude", "latitude"]
    for key in ("storage_options",):
        if dataset.get(key) is not None and not isinstance(dataset[key], dict):
            errors.append(f"dataset.{key} must be a mapping")
    if dataset.get("engine") is not None and not isinstance(dataset["engine"], str):
        errors.append("dataset.engine must be a string (e.g. 'zarr')")

    zero_weight = weights.get("zero_weight", "nan")
    if zero_weight not in ALLOWED_ZERO_WEIGHT:
        errors.append(f"weights.zero_weight {zero_weight!r} not in {sorted(ALLOWED_ZERO_WEIGHT)}")
        zero_weight = "nan"
    secondary = None
    sec_raw = weights.get("secondary")
    if sec_raw is not None:
        if not isinstance(sec_raw, dict):
            errors.append("weights.secondary must be a mapping")
        else:
            if sec_raw.get("type") not in ALLOWED_SECONDARY:
                errors.append(f"weights.secondary.type {sec_raw.get('type')!r} not in {sorted(ALLOWED_SECONDARY)}")
            if not sec_raw.get("path"):
                errors.append("weights.secondary.path is required")
            secondary = SecondaryWeightsConfig(sec_raw.get("type"), sec_raw.get("path"), sec_raw.get("crop"), sec_raw.get("feed"))

    engine = aggregate.get("engine", "auto")
    if engine not in ALLOWED_ENGINE:
        errors.append(f"aggregate.engine {engine!r} not in {sorted(ALLOWED_ENGINE)}")
    variables = aggregate.get("variables")
    if not isinstance(variables, dict) or not variables:
        errors.append("aggregate.variables must be a non-empty mapping of name -> steps")
        variables = {}
    for name, steps in variables.items():
        _check_steps(name, steps, errors)

    years = parse_years(raw.get("years"), errors)
    backend = execution.get("backend", "threads")
    if backend not in ALLOWED_BACKEND:
        errors.append(f"execution.backend {backend!r} not in {sorted(ALLOWED_BACKEND)}")

    output_path, output_format = output.get("path"), output.get("format")
    if not output_path:
        errors.append("output.path is required")
    if output_format is None and output_path:
        ext = os.path.splitext(str(output_path))[1].lstrip(".").lower()
        output_format = {"pq": "parquet"}.get(ext, ext)
    if output_format not in ALLOWED_FORMAT:
        errors.append(f"output.format {output_format!r} not in {sorted(ALLOWED_FORMAT)} "
                      "(set output.format or use a .parquet/.feather/.csv extension)")
    if dataset.get("path") and "{year}" in str(dataset["path"]) and not years:
        errors.append("dataset.path contains '{year}' but no 'years' were given (add years: 'start:end')")
    if errors:
        raise ConfigError(errors)
    return RunConfig(
        regions_path=regions["path"], regionid=regions["regionid"], region_list=regions.get("region_list"),
        dataset_path=dataset["path"], var=dataset["var"], preprocess=preprocess, preprocess_from=preprocess_from,
        lon_is_360=bool(dataset.get("lon_is_360", True)), timecoord=dataset.get("timecoord", "time"),
        xycoords=(xycoords[0], xycoords[1]), time_sel=dataset.get("time_sel"),
        clip_to_regions=bool(dataset.get("clip_to_regions", True)), project_dir=weights.get("project_dir"),
        secondary=secondary, zero_weight=zero_weight, engine=engine, variables=variables, years=years, backend=backend,
        output_path=output_path, output_format=output_format,
        extra={k: dataset.get(k) for k in ("chunks", "storage_options", "engine") if dataset.get(k) is not None})


def load_config(path: str) -> RunConfig:
    try:
        with open(path) as f:
            raw = yaml.safe_load(f)
    except FileNotFoundError:
        raise ConfigError([f"config file not found: {path}"])
    except yaml.YAMLError as e:
        raise ConfigError([f"could not parse YAML: {e}"])
    return parse_config(raw)
