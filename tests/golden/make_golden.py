"""Generate golden fixtures from the REFERENCE's own compiled arithmetic.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What it does
------------
aggfly itself cannot be imported here (no xarray/dask/geopandas), but the two halves of the
hot path are plain numba/numpy functions:

* ``aggfly/aggregate/nb_kernels.py`` -- ``resample_groups`` and the four ``_block_*`` numba
  kernels (reference lines 80-115, 121-251).  Loaded with ``importlib`` after putting empty
  stand-in modules named ``xarray``/``dask``/``dask.array`` into ``sys.modules``.
* ``aggfly/aggregate/spatial.py`` -- ``_weight_triplets`` and ``_scatter_block`` (reference lines
  157-186), pulled out of the module source and exec'd with numpy/pandas only.

Nothing from the reference is copied into this repo: only the *outputs* of those functions on
seeded inputs are stored (``tests/golden/ref_kernels.npz``), next to this script.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pandas as pd

REF = os.environ.get("AGGFLY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_kernels():
    xr = types.ModuleType("xarray")
    xr.CFTimeIndex = type("CFTimeIndex", (), {})
    xr.DataArray = type("DataArray", (), {})
    dask = types.ModuleType("dask")
    darr = types.ModuleType("dask.array")
    dask.array = darr
    saved = {k: sys.modules.get(k) for k in ("xarray", "dask", "dask.array")}
    sys.modules.update({"xarray": xr, "dask": dask, "dask.array": darr})
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_nb_kernels", os.path.join(REF, "aggfly/aggregate/nb_kernels.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_reference_spatial():
    src = open(os.path.join(REF, "aggfly/aggregate/spatial.py")).read()
    a = src.index("def _weight_triplets")
    b = src.index("def _scatter(")
    ns = {"np": np, "pd": pd}
    exec(compile(src[a:b], "ref_spatial_slice", "exec"), ns)
    return ns["_weight_triplets"], ns["_scatter_block"]


def make_cube(rng, T, Y, X, dtype, nan_frac):
    cube = rng.normal(15.0, 12.0, (T, Y, X)).astype(dtype)
    if nan_frac > 0:
        cube[rng.random((T, Y, X)) < nan_frac] = np.nan
        cube[:, 0, 0] = np.nan                      # an all-NaN ("ocean") cell
    # values sitting exactly on thresholds exercise the strict comparisons
    cube[1, -1, -1] = 10.0
    cube[2, -1, -1] = 30.0
    return cube


def main():
    nb = load_reference_kernels()
    weight_triplets, scatter_block = load_reference_spatial()
    out = {}

    # ---- group bounds (reference nb_kernels.py:80-115, datetime64 branch) -------------
    t_full = pd.date_range("2001-12-30 05:00", periods=24 * 40, freq="h")
    t_gap = t_full.delete(slice(24 * 3 + 7, 24 * 6 + 2))           # interior gap -> empty 1D bins
    t_daily = pd.date_range("1999-11-15", periods=800, freq="D")
    for tag, t in (("full", t_full), ("gap", t_gap), ("daily", t_daily)):
        out[f"time_{tag}"] = t.values.astype("datetime64[ns]").astype(np.int64)
        for freq in ("1D", "ME", "YE", "W"):
            b, lab = nb.resample_groups(t, freq)
            out[f"bounds_{tag}_{freq}"] = b
            out[f"labels_{tag}_{freq}"] = lab.values.astype("datetime64[ns]").astype(np.int64)
            # second level: the reference groups the labels of the previous step the same way
            if freq == "1D":
                for f2 in ("ME", "YE", "W"):
                    b2, lab2 = nb.resample_groups(lab, f2)
                    out[f"bounds2_{tag}_{f2}"] = b2
                    out[f"labels2_{tag}_{f2}"] = lab2.values.astype("datetime64[ns]").astype(np.int64)

    # ---- kernels (reference nb_kernels.py:121-251) -------------------------------------
    ddargs = np.array([[10.0, 30.0, 0.0], [0.1, 17.3, 1.0], [-99.0, 20.0, 0.0],
                       [20.0, 99.0, 0.0], [28.0, 29.0, 1.0]], dtype=np.float64)
    out["ddargs"] = ddargs
    T, Y, X = 24 * 9 + 5, 3, 5
    # bounds with a partial first group, an EMPTY interior group and a ragged tail
    bounds = np.array([0, 7, 31, 31, 55, 79, 103, 150, 221], dtype=np.int64)
    assert bounds[-1] == T
    out["bounds"] = bounds
    for dt in ("float32", "float64"):
        for nan_tag, nan_frac in (("clean", 0.0), ("nan", 0.02)):
            rng = np.random.default_rng(1216 + (dt == "float64") * 7 + (nan_frac > 0))
            cube = make_cube(rng, T, Y, X, dt, nan_frac)
            key = f"{dt}_{nan_tag}"
            out[f"cube_{key}"] = cube
            G = len(bounds) - 1
            for name, code in nb._STAT_CODE.items():
                o = np.empty((G, Y, X), cube.dtype)
                nb._block_stat(cube, bounds, code, o)
                out[f"stat_{name}_{key}"] = o
            for name, fn in (("dd", nb._block_dd), ("bins", nb._block_bins),
                             ("sine_dd", nb._block_sine_dd)):
                o = np.empty((G, Y, X, ddargs.shape[0]), cube.dtype)
                fn(cube, bounds, ddargs, o)
                out[f"{name}_{key}"] = o

    # ---- spatial (reference spatial.py:157-186) ----------------------------------------
    rng = np.random.default_rng(99)
    n_cells, n_t, nnz = 40, 6, 90
    wdf = pd.DataFrame({
        "cell_id": rng.integers(0, n_cells + 6, nnz),          # some cells absent from the grid
        "index_right": rng.choice([3, 5, 6, 11, 20], nnz),      # non-contiguous region ids
        "weight": rng.random(nnz),
    })
    block = rng.normal(20, 5, (n_cells, n_t))
    region_idx, cell_idx, w_vals, region_ids = weight_triplets(wdf, np.arange(n_cells))
    out["sp_cell_id"] = wdf["cell_id"].to_numpy()
    out["sp_index_right"] = wdf["index_right"].to_numpy()
    out["sp_weight"] = wdf["weight"].to_numpy()
    out["sp_block"] = block
    out["sp_region_idx"] = region_idx.astype(np.int64)
    out["sp_cell_idx"] = cell_idx.astype(np.int64)
    out["sp_w_vals"] = w_vals
    out["sp_region_ids"] = region_ids.astype(np.int64)
    out["sp_scatter"] = scatter_block(block, region_idx, cell_idx, w_vals, len(region_ids))

    path = os.path.join(HERE, "ref_kernels.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
