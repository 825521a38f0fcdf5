#!/usr/bin/env python
"""Write a small self-contained project (two synthetic ERA5-shaped years in Kelvin, a region file with
attributes, a population-like secondary raster, a config) and print the command that aggregates it:

    python examples/make_demo_project.py /tmp/aggfly_demo
    python -m aggfly_b200 validate /tmp/aggfly_demo/config.yaml
    python -m aggfly_b200 run /tmp/aggfly_demo/config.yaml -v          # needs a B200

The config is the reference's examples/era5_counties_pop.yaml with paths pointing at the demo files
(formats this package reads without xarray / geopandas: .npz rasters, .geojson regions).
"""
import json
import os
import sys

import numpy as np
import pandas as pd
import yaml


def main(out_dir: str, years=(2001, 2002)) -> str:
    os.makedirs(out_dir, exist_ok=True)
    lat = 49.75 - 0.25 * np.arange(104)                    # CONUS, descending like ERA5 files
    lon = 235.0 + 0.25 * np.arange(236)                    # 0-360 convention
    rng = np.random.default_rng(1216)
    for y in years:
        t = pd.date_range(f"{y}-01-01", f"{y}-12-31 23:00", freq="h")
        doy = np.asarray(t.dayofyear, dtype=np.float32)[:, None, None]
        hour = np.asarray(t.hour, dtype=np.float32)[:, None, None]
        base = (300.0 - 0.6 * (lat - 24.0)).astype(np.float32)[None, :, None]
        vals = base + 9 * np.sin(2 * np.pi * (doy - 110) / 365) + 4 * np.sin(2 * np.pi * (hour - 9) / 24)
        vals = (vals + rng.normal(0, 3, (len(t), len(lat), len(lon)))).astype(np.float32)       # Kelvin
        np.savez(os.path.join(out_dir, f"era5_t2m_{y}.npz"), t2m=vals, time=t.values, latitude=lat, longitude=lon)
    feats = []
    xs, ys = np.linspace(-124.5, -67.0, 13), np.linspace(25.0, 49.0, 9)
    for j in range(len(ys) - 1):
        for i in range(len(xs) - 1):
            x0, x1, y0, y1 = xs[i], xs[i + 1], ys[j], ys[j + 1]
            ring = [[x0, y0], [x1, y0], [x1, y1], [(x0 + x1) / 2, y1 + 0.3], [x0, y1], [x0, y0]]    # not just boxes
            feats.append({"type": "Feature", "properties": {"fips": f"{j:02d}{i:03d}"},
                          "geometry": {"type": "Polygon", "coordinates": [ring]}})
    json.dump({"type": "FeatureCollection", "features": feats}, open(os.path.join(out_dir, "regions.geojson"), "w"))
    plat, plon = 49.95 - 0.1 * np.arange(260), -125.05 + 0.1 * np.arange(600)
    pop = rng.lognormal(0.0, 1.5, (len(plat), len(plon)))
    np.savez(os.path.join(out_dir, "population.npz"), values=pop, latitude=plat, longitude=plon)
    cfg = {"regions": {"path": os.path.join(out_dir, "regions.geojson"), "regionid": "fips"},
           "dataset": {"path": os.path.join(out_dir, "era5_t2m_{year}.npz"), "var": "t2m",
                       "preprocess": "kelvin_to_celsius", "lon_is_360": True},
           "weights": {"project_dir": os.path.join(out_dir, "aggfly_proj"),
                       "secondary": {"type": "pop", "path": os.path.join(out_dir, "population.npz")}},
           "aggregate": {"engine": "cuda", "variables": {
               "tavg": [["aggregate", {"calc": "mean", "groupby": "date"}],
                        ["transform", {"transform": "power", "exp": [1, 2]}],
                        ["aggregate", {"calc": "sum", "groupby": "year"}]],
               "temp_bins": [["aggregate", {"calc": "mean", "groupby": "date"}],
                             ["aggregate", {"calc": "bins", "groupby": "year",
                                            "ddargs": [[0, 5, 0], [5, 10, 0], [10, 15, 0]]}]]}},
           "years": f"{years[0]}:{years[-1]}",
           "output": {"path": os.path.join(out_dir, "out", "panel.parquet")}}
    path = os.path.join(out_dir, "config.yaml")
    yaml.safe_dump(cfg, open(path, "w"), sort_keys=False)
    return path


if __name__ == "__main__":
    p = main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/aggfly_demo")
    print(f"wrote {p}\n  python -m aggfly_b200 validate {p}\n  python -m aggfly_b200 run {p} -v")
