"""End-to-end timing of aggregate_dataset fed from zarr stores (SURVEY f-1): host threads decode chunks
into pinned slots, chunks cross PCIe as stored, agf_tile_place_run builds the time-major raster.

    python tools/zarr_feed_bench.py [--grid 104x236] [--hours 8760] [--out gpurun_out/zarr_feed.json]

Writes the CONUS-sized synthetic year in three layouts under a temporary directory, runs the C1 spec from
each (and from the in-memory array, pinned) and prints one JSON object.  Not part of bench.py's contract."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="104x236")
    ap.add_argument("--hours", type=int, default=8760)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="comma list of substrings of layout names")
    a = ap.parse_args()
    import torch
    import aggfly_b200 as af
    from aggfly_b200 import stream, zarrio
    from aggfly_b200.io import _auto_chunks

    Y, X = (int(v) for v in a.grid.split("x"))
    T = a.hours
    rng = np.random.default_rng(1216)
    hours = np.arange(T)
    lat, lon = np.linspace(49.75, 24.0, Y), 235.0 + 0.25 * np.arange(X)
    arr = (27 * np.cos(np.deg2rad(lat))[None, :, None] - 6 + 9 * np.sin(2 * np.pi * (hours / 24 - 110) / 365)[:, None, None]
           + 4 * np.sin(2 * np.pi * (hours % 24 - 9) / 24)[:, None, None]).astype(np.float32)
    arr = arr + rng.normal(0, 3, (T, Y, X)).astype(np.float32)
    arr = np.round(arr * 64) / 64                                  # ~ERA5 precision after unpacking: compressible mantissas
    t = pd.date_range("2001-01-01", periods=T, freq="h")
    if a.threads:
        stream.OPTIONS["staging_threads"] = a.threads
    spec = dict(tavg=[("aggregate", {"calc": "mean", "groupby": "date"}),
                      ("transform", {"transform": "power", "exp": np.arange(1, 3)}),
                      ("aggregate", {"calc": "sum", "groupby": "year"})])
    mem = af.Dataset.from_arrays(torch.from_numpy(arr).pin_memory(), t, lat, lon, True, name="t2m")
    nr = 20
    regions = af.GeoRegions.from_rectangles([f"r{i}" for i in range(nr)], lon_min=-125 + 2.9 * np.arange(nr),
                                            lon_max=-122.1 + 2.9 * np.arange(nr), lat_min=np.full(nr, 24.0), lat_max=np.full(nr, 49.9))
    w = af.weights_from_objects(mem, regions)
    w.calculate_weights()
    tc = _auto_chunks({"time": T, "latitude": Y, "longitude": X}, 4, 256)
    layouts = {
        "time_major_24h_zstd": dict(dims=("time", "latitude", "longitude"), chunks={"time": 24}, zarr_format=3, compressor="zstd"),
        "time_major_24h_raw": dict(dims=("time", "latitude", "longitude"), chunks={"time": 24}, zarr_format=3, compressor=None),
        "reference_time_contiguous_zstd": dict(dims=("latitude", "longitude", "time"), chunks=tc, zarr_format=3, compressor="zstd"),
        "time_major_24h_blosc_lz4": dict(dims=("time", "latitude", "longitude"), chunks={"time": 24}, zarr_format=2, compressor="blosc"),
        "reference_time_contiguous_blosc_lz4": dict(dims=("latitude", "longitude", "time"), chunks=tc, zarr_format=2,
                                                    compressor="blosc"),
    }
    if a.only:
        layouts = {k: v for k, v in layouts.items() if any(w in k for w in a.only.split(","))}
    out = {"grid": [Y, X], "hours": T, "raw_gb": arr.nbytes / 1e9, "threads": stream.OPTIONS["staging_threads"],
           "host_cpus": os.cpu_count(), "layouts": {}}

    def timed(ds):
        ms = []
        for _ in range(a.reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            df = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        return df, ms[1:]

    want, ms = timed(mem)
    out["in_memory_pinned_ms"] = ms
    tmp = tempfile.mkdtemp(prefix="agf_zarr_")
    try:
        for name, lay in layouts.items():
            store = os.path.join(tmp, name + ".zarr")
            t0 = time.perf_counter()
            zarrio.write_dataset(store, arr, t, lat, lon, var="t2m", **lay)
            wsec = time.perf_counter() - t0
            size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(store) for f in fs)
            ds = af.dataset_from_path(store, var="t2m")
            for dd in ([True, False] if lay["compressor"] == "blosc" else [True]):
                stream.OPTIONS["device_decompress"] = dd
                got, ms = timed(ds)
                same = bool(np.array_equal(got[["tavg_1", "tavg_2"]].values, want[["tavg_1", "tavg_2"]].values, equal_nan=True))
                st = {k: v for k, v in stream.LAST_STATS.items() if k != "copy_events"}
                best = min(ms)
                key = name if dd else name + "_host_decode"
                out["layouts"][key] = {"chunks": list(ds.values.array.chunks), "store_gb": size / 1e9, "write_s": wsec, "ms": ms,
                                       "raw_gbs": arr.nbytes / 1e6 / best, "cell_hours_per_s": arr.size / (best / 1e3),
                                       "bitwise_equal_to_in_memory": same, "feed": st}
            stream.OPTIONS["device_decompress"] = True
            shutil.rmtree(store)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    line = json.dumps(out)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
