"""One global 0.25 degree hourly year (721 x 1440 x 8760 f32 = 36.4 GB) aggregated FROM A ZARR STORE in the
reference's own time-contiguous layout (dims latitude, longitude, time; chunks [87, 87, 8760] = 265 MB,
aggfly/dataset/zarr_convert.py:31-47) compressed with Blosc-LZ4 + byte shuffle (zarr v2's default compressor):
compressed chunks cross PCIe as stored, the GPU's decompression engine inflates them, agf_unshuffle_run /
agf_tile_place_run build the time-major raster, then the C3 spec runs.  Compared bit for bit with the same
call on the device-resident raster.  Prints one JSON object.

    python tools/zarr_global_bench.py [--quantum 0.015625] [--writer-threads 16] [--out gpurun_out/x.json]

``--quantum q`` rounds the synthetic field to multiples of q (0 = keep full-entropy float32 mantissas; ERA5
t2m unpacked from int16 carries ~16 significant bits, 1/64 K is in that range).  ``--rows`` limits the
latitude rows (smaller store, same code path)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _DeviceBlocks:
    """[lat, lon, time] view of a device raster [T, lat, lon] that hands out host blocks chunk by chunk."""
    lazy_blocks = True

    def __init__(self, raster3):
        self.r = raster3
        T, Y, X = raster3.shape
        self.shape, self.dtype, self.ndim = (Y, X, T), np.dtype(np.float32), 3

    def __getitem__(self, sl):
        ys, xs, ts = sl
        return self.r[ts, ys, xs].permute(1, 2, 0).contiguous().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quantum", type=float, default=1.0 / 64)
    ap.add_argument("--writer-threads", type=int, default=16)
    ap.add_argument("--threads", type=int, default=0, help="feed threads (default stream.OPTIONS)")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    import aggfly_b200 as af
    from aggfly_b200 import stream, synthetic as syn, zarrio
    from aggfly_b200.io import _auto_chunks

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = syn.make_workload("c3_global_bins")
    if a.rows:
        g = wl.grid
        j0 = max(0, min(int(np.argmin(np.abs(g.latitude - 40.0))), len(g.latitude) - a.rows))
        band = syn.GridDef(g.latitude[j0:j0 + a.rows], g.longitude, g.lon_is_360, g.regions)
        wl = syn.Workload(wl.name + "_band", band, wl.spec_name, wl.n_time, wl.time, wl.hourly, wl.secondary, 0.0)
    if a.threads:
        stream.OPTIONS["staging_threads"] = a.threads
    T, Y, X = wl.n_time, len(wl.grid.latitude), len(wl.grid.longitude)
    raster = wl.raster(dev, seed=1218).reshape(T, Y, X)
    if a.quantum > 0:
        for r0 in range(0, T, 512):
            raster[r0:r0 + 512] = torch.round(raster[r0:r0 + 512] / a.quantum) * a.quantum
    resident = wl.dataset(raster)
    w = wl.weights(resident)
    out = {"workload": wl.name, "grid": [Y, X], "hours": T, "raw_gb": T * Y * X * 4 / 1e9, "quantum": a.quantum,
           "feed_threads": stream.OPTIONS["staging_threads"], "host_cpus": os.cpu_count()}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    want = af.aggregate_dataset(weights=w, dataset=resident, aggregator_dict=wl.spec)
    torch.cuda.synchronize()
    out["device_resident_call_ms"] = (time.perf_counter() - t0) * 1e3

    tmp = tempfile.mkdtemp(prefix="agf_zarr_", dir=a.tmp)
    try:
        store = os.path.join(tmp, "global_tc.zarr")
        chunks = _auto_chunks({"time": T, "latitude": Y, "longitude": X}, 4, 256)
        t0 = time.perf_counter()
        zarrio.write_dataset(store, np.zeros((1, Y, X), np.float32), wl.time[:1], wl.grid.latitude, wl.grid.longitude, var="t2m",
                             compressor=None)
        shutil.rmtree(os.path.join(store, "t2m"))
        shutil.rmtree(os.path.join(store, "time"))
        hours = ((wl.time.values - wl.time.values[0]) // np.timedelta64(1, "h")).astype(np.int64)
        zarrio.write_array(os.path.join(store, "time"), hours, [-1], ["time"],
                           {"units": f"hours since {wl.time[0].strftime('%Y-%m-%d %H:%M:%S')}", "calendar": "proleptic_gregorian"},
                           zarr_format=2, compressor=None)
        zarrio.write_array(os.path.join(store, "t2m"), _DeviceBlocks(raster), [chunks["latitude"], chunks["longitude"], -1],
                           ["latitude", "longitude", "time"], None, zarr_format=2, compressor="blosc", threads=a.writer_threads)
        out["write_s"] = time.perf_counter() - t0
        out["store_gb"] = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(store) for f in fs) / 1e9
        ds = af.dataset_from_path(store, var="t2m", lon_is_360=wl.grid.lon_is_360)
        out["chunks"] = list(ds.values.array.chunks)
        out["n_chunks"] = len(ds.values.tiles())
        cols = [c for c in want.columns if c not in ("geoid", "time")]
        for dd in (True,):
            stream.OPTIONS["device_decompress"] = dd
            ms = []
            for _ in range(a.reps + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=wl.spec)
                torch.cuda.synchronize()
                ms.append((time.perf_counter() - t0) * 1e3)
            st = {k: v for k, v in stream.LAST_STATS.items() if k != "copy_events"}
            ev = stream.LAST_STATS["copy_events"]
            best = min(ms[1:])
            # the raster the feed built on the device (kept between calls) against the resident one, bit for bit
            built = next(iter(stream._DEVICE_RASTERS.values())).view(T, Y * X)
            raster_bits_equal = bool(torch.equal(built.view(torch.int32), raster.reshape(T, Y * X).view(torch.int32)))
            gv, wv = got[cols].values.astype(np.float64), want[cols].values.astype(np.float64)
            fin = np.isfinite(wv)
            rel = float(np.max(np.abs(gv[fin] - wv[fin]) / np.maximum(np.abs(wv[fin]), 1e-300))) if fin.any() else 0.0
            out["blosc_lz4_device_decompress" if dd else "blosc_lz4_host_decode"] = {
                "ms": ms, "first_call_ms": ms[0], "best_ms": best, "cell_hours_per_s": T * Y * X / (best / 1e3),
                "raw_equivalent_gbs": T * Y * X * 4 / 1e6 / best, "pcie_gbs": st["h2d_bytes"] / 1e6 / ev[0].elapsed_time(ev[1]),
                "device_raster_bitwise_equal_to_resident": raster_bits_equal,
                "panel_nan_pattern_equal": bool(len(got) == len(want) and np.array_equal(np.isnan(gv), np.isnan(wv))),
                "panel_max_rel_err_vs_single_stripe_resident_call": rel,
                "panel_note": "the streamed call scans the year in time stripes whose fp64 partial sums are merged in stripe "
                              "order; the resident call uses one stripe, so the panels agree to fp64 re-association, not bitwise",
                "panel_rows": int(len(got)), "feed": st}
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
