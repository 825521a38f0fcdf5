"""Chunked (zarr) rasters on the GPU: the placement kernel ``agf_tile_place_run`` against NumPy for every
axis order / dtype / offset, and ``aggregate_dataset`` from a store against the same call on the in-memory
array (bit-exact: the device raster must be the same bytes) and against the CPU oracle."""
import itertools

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

import aggfly_b200 as af
from aggfly_b200 import _lib, engine, stream, zarrio
from oracle import oracle as orc
from tests.test_gpu_parity import SPECS, _close, _exact, _raster, _weights_case


@pytest.fixture(autouse=True)
def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    engine.OPTIONS["target_stripes"] = 0
    yield
    engine.OPTIONS["target_stripes"] = 0


_CODES = {"float32": _lib.F32, "float64": _lib.F64, "int16": _lib.I16, "int32": _lib.I32, "uint8": _lib.U8, "int8": _lib.I8,
          "uint16": _lib.U16}


def _place(chunk, perm, ext, off, dst_shape, dst_off, dst_dtype, packed=False, scale=1.0, offset=0.0, fill=None):
    """chunk: stored C-contiguous array whose axes are perm of (t, y, x).  Returns the device raster."""
    import torch
    T, Y, X = dst_shape
    strides = [0, 0, 0]
    s = 1
    for k in reversed(range(3)):
        strides[perm[k]] = s
        s *= chunk.shape[k]
    src = torch.from_numpy(chunk.reshape(-1).copy()).cuda()
    dst = torch.full((T, Y * X), -7.0, dtype=torch.float64 if dst_dtype == "float64" else torch.float32, device="cuda")
    elem_off = sum(o * st for o, st in zip(off, strides))
    rc = _lib.lib().agf_tile_place_run(src.data_ptr() + elem_off * chunk.dtype.itemsize, _CODES[str(chunk.dtype)],
                                      ext[0], ext[1], ext[2], strides[0], strides[1], strides[2], dst.data_ptr(),
                                      _CODES[dst_dtype], Y * X, X, dst_off[0], dst_off[1], dst_off[2], int(packed),
                                      scale, offset, int(fill is not None), 0.0 if fill is None else float(fill), T,
                                      torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    torch.cuda.synchronize()
    return dst.cpu().numpy().reshape(T, Y, X)


@pytest.mark.parametrize("perm", list(itertools.permutations(range(3))))
@pytest.mark.parametrize("sdt,ddt", [("float32", "float32"), ("float64", "float64"), ("float32", "float64"),
                                     ("int16", "float64"), ("int16", "float32"), ("int32", "float64"), ("uint8", "float32"),
                                     ("int8", "float64"), ("uint16", "float32")])
def test_tile_place_matches_numpy(perm, sdt, ddt):
    rng = np.random.default_rng(100 * perm[0] + 10 * perm[1] + len(sdt))
    logical = (70, 37, 45)                                       # chunk extent along (t, y, x): ragged w.r.t. 32 x 32 tiles
    shape = tuple(logical[perm[k]] for k in range(3))           # stored axis k is logical axis perm[k]
    if sdt.startswith("float"):
        chunk = rng.normal(0, 100, shape).astype(sdt)
        chunk[rng.random(shape) < 0.01] = np.nan
    else:
        info = np.iinfo(sdt)
        chunk = rng.integers(info.min, info.max, shape, dtype=sdt, endpoint=True)
    packed = not sdt.startswith("float")
    scale, offset = (0.01, 273.15) if packed else (1.0, 0.0)
    as_tyx = np.transpose(chunk, np.argsort(perm))               # logical (t, y, x) view of the stored chunk
    off, ext = (3, 2, 5), (61, 33, 40)
    dst_shape, dst_off = (80, 50, 64), (7, 11, 9)
    block = as_tyx[off[0]:off[0] + ext[0], off[1]:off[1] + ext[1], off[2]:off[2] + ext[2]]
    fill = float(block[~np.isnan(block.astype(np.float64))][17])         # a value that occurs inside the placed block
    got = _place(chunk, perm, ext, off, dst_shape, dst_off, ddt, packed, scale, offset, fill)
    dec = (block.astype(np.float64) * scale + offset) if packed else block.astype(np.float64)
    dec = dec.astype(ddt)
    dec[block.astype(np.float64) == fill] = np.nan
    want = np.full(dst_shape, -7.0, ddt)
    want[dst_off[0]:dst_off[0] + ext[0], dst_off[1]:dst_off[1] + ext[1], dst_off[2]:dst_off[2] + ext[2]] = dec
    assert (block.astype(np.float64) == fill).any()
    _exact(got, want)


def test_tile_place_rejects_bad_arguments():
    import torch
    a = torch.zeros(64, device="cuda")
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    assert L.agf_tile_place_run(a.data_ptr(), 0, 1, 2, 9, 16, 8, 1, a.data_ptr(), 0, 16, 8, 0, 0, 0, 0, 1.0, 0.0, 0, 0.0, 1, s) == -1
    assert b"does not fit" in L.agf_last_error()
    assert L.agf_tile_place_run(a.data_ptr(), 1, 1, 1, 1, 1, 1, 1, a.data_ptr(), 0, 16, 8, 0, 0, 0, 0, 1.0, 0.0, 0, 0.0, 1, s) == -1
    assert L.agf_tile_place_run(a.data_ptr(), 9, 1, 1, 1, 1, 1, 1, a.data_ptr(), 0, 16, 8, 0, 0, 0, 0, 1.0, 0.0, 0, 0.0, 1, s) == -1
    assert L.agf_tile_place_run(None, 0, 1, 1, 1, 1, 1, 1, a.data_ptr(), 0, 16, 8, 0, 0, 0, 0, 1.0, 0.0, 0, 0.0, 1, s) == -1
    assert L.agf_tile_place_run(a.data_ptr(), 0, 0, 1, 1, 1, 1, 1, a.data_ptr(), 0, 16, 8, 0, 0, 0, 0, 1.0, 0.0, 0, 0.0, 1, s) == 0
    # a time offset past the raster's rows is rejected instead of written (ABI 5)
    assert L.agf_tile_place_run(a.data_ptr(), 0, 2, 1, 1, 1, 1, 1, a.data_ptr(), 0, 16, 8, 3, 0, 0, 0, 1.0, 0.0, 0, 0.0, 4, s) == -1
    assert b"do not fit a raster of 4 rows" in L.agf_last_error()


LAYOUTS = {
    "time_major_daily_chunks": dict(dims=("time", "latitude", "longitude"), chunks={"time": 24}, fmt=3, comp="zstd", order="C"),
    "time_contiguous_tiles": dict(dims=("latitude", "longitude", "time"), chunks={"latitude": 2, "longitude": 5}, fmt=3,
                                  comp="zstd", order="C"),
    "ragged_3d_chunks_v2_F": dict(dims=("time", "latitude", "longitude"), chunks={"time": 100, "latitude": 3, "longitude": 7},
                                  fmt=2, comp="zlib", order="F"),
    "lon_time_lat_lz4": dict(dims=("longitude", "time", "latitude"), chunks={"time": 250, "longitude": 8}, fmt=2, comp="lz4",
                             order="C"),
}


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily", "monthly_mix"])
def test_aggregate_dataset_from_zarr_is_bitwise_the_in_memory_result(tmp_path, name, layout):
    import torch
    lay = LAYOUTS[layout]
    arr, t, lat, lon = _raster("float32", True, T=24 * 40 + 5, seed=21)
    rng = np.random.default_rng(4)
    wdf, shp = _weights_case(lat, lon, rng)
    store = zarrio.write_dataset(str(tmp_path / "r.zarr"), arr, t, lat, lon, var="t2m", dims=lay["dims"], chunks=lay["chunks"],
                                 zarr_format=lay["fmt"], compressor=lay["comp"], order=lay["order"])

    def run(ds):
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    engine.OPTIONS["target_stripes"] = 11
    old = dict(stream.OPTIONS)
    try:
        stream.OPTIONS.update(staging_slots=3, staging_threads=2)
        resident = run(af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, True))
        before = stream.LAST_STATS
        ds = af.dataset_from_path(store, var="t2m")
        assert getattr(ds.values, "is_chunked_raster", False)
        got = run(ds)
        st = stream.LAST_STATS
        assert st is not before and st["chunked"] and st["chunks"] == len(ds.values.tiles()) and st["absent_chunks"] == 0
        if layout == "time_major_daily_chunks":
            assert st["direct_copies"] == st["chunks"] and st["place_launches"] == 0 and st["k1_launches"] > 1
        else:
            assert st["place_launches"] == st["chunks"]
    finally:
        stream.OPTIONS.update(old)
    want = orc.aggregate_dataset(orc.OWeights(wdf, np.arange(arr.shape[1] * arr.shape[2]), shp, "geoid", "nan"),
                                 orc.ODataset(arr, t, lat, lon, True), aggregator_dict=SPECS[name])
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    _exact(got[vals].values, resident[vals].values)
    _close(got[vals].values, want[vals].values, 1e-11)


def test_packed_store_with_missing_chunks_and_time_sel(tmp_path):
    """int16 + scale_factor / add_offset / _FillValue (the usual ERA5 packing) decoded on the device; a time
    chunk without files is all fill; ``time_sel`` windows the store lazily."""
    T, Y, X = 24 * 9, 6, 10
    rng = np.random.default_rng(8)
    packed = rng.integers(-2000, 4000, (T, Y, X)).astype(np.int16)
    packed[rng.random((T, Y, X)) < 0.01] = -32767
    packed[24 * 4: 24 * 5] = -32767                                        # day 5 missing entirely
    t = pd.date_range("2000-12-30", periods=T, freq="h")
    lat, lon = np.linspace(49.75, 48.5, Y), np.linspace(235.0, 237.25, X)
    root = zarrio.write_dataset(str(tmp_path / "p.zarr"), np.zeros((T, Y, X), np.float32), t, lat, lon, var="t2m")
    import shutil
    shutil.rmtree(root + "/t2m")
    zarrio.write_array(root + "/t2m", packed, [24, 6, 4], ["time", "latitude", "longitude"],
                       {"scale_factor": 0.01, "add_offset": 273.15, "_FillValue": -32767}, zarr_format=2, compressor="zlib",
                       fill_value=-32767, skip_fill_chunks=True)
    ds = af.dataset_from_path(root, var="t2m", preprocess="kelvin_to_celsius", time_sel="2001")
    assert ds.shape == (24 * 7, Y, X) and ds.dtype == np.float64
    host = packed.astype(np.float64) * 0.01 + 273.15
    host[packed == -32767] = np.nan
    host = host[48:]
    dev = engine.to_device(ds.values).cpu().numpy()
    assert stream.LAST_STATS["absent_chunks"] == 3
    _exact(dev, host)
    spec = dict(tavg=[("aggregate", {"calc": "nanmean", "groupby": "date"})],
                cnt=[("aggregate", {"calc": "bins", "groupby": "date", "ddargs": [0, 40, 0]})])
    mem = af.Dataset.from_arrays(host, ds.time, lat, lon, True, preprocess="kelvin_to_celsius")
    regions = af.GeoRegions.from_rectangles(["a", "b"], lon_min=[-125.0, -124.0], lon_max=[-124.0, -122.5],
                                            lat_min=[48.4, 48.4], lat_max=[49.9, 49.9])
    w = af.weights_from_objects(mem, regions)
    w.calculate_weights()
    got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    want = af.aggregate_dataset(weights=w, dataset=mem, aggregator_dict=spec)
    assert list(got.columns) == list(want.columns) and len(got) == len(want) > 0
    _exact(got[["tavg", "cnt"]].values, want[["tavg", "cnt"]].values)
    got_t = af.aggregate_time(dataset=ds, aggregator_dict=spec)
    want_t = af.aggregate_time(dataset=mem, aggregator_dict=spec)
    for k in want_t:
        _exact(np.asarray(got_t[k].values), np.asarray(want_t[k].values))


def test_reference_converter_layout_end_to_end(tmp_path):
    """dataset_to_zarr (the reference's time-contiguous layout, zarr v3 + zstd) -> aggregate_dataset."""
    import torch
    arr, t, lat, lon = _raster("float32", False, T=24 * 20, Y=8, X=12, seed=5)
    mem = af.Dataset.from_arrays(arr, t, lat, lon, True, name="t2m")
    ds = af.dataset_to_zarr(mem, str(tmp_path / "tc.zarr"), chunking={"time": -1, "latitude": 3, "longitude": 5})
    regions = af.GeoRegions.from_rectangles([f"r{i}" for i in range(4)], lon_min=-125.0 + 10 * np.arange(4),
                                            lon_max=-115.0 + 10 * np.arange(4), lat_min=np.full(4, 24.0), lat_max=np.full(4, 50.0))
    w = af.weights_from_objects(mem, regions)
    w.calculate_weights()
    spec = SPECS["c1_tavg_poly"]
    got = af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=spec)
    want = af.aggregate_dataset(weights=w, dataset=af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, True),
                                aggregator_dict=spec)
    assert len(got) == len(want) > 0
    _exact(got[["tavg_1", "tavg_2"]].values, want[["tavg_1", "tavg_2"]].values)


# ---- Blosc-LZ4 chunks inflated by the decompression engine -----------------------------------------------
def _engine():
    mask, max_len = stream.device_decompress_caps()
    return bool(mask & 4) and max_len > 0, max_len


@pytest.mark.parametrize("ts", [2, 4, 8])
def test_unshuffle_and_segment_copy_match_numpy(ts):
    import torch
    rng = np.random.default_rng(ts)
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    for nbytes, bs in [(10_007 * ts + 3, 512 * ts), (4096 * ts, 4096 * ts), (77, 64 * ts)]:
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        want = np.concatenate([zarrio._unshuffle(data[b:b + bs], ts) for b in range(0, nbytes, bs)])
        src, dst = torch.from_numpy(data).cuda(), torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        _lib.check(L.agf_unshuffle_run(src.data_ptr(), dst.data_ptr(), nbytes, ts, bs, s))
        assert np.array_equal(dst.cpu().numpy(), want)
    assert L.agf_unshuffle_run(src.data_ptr(), dst.data_ptr(), 10, 3, 64, s) == -2
    table = np.array([[3, 500, 77], [2100, 0, 5500], [3000, 2000, 1]], np.int64)                # src_off, dst_off, len
    src = torch.from_numpy(rng.integers(0, 256, 6000, dtype=np.uint8)).cuda()
    dst = torch.zeros(6000, dtype=torch.uint8, device="cuda")
    _lib.check(L.agf_copy_segments_run(src.data_ptr(), dst.data_ptr(), torch.from_numpy(table).cuda().data_ptr(), 3, s))
    want = np.zeros(6000, np.uint8)
    h = src.cpu().numpy()
    for a, d, m in table.T:
        want[d:d + m] = h[a:a + m]
    assert np.array_equal(dst.cpu().numpy(), want)


def test_decompression_engine_inflates_blosc_frames():
    """agf_decompress_lz4_run + agf_copy_segments_run + agf_unshuffle_run on a Blosc frame == the host decoder."""
    import ctypes as C
    import torch
    have, max_len = _engine()
    if not have:
        pytest.skip("no LZ4 decompression engine on this device / driver")
    rng = np.random.default_rng(3)
    vals = (20 + np.cumsum(rng.normal(0, 0.05, 700_001))).astype(np.float32)               # noisy low mantissa bytes
    data = vals.tobytes()
    frame = zarrio.blosc_compress(data, 4, "lz4", 1, 64 << 10)
    u8 = np.frombuffer(frame, np.uint8)
    plan = zarrio.blosc_device_plan(u8, max_len)
    assert plan is not None and plan.kind == "lz4" and len(plan.src_off) > 0 and plan.raw.shape[1] > 0
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    d_frame = torch.from_numpy(u8.copy()).cuda()
    infl = torch.zeros(len(data), dtype=torch.uint8, device="cuda")
    plain = torch.zeros(len(data), dtype=torch.uint8, device="cuda")
    act = torch.zeros(len(plan.src_off), dtype=torch.int32, device="cuda")
    p64 = C.POINTER(C.c_int64)
    _lib.check(L.agf_decompress_lz4_run(d_frame.data_ptr(), plan.src_off.ctypes.data_as(p64), plan.src_len.ctypes.data_as(p64),
                                        infl.data_ptr(), plan.dst_off.ctypes.data_as(p64), plan.dst_len.ctypes.data_as(p64),
                                        len(plan.src_off), act.data_ptr(), s))
    _lib.check(L.agf_copy_segments_run(d_frame.data_ptr(), infl.data_ptr(), torch.from_numpy(plan.raw).cuda().data_ptr(),
                                       plan.raw.shape[1], s))
    _lib.check(L.agf_unshuffle_run(infl.data_ptr(), plain.data_ptr(), plan.nbytes, plan.typesize, plan.blocksize, s))
    torch.cuda.synchronize()
    assert np.array_equal(act.cpu().numpy(), plan.dst_len.astype(np.int32))
    assert plain.cpu().numpy().tobytes() == data


@pytest.mark.parametrize("device_decompress", [True, False, "tables_uploaded"])
@pytest.mark.parametrize("layout", ["time_major_daily_chunks", "time_contiguous_tiles", "lon_time_lat_lz4"])
def test_blosc_store_matches_in_memory_result(tmp_path, layout, device_decompress):
    # "tables_uploaded": the per-chunk tables as separate pageable uploads instead of riding the chunk's own copy (the default)
    inline, device_decompress = device_decompress != "tables_uploaded", bool(device_decompress)
    import torch
    lay = LAYOUTS[layout]
    arr, t, lat, lon = _raster("float32", True, T=24 * 40 + 5, seed=23)
    rng = np.random.default_rng(6)
    wdf, shp = _weights_case(lat, lon, rng)
    store = zarrio.write_dataset(str(tmp_path / "b.zarr"), arr, t, lat, lon, var="t2m", dims=lay["dims"], chunks=lay["chunks"],
                                 zarr_format=lay["fmt"], compressor="blosc", order=lay["order"])
    name = "c3_bins_and_poly"

    def run(ds):
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    old = dict(stream.OPTIONS)
    try:
        stream.OPTIONS.update(staging_slots=3, staging_threads=2, device_decompress=device_decompress, inline_chunk_tables=inline)
        resident = run(af.Dataset.from_arrays(torch.from_numpy(arr).cuda(), t, lat, lon, True))
        ds = af.dataset_from_path(store, var="t2m")
        assert ds.values.array.blosc_only
        got = run(ds)
        st = stream.LAST_STATS
        have, _ = _engine()
        assert st["device_decompress"] == (device_decompress and have)
        if st["device_decompress"]:
            assert st["inflated_on_device"] == st["chunks"] and st["engine_ops"] > 0 and st["h2d_bytes"] < arr.nbytes
        else:
            assert st["inflated_on_device"] == 0
        dev_raster = engine.to_device(ds.values).cpu().numpy()
    finally:
        stream.OPTIONS.update(old)
    _exact(dev_raster, arr)
    vals = [c for c in resident.columns if c not in ("geoid", "time")]
    assert len(got) == len(resident) > 0
    _exact(got[vals].values, resident[vals].values)


def test_multi_file_dataset_matches_single_array(tmp_path):
    import torch
    arr, t, lat, lon = _raster("float32", True, T=24 * 30, seed=31)
    d = str(tmp_path)
    np.savez(d + "/p2.npz", t2m=arr[240:480], time=t[240:480].values, latitude=lat, longitude=lon)
    zarrio.write_dataset(d + "/p1.zarr", arr[:240], t[:240], lat, lon, var="t2m", chunks={"time": 24}, compressor="zstd", zarr_format=3)
    zarrio.write_dataset(d + "/p3.zarr", arr[480:], t[480:], lat, lon, var="t2m", chunks={"time": 48}, compressor="blosc")
    ds = af.dataset_from_path([d + "/p1.zarr", d + "/p2.npz", d + "/p3.zarr"], var="t2m")
    rng = np.random.default_rng(9)
    wdf, shp = _weights_case(lat, lon, rng)

    def run(dset):
        w = af.weights_from_objects(dset, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=dset, aggregator_dict=SPECS["monthly_mix"])

    got, want = run(ds), run(af.Dataset.from_arrays(arr, t, lat, lon, True))
    vals = [c for c in want.columns if c not in ("geoid", "time")]
    assert len(got) == len(want) > 0
    _exact(got[vals].values, want[vals].values)                       # same host feed, same stripes
    _exact(engine.to_device(ds.values).cpu().numpy(), arr)


# ---- NetCDF-4 (HDF5) files: chunks inflated + un-shuffled by host threads, placed / unpacked on the device (hdf5io.py) ----
@pytest.mark.parametrize("layout", ["time_major_packed_deflate", "lat_lon_time_packed_deflate", "float32_contiguous",
                                    "float32_chunked_shuffle"])
@pytest.mark.parametrize("name", ["c3_bins_and_poly", "c3b_daily"])
def test_netcdf4_file_matches_in_memory_result(tmp_path, name, layout):
    import torch
    from aggfly_b200 import hdf5io
    arr, t, lat, lon = _raster("float32", False, T=24 * 21 + 3, seed=41)
    T, Y, X = arr.shape
    hours = ((t - t[0]) / np.timedelta64(1, "h")).astype(np.int32)
    units = f"hours since {t[0]:%Y-%m-%d %H:%M:%S}"
    path = str(tmp_path / "era5.nc")
    if "packed" in layout:
        scale, offset = 0.002, float(np.nanmean(arr))
        q = np.clip(np.rint((arr.astype(np.float64) - offset) / scale), -32000, 32000).astype(np.int16)
        q[7:11, 1, 2] = -32767
        decoded = q.astype(np.float64) * scale + offset
        decoded[q == -32767] = np.nan
        attrs = {"scale_factor": scale, "add_offset": offset, "_FillValue": np.int16(-32767), "units": "K"}
        if layout.startswith("time_major"):
            hdf5io.write_netcdf4(path, q, hours, units, lat, lon, var="t2m", chunks=(24, Y, X), attrs=attrs)
        else:
            hdf5io.write_netcdf4(path, np.ascontiguousarray(np.transpose(q, (1, 2, 0))), hours, units, lat, lon, var="t2m",
                                 dims=("latitude", "longitude", "time"), chunks=(3, 5, T), attrs=attrs)
    else:
        decoded = arr
        hdf5io.write_netcdf4(path, arr, hours, units, lat, lon, var="t2m", chunks=None if "contiguous" in layout else (48, 4, X),
                             deflate=None, shuffle="shuffle" in layout)
    rng = np.random.default_rng(12)
    wdf, shp = _weights_case(lat, lon, rng)

    def run(ds):
        w = af.weights_from_objects(ds, af.GeoRegions(shp, "geoid"), zero_weight="nan")
        w.weights = wdf
        return af.aggregate_dataset(weights=w, dataset=ds, aggregator_dict=SPECS[name])

    engine.OPTIONS["target_stripes"] = 5                       # same stripes (same merge order) in both runs
    old = dict(stream.OPTIONS)
    try:
        stream.OPTIONS.update(staging_slots=3, staging_threads=2)
        resident = run(af.Dataset.from_arrays(torch.from_numpy(decoded).cuda(), t, lat, lon, True))
        ds = af.dataset_from_path(path, var="t2m")
        assert getattr(ds.values, "is_chunked_raster", False) and ds.dtype == decoded.dtype
        assert len(ds.time) == T and ds.time[0] == t[0] and ds.time[-1] == t[-1]
        got = run(ds)
        assert stream.LAST_STATS.get("chunked") and stream.LAST_STATS["chunks"] >= 1
        dev_raster = engine.to_device(ds.values).cpu().numpy()
    finally:
        stream.OPTIONS.update(old)
    _exact(dev_raster, decoded)
    vals = [c for c in resident.columns if c not in ("geoid", "time")]
    assert len(got) == len(resident) > 0
    _exact(got[vals].values, resident[vals].values)
