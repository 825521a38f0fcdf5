"""Zarr reader (aggfly_b200/zarrio.py) on the CPU: stores assembled BY HAND from the published v2 / v3
layouts (independent of the package's own writer), every codec the reader claims, CF decoding, the
lazy raster's tiles.  The reference reaches the same data through xarray's zarr backend
(aggfly/dataset/dataset.py:585-615, 697-707; aggfly/dataset/zarr_convert.py:50-121)."""
import gzip
import itertools
import json
import os
import struct
import zlib

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import aggfly_b200 as af
from aggfly_b200 import zarrio
from aggfly_b200.io import _auto_chunks


def _dump(path, obj):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f)


def _put(path, raw: bytes):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(raw)


def test_hand_built_v2_store(tmp_path):
    """A 3 x 4 int32 array in 2 x 3 chunks, big-endian, '/' separator, zlib; one chunk left out."""
    root = str(tmp_path / "a")
    _dump(root + "/.zarray", {"zarr_format": 2, "shape": [3, 4], "chunks": [2, 3], "dtype": ">i4", "order": "C",
                             "compressor": {"id": "zlib", "level": 1}, "fill_value": -1, "filters": None,
                             "dimension_separator": "/"})
    _dump(root + "/.zattrs", {"_ARRAY_DIMENSIONS": ["y", "x"], "units": "K"})
    full = np.arange(12, dtype=np.int32).reshape(3, 4)
    for (i, j) in [(0, 0), (0, 1), (1, 0)]:                                   # chunk (1, 1) is absent -> fill value
        block = np.full((2, 3), -1, ">i4")
        part = full[2 * i: 2 * i + 2, 3 * j: 3 * j + 3]
        block[: part.shape[0], : part.shape[1]] = part
        _put(f"{root}/{i}/{j}", zlib.compress(block.tobytes()))
    a = zarrio.ZarrArray(root)
    assert a.shape == (3, 4) and a.chunks == (2, 3) and a.dims == ("y", "x") and a.attrs["units"] == "K"
    want = full.copy()
    want[2:, 3:] = -1
    assert np.array_equal(a.read(), want) and a.read().dtype == np.int32
    assert np.array_equal(a[1:3, 2:4], want[1:3, 2:4]) and a[2, 0] == 8 and np.array_equal(a[-1], want[-1])


def test_hand_built_v3_store(tmp_path):
    """A 4 x 3 float32 array, chunks 2 x 3, codecs transpose([1, 0]) -> bytes(little) -> gzip -> crc32c."""
    root = str(tmp_path / "g")
    _dump(root + "/zarr.json", {"zarr_format": 3, "node_type": "group", "attributes": {"title": "t"}})
    _dump(root + "/v/zarr.json", {
        "zarr_format": 3, "node_type": "array", "shape": [4, 3], "data_type": "float32",
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [2, 3]}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}}, "fill_value": "NaN",
        "codecs": [{"name": "transpose", "configuration": {"order": [1, 0]}},
                   {"name": "bytes", "configuration": {"endian": "little"}},
                   {"name": "gzip", "configuration": {"level": 1}}, {"name": "crc32c"}],
        "attributes": {"long_name": "x"}, "dimension_names": ["a", "b"]})
    full = np.arange(12, dtype="<f4").reshape(4, 3) / 4
    for i in range(2):
        stored = np.ascontiguousarray(full[2 * i: 2 * i + 2].T)                # transpose codec: axis order [1, 0]
        _put(f"{root}/v/c/{i}/0", gzip.compress(stored.tobytes()) + b"\0\0\0\0")
    g = zarrio.ZarrGroup(root)
    assert g.names() == ["v"] and g.attrs == {"title": "t"}
    a = g["v"]
    assert a.storage_axes == (1, 0) and a.dims == ("a", "b") and np.isnan(a.fill_value)
    assert np.array_equal(a.read(), full)
    # v2-style keys inside a v3 array
    meta = json.load(open(root + "/v/zarr.json"))
    meta["chunk_key_encoding"] = {"name": "v2", "configuration": {"separator": "."}}
    meta["codecs"] = [{"name": "bytes", "configuration": {"endian": "big"}}]
    _dump(root + "/w/zarr.json", meta)
    for i in range(2):
        _put(f"{root}/w/{i}.0", full[2 * i: 2 * i + 2].astype(">f4").tobytes())
    assert np.array_equal(zarrio.ZarrArray(root + "/w").read(), full)


def _blosc_frame(data: bytes, typesize: int, blocksize: int, cname: str, shuffle: int, split: bool = True) -> bytes:
    """A Blosc-1 frame assembled from the published layout (header, block offsets, per-split streams)."""
    comp_id = {"lz4": 1, "zlib": 3, "zstd": 4}[cname]
    enc = {"lz4": lambda b: pa.Codec("lz4_raw").compress(b, asbytes=True), "zlib": lambda b: zlib.compress(b, 1),
           "zstd": lambda b: pa.Codec("zstd").compress(b, asbytes=True)}[cname]
    nbytes = len(data)
    nblocks = -(-nbytes // blocksize)
    can_split = split and typesize <= 16 and blocksize // typesize >= 128
    flags = (comp_id << 5) | (0x10 if not can_split else 0) | {0: 0, 1: 0x1, 2: 0x4}[shuffle]
    body, bstarts = b"", []
    arr = np.frombuffer(data, np.uint8)
    for b in range(nblocks):
        blk = arr[b * blocksize: (b + 1) * blocksize]
        n = blk.size // typesize
        if shuffle == 1 and typesize > 1:
            blk = np.concatenate([blk[: n * typesize].reshape(n, typesize).T.reshape(-1), blk[n * typesize:]])
        elif shuffle == 2:
            n8 = n & ~7
            bits = np.unpackbits(blk[: n8 * typesize].reshape(n8, typesize), axis=1, bitorder="little")
            blk = np.concatenate([np.packbits(bits.T, axis=1, bitorder="little").reshape(-1), blk[n8 * typesize:]])
        leftover = blk.size != blocksize
        nsplits = typesize if (can_split and not leftover) else 1
        ne = blk.size // nsplits
        bstarts.append(16 + 4 * nblocks + len(body))
        for s in range(nsplits):
            raw = blk[s * ne: (s + 1) * ne].tobytes()
            c = enc(raw)
            if len(c) >= len(raw):
                c = raw                                                         # stored uncompressed
            body += struct.pack("<i", len(c)) + c
    head = struct.pack("<4B3I", 2, 1, flags, typesize, nbytes, blocksize, 16 + 4 * nblocks + len(body))
    return head + struct.pack(f"<{nblocks}i", *bstarts) + body


@pytest.mark.parametrize("cname,shuffle,split", [("lz4", 1, True), ("lz4", 0, False), ("zstd", 1, True), ("zlib", 2, False),
                                                 ("zstd", 2, True)])
def test_blosc_frames(cname, shuffle, split):
    rng = np.random.default_rng(5)
    data = (np.cumsum(rng.integers(-3, 4, 5000)) / 8).astype(np.float32).tobytes() + b"xyz"      # ragged tail
    frame = _blosc_frame(data, 4, 4096, cname, shuffle, split)
    assert zarrio.blosc_decompress(frame).tobytes() == data
    memcpyed = struct.pack("<4B3I", 2, 1, 0x2, 4, len(data), len(data), len(data) + 16) + data
    assert zarrio.blosc_decompress(memcpyed).tobytes() == data


@pytest.mark.parametrize("n,ts,bs,shuffle,split", [(1_000_003, 4, 65536, 1, True), (300_000, 4, 0, 1, True), (5000, 2, 256, 1, True),
                                                   (40000, 8, 4096, 0, False), (100, 4, 64, 1, True), (70000, 4, 8192, 1, False)])
def test_blosc_device_plan_describes_the_frame(n, ts, bs, shuffle, split):
    """``blosc_device_plan`` is what the GPU path executes (LZ4 streams -> decompression engine, raw streams ->
    segment copies, then the byte un-shuffle): emulate those three steps with NumPy and compare with the data.
    Smooth high bytes + noisy low bytes, so that some byte planes are stored uncompressed like real float data."""
    rng = np.random.default_rng(n)
    smooth = (np.cumsum(rng.integers(-3, 4, n // 4 + 1)) * 1024 + rng.integers(0, 256, n // 4 + 1)).astype("<i4")
    data = smooth.tobytes()[:n]
    frame = zarrio.blosc_compress(data, ts, "lz4", shuffle, bs, split)
    assert zarrio.blosc_decompress(frame).tobytes() == data
    assert frame == _blosc_frame(data, ts, bs or (256 << 10), "lz4", shuffle, split) or bs == 0 or n < bs
    u8 = np.frombuffer(frame, np.uint8)
    plan = zarrio.blosc_device_plan(u8, 4 << 20)
    assert plan is not None and plan.kind == "lz4" and plan.nbytes == n and plan.typesize == ts
    out = np.zeros(n, np.uint8)
    for a, l, d, m in zip(plan.src_off, plan.src_len, plan.dst_off, plan.dst_len):
        out[d:d + m] = np.frombuffer(pa.Codec("lz4_raw").decompress(u8[a:a + l].tobytes(), decompressed_size=int(m)), np.uint8)
    for a, d, m in plan.raw.T:
        out[d:d + m] = u8[a:a + m]
    if ts == 4 and n > 1000 and split:
        assert plan.raw.shape[1] > 0                                     # the noisy byte plane
    if plan.shuffled:
        out = np.concatenate([zarrio._unshuffle(out[b:b + plan.blocksize], ts) for b in range(0, n, plan.blocksize)])
    assert out.tobytes() == data
    assert zarrio.blosc_device_plan(u8, 16) is None                      # streams above the engine's limit: host decode
    assert zarrio.blosc_device_plan(np.frombuffer(zarrio.blosc_compress(data, ts, "zstd", shuffle, bs, split), np.uint8), 4 << 20) is None
    mem = struct.pack("<4B3I", 2, 1, 0x2, ts, n, n, n + 16) + data
    assert zarrio.blosc_device_plan(np.frombuffer(mem, np.uint8), 4 << 20).kind == "memcpy"
    assert zarrio.blosc_device_plan(u8[: u8.size // 2], 4 << 20) is None   # truncated


def test_v2_blosc_and_filters(tmp_path):
    vals = np.arange(6 * 5, dtype="<i2").reshape(6, 5) * 3
    root = str(tmp_path / "b")
    _dump(root + "/.zarray", {"zarr_format": 2, "shape": [6, 5], "chunks": [6, 5], "dtype": "<i2", "order": "F",
                             "compressor": {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0},
                             "fill_value": 0, "filters": [{"id": "delta", "dtype": "<i2"}]})
    stored = vals.tobytes(order="F")
    delta = np.diff(np.frombuffer(stored, "<i2"), prepend=np.int16(0)).astype("<i2")
    _put(root + "/0.0", _blosc_frame(delta.tobytes(), 2, 64, "lz4", 1))
    assert np.array_equal(zarrio.ZarrArray(root).read(), vals)
    # numcodecs Shuffle filter + bz2
    root = str(tmp_path / "c")
    _dump(root + "/.zarray", {"zarr_format": 2, "shape": [30], "chunks": [30], "dtype": "<f8", "order": "C",
                             "compressor": {"id": "bz2", "level": 1}, "fill_value": "NaN",
                             "filters": [{"id": "shuffle", "elementsize": 8}]})
    import bz2
    v = np.linspace(0, 1, 30)
    _put(root + "/0", bz2.compress(np.frombuffer(v.tobytes(), np.uint8).reshape(30, 8).T.tobytes()))
    assert np.array_equal(zarrio.ZarrArray(root).read(), v)
    _dump(root + "/.zarray", dict(json.load(open(root + "/.zarray")), filters=[{"id": "quantize", "digits": 2}]))
    with pytest.raises(NotImplementedError):
        zarrio.ZarrArray(root)


def _raster(T=50, Y=7, X=9, seed=0):
    rng = np.random.default_rng(seed)
    vals = rng.normal(10, 5, (T, Y, X)).astype(np.float32)
    vals[3, 2, :] = np.nan
    return vals, pd.date_range("2001-01-01", periods=T, freq="h"), np.linspace(40, 37, Y), np.linspace(230, 234, X)


@pytest.mark.parametrize("fmt,comp", [(2, None), (2, "zlib"), (2, "zstd"), (2, "lz4"), (2, "gzip"), (2, "blosc"), (3, None),
                                      (3, "zstd"), (3, "gzip"), (3, "blosc")])
def test_dataset_roundtrip_all_layouts(tmp_path, fmt, comp):
    vals, t, lat, lon = _raster()
    T, Y, X = vals.shape
    dim_orders = [("time", "latitude", "longitude"), ("latitude", "longitude", "time"), ("longitude", "time", "latitude")]
    chunkings = [{"time": 24}, {"latitude": 3, "longitude": 4}, {"time": 7, "latitude": 2, "longitude": 5}]
    for n, (order, dims, chunks) in enumerate(itertools.product("CF", dim_orders, chunkings)):
        store = zarrio.write_dataset(str(tmp_path / f"s{n}.zarr"), vals, t, lat, lon, var="t2m", dims=dims, chunks=chunks,
                                     zarr_format=fmt, compressor=comp, order=order)
        ds = af.dataset_from_path(store, var="t2m")
        assert ds.shape == (T, Y, X) and ds.dtype == np.float32 and getattr(ds.values, "is_chunked_raster", False)
        assert np.array_equal(np.asarray(ds.values), vals, equal_nan=True)
        assert (ds.time == t).all() and np.array_equal(ds.latitude, lat) and np.array_equal(ds.longitude, lon)
        cover = np.zeros(vals.shape, int)
        buf = np.empty(ds.values.slot_elems, np.float32)
        for tl in ds.values.tiles():                                  # what stream.feed_chunked hands to the device
            cover[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1] += 1
            assert ds.values.load(tl, buf)
            nt, ny, nx = tl.extent
            blk = np.lib.stride_tricks.as_strided(buf[tl.offset:], (nt, ny, nx), (tl.st * 4, tl.sy * 4, tl.sx * 4))
            assert np.array_equal(blk, vals[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1], equal_nan=True)
        assert (cover == 1).all()
        t0s = [tl.t0 for tl in ds.values.tiles()]
        assert t0s == sorted(t0s)                                     # time chunk by time chunk
        win = ds.values[:, 1:5, 2:8][3:, :, 1:]                       # lazy windows compose
        assert getattr(win, "is_chunked_raster", False) and np.array_equal(np.asarray(win), vals[3:, 1:5, 3:8], equal_nan=True)


def test_time_sel_and_clip_stay_lazy(tmp_path):
    vals, _, lat, lon = _raster(T=72)
    t = pd.date_range("2001-12-31", periods=72, freq="h")
    store = zarrio.write_dataset(str(tmp_path / "y.zarr"), vals, t, lat, lon, var="t2m", chunks={"time": 24})
    ds = af.dataset_from_path(store, var="t2m", time_sel="2002")
    assert ds.shape[0] == 48 and getattr(ds.values, "is_chunked_raster", False)
    assert np.array_equal(np.asarray(ds.values), vals[24:], equal_nan=True)
    from aggfly_b200.io import clip_to_extent
    cl = clip_to_extent(ds, -129.0, -128.0, 38.0, 39.0)
    assert getattr(cl.values, "is_chunked_raster", False) and cl.shape[1] < 7 and cl.shape[2] < 9
    iy = [int(np.argmin(np.abs(lat - v))) for v in cl.latitude]
    ix = [int(np.argmin(np.abs(lon - v))) for v in cl.longitude]
    assert np.array_equal(np.asarray(cl.values), vals[24:][:, iy][:, :, ix], equal_nan=True)
    for layout in (dict(dims=("latitude", "longitude", "time"), chunks={"latitude": 2, "longitude": 4}),
                   dict(dims=("time", "latitude", "longitude"), chunks={"time": 24, "latitude": 3, "longitude": 5})):
        store = zarrio.write_dataset(str(tmp_path / "w.zarr"), vals, t, lat, lon, var="t2m", compressor="blosc", **layout)
        win = clip_to_extent(af.dataset_from_path(store, var="t2m", time_sel="2002"), -129.0, -128.0, 38.0, 39.0).values
        want = vals[24:][:, iy][:, :, ix]
        cover = np.zeros(want.shape, int)
        buf = np.empty(win.slot_elems, np.float32)
        for tl in win.tiles():                                        # the windowed tiles the device feed uses
            assert win.load(tl, buf)
            blk = np.lib.stride_tricks.as_strided(buf[tl.offset:], tl.extent, (tl.st * 4, tl.sy * 4, tl.sx * 4))
            assert np.array_equal(blk, want[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1], equal_nan=True)
            cover[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1] += 1
        assert (cover == 1).all() and win.single_time_chunk == (layout["dims"][0] == "latitude")
        import shutil
        shutil.rmtree(store)


def test_cf_packing_fill_and_missing_chunks(tmp_path):
    T, Y, X = 10, 4, 6
    rng = np.random.default_rng(2)
    packed = rng.integers(-3000, 3000, (T, Y, X)).astype(np.int16)
    packed[2, 1, 1] = -32767
    packed[5:] = -32767                                               # a whole time chunk of fill -> chunk files skipped
    root = str(tmp_path / "p.zarr")
    t = pd.date_range("2001-01-01", periods=T, freq="D")
    zarrio.write_dataset(root, np.zeros((T, Y, X), np.float32), t, np.arange(Y), np.arange(X), var="t2m", time_units="days")
    import shutil
    shutil.rmtree(root + "/t2m")
    zarrio.write_array(root + "/t2m", packed, [5, 2, 6], ["time", "latitude", "longitude"],
                       {"scale_factor": 0.01, "add_offset": 273.15, "_FillValue": -32767}, zarr_format=2, compressor="zlib",
                       fill_value=-32767, skip_fill_chunks=True)
    assert not os.path.exists(root + "/t2m/1.0.0")
    ds = af.dataset_from_path(root, var="t2m")
    want = packed.astype(np.float64) * 0.01 + 273.15
    want[packed == -32767] = np.nan
    assert ds.dtype == np.float64 and ds.values.packed and ds.values.fill == -32767.0
    assert np.array_equal(np.asarray(ds.values), want, equal_nan=True)
    assert (ds.time == t).all()
    loaded = [ds.values.load(tl, np.empty(ds.values.slot_elems, np.int16)) for tl in ds.values.tiles()]
    assert loaded == [True, True, False, False]
    # a v2 float array whose fill_value is a number and no _FillValue attribute: xarray masks it
    zarrio.write_array(root + "/f", np.array([[1.0, -9.0], [3.0, 4.0]], np.float32)[None], [1, 2, 2],
                       ["time", "latitude", "longitude"], None, zarr_format=2, compressor=None, fill_value=-9.0)
    r = zarrio.ChunkedRaster(zarrio.ZarrArray(root + "/f"), (0, 1, 2))
    assert np.array_equal(np.asarray(r)[0], [[1.0, np.nan], [3.0, 4.0]], equal_nan=True)


def test_xarray_style_v3_fill_value_attribute(tmp_path):
    """xarray stores a float variable's ``_FillValue`` in zarr v3 attributes as base64 of the float64."""
    import base64
    root = str(tmp_path / "x.zarr")
    data = np.array([[[1.0, -999.0], [3.0, np.nan]]], np.float32)
    for enc, masked in ((base64.standard_b64encode(struct.pack("<d", -999.0)).decode(), True),
                        (base64.standard_b64encode(struct.pack("<d", np.nan)).decode(), False), ("NaN", False), (-999.0, True)):
        zarrio.write_array(root + "/v", data, [1, 2, 2], ["time", "latitude", "longitude"], {"_FillValue": enc}, zarr_format=3,
                           compressor="zstd")
        r = zarrio.ChunkedRaster(zarrio.ZarrArray(root + "/v"), (0, 1, 2))
        assert (r.fill == -999.0) if masked else (r.fill is None)
        want = data.copy()
        if masked:
            want[0, 0, 1] = np.nan
        assert np.array_equal(np.asarray(r), want, equal_nan=True)


def test_calendar_time_axes(tmp_path):
    vals = np.zeros((800, 2, 2), np.float32)
    for cal in ("noleap", "360_day"):
        t = af.CalendarIndex.range(cal, 2001, 800, "D")
        store = zarrio.write_dataset(str(tmp_path / f"{cal}.zarr"), vals, t, [0.0, 1.0], [0.0, 1.0], var="tas", chunks={"time": 365})
        ds = af.dataset_from_path(store, var="tas")
        assert isinstance(ds.time, af.CalendarIndex) and ds.time.calendar == cal
        for f in ("year", "month", "day", "hour"):
            assert np.array_equal(getattr(ds.time, f), getattr(t, f))
    # fractional days since an origin on the standard calendar
    root = str(tmp_path / "frac.zarr")
    zarrio.write_dataset(root, vals[:4], pd.date_range("2001-01-01", periods=4, freq="6h"), [0.0, 1.0], [0.0, 1.0], var="tas",
                         time_units="days")
    a = zarrio.ZarrArray(root + "/time")
    assert a.dtype.kind == "f" and a.attrs["units"].startswith("days since 2001-01-01")
    assert (zarrio.decode_time(a) == pd.date_range("2001-01-01", periods=4, freq="6h")).all()


def test_errors(tmp_path):
    vals, t, lat, lon = _raster(T=5)
    store = zarrio.write_dataset(str(tmp_path / "e.zarr"), vals, t, lat, lon, var="t2m")
    with pytest.raises(KeyError):
        af.dataset_from_path(store, var="nope")
    with pytest.raises(ValueError):
        af.dataset_from_path(store, var="t2m", xycoords=("lon", "lat"))
    assert zarrio.looks_like_zarr(store) and not zarrio.looks_like_zarr(str(tmp_path))
    os.rename(store, str(tmp_path / "plain"))
    assert zarrio.looks_like_zarr(str(tmp_path / "plain"))                      # no ".zarr" in the name: the markers decide
    meta = json.load(open(str(tmp_path / "plain/t2m/.zarray")))
    meta["compressor"] = {"id": "pcodec"}
    _dump(str(tmp_path / "plain/t2m/.zarray"), meta)
    with pytest.raises(NotImplementedError):
        af.dataset_from_path(str(tmp_path / "plain"), var="t2m")


def test_auto_chunks_policy():
    """aggfly/dataset/zarr_convert.py:31-47."""
    assert _auto_chunks({"time": 8760, "latitude": 721, "longitude": 1440}, 4, 256) == {"time": -1, "latitude": 87, "longitude": 87}
    assert _auto_chunks({"time": 8760, "latitude": 20, "longitude": 30}, 4, 256) == {"time": -1, "latitude": 20, "longitude": 20}
    got = _auto_chunks({"time": 350640, "latitude": 721, "longitude": 1440}, 4, 64)
    assert got == {"time": 1024, "latitude": 128, "longitude": 128}


def test_dataset_to_zarr_reopens_time_contiguous(tmp_path):
    vals, t, lat, lon = _raster()
    ds = af.Dataset.from_arrays(vals, t, lat, lon, name="t2m", preprocess="x - 273.15")
    out = af.dataset_to_zarr(ds, str(tmp_path / "tc.zarr"), chunking={"time": -1, "latitude": 4, "longitude": 4})
    arr = out.values.array
    assert arr.dims == ("latitude", "longitude", "time") and arr.chunks == (4, 4, 50) and arr.zarr_format == 3
    assert np.array_equal(np.asarray(out.values), vals, equal_nan=True) and out.pre_ops == ds.pre_ops
    with pytest.raises(FileExistsError):
        af.dataset_to_zarr(ds, str(tmp_path / "tc.zarr"))
    assert af.dataset_to_zarr(ds, str(tmp_path / "tc.zarr"), overwrite=True, return_dataset=False) is None


def test_writer_takes_lazy_block_sources_and_threads(tmp_path):
    """write_array pulls blocks chunk by chunk from any object with shape / dtype / __getitem__ (the global
    bench hands it a device tensor) and may compress chunks on several threads."""
    vals, t, lat, lon = _raster(T=40, Y=9, X=11, seed=3)

    class Blocks:
        lazy_blocks = True
        shape, dtype, ndim = (9, 11, 40), np.dtype(np.float32), 3
        asked = []

        def __getitem__(self, sl):
            self.asked.append(sl)
            return np.ascontiguousarray(np.transpose(vals, (1, 2, 0))[sl])

    store = zarrio.write_dataset(str(tmp_path / "l.zarr"), vals[:1], t[:1], lat, lon, var="t2m", compressor=None)
    import shutil
    shutil.rmtree(store + "/t2m")
    shutil.rmtree(store + "/time")
    zarrio.write_array(store + "/time", np.arange(40, dtype=np.int64), [-1], ["time"],
                       {"units": "hours since 2001-01-01 00:00:00", "calendar": "proleptic_gregorian"}, compressor=None)
    src = Blocks()
    zarrio.write_array(store + "/t2m", src, [4, 5, -1], ["latitude", "longitude", "time"], None, zarr_format=2,
                       compressor="blosc", threads=3)
    assert len(src.asked) == 9 and json.load(open(store + "/time/.zarray"))["fill_value"] is None
    ds = af.dataset_from_path(store, var="t2m")
    assert ds.values.array.blosc_only and (ds.time == t).all()
    assert np.array_equal(np.asarray(ds.values), vals, equal_nan=True)
    tile = ds.values.tiles()[-1]                                      # edge chunk, stored full-size
    slot = np.zeros(ds.values.slot_elems * 4 + 70000, np.uint8)
    kind, n, plan = ds.values.load_stored(tile, slot, 4 << 20)
    assert kind == "de" and n == os.path.getsize(ds.values.array.chunk_path(tile.index)) and plan.nbytes == 4 * 5 * 40 * 4
    kind, n2, plan2 = ds.values.load_stored(tile, slot, 4 << 20, inline_tables=True)    # tables appended behind the frame
    a, b = plan2.inline
    assert a % 16 == 0 and b % 16 == 0 and n <= a < b <= n2 <= slot.size
    assert np.array_equal(slot[a:a + 4 * len(plan2.dst_len)].view(np.int32), plan2.dst_len)
    assert np.array_equal(slot[b:n2].view(np.int64).reshape(3, -1), plan2.raw)
    assert ds.values.load_stored(tile, slot[: n + 8], 4 << 20, inline_tables=True)[2].inline is None      # no room: not inlined
    assert ds.values.load_stored(tile, slot, 8) == ("host",)          # engine limit too small: decoded on the host instead
    blk = np.lib.stride_tricks.as_strided(slot[: plan.nbytes].view(np.float32)[tile.offset:], tile.extent,
                                          (tile.st * 4, tile.sy * 4, tile.sx * 4))
    assert np.array_equal(blk, vals[tile.t0:tile.t1, tile.y0:tile.y1, tile.x0:tile.x1], equal_nan=True)


def test_tiles_property_random_shapes_chunks_orders_windows(tmp_path):
    """Property check of ChunkedRaster.tiles(): for random shapes / chunkings / axis orders / C-F storage / windows
    the tiles partition the window and (offset, strides) address exactly the window's elements in the stored chunk."""
    from hypothesis import given, settings, strategies as st

    counter = [0]

    @settings(max_examples=60, deadline=None)
    @given(st.tuples(st.integers(1, 9), st.integers(1, 7), st.integers(1, 8)), st.permutations(["time", "latitude", "longitude"]),
           st.sampled_from("CF"), st.sampled_from([2, 3]), st.data())
    def check(shape, dims, order, fmt, data):
        T, Y, X = shape
        chunks = {d: data.draw(st.integers(1, n + 1)) for d, n in zip(("time", "latitude", "longitude"), shape)}
        vals = np.arange(T * Y * X, dtype=np.float32).reshape(T, Y, X)
        counter[0] += 1
        store = zarrio.write_dataset(str(tmp_path / f"h{counter[0]}.zarr"), vals, pd.date_range("2001-01-01", periods=T, freq="h"),
                                     np.arange(Y, dtype=float), np.arange(X, dtype=float), var="v", dims=tuple(dims), chunks=chunks,
                                     zarr_format=fmt, compressor=None, order=order)
        r = af.dataset_from_path(store, var="v").values
        lo = [data.draw(st.integers(0, n - 1)) for n in shape]
        hi = [data.draw(st.integers(l + 1, n)) for l, n in zip(lo, shape)]
        win = r[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        want = vals[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        assert win.shape == want.shape and np.array_equal(np.asarray(win), want)
        cover = np.zeros(want.shape, int)
        buf = np.empty(win.slot_elems, np.float32)
        for tl in win.tiles():
            assert win.load(tl, buf)
            blk = np.lib.stride_tricks.as_strided(buf[tl.offset:], tl.extent, (tl.st * 4, tl.sy * 4, tl.sx * 4))
            assert np.array_equal(blk, want[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1])
            cover[tl.t0:tl.t1, tl.y0:tl.y1, tl.x0:tl.x1] += 1
            if win.direct_rows(tl):
                assert tl.extent[1:] == want.shape[1:] and np.array_equal(
                    buf[tl.offset: tl.offset + blk.size].reshape(blk.shape), blk)
        assert (cover == 1).all()

    check()


def test_multi_file_dataset_is_a_lazy_time_concat(tmp_path):
    """A list / glob of files = one dataset along time (the reference's open_mfdataset branch,
    aggfly/dataset/dataset.py:686-695); parts of different formats, opened lazily, ordered by time."""
    import torch
    from aggfly_b200 import stream
    from aggfly_b200.dataset import TimeConcat
    rng = np.random.default_rng(0)
    lat, lon = np.linspace(40, 38, 5), np.linspace(250, 253, 7)
    full = rng.normal(280, 5, (24 * 9, 5, 7)).astype(np.float32)
    t = pd.date_range("2001-01-30", periods=24 * 9, freq="h")
    d = str(tmp_path)
    np.savez(d + "/b_2.npz", t2m=full[72:144], time=t[72:144].values, latitude=lat, longitude=lon)
    zarrio.write_dataset(d + "/a_1.zarr", full[:72], t[:72], lat, lon, var="t2m", chunks={"time": 24}, compressor="zstd", zarr_format=3)
    zarrio.write_dataset(d + "/c_3.zarr", full[144:], t[144:], lat, lon, var="t2m", chunks={"time": 24}, compressor="blosc")
    ds = af.dataset_from_path([d + "/c_3.zarr", d + "/b_2.npz", d + "/a_1.zarr"], var="t2m", preprocess="kelvin_to_celsius")
    assert isinstance(ds.values, TimeConcat) and ds.shape == full.shape and ds.dtype == np.float32 and len(ds.pre_ops) == 1
    assert (ds.time == t).all() and np.array_equal(np.asarray(ds.values), full)
    assert isinstance(ds.values[60:100], TimeConcat) and isinstance(ds.values[80:100], np.ndarray)
    assert getattr(ds.values[10:30], "is_chunked_raster", False)                       # inside one zarr part: that part's window
    assert np.array_equal(np.asarray(ds.values[:, 1:4, 2:6][100:200]), full[100:200, 1:4, 2:6])
    assert af.dataset_from_path(d + "/*_?.zarr", var="t2m").shape == (144, 5, 7)       # glob: the two zarr parts
    feb = af.dataset_from_path(d + "/*", var="t2m", time_sel="2001-02")
    assert feb.shape[0] == 168 and np.array_equal(np.asarray(feb.values), full[48:])
    with pytest.raises(ValueError, match="overlap"):
        af.dataset_from_path([d + "/a_1.zarr", d + "/a_1.zarr"], var="t2m")
    with pytest.raises(FileNotFoundError):
        af.dataset_from_path(d + "/nothing_*.zarr", var="t2m")
    # what the host feed does with it: row chunks staged by worker threads through np.copyto (stream._Staging.fill)
    host, src = stream._host_source(ds.values)
    assert host is None and src is ds.values
    ring = stream._Staging.__new__(stream._Staging)
    ring.slots, ring.events = [torch.empty(40 * 35, dtype=torch.float32)], [None]
    for r0, r1 in stream.chunk_rows(full.shape[0], 35 * 4, 40 * 35 * 4):
        got = ring.fill(0, src[r0:r1]).view(r1 - r0, 35).numpy()
        assert np.array_equal(got, full[r0:r1].reshape(r1 - r0, 35))


def test_decode_dtype_policy_is_pinned(tmp_path):
    """Integer and packed sources decode to float64, float sources keep their width.  xarray's CF decoding (which the
    reference reads through) picks float32 for integers of <= 2 bytes in older releases and follows the dtype of
    ``scale_factor`` in newer ones; this engine always takes the widest of those choices, so a panel from a packed store is
    computed at least as precisely as the reference's ("Decode dtype").  ``dataset.PackedRaster`` lets the caller
    choose float32 explicitly."""
    t = pd.date_range("2001-01-01", periods=6, freq="h")
    lat, lon = np.array([1.0, 0.0]), np.array([10.0, 11.0, 12.0])
    cases = {"int16_packed": (np.arange(36, dtype=np.int16).reshape(6, 2, 3), {"scale_factor": 0.5, "add_offset": 1.0}, np.float64),
             "int16_plain": (np.arange(36, dtype=np.int16).reshape(6, 2, 3), {}, np.float64),
             "uint8_fill_only": (np.arange(36, dtype=np.uint8).reshape(6, 2, 3), {"_FillValue": 7}, np.float64),
             "float32": (np.arange(36, dtype=np.float32).reshape(6, 2, 3), {}, np.float32),
             "float32_packed": (np.arange(36, dtype=np.float32).reshape(6, 2, 3), {"scale_factor": 2.0}, np.float64)}
    for name, (vals, attrs, want) in cases.items():
        store = zarrio.write_dataset(str(tmp_path / f"{name}.zarr"), vals, t, lat, lon, var="v", attrs=attrs, compressor=None)
        ds = af.dataset_from_path(store, var="v")
        assert ds.dtype == np.dtype(want), name
