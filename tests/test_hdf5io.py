"""NetCDF-4 / HDF5 reader (aggfly_b200/hdf5io.py) against files of its own spec-following writer and a hand-assembled
new-style (superblock v2, "OHDR" object header, link messages) root group.  No libhdf5 exists in the image: see the
module's validation note."""
import struct

import numpy as np
import pandas as pd
import pytest

import aggfly_b200 as af
from aggfly_b200 import hdf5io


def _cube(T=53, Y=7, X=12, seed=0):
    rng = np.random.default_rng(seed)
    return (280 + 10 * rng.normal(size=(T, Y, X))).astype(np.float32)


def _axes(T, Y, X):
    return np.arange(T, dtype=np.int32) + 24 * 11, np.linspace(49.75, 48.25, Y), np.linspace(235.0, 237.75, X)


@pytest.mark.parametrize("layout", ["chunked_deflate", "chunked_shuffle_only", "chunked_plain", "contiguous"])
def test_float_variable_round_trips(tmp_path, layout):
    cube = _cube()
    t, lat, lon = _axes(*cube.shape)
    chunks = None if layout == "contiguous" else (24, 4, 5)
    path = str(tmp_path / "t2m.nc")
    hdf5io.write_netcdf4(path, cube, t, "hours since 2001-01-01 00:00:00", lat, lon, var="t2m", chunks=chunks,
                         deflate=4 if layout == "chunked_deflate" else None, shuffle=layout in ("chunked_deflate", "chunked_shuffle_only"),
                         attrs={"units": "K", "long_name": "2 metre temperature"})
    assert hdf5io.looks_like_hdf5(path)
    f = hdf5io.Hdf5File(path)
    assert f.names() == ["latitude", "longitude", "t2m", "time"]
    a = f["t2m"]
    assert a.shape == cube.shape and a.dtype == np.dtype("<f4") and a.attrs["units"] == "K"
    assert a.attrs["long_name"] == "2 metre temperature"
    if chunks:
        assert a.chunks == chunks and [fid for fid, _ in a.filters] == {"chunked_deflate": [2, 1], "chunked_shuffle_only": [2],
                                                                          "chunked_plain": []}[layout]
    assert np.array_equal(a.read(), cube)
    assert np.array_equal(a[5:30, 2:6, 3], cube[5:30, 2:6, 3])
    assert f["time"].attrs["units"].startswith("hours since") and np.array_equal(f["time"].read(), t)
    raster, time, la, lo = hdf5io.open_raster(path, None)
    assert raster.shape == cube.shape and raster.dtype == np.float32
    assert np.array_equal(np.asarray(raster), cube) and np.array_equal(np.asarray(raster[24:48]), cube[24:48])
    assert time[0] == pd.Timestamp("2001-01-12") and len(time) == cube.shape[0]
    assert np.allclose(la, lat) and np.allclose(lo, lon)


def test_packed_int16_with_fill_and_a_missing_chunk_in_lat_lon_time_order(tmp_path):
    cube = _cube(48, 6, 10, seed=3)
    t, lat, lon = _axes(*cube.shape)
    scale, offset = 0.01, 280.0
    q = np.clip(np.rint((cube.astype(np.float64) - offset) / scale), -32000, 32000).astype(np.int16)
    q[3:6, 1, 2] = -32767
    stored = np.ascontiguousarray(np.transpose(q, (1, 2, 0)))                      # (latitude, longitude, time) like zarr_convert
    path = str(tmp_path / "packed.nc")
    hdf5io.write_netcdf4(path, stored, t, "hours since 2001-01-01", lat, lon, var="t2m", dims=("latitude", "longitude", "time"),
                         chunks=(4, 4, 48), attrs={"scale_factor": scale, "add_offset": offset, "_FillValue": np.int16(-32767)},
                         skip_chunks=[(1, 2, 0)])
    f = hdf5io.Hdf5File(path)
    a = f["t2m"]
    assert a.dtype == np.dtype("<i2") and a.fill_value == -32767 and a.attrs["scale_factor"] == scale
    want = stored.copy()
    want[4:8, 8:12, :] = -32767                                                    # the unwritten chunk reads as fill
    assert np.array_equal(a.read(), want)
    # the raster view: axes found by name / length, decoded like xarray's mask_and_scale (fill -> NaN)
    raster, time, la, lo = hdf5io.open_raster(path, "t2m")
    assert raster.axes == (2, 0, 1) and raster.shape == cube.shape and raster.dtype == np.float64
    dec = np.transpose(want, (2, 0, 1)).astype(np.float64) * scale + offset
    dec[np.transpose(want, (2, 0, 1)) == -32767] = np.nan
    assert np.array_equal(np.asarray(raster), dec, equal_nan=True)
    assert len(time) == 48 and time[0] == pd.Timestamp("2001-01-12")


def test_dataset_from_path_reads_netcdf4(tmp_path):
    cube = _cube(24 * 3, 6, 8, seed=5)
    t, lat, lon = _axes(*cube.shape)
    t = np.arange(cube.shape[0], dtype=np.int32)
    path = str(tmp_path / "era5_2001.nc")
    q = np.clip(np.rint((cube.astype(np.float64) - 280.0) / 0.005), -32000, 32000).astype(np.int16)
    hdf5io.write_netcdf4(path, q, t, "hours since 2001-03-01 00:00:00", lat, lon, var="t2m", chunks=(24, 6, 8),
                         attrs={"scale_factor": 0.005, "add_offset": 280.0, "_FillValue": np.int16(-32767)})
    ds = af.dataset_from_path(path, var="t2m", preprocess="kelvin_to_celsius")
    assert getattr(ds.values, "is_chunked_raster", False) and ds.shape == cube.shape
    assert ds.time[0] == pd.Timestamp("2001-03-01") and ds.time[-1] == pd.Timestamp("2001-03-03 23:00")
    assert np.array_equal(np.asarray(ds.values), q.astype(np.float64) * 0.005 + 280.0)
    sub = af.dataset_from_path(path, var="t2m", time_sel="2001-03-02")
    assert sub.shape[0] == 24 and np.array_equal(np.asarray(sub.values), (q.astype(np.float64) * 0.005 + 280.0)[24:48])


def test_new_style_root_group_with_link_messages_and_v2_headers(tmp_path):
    """Superblock v2 + an "OHDR" root object header holding compact link messages (what libhdf5 writes for files created
    with link creation order tracked, as netcdf-c does), assembled by hand over the datasets of the fixture writer."""
    cube = _cube(30, 5, 6, seed=9)
    t, lat, lon = _axes(*cube.shape)
    path = str(tmp_path / "v2.nc")
    hdf5io.write_netcdf4(path, cube, t, "hours since 2001-01-01", lat, lon, var="t2m", chunks=(10, 5, 6))
    old = hdf5io.Hdf5File(path)
    links = dict(old.links)
    old.buf.mm.close()
    with open(path, "r+b") as fh:
        fh.seek(0, 2)
        end = (fh.tell() + 7) & ~7
        body = b""
        for k, (name, addr) in enumerate(sorted(links.items())):
            nm = name.encode()
            data = struct.pack("<BB", 1, 0x04) + struct.pack("<Q", k) + struct.pack("<B", len(nm)) + nm + struct.pack("<Q", addr)
            body += struct.pack("<BHB", 6, len(data), 0) + struct.pack("<H", k) + data
        # link info message: no fractal heap (compact storage)
        li = struct.pack("<BB", 0, 0) + struct.pack("<QQ", hdf5io.UNDEF, hdf5io.UNDEF)
        body = struct.pack("<BHB", 2, len(li), 0) + struct.pack("<H", 0) + li + body
        ohdr = b"OHDR" + struct.pack("<BB", 2, 0x04 | 0x01) + struct.pack("<H", len(body)) + body + bytes(4)     # checksum not verified
        fh.seek(end)
        fh.write(ohdr)
        eof = end + len(ohdr)
        sb = hdf5io.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, hdf5io.UNDEF, eof, end) + bytes(4)
        fh.seek(0)
        fh.write(sb)
    new = hdf5io.Hdf5File(path)
    assert new.names() == ["latitude", "longitude", "t2m", "time"]
    assert np.array_equal(new["t2m"].read(), cube)


def test_unsupported_structures_are_named(tmp_path):
    p = tmp_path / "not.nc"
    p.write_bytes(b"CDF\x01" + bytes(100))
    assert not hdf5io.looks_like_hdf5(str(p))
    with pytest.raises(IOError, match="not an HDF5 file"):
        hdf5io.Hdf5File(str(p))


def _v2_header(messages, flags=0x04 | 0x20 | 0x10 | 0x01):
    """A version-2 object header ("OHDR") as the specification lays it out: optional timestamps (bit 5), attribute phase
    change values (bit 4), creation order per message (bit 2), 2-byte chunk size (bits 0-1 = 1); checksum left zero."""
    body = b""
    for k, (mtype, data) in enumerate(messages):
        body += struct.pack("<BHB", mtype, len(data), 0) + (struct.pack("<H", k) if flags & 0x04 else b"") + data
    body += bytes(5)                                                    # a gap the parser must not read as a message
    head = b"OHDR" + struct.pack("<BB", 2, flags)
    if flags & 0x20:
        head += struct.pack("<IIII", 1, 2, 3, 4)
    if flags & 0x10:
        head += struct.pack("<HH", 8, 6)
    head += struct.pack("<H", len(body))
    return head + body + bytes(4)


def test_new_style_dataset_headers_filters_attributes_and_vlen_strings(tmp_path):
    """What libhdf5 >= 1.8 writes for a NetCDF-4 variable, assembled by hand: "OHDR" headers with timestamps and creation
    order, dataspace v2, fill value v3, filter pipeline v2 (no names for predefined filters), attribute v3 -- one of them a
    variable-length string kept in a global heap collection -- over the v1 B-tree chunk index of the fixture writer."""
    import zlib
    rng = np.random.default_rng(11)
    T, Y, X = 50, 6, 8
    q = rng.integers(-2000, 2000, (T, Y, X)).astype(np.int16)
    chunks = (24, 3, 8)
    path = str(tmp_path / "new_style.nc")
    out = hdf5io._Out(path)
    # chunks: shuffle + deflate, the last time chunk partial
    entries = []
    for idx in np.ndindex(*[-(-s // c) for s, c in zip(q.shape, chunks)]):
        block = np.full(chunks, -32767, "<i2")
        sl = tuple(slice(i * c, min(s, (i + 1) * c)) for i, c, s in zip(idx, chunks, q.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = q[sl]
        raw = zlib.compress(np.frombuffer(block.tobytes(), np.uint8).reshape(-1, 2).T.tobytes(), 5)
        entries.append((tuple(i * c for i, c in zip(idx, chunks)), len(raw), out.put(raw)))
    btree = hdf5io._write_chunk_btree(out, entries, q.shape, chunks)
    # a global heap collection holding the string of a variable-length attribute
    text = b"2 metre temperature"
    gcol_body = struct.pack("<HHIQ", 1, 1, 0, len(text)) + text + bytes(hdf5io._pad8(len(text)) - len(text)) + struct.pack("<HHIQ", 0, 0, 0, 0)
    gcol = out.put(b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(gcol_body)) + gcol_body)

    def attr_v3(name, dt, space, data):
        nm = name.encode() + b"\0"
        return struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(space), 0) + nm + dt + space + data

    scalar = struct.pack("<BBBB", 2, 0, 0, 0)                           # dataspace v2, scalar
    f64 = hdf5io._dt_message(np.dtype("<f8"))
    vlen_str = struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16) + hdf5io._dt_message(np.dtype("S1"))
    msgs = [
        (0x01, struct.pack("<BBBB", 2, 3, 0, 1) + b"".join(struct.pack("<Q", s) for s in q.shape)),      # dataspace v2, simple
        (0x03, hdf5io._dt_message(np.dtype("<i2"))),
        (0x05, struct.pack("<BBI", 3, 0x20 | 0x09, 2) + np.int16(-32767).tobytes()),                       # fill value v3, defined
        (0x08, struct.pack("<BBBQ", 3, 2, 4, btree) + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", 2)),
        (0x0B, struct.pack("<BB", 2, 2) + struct.pack("<HHHI", 2, 0, 1, 2) + struct.pack("<HHHI", 1, 0, 1, 5)),   # shuffle(2), deflate(5)
        (0x0C, attr_v3("scale_factor", f64, scalar, struct.pack("<d", 0.01))),
        (0x0C, attr_v3("add_offset", f64, scalar, struct.pack("<d", 273.15))),
        (0x0C, attr_v3("_FillValue", hdf5io._dt_message(np.dtype("<i2")), scalar, np.int16(-32767).tobytes())),
        (0x0C, attr_v3("long_name", vlen_str, scalar, struct.pack("<IQI", len(text), gcol, 1))),
    ]
    var_addr = out.put(_v2_header(msgs))
    addrs = {"t2m": var_addr}
    coords = {"time": np.arange(T, dtype="<i4"), "latitude": np.linspace(50, 49, Y), "longitude": np.linspace(10, 12, X)}
    for name, vals in coords.items():
        vals = np.ascontiguousarray(vals)
        data = out.put(vals.tobytes())
        m = [(0x01, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", len(vals))), (0x03, hdf5io._dt_message(vals.dtype)),
             (0x05, struct.pack("<BB", 3, 0x10 | 0x05)),                                                    # fill value v3, undefined
             (0x08, struct.pack("<BBQQ", 3, 1, data, vals.nbytes))]
        if name == "time":
            units = b"hours since 2001-01-01 00:00:00\0"
            m.append((0x0C, attr_v3("units", hdf5io._dt_message(np.dtype(f"S{len(units)}")), scalar, units)))
        addrs[name] = out.put(_v2_header(m, flags=0x01))                                                    # plain flags on these
    links = []
    for k, (name, addr) in enumerate(sorted(addrs.items())):
        nm = name.encode()
        links.append((0x06, struct.pack("<BB", 1, 0x04 | 0x08 | 0x10) + struct.pack("<B", 0) + struct.pack("<Q", k) + struct.pack("<B", 1)
                      + struct.pack("<B", len(nm)) + nm + struct.pack("<Q", addr)))
    root = out.put(_v2_header([(0x02, struct.pack("<BBQ", 0, 0x01, 3) + struct.pack("<QQ", hdf5io.UNDEF, hdf5io.UNDEF))] + links))
    eof = out.f.tell()
    out.f.seek(0)
    out.f.write(hdf5io.SIGNATURE + struct.pack("<BBBB", 3, 8, 8, 0) + struct.pack("<QQQQ", 0, hdf5io.UNDEF, eof, root) + bytes(4))
    out.f.close()

    f = hdf5io.Hdf5File(path)
    assert f.names() == ["latitude", "longitude", "t2m", "time"]
    a = f["t2m"]
    assert a.shape == q.shape and a.chunks == chunks and a.dtype == np.dtype("<i2") and a.fill_value == -32767
    assert [fid for fid, _ in a.filters] == [2, 1]
    assert a.attrs["scale_factor"] == 0.01 and a.attrs["add_offset"] == 273.15 and a.attrs["_FillValue"] == -32767
    assert a.attrs["long_name"] == "2 metre temperature"
    assert np.array_equal(a.read(), q)
    raster, time, lat, lon = hdf5io.open_raster(path, "t2m")
    assert raster.packed and raster.fill == -32767.0 and len(time) == T and time[1] == pd.Timestamp("2001-01-01 01:00")
    assert np.array_equal(np.asarray(raster), q.astype(np.float64) * 0.01 + 273.15)
