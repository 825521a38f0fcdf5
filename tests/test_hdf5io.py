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
