"""Zarr stores without zarr-python / xarray: metadata, chunk decoding, CF coordinates, a small writer.

The reference opens rasters with ``xr.open_dataset(path, engine="zarr", chunks=...)``
(aggfly/dataset/dataset.py:585-615, 697-707) and recommends converting everything to a
*time-contiguous* store first (``dataset_to_zarr``, aggfly/dataset/zarr_convert.py:50-121: dims
``(latitude, longitude, time)``, chunks ``[s, s, T]``).  Neither zarr nor xarray exists in this
image, and the engine wants storage chunks, not a dask graph: each chunk is decoded by a host thread
straight into a pinned staging slot, copied to the device as stored, and placed into the time-major
raster by ``agf_tile_place_run`` (``stream.feed_chunked``), so no transposed or concatenated host
copy of the raster is ever made.

Supported (directory stores on a local filesystem):

===========  =====================================================================================
zarr v2      ``.zgroup`` / ``.zarray`` / ``.zattrs``; ``order`` C / F; ``dimension_separator`` . or /;
             compressors ``null, zlib, gzip, zstd, lz4, bz2, lzma, blosc``; filters ``shuffle, delta``
zarr v3      ``zarr.json``; codecs ``transpose, bytes, gzip, zstd, blosc, crc32c`` (checksum not
             verified); chunk key encodings ``default`` and ``v2``; ``sharding_indexed`` is refused
CF decoding  ``units = "<unit> since <origin>"`` time axes on the standard, noleap and 360_day
             calendars, ``_FillValue`` / ``missing_value`` -> NaN, ``scale_factor`` / ``add_offset``
===========  =====================================================================================

zstd / lz4 come from pyarrow's codecs (a dependency of the panel writer already), the rest from the
standard library.  The Blosc container is decoded here from its published frame layout (16-byte
header, block offsets, per-split streams, byte / bit shuffle) with the inner codecs above; ``blosclz``
is not implemented.  No Blosc encoder exists in this image, so that decoder is checked against frames
assembled by the tests from the same layout -- not against c-blosc output.
"""
from __future__ import annotations

import json
import os
import struct
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_ZARR_MARKERS = ("zarr.json", ".zmetadata", ".zgroup", ".zarray")          # dataset.py:585-586


def looks_like_zarr(path: str) -> bool:
    """aggfly/dataset/dataset.py:589-615 for local paths: the name, else zarr's root metadata files."""
    if not isinstance(path, str):
        return False
    if ".zarr" in path.lower().rstrip("/"):
        return True
    return os.path.isdir(path) and any(os.path.exists(os.path.join(path, m)) for m in _ZARR_MARKERS)


# ---------------------------------------------------------------------------------------------
# bytes -> bytes codecs
# ---------------------------------------------------------------------------------------------
def _arrow(name: str):
    import pyarrow as pa
    return pa.Codec(name)


def _zstd(buf, nbytes: int):
    return memoryview(_arrow("zstd").decompress(buf, decompressed_size=nbytes))


def _lz4_raw(buf, nbytes: int):
    return memoryview(_arrow("lz4_raw").decompress(buf, decompressed_size=nbytes))


def _zlib_any(buf, nbytes: int):
    return zlib.decompress(buf, 47)                  # zlib or gzip container, detected from the header


def _numcodecs_lz4(buf, nbytes: int):
    n = struct.unpack_from("<I", buf, 0)[0]          # numcodecs.LZ4: uint32 decoded size + one raw block
    return _lz4_raw(memoryview(buf)[4:], n)


def _bz2(buf, nbytes: int):
    import bz2
    return bz2.decompress(buf)


def _lzma(buf, nbytes: int):
    import lzma
    return lzma.decompress(buf)


_BLOSC_INNER = {1: _lz4_raw, 3: _zlib_any, 4: _zstd}          # header flags >> 5 (0 blosclz, 2 snappy)


def _unshuffle(block: np.ndarray, typesize: int) -> np.ndarray:
    """Inverse of Blosc's byte shuffle on one block: byte j of every element is stored together."""
    n = block.size // typesize
    out = np.empty_like(block)
    out[: n * typesize] = block[: n * typesize].reshape(typesize, n).T.reshape(-1)
    out[n * typesize:] = block[n * typesize:]
    return out


def _bitunshuffle(block: np.ndarray, typesize: int) -> np.ndarray:
    """Inverse of Blosc's bit shuffle: bit k of every element is stored together (multiples of 8
    elements; the tail is stored unshuffled)."""
    n = (block.size // typesize) & ~7
    out = block.copy()
    if n:
        planes = np.unpackbits(block[: n * typesize].reshape(typesize * 8, n // 8), axis=1, bitorder="little")
        out[: n * typesize] = np.packbits(planes.T, axis=1, bitorder="little").reshape(-1)
    return out


def blosc_decompress(buf, nbytes_hint: int = 0) -> np.ndarray:
    """One Blosc (v1 frame) buffer -> uint8 array.  Layout: version, versionlz, flags, typesize,
    nbytes, blocksize, cbytes (u8 x 4, u32 x 3 little endian), then int32 block offsets unless the
    frame is a plain copy (flag 0x2)."""
    mv = memoryview(buf).cast("B") if not isinstance(buf, memoryview) else buf.cast("B")
    _, _, flags, typesize = struct.unpack_from("<4B", mv, 0)
    nbytes, blocksize, cbytes = struct.unpack_from("<3I", mv, 4)
    if cbytes > len(mv):
        raise ValueError(f"blosc frame says {cbytes} bytes, buffer has {len(mv)}")
    if flags & 0x2:                                                      # memcpyed
        return np.frombuffer(mv, np.uint8, nbytes, 16).copy()
    comp = flags >> 5
    inner = _BLOSC_INNER.get(comp)
    if inner is None:
        raise NotImplementedError(f"blosc inner codec {['blosclz', 'lz4', 'snappy', 'zlib', 'zstd'][comp] if comp < 5 else comp}")
    out = np.empty(nbytes, np.uint8)
    if nbytes == 0:
        return out
    nblocks = -(-nbytes // blocksize)
    bstarts = struct.unpack_from(f"<{nblocks}i", mv, 16)
    dont_split = bool(flags & 0x10)
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split) and typesize <= 16 and blocksize // typesize >= 128 and not leftover
        nsplits = typesize if split else 1
        ne = bsize // nsplits
        pos = bstarts[b]
        block = out[b * blocksize: b * blocksize + bsize]
        for s in range(nsplits):
            c = struct.unpack_from("<i", mv, pos)[0]
            pos += 4
            raw = mv[pos: pos + c]
            block[s * ne: (s + 1) * ne] = np.frombuffer(raw if c == ne else inner(raw, ne), np.uint8, ne)
            pos += c
        if typesize > 1 and flags & 0x1:
            block[:] = _unshuffle(block, typesize)
        elif flags & 0x4:
            block[:] = _bitunshuffle(block, typesize)
    return out


class BloscPlan:
    """What the device needs to decode one Blosc frame whose bytes sit in a buffer: raw-LZ4 streams for the
    decompression engine (offsets into the frame / into the decoded chunk), streams stored uncompressed (plain
    copies), and the shuffle parameters.  ``kind == "memcpy"``: the frame is a 16-byte header + plain data."""
    __slots__ = ("kind", "nbytes", "typesize", "blocksize", "shuffled", "src_off", "src_len", "dst_off", "dst_len", "raw",
                 "inline")

    def __init__(self):
        self.inline = None          # (offset of int32 dst_len[n_ops], offset of int64 raw[3, n_raw]) inside the staged slot


def blosc_device_plan(u8: np.ndarray, max_len: int) -> Optional[BloscPlan]:
    """Parse a Blosc-1 frame (uint8 view) without touching its payload.  None: not decodable on the device
    (inner codec other than LZ4, bit shuffle, streams above the engine's limit, malformed) -- the caller
    decodes on the host instead.  Vectorised over blocks: a 256 MB chunk has ~1000 blocks x typesize streams."""
    if u8.size < 16:
        return None
    flags, typesize = int(u8[2]), int(u8[3])
    nbytes, blocksize, cbytes = (int(v) for v in np.frombuffer(u8[4:16], "<u4"))
    plan = BloscPlan()
    plan.nbytes, plan.typesize, plan.blocksize = nbytes, typesize, blocksize
    plan.shuffled = bool(flags & 0x1) and typesize > 1
    if cbytes > u8.size or blocksize <= 0 or typesize <= 0:
        return None
    if flags & 0x2:
        plan.kind = "memcpy"
        return plan if 16 + nbytes <= u8.size else None
    if (flags >> 5) != 1 or flags & 0x4 or (plan.shuffled and typesize not in (2, 4, 8)):
        return None
    nblocks = -(-nbytes // blocksize)
    if 16 + 4 * nblocks > cbytes:
        return None
    bstarts = np.frombuffer(u8[16:16 + 4 * nblocks], "<i4").astype(np.int64)
    leftover = nbytes % blocksize
    nfull = nblocks - (1 if leftover else 0)
    split = (not flags & 0x10) and typesize <= 16 and blocksize // typesize >= 128
    nsplits = typesize if split else 1
    if blocksize % nsplits:
        return None
    four = np.arange(4, dtype=np.int64)

    def lengths(pos):
        if (pos < 0).any() or (pos + 4 > cbytes).any():
            return None
        return np.ascontiguousarray(u8[pos[:, None] + four]).view("<i4").ravel().astype(np.int64)

    so, sl, do, dl = [], [], [], []
    if nfull:
        ne = blocksize // nsplits
        pos = bstarts[:nfull].copy()
        base = np.arange(nfull, dtype=np.int64) * blocksize
        for k in range(nsplits):
            c = lengths(pos)
            if c is None or (c <= 0).any() or (pos + 4 + c > cbytes).any():
                return None
            so.append(pos + 4), sl.append(c), do.append(base + k * ne), dl.append(np.full(nfull, ne, np.int64))
            pos = pos + 4 + c
    if leftover:
        pos = bstarts[nfull:nfull + 1]
        c = lengths(pos)
        if c is None or c[0] <= 0 or pos[0] + 4 + c[0] > cbytes:
            return None
        so.append(pos + 4), sl.append(c), do.append(np.array([nfull * blocksize], np.int64)), dl.append(np.array([leftover], np.int64))
    if not so:
        return None
    src_off, src_len, dst_off, dst_len = (np.concatenate(v) for v in (so, sl, do, dl))
    raw = src_len == dst_len                                             # stream stored uncompressed
    if (src_len > dst_len).any() or (dst_len[~raw] > max_len).any():
        return None
    plan.kind = "lz4"
    plan.raw = np.ascontiguousarray(np.stack([src_off[raw], dst_off[raw], dst_len[raw]]))      # int64[3, n_raw]
    keep = ~raw
    plan.src_off, plan.src_len = np.ascontiguousarray(src_off[keep]), np.ascontiguousarray(src_len[keep])
    plan.dst_off, plan.dst_len = np.ascontiguousarray(dst_off[keep]), np.ascontiguousarray(dst_len[keep])
    return plan


def blosc_compress(data: bytes, typesize: int, cname: str = "lz4", shuffle: int = 1, blocksize: int = 0,
                   split: bool = True) -> bytes:
    """A Blosc-1 frame assembled from the published layout (the inverse of ``blosc_decompress``): blocks of
    ``blocksize`` bytes (0: 256 KB), optional byte (1) / bit (2) shuffle, each full block split into
    ``typesize`` streams when ``split``.  For fixtures and for stores this package writes; not c-blosc."""
    comp_id = {"lz4": 1, "zlib": 3, "zstd": 4}[cname]
    enc = {"lz4": lambda b: _arrow("lz4_raw").compress(b, asbytes=True), "zlib": lambda b: zlib.compress(b, 1),
           "zstd": lambda b: _arrow("zstd").compress(b, asbytes=True)}[cname]
    nbytes = len(data)
    blocksize = int(blocksize) or (256 << 10)
    blocksize = max(typesize, min(blocksize, max(nbytes, typesize)) // typesize * typesize)
    nblocks = -(-nbytes // blocksize) if nbytes else 0
    can_split = split and typesize <= 16 and blocksize // typesize >= 128
    flags = (comp_id << 5) | (0 if can_split else 0x10) | {0: 0, 1: 0x1, 2: 0x4}[shuffle]
    arr = np.frombuffer(data, np.uint8)
    body, bstarts, pos = [], [], 16 + 4 * nblocks
    for b in range(nblocks):
        blk = arr[b * blocksize: (b + 1) * blocksize]
        n = blk.size // typesize
        if shuffle == 1 and typesize > 1:
            blk = np.concatenate([blk[: n * typesize].reshape(n, typesize).T.reshape(-1), blk[n * typesize:]])
        elif shuffle == 2:
            n8 = n & ~7
            bits = np.unpackbits(blk[: n8 * typesize].reshape(n8, typesize), axis=1, bitorder="little")
            blk = np.concatenate([np.packbits(bits.T, axis=1, bitorder="little").reshape(-1), blk[n8 * typesize:]])
        nsplits = typesize if (can_split and blk.size == blocksize) else 1
        ne = blk.size // nsplits
        bstarts.append(pos)
        for k in range(nsplits):
            raw = blk[k * ne: (k + 1) * ne].tobytes()
            c = enc(raw)
            if len(c) >= len(raw):
                c = raw                                                         # stored uncompressed
            body.append(struct.pack("<i", len(c)) + c)
            pos += 4 + len(c)
    head = struct.pack("<4B3I", 2, 1, flags, typesize, nbytes, blocksize, pos)
    return head + struct.pack(f"<{nblocks}i", *bstarts) + b"".join(body)


def _blosc(buf, nbytes: int):
    return memoryview(blosc_decompress(buf, nbytes))


_V2_COMPRESSORS = {"zlib": _zlib_any, "gzip": _zlib_any, "zstd": _zstd, "lz4": _numcodecs_lz4, "bz2": _bz2,
                   "lzma": _lzma, "blosc": _blosc}
_V3_BYTES_CODECS = {"gzip": _zlib_any, "zstd": _zstd, "blosc": _blosc}

_V3_DTYPES = {"bool": "?", "int8": "i1", "int16": "i2", "int32": "i4", "int64": "i8", "uint8": "u1", "uint16": "u2",
              "uint32": "u4", "uint64": "u8", "float32": "f4", "float64": "f8"}


def _fill(value, dtype: np.dtype):
    """zarr's JSON spelling of a fill value -> scalar of ``dtype`` (None: no fill value)."""
    if value is None:
        return None
    if isinstance(value, str):
        value = {"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(value, value)
        if isinstance(value, str):
            if dtype.kind in "SUO":
                return value
            value = float(value) if dtype.kind == "f" else int(value, 0)
    return np.asarray(value).astype(dtype)[()]


def _attr_fill(value: str, dtype: np.dtype):
    """A ``_FillValue`` ATTRIBUTE spelled as a string: zarr's "NaN" / "Infinity" / a number, or -- what xarray
    writes into zarr v3 stores for floating-point variables -- base64 of the little-endian float64."""
    try:
        return _fill(value, dtype)
    except ValueError:
        import base64
        raw = base64.standard_b64decode(value)
        if dtype.kind == "f" and len(raw) == 8:
            return struct.unpack("<d", raw)[0]
        if dtype.kind == "f" and len(raw) == 4:
            return struct.unpack("<f", raw)[0]
        raise


# ---------------------------------------------------------------------------------------------
# arrays
# ---------------------------------------------------------------------------------------------
class ZarrArray:
    """One array of a store: metadata + ``read_chunk``.  ``storage_axes[k]`` is the logical axis that
    is k-th slowest in a stored chunk (identity for C order; reversed for F order / a v3 transpose)."""

    def __init__(self, path: str):
        self.path = path
        if os.path.exists(os.path.join(path, "zarr.json")):
            self._init_v3(json.load(open(os.path.join(path, "zarr.json"))))
        elif os.path.exists(os.path.join(path, ".zarray")):
            attrs_path = os.path.join(path, ".zattrs")
            self._init_v2(json.load(open(os.path.join(path, ".zarray"))),
                          json.load(open(attrs_path)) if os.path.exists(attrs_path) else {})
        else:
            raise FileNotFoundError(f"{path}: neither zarr.json nor .zarray")
        self.ndim = len(self.shape)
        self.grid = tuple(-(-s // c) for s, c in zip(self.shape, self.chunks))

    # -- metadata ------------------------------------------------------------------------------
    def _init_v2(self, meta: dict, attrs: dict):
        self.zarr_format = 2
        self.shape = tuple(int(s) for s in meta["shape"])
        self.chunks = tuple(int(c) for c in meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        if self.dtype.kind in "OV":
            raise NotImplementedError(f"{self.path}: dtype {meta['dtype']!r}")
        self.fill_value = _fill(meta.get("fill_value"), self.dtype)
        n = len(self.shape)
        self.storage_axes = tuple(range(n)) if meta.get("order", "C") == "C" else tuple(reversed(range(n)))
        self._sep = meta.get("dimension_separator", ".")
        self._prefix = ""
        self.attrs = dict(attrs)
        self.dims = tuple(attrs.get("_ARRAY_DIMENSIONS", ())) or None
        self._decoders = []                                      # applied in this order to the stored bytes
        comp = meta.get("compressor")
        if comp is not None:
            if comp["id"] not in _V2_COMPRESSORS:
                raise NotImplementedError(f"{self.path}: compressor {comp['id']!r}")
            self._decoders.append(_V2_COMPRESSORS[comp["id"]])
        self._filters = []                                       # decoded last to first
        for f in (meta.get("filters") or []):
            if f["id"] == "shuffle":
                self._filters.append(("shuffle", int(f.get("elementsize", self.dtype.itemsize))))
            elif f["id"] == "delta":
                self._filters.append(("delta", np.dtype(f.get("astype", f["dtype"])), np.dtype(f["dtype"])))
            else:
                raise NotImplementedError(f"{self.path}: filter {f['id']!r}")
        # xarray treats a v2 array's fill_value as the CF _FillValue
        if self.fill_value is not None and "_FillValue" not in self.attrs and self.dtype.kind in "fiu":
            self._cf_fill_default = self.fill_value
        else:
            self._cf_fill_default = None

    def _init_v3(self, meta: dict):
        if meta.get("node_type") != "array":
            raise ValueError(f"{self.path}: zarr.json is a {meta.get('node_type')!r}, not an array")
        self.zarr_format = 3
        self.shape = tuple(int(s) for s in meta["shape"])
        dt = meta["data_type"]
        if not isinstance(dt, str) or dt not in _V3_DTYPES:
            raise NotImplementedError(f"{self.path}: data_type {dt!r}")
        grid = meta["chunk_grid"]
        if grid["name"] != "regular":
            raise NotImplementedError(f"{self.path}: chunk grid {grid['name']!r}")
        self.chunks = tuple(int(c) for c in grid["configuration"]["chunk_shape"])
        enc = meta.get("chunk_key_encoding", {"name": "default"})
        conf = enc.get("configuration") or {}
        if enc["name"] == "default":
            self._sep, self._prefix = conf.get("separator", "/"), "c"
        elif enc["name"] == "v2":
            self._sep, self._prefix = conf.get("separator", "."), ""
        else:
            raise NotImplementedError(f"{self.path}: chunk key encoding {enc['name']!r}")
        n = len(self.shape)
        axes, endian = tuple(range(n)), "<"
        self._decoders, self._filters = [], []
        seen_bytes = False
        for c in meta["codecs"]:
            name, conf = c["name"], c.get("configuration") or {}
            if name == "transpose":
                order = conf["order"]
                order = {"C": list(range(n)), "F": list(reversed(range(n)))}.get(order, order) if isinstance(order, str) else order
                axes = tuple(axes[o] for o in order)
            elif name == "bytes":
                endian, seen_bytes = {"little": "<", "big": ">"}[conf.get("endian", "little")], True
            elif name in _V3_BYTES_CODECS:
                self._decoders.insert(0, _V3_BYTES_CODECS[name])
            elif name == "crc32c":
                self._decoders.insert(0, lambda buf, nbytes: memoryview(buf)[:-4])      # trailing checksum, not verified
            else:
                raise NotImplementedError(f"{self.path}: codec {name!r}")
        if not seen_bytes:
            raise ValueError(f"{self.path}: no array -> bytes codec")
        self.storage_axes = axes
        self.dtype = np.dtype(endian + _V3_DTYPES[dt]) if _V3_DTYPES[dt] not in ("?", "i1", "u1") else np.dtype(_V3_DTYPES[dt])
        self.fill_value = _fill(meta.get("fill_value"), self.dtype)
        self.attrs = dict(meta.get("attributes") or {})
        self.dims = tuple(meta["dimension_names"]) if meta.get("dimension_names") else None
        self._cf_fill_default = None

    # -- chunks --------------------------------------------------------------------------------
    @property
    def storage_shape(self) -> Tuple[int, ...]:
        """Shape of a stored (always full-size) chunk, slowest axis first."""
        return tuple(self.chunks[a] for a in self.storage_axes)

    @property
    def chunk_nbytes(self) -> int:
        return int(np.prod(self.chunks)) * self.dtype.itemsize

    def chunk_path(self, idx: Sequence[int]) -> str:
        key = self._sep.join(str(int(i)) for i in idx) if len(idx) else ("0" if self._prefix == "" else "")
        if self._prefix:
            key = self._prefix + (self._sep + key if key else "")
        return os.path.join(self.path, *key.split("/"))

    def read_chunk_bytes(self, idx: Sequence[int], into: Optional[np.ndarray] = None):
        """The chunk file as stored: ``bytes``, or -- when ``into`` (uint8 array, e.g. a pinned slot) is
        large enough -- the number of bytes read straight into it.  None: no such file (all fill value)."""
        p = self.chunk_path(idx)
        try:
            with open(p, "rb") as f:
                if into is not None:
                    size = os.fstat(f.fileno()).st_size
                    if size <= into.size:
                        got = f.readinto(memoryview(into)[:size])
                        if got != size:
                            raise IOError(f"{p}: short read ({got} of {size} bytes)")
                        return size
                return f.read()
        except FileNotFoundError:
            return None

    def decode_chunk(self, buf, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Stored bytes -> C-contiguous array of ``storage_shape`` (native byte order), written into ``out``
        (flat, same dtype, e.g. a pinned staging slot) when given."""
        n = self.chunk_nbytes
        for dec in self._decoders:
            buf = dec(buf, n)
        for flt in reversed(self._filters):
            if flt[0] == "shuffle":
                a = np.frombuffer(buf, np.uint8)
                buf = memoryview(np.ascontiguousarray(a[: (a.size // flt[1]) * flt[1]].reshape(flt[1], -1).T).reshape(-1))
            else:                                                        # delta
                buf = memoryview(np.cumsum(np.frombuffer(buf, flt[1]), dtype=flt[2]))
        arr = np.frombuffer(buf, self.dtype, int(np.prod(self.chunks)))
        native = self.dtype.newbyteorder("=")
        if out is None:
            out = np.empty(arr.size, native)
        dst = out[: arr.size]
        np.copyto(dst, arr, casting="unsafe" if not self.dtype.isnative else "same_kind")   # byteswaps if needed
        return dst.reshape(self.storage_shape)

    def read_chunk_storage(self, idx: Sequence[int], out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """Decoded chunk as stored, or None when the chunk file does not exist (all fill value)."""
        buf = self.read_chunk_bytes(idx)
        return None if buf is None else self.decode_chunk(buf, out)

    @property
    def blosc_only(self) -> bool:
        """The stored bytes of a chunk are exactly one Blosc frame of native little-endian values: the
        frame's LZ4 streams can be inflated on the device (``blosc_device_plan``)."""
        return (len(self._decoders) == 1 and self._decoders[0] is _blosc and not self._filters
                and (self.dtype.byteorder in "<|" or (self.dtype.byteorder == "=" and np.little_endian)))

    def read_chunk(self, idx: Sequence[int]) -> np.ndarray:
        """Chunk in logical axis order (a transposed view of the stored chunk), padded to ``chunks``."""
        st = self.read_chunk_storage(idx)
        if st is None:
            fv = self.fill_value if self.fill_value is not None else 0
            return np.full(self.chunks, fv, self.dtype.newbyteorder("="))
        return np.transpose(st, np.argsort(self.storage_axes))

    def __getitem__(self, key) -> np.ndarray:
        """Basic (step-1 slices / integers) region read to a NumPy array."""
        if not isinstance(key, tuple):
            key = (key,)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        lo, hi, squeeze = [], [], []
        for k, n in zip(key, self.shape):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (n if k < 0 else 0)
                if not 0 <= k < n:
                    raise IndexError(f"index {k} out of range for axis of length {n}")
                lo.append(k), hi.append(k + 1), squeeze.append(True)
            else:
                a, b, step = k.indices(n)
                if step != 1:
                    raise IndexError("ZarrArray supports step-1 slices")
                lo.append(a), hi.append(max(a, b)), squeeze.append(False)
        out = np.empty([h - l for l, h in zip(lo, hi)], self.dtype.newbyteorder("="))
        if out.size:
            ranges = [range(l // c, (h - 1) // c + 1) for l, h, c in zip(lo, hi, self.chunks)]
            for idx in np.ndindex(*[len(r) for r in ranges]):
                cidx = [r[i] for r, i in zip(ranges, idx)]
                chunk = self.read_chunk(cidx)
                src, dst = [], []
                for ci, c, l, h in zip(cidx, self.chunks, lo, hi):
                    a, b = max(l, ci * c), min(h, (ci + 1) * c)
                    src.append(slice(a - ci * c, b - ci * c)), dst.append(slice(a - l, b - l))
                out[tuple(dst)] = chunk[tuple(src)]
        return out[tuple(0 if s else slice(None) for s in squeeze)]

    def read(self) -> np.ndarray:
        return self[()] if self.ndim == 0 else self[(slice(None),) * self.ndim]

    # -- CF -----------------------------------------------------------------------------------
    def cf_packing(self) -> Tuple[Optional[float], float, float]:
        """(fill value or None, scale, offset) of xarray's ``decode_cf`` for this variable."""
        fv = self.attrs.get("_FillValue", self.attrs.get("missing_value", self._cf_fill_default))
        if isinstance(fv, list):
            fv = fv[0] if fv else None
        if isinstance(fv, str):
            fv = _attr_fill(fv, self.dtype)
        fv = None if fv is None else float(fv)
        if fv is not None and np.isnan(fv):
            fv = None                                                     # NaN stays NaN by itself
        return fv, float(self.attrs.get("scale_factor", 1.0)), float(self.attrs.get("add_offset", 0.0))

    def __repr__(self):
        return (f"<ZarrArray v{self.zarr_format} {os.path.basename(self.path)!r} shape={self.shape} chunks={self.chunks} "
                f"dtype={self.dtype} dims={self.dims}>")


class ZarrGroup:
    """A (flat) group: the arrays found one level below ``path``."""

    def __init__(self, path: str):
        self.path = path.rstrip("/")
        if not os.path.isdir(self.path):
            raise FileNotFoundError(self.path)
        self.attrs: Dict = {}
        root = os.path.join(self.path, "zarr.json")
        if os.path.exists(root):
            self.attrs = dict(json.load(open(root)).get("attributes") or {})
        elif os.path.exists(os.path.join(self.path, ".zattrs")):
            self.attrs = json.load(open(os.path.join(self.path, ".zattrs")))
        self._arrays: Dict[str, ZarrArray] = {}

    def names(self) -> List[str]:
        return sorted(d for d in os.listdir(self.path)
                      if os.path.exists(os.path.join(self.path, d, ".zarray")) or self._is_v3_array(d))

    def _is_v3_array(self, d: str) -> bool:
        p = os.path.join(self.path, d, "zarr.json")
        if not os.path.exists(p):
            return False
        try:
            return json.load(open(p)).get("node_type") == "array"
        except Exception:
            return False

    def __contains__(self, name: str) -> bool:
        return name in self.names()

    def __getitem__(self, name: str) -> ZarrArray:
        if name not in self._arrays:
            p = os.path.join(self.path, name)
            if not os.path.isdir(p):
                raise KeyError(f"{self.path}: no array {name!r} (have {self.names()})")
            self._arrays[name] = ZarrArray(p)
        return self._arrays[name]


# ---------------------------------------------------------------------------------------------
# the lazy raster the engine streams from
# ---------------------------------------------------------------------------------------------
class Tile:
    """One storage chunk's share of a raster view: the block ``[t0:t1, y0:y1, x0:x1]`` of the VIEW comes
    from the chunk ``index`` at element ``offset`` with element strides ``(st, sy, sx)``."""
    __slots__ = ("index", "t0", "t1", "y0", "y1", "x0", "x1", "offset", "st", "sy", "sx")

    def __init__(self, index, t0, t1, y0, y1, x0, x1, offset, st, sy, sx):
        self.index, self.t0, self.t1, self.y0, self.y1, self.x0, self.x1 = index, t0, t1, y0, y1, x0, x1
        self.offset, self.st, self.sy, self.sx = offset, st, sy, sx

    @property
    def extent(self):
        return self.t1 - self.t0, self.y1 - self.y0, self.x1 - self.x0


class ChunkedRaster:
    """Lazy ``values[time, lat, lon]`` over a chunked array whose dims may be in any order.  NumPy sees it
    through ``__array__`` (decoded on the host, for small uses); the engine asks for ``tiles()`` and
    ``load(tile, out)`` and does the axis permutation and CF decoding on the device."""

    is_chunked_raster = True

    def __init__(self, array: ZarrArray, axes: Tuple[int, int, int], window=None, decode: bool = True):
        if array.ndim != 3:
            raise ValueError(f"{array!r}: the raster variable must have exactly (time, lat, lon) dimensions")
        self.array, self.axes = array, tuple(int(a) for a in axes)            # positions of time / lat / lon in array.shape
        full = tuple((0, array.shape[a]) for a in self.axes)
        self.window = tuple(window) if window is not None else full            # ((t_lo, t_hi), (y_lo, y_hi), (x_lo, x_hi))
        fv, scale, offset = array.cf_packing() if decode else (None, 1.0, 0.0)
        self.fill, self.scale, self.offset = fv, scale, offset
        self.packed = scale != 1.0 or offset != 0.0
        src = array.dtype
        if src.kind == "f" and src.itemsize in (4, 8) and not self.packed:
            self.dtype = np.dtype(src.newbyteorder("="))
        elif src.kind == "f" and src.itemsize == 2:
            raise NotImplementedError("float16 rasters")
        elif src.kind in "iu" and src.itemsize > 4:
            raise NotImplementedError(f"{src} rasters (no 64-bit integer tile decoder)")
        else:
            self.dtype = np.dtype(np.float64)                                    # packed / integer -> float64 like xarray

    @property
    def shape(self):
        return tuple(h - l for l, h in self.window)

    ndim = 3

    def __len__(self):
        return self.shape[0]

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    @property
    def single_time_chunk(self) -> bool:
        """Every tile spans the view's whole time range (time-contiguous stores): no row is complete before the
        last tile has landed, so there is nothing to gain from time stripes."""
        (lo, hi), c = self.window[0], self.array.chunks[self.axes[0]]
        return hi > lo and lo // c == (hi - 1) // c

    @property
    def nbytes_stored(self) -> int:
        return int(np.prod(self.shape)) * self.array.dtype.itemsize

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) <= 3 and all(isinstance(k, slice) and k.step in (None, 1) for k in key):
            win = list(self.window)
            for i, k in enumerate(key):
                a, b, _ = k.indices(win[i][1] - win[i][0])
                win[i] = (win[i][0] + a, win[i][0] + max(a, b))
            new = ChunkedRaster.__new__(ChunkedRaster)
            new.__dict__.update(self.__dict__)
            new.window = tuple(win)
            return new
        return np.asarray(self)[key]

    def decode_host(self, block: np.ndarray) -> np.ndarray:
        """The decoding ``agf_tile_place_run`` does on the device, in NumPy (host materialisation only)."""
        out = block.astype(self.dtype, copy=True)
        if self.packed:
            out = (block.astype(np.float64) * self.scale + self.offset).astype(self.dtype)
        if self.fill is not None:
            out[block.astype(np.float64) == self.fill] = np.nan
        return out

    def __array__(self, dtype=None, copy=None):
        key = [slice(None)] * 3
        for pos, (l, h) in zip(self.axes, self.window):
            key[pos] = slice(l, h)
        raw = np.transpose(self.array[tuple(key)], self.axes)
        out = self.decode_host(raw)
        return out if dtype is None else out.astype(dtype, copy=False)

    # -- engine side ----------------------------------------------------------------------------
    def tiles(self) -> List[Tile]:
        """Tiles of the view, ordered by time chunk first (so a time-major store completes rows early)."""
        arr = self.array
        stor_strides = {}
        stride = 1
        for a in reversed(arr.storage_axes):
            stor_strides[a] = stride
            stride *= arr.chunks[a]
        ranges = []
        for pos, (l, h) in zip(self.axes, self.window):
            c = arr.chunks[pos]
            ranges.append(range(l // c, (h - 1) // c + 1) if h > l else range(0))
        out = []
        for it in ranges[0]:
            for iy in ranges[1]:
                for ix in ranges[2]:
                    idx = [0, 0, 0]
                    ext, off = [], 0
                    for pos, ci, (l, h) in zip(self.axes, (it, iy, ix), self.window):
                        c = arr.chunks[pos]
                        idx[pos] = ci
                        a, b = max(l, ci * c), min(h, (ci + 1) * c)
                        ext.append((a - l, b - l))
                        off += (a - ci * c) * stor_strides[pos]
                    out.append(Tile(tuple(idx), ext[0][0], ext[0][1], ext[1][0], ext[1][1], ext[2][0], ext[2][1], off,
                                    stor_strides[self.axes[0]], stor_strides[self.axes[1]], stor_strides[self.axes[2]]))
        return out

    @property
    def slot_elems(self) -> int:
        return int(np.prod(self.array.chunks))

    def load_stored(self, tile: Tile, slot_u8: np.ndarray, max_len: int, inline_tables: bool = False):
        """Device-decode variant of ``load``: read the chunk FILE into ``slot_u8`` (pinned bytes) and plan its
        decoding on the device.  Returns None (chunk absent), ("de", n_bytes, BloscPlan), or ("host",) after
        decoding on the host into the same slot (frames the engine cannot take).  ``inline_tables``: the plan's
        small device-side tables (expected stream lengths, raw-segment table) are appended to the slot behind the
        frame (``plan.inline``), so that they travel with the chunk's one host-to-device copy."""
        arr = self.array
        got = arr.read_chunk_bytes(tile.index, into=slot_u8)
        if got is None:
            return None
        sdt = arr.dtype.newbyteorder("=")
        if isinstance(got, int):
            plan = blosc_device_plan(slot_u8[:got], max_len)
            if plan is not None and plan.nbytes == arr.chunk_nbytes and plan.typesize in (1, sdt.itemsize):
                if inline_tables and plan.kind == "lz4":
                    n_ops, n_raw = len(plan.src_off), plan.raw.shape[1]
                    a = (got + 15) // 16 * 16
                    b = (a + 4 * n_ops + 15) // 16 * 16
                    if b + 24 * n_raw <= slot_u8.size:
                        slot_u8[a:a + 4 * n_ops].view(np.int32)[:] = plan.dst_len
                        slot_u8[b:b + 24 * n_raw].view(np.int64)[:] = plan.raw.ravel()
                        plan.inline = (a, b)
                        got = b + 24 * n_raw
                return ("de", got, plan)
            got = bytes(slot_u8[:got])                                       # the decoder writes into the same slot
        arr.decode_chunk(got, slot_u8[: arr.chunk_nbytes].view(sdt))
        return ("host",)

    def load(self, tile: Tile, out: np.ndarray) -> bool:
        """Decode the tile's chunk into ``out`` (flat array of the STORED dtype, >= slot_elems).  False: the
        chunk is absent from the store (every value is the fill value)."""
        return self.array.read_chunk_storage(tile.index, out) is not None

    def direct_rows(self, tile: Tile) -> bool:
        """True when the tile is a run of whole raster rows that needs no decoding: it can be copied into
        the raster without the placement kernel."""
        nt, ny, nx = tile.extent
        T, Y, X = self.shape
        return (not self.packed and self.fill is None and self.array.dtype.newbyteorder("=") == self.dtype
                and ny == Y and nx == X and tile.sx == 1 and tile.sy == X and tile.st == Y * X)

    def __repr__(self):
        return f"<ChunkedRaster {self.shape} {self.dtype} over {self.array!r}>"


# ---------------------------------------------------------------------------------------------
# CF coordinates
# ---------------------------------------------------------------------------------------------
def decode_time(arr: ZarrArray):
    """CF time coordinate -> DatetimeIndex, or CalendarIndex on the noleap / 360_day calendars."""
    import pandas as pd
    from .timeaxis import CalendarIndex
    vals = arr.read()
    if vals.dtype.kind == "M":
        return pd.DatetimeIndex(vals.astype("datetime64[ns]"))
    units = arr.attrs.get("units")
    if units is None or " since " not in units:
        raise ValueError(f"{arr.path}: time coordinate without CF units (got {units!r})")
    calendar = str(arr.attrs.get("calendar", "standard")).lower()
    unit, _, origin = units.partition(" since ")
    unit = unit.strip().lower().rstrip("s")
    per_hour = {"second": 1 / 3600.0, "minute": 1 / 60.0, "hour": 1.0, "day": 24.0}.get(unit)
    if per_hour is None:
        raise ValueError(f"unsupported CF time unit {units!r}")
    if calendar in ("standard", "gregorian", "proleptic_gregorian"):
        o = pd.Timestamp(origin.strip())
        if o.tzinfo is not None:
            o = o.tz_convert(None)
        if vals.dtype.kind in "iu":
            ns = np.asarray(vals, np.int64) * np.int64(round(per_hour * 3600e9))
        else:
            ns = np.round(np.asarray(vals, np.float64) * (per_hour * 3600e9)).astype(np.int64)
        return pd.DatetimeIndex(o.value + ns)
    if calendar in ("noleap", "365_day", "360_day"):
        tmp = CalendarIndex(calendar, [1], [1], [1])
        ymd = origin.strip().replace("T", " ").split(" ")
        y, m, d = (int(v) for v in ymd[0].split("-"))
        h0 = int(ymd[1].split(":")[0]) if len(ymd) > 1 and ymd[1] else 0
        cum = np.concatenate([[0], np.cumsum(tmp.month_lengths)])
        origin_h = (y * tmp.year_length + cum[m - 1] + (d - 1)) * 24 + h0
        hours = origin_h + np.round(np.asarray(vals, np.float64) * per_hour).astype(np.int64)
        days, hour = hours // 24, hours % 24
        year, doy = days // tmp.year_length, days % tmp.year_length
        month = np.searchsorted(cum, doy, side="right")
        return CalendarIndex(calendar, year, month, doy - cum[month - 1] + 1, hour)
    raise NotImplementedError(f"calendar {calendar!r}")


def open_raster(path: str, var: Optional[str], xycoords=("longitude", "latitude"), timecoord: str = "time"):
    """(ChunkedRaster, time, latitude, longitude) of variable ``var`` of the store at ``path``."""
    g = ZarrGroup(path)
    names = g.names()
    xdim, ydim = xycoords
    if var is None:
        def _ndim(n):
            try:
                return g[n].ndim
            except NotImplementedError:                      # e.g. a string-typed auxiliary variable
                return -1
        cands = [n for n in names if n not in (xdim, ydim, timecoord) and _ndim(n) == 3]
        if len(cands) != 1:
            raise KeyError(f"{path}: pass var= (3-D arrays: {cands})")
        var = cands[0]
    if var not in names:
        raise KeyError(f"{path}: variable {var!r} not found (have {names})")
    arr = g[var]
    if arr.dims is None:
        raise ValueError(f"{path}/{var}: no dimension names (_ARRAY_DIMENSIONS / dimension_names)")
    for need in (timecoord, ydim, xdim):
        if need not in arr.dims:
            raise ValueError(f"dimension {need!r} not found in {list(arr.dims)}")
    axes = (arr.dims.index(timecoord), arr.dims.index(ydim), arr.dims.index(xdim))
    time = decode_time(g[timecoord])
    lat = np.asarray(g[ydim].read(), dtype=float)
    lon = np.asarray(g[xdim].read(), dtype=float)
    return ChunkedRaster(arr, axes), time, lat, lon


# ---------------------------------------------------------------------------------------------
# writer (fixtures, examples, dataset_to_zarr)
# ---------------------------------------------------------------------------------------------
def _compress(buf: bytes, compressor: Optional[str], level: int, typesize: int = 1) -> bytes:
    if compressor is None:
        return buf
    if compressor == "blosc":                                              # Blosc(cname="lz4", shuffle=SHUFFLE): zarr v2's default
        return blosc_compress(buf, typesize, "lz4", 1)
    if compressor in ("zlib",):
        return zlib.compress(buf, level)
    if compressor == "gzip":
        import gzip
        return gzip.compress(buf, compresslevel=level, mtime=0)
    if compressor == "zstd":
        return _arrow("zstd").compress(buf, asbytes=True)
    if compressor == "lz4":
        return struct.pack("<I", len(buf)) + _arrow("lz4_raw").compress(buf, asbytes=True)
    raise ValueError(f"compressor {compressor!r} not in [None, 'zlib', 'gzip', 'zstd', 'lz4', 'blosc']")


def write_array(path: str, data: np.ndarray, chunks: Sequence[int], dims: Sequence[str], attrs: Optional[dict] = None,
                zarr_format: int = 2, compressor: Optional[str] = "zlib", level: int = 1, order: str = "C",
                fill_value=None, skip_fill_chunks: bool = False, threads: int = 1) -> None:
    """One array of a directory store (v2 or v3).  ``order="F"`` stores chunks with the first axis fastest
    (v2 ``order``; a v3 ``transpose`` codec).  ``data``: an array, or any object with ``shape``, ``dtype`` and
    ``__getitem__(tuple of slices) -> ndarray`` whose ``lazy_blocks`` attribute is true (blocks are then fetched
    chunk by chunk, e.g. from a device tensor); ``threads`` > 1 compresses / writes chunks concurrently."""
    if not getattr(data, "lazy_blocks", False):
        data = np.asarray(data)
    chunks = tuple(int(min(max(1, c), max(1, s))) if c > 0 else max(1, int(s)) for c, s in zip(chunks, data.shape))
    os.makedirs(path, exist_ok=True)
    attrs = dict(attrs or {})
    n = data.ndim
    no_fill = fill_value is None and data.dtype.kind != "f"          # integers: no fill value (xarray would mask it)
    if fill_value is None:
        fill_value = np.nan if data.dtype.kind == "f" else 0
    fv_json = ("NaN" if np.isnan(fill_value) else float(fill_value)) if data.dtype.kind == "f" else int(fill_value)
    if no_fill and zarr_format == 2:
        fv_json = None
    if zarr_format == 2:
        if compressor == "lz4":
            comp = {"id": "lz4", "acceleration": 1}
        elif compressor == "zstd":
            comp = {"id": "zstd", "level": level}
        elif compressor == "blosc":
            comp = {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0}
        else:
            comp = None if compressor is None else {"id": compressor, "level": level}
        meta = {"zarr_format": 2, "shape": list(data.shape), "chunks": list(chunks), "dtype": data.dtype.str,
                "compressor": comp, "fill_value": fv_json, "order": order, "filters": None}
        json.dump(meta, open(os.path.join(path, ".zarray"), "w"))
        json.dump(dict(attrs, _ARRAY_DIMENSIONS=list(dims)), open(os.path.join(path, ".zattrs"), "w"))
        key = lambda idx: os.path.join(path, ".".join(map(str, idx)) if n else "0")          # noqa: E731
    elif zarr_format == 3:
        if compressor not in (None, "gzip", "zstd", "blosc"):
            raise ValueError("zarr v3 stores written here use None, 'gzip', 'zstd' or 'blosc'")
        inv = {"?": "bool", "i1": "int8", "i2": "int16", "i4": "int32", "i8": "int64", "u1": "uint8", "u2": "uint16",
               "u4": "uint32", "u8": "uint64", "f4": "float32", "f8": "float64"}
        codecs = []
        if order == "F" and n > 1:
            codecs.append({"name": "transpose", "configuration": {"order": list(reversed(range(n)))}})
        codecs.append({"name": "bytes", "configuration": {"endian": "little"}})
        if compressor == "zstd":
            codecs.append({"name": "zstd", "configuration": {"level": level, "checksum": False}})
        elif compressor == "gzip":
            codecs.append({"name": "gzip", "configuration": {"level": level}})
        elif compressor == "blosc":
            codecs.append({"name": "blosc", "configuration": {"cname": "lz4", "clevel": 5, "shuffle": "shuffle",
                                                              "typesize": data.dtype.itemsize, "blocksize": 0}})
        meta = {"zarr_format": 3, "node_type": "array", "shape": list(data.shape), "data_type": inv[data.dtype.str[1:]],
                "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": list(chunks)}},
                "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
                "fill_value": fv_json, "codecs": codecs, "attributes": attrs, "dimension_names": list(dims)}
        json.dump(meta, open(os.path.join(path, "zarr.json"), "w"))
        key = lambda idx: os.path.join(path, "c", *map(str, idx))                                 # noqa: E731
        if isinstance(data, np.ndarray):
            data = data.astype(data.dtype.newbyteorder("<"), copy=False)
    else:
        raise ValueError("zarr_format must be 2 or 3")
    grid = [-(-s // c) for s, c in zip(data.shape, chunks)]

    def one(idx):
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, data.shape))
        part = np.asarray(data[sl])
        if part.shape == tuple(chunks):
            block = part
        else:                                                            # edge chunks are stored full-size
            block = np.full(chunks, fill_value, data.dtype)
            block[tuple(slice(0, p) for p in part.shape)] = part
        if skip_fill_chunks and (np.isnan(block).all() if data.dtype.kind == "f" and np.isnan(fill_value)
                                 else (block == fill_value).all()):
            return
        raw = block.tobytes(order="F" if order == "F" else "C")
        p = key(idx)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(_compress(raw, compressor, level, data.dtype.itemsize))

    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=int(threads)) as ex:
            list(ex.map(one, np.ndindex(*grid)))
    else:
        for idx in np.ndindex(*grid):
            one(idx)


def write_dataset(store: str, values: np.ndarray, time, latitude, longitude, var: str = "variable",
                  dims: Sequence[str] = ("time", "latitude", "longitude"), chunks: Optional[Dict[str, int]] = None,
                  zarr_format: int = 2, compressor: Optional[str] = "zlib", level: int = 1, order: str = "C",
                  attrs: Optional[dict] = None, time_units: Optional[str] = None, calendar: str = "proleptic_gregorian",
                  skip_fill_chunks: bool = False, xycoords=("longitude", "latitude"), timecoord: str = "time") -> str:
    """A CF-style store xarray could open: ``var[dims]`` + coordinate arrays.  ``values`` is given as
    ``[time, lat, lon]`` and stored with the axis order of ``dims`` (the reference's converter stores
    ``(latitude, longitude, time)``, aggfly/dataset/zarr_convert.py:109)."""
    import pandas as pd
    from .timeaxis import CalendarIndex
    xdim, ydim = xycoords
    os.makedirs(store, exist_ok=True)
    if zarr_format == 2:
        json.dump({"zarr_format": 2}, open(os.path.join(store, ".zgroup"), "w"))
    else:
        json.dump({"zarr_format": 3, "node_type": "group", "attributes": {}}, open(os.path.join(store, "zarr.json"), "w"))
    names = {timecoord: 0, ydim: 1, xdim: 2}
    perm = [names[d] for d in dims]
    data = np.transpose(np.asarray(values), perm)
    chunks = chunks or {}
    cshape = [int(chunks.get(d, -1)) for d in dims]
    write_array(os.path.join(store, var), data, cshape, dims, attrs, zarr_format, compressor, level, order,
                skip_fill_chunks=skip_fill_chunks)
    if isinstance(time, CalendarIndex):
        t0 = CalendarIndex(time.calendar, time.year[:1], time.month[:1], time.day[:1], time.hour[:1])
        tvals = (time.ordinal_hours() - t0.ordinal_hours()[0]).astype(np.int64)
        tattrs = {"units": f"hours since {int(time.year[0]):04d}-{int(time.month[0]):02d}-{int(time.day[0]):02d} "
                           f"{int(time.hour[0]):02d}:00:00", "calendar": time.calendar}
    else:
        t = pd.DatetimeIndex(time)
        origin = t[0] if len(t) else pd.Timestamp("1970-01-01")
        unit = (time_units or "hours").rstrip("s")
        step = {"second": 10 ** 9, "minute": 60 * 10 ** 9, "hour": 3600 * 10 ** 9, "day": 86400 * 10 ** 9}[unit]
        delta = (t.values.astype("datetime64[ns]").astype(np.int64) - origin.value)
        tvals = delta // step if (delta % step == 0).all() else delta / step
        tattrs = {"units": f"{unit}s since {origin.strftime('%Y-%m-%d %H:%M:%S')}", "calendar": calendar}
    write_array(os.path.join(store, timecoord), np.asarray(tvals), [-1], [timecoord], tattrs, zarr_format, compressor, level)
    write_array(os.path.join(store, ydim), np.asarray(latitude, np.float64), [-1], [ydim], None, zarr_format, compressor, level)
    write_array(os.path.join(store, xdim), np.asarray(longitude, np.float64), [-1], [xdim], None, zarr_format, compressor, level)
    return store
