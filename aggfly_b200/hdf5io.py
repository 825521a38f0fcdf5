"""NetCDF-4 (HDF5) files read natively: what ``xr.open_dataset(path, chunks={"time": 24, ...})`` is to the reference
(aggfly/dataset/dataset.py:636-740 -- ERA5 bricks are NetCDF-4 files with shuffle + deflate chunks).  Neither h5py nor
netCDF4 nor xarray exists in the image, so the HDF5 structures a NetCDF-4 file is made of are parsed here from the
published file-format specification (HDF5 File Format Specification v3):

* superblock v0 / v1 (symbol-table root group) and v2 / v3 (root object header address);
* object headers v1 and v2 ("OHDR" / "OCHK" continuation blocks), header-message continuations;
* groups: old style (symbol table message -> v1 B-tree of "SNOD" nodes + local heap) and new style with COMPACT
  link messages (what netcdf-c writes for a file with a handful of variables); dense link storage (fractal heap) is
  rejected with a clear message;
* messages: dataspace v1 / v2, datatype (fixed point, IEEE float, fixed and variable-length strings), fill value
  (old, v1-v3), data layout v3 (compact / contiguous / chunked through a v1 B-tree) and v4 (contiguous, single chunk,
  implicit index, un-paged fixed array), filter pipeline v1 / v2, attribute v1-v3 (variable-length strings through the
  global heap); dense attribute storage is skipped with a warning;
* filters: deflate (1), shuffle (2), fletcher32 (3, stripped), per-chunk filter masks.

A dataset is exposed with the interface of ``zarrio.ZarrArray`` (``Hdf5Array``), so ``zarrio.ChunkedRaster`` and
``stream.feed_chunked`` drive it unchanged: host threads ``pread`` a chunk, undo deflate + shuffle into a pinned slot
(zlib releases the GIL), the chunk crosses PCIe as stored (int16 for packed ERA5 variables) and
``agf_tile_place_run`` transposes / unpacks it on the device.  Contiguous (unchunked) variables are cut into virtual
chunks of whole rows of the slowest axis.

Validation status: there is no libhdf5 here to write or cross-check files with.  The reader is tested against the
files of ``write_netcdf4`` below -- a writer that emits the OLD-style structures exactly as the specification lays
them out (superblock v0, v1 object headers, symbol-table group, v1 B-tree chunk index, v1 attributes) -- and against
hand-assembled v2 object headers / link messages in ``tests/test_hdf5io.py``.  Files written by libhdf5 follow the
same specification, but have not been seen by this code.
"""
from __future__ import annotations

import os
import struct
import warnings
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .zarrio import ChunkedRaster, ZarrArray, decode_time

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def looks_like_hdf5(path: str) -> bool:
    try:
        with open(path, "rb") as f:
            for off in (0, 512, 1024, 2048):
                f.seek(off)
                if f.read(8) == SIGNATURE:
                    return True
    except OSError:
        pass
    return False


# ---------------------------------------------------------------------------------------------
# low level
# ---------------------------------------------------------------------------------------------
class _Buf:
    """Random access into the file's bytes (memory mapped; metadata only -- chunks are read with pread)."""

    def __init__(self, path: str):
        import mmap
        self.path = path
        self.fd = os.open(path, os.O_RDONLY)
        self.size = os.fstat(self.fd).st_size
        self.mm = mmap.mmap(self.fd, 0, access=mmap.ACCESS_READ) if self.size else b""
        self.O = self.L = 8
        self.base = 0

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(self.mm[off:off + n], "little")

    def addr(self, off: int) -> int:
        v = self.u(off, self.O)
        return UNDEF if v == (1 << (8 * self.O)) - 1 else v + self.base

    def length(self, off: int) -> int:
        return self.u(off, self.L)

    def bytes(self, off: int, n: int) -> bytes:
        if off < 0 or off + n > self.size:
            raise IOError(f"{self.path}: read of {n} bytes at {off} past the end of the file")
        return bytes(self.mm[off:off + n])


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Datatype:
    def __init__(self, cls: int, size: int, dtype: Optional[np.dtype], vlen_string: bool = False, nbytes: int = 0):
        self.cls, self.size, self.dtype, self.vlen_string, self.nbytes = cls, size, dtype, vlen_string, nbytes


def _parse_datatype(b: bytes) -> _Datatype:
    """Datatype message -> (class, element size, NumPy dtype | None).  ``nbytes`` = bytes the message occupies."""
    cls, bits0 = b[0] & 0x0F, b[1]
    size = int.from_bytes(b[4:8], "little")
    if cls == 0:                                                      # fixed point
        dt = np.dtype(f"{'>' if bits0 & 1 else '<'}{'i' if bits0 & 8 else 'u'}{size}")
        return _Datatype(cls, size, dt, nbytes=12)
    if cls == 1:                                                      # floating point (IEEE layouts only)
        if size not in (2, 4, 8):
            raise NotImplementedError(f"HDF5 float of {size} bytes")
        return _Datatype(cls, size, np.dtype(f"{'>' if bits0 & 1 else '<'}f{size}"), nbytes=20)
    if cls == 3:                                                      # fixed-length string
        return _Datatype(cls, size, np.dtype(f"S{size}"), nbytes=8)
    if cls == 9:                                                      # variable length: strings only
        base = _parse_datatype(b[8:])
        return _Datatype(cls, size, None, vlen_string=(bits0 & 0x0F) == 1, nbytes=8 + base.nbytes)
    if cls == 7:                                                      # object reference
        return _Datatype(cls, size, np.dtype(f"V{size}"), nbytes=8)
    return _Datatype(cls, size, None, nbytes=8)


def _parse_dataspace(b: bytes, L: int) -> Tuple[Tuple[int, ...], int]:
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        p = 8
    elif ver == 2:
        if b[3] == 2:                                                 # null dataspace
            return (0,), 4
        p = 4
    else:
        raise NotImplementedError(f"dataspace message version {ver}")
    dims = tuple(int.from_bytes(b[p + i * L:p + (i + 1) * L], "little") for i in range(rank))
    n = p + rank * L * (2 if flags & 1 else 1)
    return dims, n


class _Object:
    """One object header, its messages collected: ``msgs`` = [(type, flags, data bytes)]."""

    def __init__(self, buf: _Buf, addr: int):
        self.buf, self.addr = buf, addr
        self.msgs: List[Tuple[int, int, bytes]] = []
        if buf.bytes(addr, 4) == b"OHDR":
            self._parse_v2(addr)
        else:
            self._parse_v1(addr)

    def _parse_v1(self, addr: int):
        b = self.buf
        if b.u(addr, 1) != 1:
            raise IOError(f"{b.path}: no object header at {addr}")
        blocks = [(addr + 16, b.u(addr + 8, 4))]
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end:
                mtype, size, flags = b.u(p, 2), b.u(p + 2, 2), b.u(p + 4, 1)
                data = b.bytes(p + 8, size)
                if mtype == 0x10:
                    blocks.append((b.addr(p + 8), b.length(p + 8 + b.O)))
                elif mtype != 0:
                    self.msgs.append((mtype, flags, data))
                p += 8 + size

    def _parse_v2(self, addr: int):
        b = self.buf
        flags = b.u(addr + 5, 1)
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        size0 = b.u(p, nsz)
        p += nsz
        order = 2 if flags & 0x04 else 0
        blocks = [(p, p + size0)]
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 + order <= end:
                mtype, size, mflags = b.u(p, 1), b.u(p + 1, 2), b.u(p + 3, 1)
                q = p + 4 + order
                if q + size > end:
                    break                                              # gap before the checksum
                data = b.bytes(q, size)
                if mtype == 0x10:
                    caddr, clen = b.addr(q), b.length(q + b.O)
                    if b.bytes(caddr, 4) != b"OCHK":
                        raise IOError(f"{b.path}: continuation block at {caddr} without OCHK signature")
                    blocks.append((caddr + 4, caddr + clen - 4))
                elif mtype != 0:
                    self.msgs.append((mtype, mflags, data))
                p = q + size

    def first(self, mtype: int) -> Optional[bytes]:
        for t, _, d in self.msgs:
            if t == mtype:
                return d
        return None

    # -- attributes -------------------------------------------------------------------------------
    def attributes(self) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for t, _, d in self.msgs:
            if t == 0x0C:
                try:
                    name, value = self._attribute(d)
                    out[name] = value
                except NotImplementedError:
                    continue
            elif t == 0x15:
                if self.buf.u_from(d, 2 + (2 if d[1] & 1 else 0), self.buf.O) != (1 << (8 * self.buf.O)) - 1:
                    warnings.warn(f"{self.buf.path}: an object keeps its attributes in dense storage (fractal heap), which "
                                  "this reader does not parse; attributes of that object are missing", stacklevel=3)
        return out

    def _attribute(self, d: bytes):
        ver = d[0]
        nsz, tsz, ssz = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
        p = 8 if ver < 3 else 9
        pad = _pad8 if ver == 1 else (lambda n: n)
        name = d[p:p + nsz].split(b"\0")[0].decode("utf-8", "replace")
        p += pad(nsz)
        dt = _parse_datatype(d[p:p + tsz])
        p += pad(tsz)
        dims, _ = _parse_dataspace(d[p:p + ssz], self.buf.L) if ssz else ((), 0)
        p += pad(ssz)
        count = int(np.prod(dims)) if dims else 1
        raw = d[p:]
        if dt.vlen_string:
            vals = [self._vlen_string(raw[i * dt.size:(i + 1) * dt.size]) for i in range(count)]
            return name, (vals[0] if not dims else vals)
        if dt.dtype is None or dt.cls == 7:
            raise NotImplementedError
        arr = np.frombuffer(raw, dt.dtype, count)
        if dt.cls == 3:
            vals = [v.split(b"\0")[0].decode("utf-8", "replace") for v in arr.tolist()]
            return name, (vals[0] if not dims else vals)
        arr = arr.astype(dt.dtype.newbyteorder("="))
        return name, (arr[0].item() if not dims else arr.copy())

    def _vlen_string(self, ref: bytes) -> str:
        b = self.buf
        n = int.from_bytes(ref[0:4], "little")
        caddr = int.from_bytes(ref[4:4 + b.O], "little") + b.base
        index = int.from_bytes(ref[4 + b.O:8 + b.O], "little")
        if b.bytes(caddr, 4) != b"GCOL":
            raise IOError(f"{b.path}: no global heap collection at {caddr}")
        end = caddr + b.length(caddr + 8)
        p = caddr + 8 + b.L
        while p + 8 + b.L <= end:
            idx, size = b.u(p, 2), b.length(p + 8)
            if idx == 0:
                break
            if idx == index:
                return b.bytes(p + 8 + b.L, min(n, size)).split(b"\0")[0].decode("utf-8", "replace")
            p += 8 + b.L + _pad8(size)
        raise IOError(f"{b.path}: global heap object {index} not found")


def _u_from(self, data: bytes, off: int, n: int) -> int:
    return int.from_bytes(data[off:off + n], "little")


_Buf.u_from = _u_from


class Hdf5File:
    """The root group of an HDF5 file: ``names()`` and ``file[name] -> Hdf5Array``."""

    def __init__(self, path: str):
        self.path = path
        b = self.buf = _Buf(path)
        sb = next((o for o in (0, 512, 1024, 2048, 4096) if o + 8 <= b.size and b.bytes(o, 8) == SIGNATURE), None)
        if sb is None:
            raise IOError(f"{path}: not an HDF5 file")
        ver = b.u(sb + 8, 1)
        if ver in (0, 1):
            b.O, b.L = b.u(sb + 13, 1), b.u(sb + 14, 1)
            p = sb + 24 + (4 if ver == 1 else 0)
            b.base = b.u(p, b.O)
            p += 4 * b.O                                               # base, free space, end of file, driver info
            root = b.addr(p + b.O)                                     # root symbol table entry: name offset, header address
        elif ver in (2, 3):
            b.O, b.L = b.u(sb + 9, 1), b.u(sb + 10, 1)
            b.base = b.u(sb + 12, b.O)
            root = b.addr(sb + 12 + 3 * b.O)
        else:
            raise NotImplementedError(f"{path}: HDF5 superblock version {ver}")
        self.links = self._group_links(_Object(b, root))
        self._objects: Dict[str, _Object] = {}

    def _group_links(self, obj: _Object) -> Dict[str, int]:
        b = self.buf
        links: Dict[str, int] = {}
        st = obj.first(0x11)
        if st is not None:                                             # old style: v1 B-tree of symbol nodes + local heap
            btree, heap = int.from_bytes(st[:b.O], "little") + b.base, int.from_bytes(st[b.O:2 * b.O], "little") + b.base
            if b.bytes(heap, 4) != b"HEAP":
                raise IOError(f"{b.path}: no local heap at {heap}")
            data = b.addr(heap + 8 + 2 * b.L)
            for snod in self._btree_leaves(btree, 0):
                if b.bytes(snod, 4) != b"SNOD":
                    raise IOError(f"{b.path}: no symbol node at {snod}")
                for i in range(b.u(snod + 6, 2)):
                    e = snod + 8 + i * (2 * b.O + 24)
                    noff = b.u(e, b.O)
                    end = b.mm.find(b"\0", data + noff)
                    links[bytes(b.mm[data + noff:end]).decode("utf-8", "replace")] = b.addr(e + b.O)
            return links
        for t, _, d in obj.msgs:
            if t == 0x02:                                              # link info: a defined fractal heap = dense storage
                p = 2 + (8 if d[1] & 1 else 0)
                if int.from_bytes(d[p:p + b.O], "little") != (1 << (8 * b.O)) - 1:
                    raise NotImplementedError(f"{b.path}: the group stores its links densely (fractal heap); files with more "
                                              "than eight objects per group are not supported by this reader")
            elif t == 0x06:
                flags = d[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = d[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nsz = 1 << (flags & 3)
                n = int.from_bytes(d[p:p + nsz], "little")
                p += nsz
                name = d[p:p + n].decode("utf-8", "replace")
                p += n
                if ltype == 0:
                    links[name] = int.from_bytes(d[p:p + b.O], "little") + b.base
        return links

    def _btree_leaves(self, addr: int, node_type: int, ndim: int = 0):
        """Children of the level-0 nodes of a v1 B-tree, left to right: symbol node addresses (type 0) or
        (chunk offsets, size, filter mask, address) (type 1)."""
        b = self.buf
        if b.bytes(addr, 4) != b"TREE" or b.u(addr + 4, 1) != node_type:
            raise IOError(f"{b.path}: no v1 B-tree node of type {node_type} at {addr}")
        level, n = b.u(addr + 5, 1), b.u(addr + 6, 2)
        p = addr + 8 + 2 * b.O
        key = b.L if node_type == 0 else 8 + 8 * ndim
        for i in range(n):
            k = p + i * (key + b.O)
            child = b.addr(k + key)
            if level > 0:
                yield from self._btree_leaves(child, node_type, ndim)
            elif node_type == 0:
                yield child
            else:
                offs = tuple(b.u(k + 8 + 8 * j, 8) for j in range(ndim - 1))
                yield offs, b.u(k, 4), b.u(k + 4, 4), child

    def names(self) -> List[str]:
        return sorted(self.links)

    def __contains__(self, name: str) -> bool:
        return name in self.links

    def object(self, name: str) -> _Object:
        if name not in self._objects:
            if name not in self.links:
                raise KeyError(f"{self.path}: no object {name!r} (have {self.names()})")
            self._objects[name] = _Object(self.buf, self.links[name])
        return self._objects[name]

    def is_dataset(self, name: str) -> bool:
        o = self.object(name)
        return o.first(0x08) is not None and o.first(0x01) is not None

    def __getitem__(self, name: str) -> "Hdf5Array":
        return Hdf5Array(self, name)


class Hdf5Array(ZarrArray):
    """An HDF5 dataset behind the ``ZarrArray`` interface (shape / chunks / dtype / attrs / fill_value, chunk reads):
    everything built on that interface -- ``ChunkedRaster``, the chunked host feed -- works on it unchanged."""

    zarr_format = 0

    def __init__(self, file: Hdf5File, name: str):                      # noqa: super().__init__ is zarr-specific
        self.file, self.name = file, name
        self.path = f"{file.path}:{name}"
        b = file.buf
        obj = file.object(name)
        space, dtype, layout = obj.first(0x01), obj.first(0x03), obj.first(0x08)
        if space is None or dtype is None or layout is None:
            raise ValueError(f"{self.path}: not a dataset")
        self.shape, _ = _parse_dataspace(space, b.L)
        self.ndim = len(self.shape)
        dt = _parse_datatype(dtype)
        if dt.dtype is None or dt.cls not in (0, 1):
            raise NotImplementedError(f"{self.path}: HDF5 datatype class {dt.cls}")
        self.dtype = dt.dtype
        self.attrs = obj.attributes()
        self.storage_axes = tuple(range(self.ndim))                    # HDF5 chunks are C-ordered
        self.dims = None
        self._cf_fill_default = None
        self._decoders, self._filters = [], []
        # ---- fill value ----
        self.fill_value = None
        fv = obj.first(0x05)
        raw = None
        if fv is not None:
            if fv[0] in (1, 2):
                if fv[0] == 1 or fv[3]:
                    n = int.from_bytes(fv[4:8], "little")
                    raw = fv[8:8 + n] if n else None
            elif fv[0] == 3 and fv[1] & 0x20:
                n = int.from_bytes(fv[2:6], "little")
                raw = fv[6:6 + n] if n else None
        elif obj.first(0x04) is not None:
            old = obj.first(0x04)
            n = int.from_bytes(old[0:4], "little")
            raw = old[4:4 + n] if n else None
        if raw is not None and len(raw) == self.dtype.itemsize:
            self.fill_value = np.frombuffer(raw, self.dtype)[0].astype(self.dtype.newbyteorder("=")).item()
        # ---- filters ----
        self.filters: List[Tuple[int, Tuple[int, ...]]] = []
        fp = obj.first(0x0B)
        if fp is not None:
            ver, nf = fp[0], fp[1]
            p = 8 if ver == 1 else 2
            for _ in range(nf):
                fid = int.from_bytes(fp[p:p + 2], "little")
                p += 2
                nlen = 0
                if ver == 1 or fid >= 256:
                    nlen = int.from_bytes(fp[p:p + 2], "little")
                    p += 2
                p += 2                                                 # flags
                ncd = int.from_bytes(fp[p:p + 2], "little")
                p += 2
                p += _pad8(nlen) if ver == 1 else nlen
                cd = tuple(int.from_bytes(fp[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd))
                p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                if fid not in (1, 2, 3):
                    raise NotImplementedError(f"{self.path}: HDF5 filter {fid} (deflate, shuffle and fletcher32 are read)")
                self.filters.append((fid, cd))
        # ---- layout ----
        self._chunk_index: Optional[Dict[Tuple[int, ...], Tuple[int, int, int]]] = None
        self._contiguous: Optional[Tuple[int, int]] = None
        self._compact: Optional[bytes] = None
        ver, cls = layout[0], layout[1]
        if ver not in (3, 4):
            raise NotImplementedError(f"{self.path}: data layout message version {ver}")
        if cls == 0:
            n = int.from_bytes(layout[2:4], "little")
            self._compact = layout[4:4 + n]
            self.chunks = tuple(max(1, s) for s in self.shape)
        elif cls == 1:
            addr = int.from_bytes(layout[2:2 + b.O], "little")
            self._contiguous = (UNDEF if addr == (1 << (8 * b.O)) - 1 else addr + b.base, int.from_bytes(layout[2 + b.O:2 + b.O + b.L], "little"))
            # virtual chunks: whole rows of the slowest axis, about 32 MB each
            row = int(np.prod(self.shape[1:])) * self.dtype.itemsize if self.ndim else self.dtype.itemsize
            rows = max(1, min(self.shape[0] if self.ndim else 1, (32 << 20) // max(1, row)))
            self.chunks = ((rows,) + tuple(self.shape[1:])) if self.ndim else ()
        elif cls == 2 and ver == 3:
            nd = layout[2]
            btree = int.from_bytes(layout[3:3 + b.O], "little")
            self.chunks = tuple(int.from_bytes(layout[3 + b.O + 4 * i:7 + b.O + 4 * i], "little") for i in range(nd - 1))
            self._btree = (UNDEF if btree == (1 << (8 * b.O)) - 1 else btree + b.base, nd)
        elif cls == 2:
            self._layout_v4(layout)
        else:
            raise NotImplementedError(f"{self.path}: data layout class {cls}")
        self._fd = b.fd

    def _layout_v4(self, d: bytes):
        b = self.file.buf
        flags, nd, enc = d[2], d[3], d[4]
        self.chunks = tuple(int.from_bytes(d[5 + enc * i:5 + enc * (i + 1)], "little") for i in range(nd - 1))
        p = 5 + enc * nd
        itype = d[p]
        p += 1
        grid = [-(-s // c) for s, c in zip(self.shape, self.chunks)]
        nbytes = int(np.prod(self.chunks)) * self.dtype.itemsize
        index: Dict[Tuple[int, ...], Tuple[int, int, int]] = {}
        if itype == 1:                                                 # single chunk
            size, mask = nbytes, 0
            if flags & 2:
                size, mask = int.from_bytes(d[p:p + b.L], "little"), int.from_bytes(d[p + b.L:p + b.L + 4], "little")
                p += b.L + 4
            index[(0,) * (nd - 1)] = (int.from_bytes(d[p:p + b.O], "little") + b.base, size, mask)
        elif itype == 2:                                               # implicit: unfiltered chunks back to back
            addr = int.from_bytes(d[p:p + b.O], "little") + b.base
            for k, idx in enumerate(np.ndindex(*grid)):
                index[tuple(int(i) for i in idx)] = (addr + k * nbytes, nbytes, 0)
        elif itype == 3:                                               # fixed array (un-paged)
            hdr = int.from_bytes(d[p + 1:p + 1 + b.O], "little") + b.base
            if b.bytes(hdr, 4) != b"FAHD":
                raise IOError(f"{self.path}: no fixed array header at {hdr}")
            client, esz, page_bits = b.u(hdr + 5, 1), b.u(hdr + 6, 1), b.u(hdr + 7, 1)
            n = b.length(hdr + 8)
            dblk = b.addr(hdr + 8 + b.L)
            if n > (1 << page_bits):
                raise NotImplementedError(f"{self.path}: paged fixed-array chunk index")
            if b.bytes(dblk, 4) != b"FADB":
                raise IOError(f"{self.path}: no fixed array data block at {dblk}")
            q = dblk + 6 + b.O
            for k, idx in enumerate(np.ndindex(*grid)):
                if k >= n:
                    break
                e = q + k * esz
                a = b.addr(e)
                if client == 1:
                    size, mask = b.u(e + b.O, esz - b.O - 4), b.u(e + esz - 4, 4)
                else:
                    size, mask = nbytes, 0
                if a != UNDEF:
                    index[tuple(int(i) for i in idx)] = (a, size, mask)
        else:
            raise NotImplementedError(f"{self.path}: chunk index type {itype} (extensible array / v2 B-tree)")
        self._chunk_index = index
        self._btree = (UNDEF, nd)

    # -- chunk access -------------------------------------------------------------------------------
    def _index(self) -> Dict[Tuple[int, ...], Tuple[int, int, int]]:
        if self._chunk_index is None:
            idx: Dict[Tuple[int, ...], Tuple[int, int, int]] = {}
            addr, nd = self._btree
            if addr != UNDEF:
                for offs, size, mask, child in self.file._btree_leaves(addr, 1, nd):
                    idx[tuple(o // c for o, c in zip(offs, self.chunks))] = (child, size, mask)
            self._chunk_index = idx
        return self._chunk_index

    def chunk_path(self, idx: Sequence[int]) -> str:
        return f"{self.path}[chunk {tuple(int(i) for i in idx)}]"

    def read_chunk_bytes(self, idx: Sequence[int], into: Optional[np.ndarray] = None):
        """The chunk as stored (filtered bytes); None when it was never written (all fill value)."""
        idx = tuple(int(i) for i in idx)
        if self._compact is not None:
            return self._compact
        if self._contiguous is not None:
            addr, _size = self._contiguous
            if addr == UNDEF:
                return None
            row = int(np.prod(self.shape[1:])) * self.dtype.itemsize if self.ndim else self.dtype.itemsize
            r0 = idx[0] * self.chunks[0] if self.ndim else 0
            n = (min(self.shape[0], r0 + self.chunks[0]) - r0) * row if self.ndim else row
            got = os.pread(self._fd, n, addr + r0 * row)
            if len(got) != n:
                raise IOError(f"{self.chunk_path(idx)}: short read")
            return got + bytes(self.chunk_nbytes - n)                 # the last virtual chunk is padded to full size
        ent = self._index().get(idx)
        if ent is None:
            return None
        addr, size, mask = ent
        got = os.pread(self._fd, size, addr)
        if len(got) != size:
            raise IOError(f"{self.chunk_path(idx)}: short read ({len(got)} of {size} bytes)")
        return (mask, got) if mask else got

    def decode_chunk(self, buf, out: Optional[np.ndarray] = None) -> np.ndarray:
        mask = 0
        if isinstance(buf, tuple):
            mask, buf = buf
        if self._contiguous is None and self._compact is None:
            for k in range(len(self.filters) - 1, -1, -1):             # undo the pipeline back to front
                if mask & (1 << k):
                    continue
                fid, cd = self.filters[k]
                if fid == 1:
                    buf = zlib.decompress(buf)
                elif fid == 3:
                    buf = buf[:-4]
                elif fid == 2:
                    es = cd[0] if cd else self.dtype.itemsize
                    a = np.frombuffer(buf, np.uint8)
                    n = a.size // es
                    body = np.ascontiguousarray(a[: n * es].reshape(es, n).T).reshape(-1)
                    buf = body.tobytes() + bytes(a[n * es:]) if a.size != n * es else memoryview(body)
        arr = np.frombuffer(buf, self.dtype, int(np.prod(self.chunks)))
        native = self.dtype.newbyteorder("=")
        if out is None:
            out = np.empty(arr.size, native)
        dst = out[: arr.size]
        np.copyto(dst, arr, casting="unsafe" if not self.dtype.isnative else "same_kind")
        return dst.reshape(self.storage_shape)

    @property
    def blosc_only(self) -> bool:
        return False

    def __repr__(self):
        return f"<Hdf5Array {self.path!r} shape={self.shape} chunks={self.chunks} dtype={self.dtype} filters={self.filters}>"


# ---------------------------------------------------------------------------------------------
# NetCDF-4 view
# ---------------------------------------------------------------------------------------------
def open_raster(path: str, var: Optional[str], xycoords=("longitude", "latitude"), timecoord: str = "time"):
    """(ChunkedRaster, time, latitude, longitude) of variable ``var`` of the NetCDF-4 file at ``path``.  The variable's
    axes are matched to the coordinate variables by NAME where the file says so (``_Netcdf4Coordinates`` is not needed:
    the coordinate variables are 1-D datasets called like the dimensions) and by LENGTH."""
    f = Hdf5File(path)
    xdim, ydim = xycoords
    for need in (timecoord, ydim, xdim):
        if need not in f:
            raise ValueError(f"{path}: coordinate variable {need!r} not found (have {f.names()})")
    if var is None:
        cands = [n for n in f.names() if n not in (xdim, ydim, timecoord) and f.is_dataset(n) and f[n].ndim == 3]
        if len(cands) != 1:
            raise KeyError(f"{path}: pass var= (3-D variables: {cands})")
        var = cands[0]
    if var not in f:
        raise KeyError(f"{path}: variable {var!r} not found (have {f.names()})")
    arr = f[var]
    if arr.ndim != 3:
        raise ValueError(f"{path}:{var}: the raster variable must have exactly (time, lat, lon) dimensions, has shape {arr.shape}")
    coords = {n: f[n] for n in (timecoord, ydim, xdim)}
    sizes = {n: int(c.shape[0]) for n, c in coords.items()}
    axes = []
    free = list(range(3))
    for n in (timecoord, ydim, xdim):
        match = [a for a in free if arr.shape[a] == sizes[n]]
        if not match:
            raise ValueError(f"{path}:{var}: no axis of length {sizes[n]} for dimension {n!r} (shape {arr.shape})")
        if len(match) > 1:                                             # equal lengths: NetCDF order is (time, lat, lon)
            match = [a for a in match if a == (timecoord, ydim, xdim).index(n)] or match
        axes.append(match[0])
        free.remove(match[0])
    arr.dims = tuple(n for _, n in sorted(zip(axes, (timecoord, ydim, xdim))))
    time = decode_time(coords[timecoord])
    lat = np.asarray(coords[ydim].read(), dtype=float)
    lon = np.asarray(coords[xdim].read(), dtype=float)
    return ChunkedRaster(arr, tuple(axes)), time, lat, lon


# ---------------------------------------------------------------------------------------------
# writer (fixtures, examples): old-style structures, exactly as the specification lays them out
# ---------------------------------------------------------------------------------------------
def _dt_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, be | (8 if dt.kind == "i" else 0), 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        e, m, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20 | be, dt.itemsize * 8 - 1, 0, dt.itemsize, 0, dt.itemsize * 8, m, e, 0, m, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)
    raise TypeError(f"no HDF5 datatype for {dt}")


def _space_message(shape: Sequence[int]) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _attr_message(name: str, value) -> bytes:
    if isinstance(value, str):
        data = value.encode() + b"\0"
        dt, sp = _dt_message(np.dtype(f"S{len(data)}")), _space_message(())
    else:
        a = np.atleast_1d(np.asarray(value))
        if a.dtype.kind not in "iuf":
            raise TypeError(f"attribute {name!r}: {a.dtype}")
        a = a.astype(a.dtype.newbyteorder("<"))
        data, dt = a.tobytes(), _dt_message(a.dtype)
        sp = _space_message(() if np.ndim(value) == 0 else a.shape)
    nm = name.encode() + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp))
    for part in (nm, dt, sp):
        body += part + bytes(_pad8(len(part)) - len(part))
    return body + data


def _object_header(messages: List[Tuple[int, bytes]]) -> bytes:
    body = b""
    for mtype, data in messages:
        body += struct.pack("<HHB3x", mtype, _pad8(len(data)), 0) + data + bytes(_pad8(len(data)) - len(data))
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Out:
    def __init__(self, path: str):
        self.f = open(path, "wb")
        self.f.write(bytes(96))                                        # the superblock goes here at the end

    def put(self, data: bytes) -> int:
        pos = self.f.tell()
        pad = _pad8(pos) - pos
        if pad:
            self.f.write(bytes(pad))
        addr = self.f.tell()
        self.f.write(data)
        return addr


def _write_chunk_btree(out: _Out, entries: List[Tuple[Tuple[int, ...], int, int]], shape, chunks, K: int = 32) -> int:
    """v1 B-tree (node type 1) over (chunk offsets, stored size, address), offsets ascending.  Returns its address."""
    nd = len(shape) + 1
    end_key = struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", -(-s // c) * c) for s, c in zip(shape, chunks)) + struct.pack("<Q", 0)

    def key(e):
        return struct.pack("<II", e[1], 0) + b"".join(struct.pack("<Q", o) for o in e[0]) + struct.pack("<Q", 0)

    level = 0
    nodes = [(e[0], key(e), e[2]) for e in entries]                    # (first offsets, key bytes, child address)
    while True:
        groups = [nodes[i:i + 2 * K] for i in range(0, len(nodes), 2 * K)] or [[]]
        addrs, pos = [], out.f.tell()
        pos = _pad8(pos)
        node_size = 8 + 16 + 2 * K * (8 + 8 * nd + 8) + (8 + 8 * nd)
        for g in range(len(groups)):
            addrs.append(pos + g * _pad8(node_size))
        new_nodes = []
        for g, grp in enumerate(groups):
            left = addrs[g - 1] if g > 0 else UNDEF
            right = addrs[g + 1] if g + 1 < len(groups) else UNDEF
            body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
            for _, kb, child in grp:
                body += kb + struct.pack("<Q", child)
            nxt = groups[g + 1][0][1] if g + 1 < len(groups) else end_key
            body += nxt
            body += bytes(node_size - len(body))
            a = out.put(body)
            assert a == addrs[g]
            if grp:
                new_nodes.append((grp[0][0], grp[0][1], a))
        if len(groups) == 1:
            return addrs[0]
        nodes, level = new_nodes, level + 1


def write_netcdf4(path: str, values: np.ndarray, time_values: np.ndarray, time_units: str, latitude, longitude,
                  var: str = "t2m", dims: Sequence[str] = ("time", "latitude", "longitude"), chunks: Optional[Sequence[int]] = None,
                  deflate: Optional[int] = 4, shuffle: bool = True, attrs: Optional[dict] = None, calendar: Optional[str] = None,
                  fill_value=None, skip_chunks: Sequence[Tuple[int, ...]] = (),
                  coord_names: Sequence[str] = ("time", "latitude", "longitude")) -> None:
    """A NetCDF-4-shaped HDF5 file: the coordinate variables ``coord_names`` = (time, latitude, longitude) (contiguous) + the
    variable ``var`` whose ``values`` are laid out in the order ``dims`` (a permutation of ``coord_names``) -- chunked (v1 B-tree index) with optional shuffle + deflate filters when ``chunks`` is
    given, contiguous otherwise.  ``values`` may be float32 / float64 or CF-packed integers (pass ``scale_factor`` /
    ``add_offset`` / ``_FillValue`` in ``attrs``).  ``skip_chunks``: chunk grid indices left unwritten (they read back as
    the fill value).  Old-style HDF5 structures only (see the module docstring)."""
    values = np.asarray(values)
    coords = {coord_names[0]: np.asarray(time_values), coord_names[1]: np.asarray(latitude), coord_names[2]: np.asarray(longitude)}
    if sorted(dims) != sorted(coord_names) or tuple(len(coords[d]) for d in dims) != values.shape:
        raise ValueError(f"values of shape {values.shape} do not match dims {tuple(dims)}")
    out = _Out(path)
    headers: Dict[str, int] = {}

    def dataset(name: str, arr: np.ndarray, a: dict, chunked: Optional[Sequence[int]]):
        arr = np.ascontiguousarray(arr)
        le = arr.astype(arr.dtype.newbyteorder("<"), copy=False)
        msgs = [(0x01, _space_message(arr.shape)), (0x03, _dt_message(le.dtype))]
        fv = a.get("_FillValue", fill_value if name == var else None)
        if fv is not None:
            msgs.append((0x05, struct.pack("<BBBBI", 2, 2, 0, 1, le.dtype.itemsize) + np.asarray(fv, le.dtype).tobytes()))
        else:
            msgs.append((0x05, struct.pack("<BBBB", 2, 2, 0, 0)))
        if chunked is None:
            addr = out.put(le.tobytes())
            msgs.append((0x08, struct.pack("<BBQQ", 3, 1, addr, le.nbytes)))
        else:
            chunked = tuple(int(c) for c in chunked)
            grid = [-(-s // c) for s, c in zip(arr.shape, chunked)]
            entries = []
            for idx in np.ndindex(*grid):
                if tuple(int(i) for i in idx) in set(tuple(s) for s in skip_chunks):
                    continue
                block = np.zeros(chunked, le.dtype)
                if fv is not None:
                    block[...] = fv
                sl = tuple(slice(i * c, min(s, (i + 1) * c)) for i, c, s in zip(idx, chunked, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = le[sl]
                raw = block.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, le.dtype.itemsize).T.tobytes()
                if deflate is not None:
                    raw = zlib.compress(raw, int(deflate))
                entries.append((tuple(i * c for i, c in zip(idx, chunked)), len(raw), out.put(raw)))
            btree = _write_chunk_btree(out, entries, arr.shape, chunked)
            msgs.append((0x08, struct.pack("<BBBQ", 3, 2, arr.ndim + 1, btree) + b"".join(struct.pack("<I", c) for c in chunked)
                         + struct.pack("<I", le.dtype.itemsize)))
            flt = []
            if shuffle:
                flt.append((2, b"shuffle\0", (le.dtype.itemsize,)))
            if deflate is not None:
                flt.append((1, b"deflate\0", (int(deflate),)))
            if flt:
                body = struct.pack("<BB6x", 1, len(flt))
                for fid, nm, cd in flt:
                    body += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm + bytes(_pad8(len(nm)) - len(nm))
                    body += b"".join(struct.pack("<I", c) for c in cd) + (bytes(4) if len(cd) % 2 else b"")
                msgs.append((0x0B, body))
        for k, v in a.items():
            if k == "_FillValue":
                v = np.asarray(v, le.dtype)
            msgs.append((0x0C, _attr_message(k, v)))
        headers[name] = out.put(_object_header(msgs))

    for i, d in enumerate(dims):
        a = {"CLASS": "DIMENSION_SCALE", "NAME": d, "_Netcdf4Dimid": np.int32(i)}
        if d == coord_names[0] and time_units:
            a["units"] = time_units
            if calendar:
                a["calendar"] = calendar
        dataset(d, coords[d], a, None)
    dataset(var, values, dict(attrs or {}), chunks)
    # ---- root group: local heap with the names, one symbol node, a one-node B-tree, the group's object header ----
    names = sorted(headers)
    heap_data = b"\0" * 8
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        heap_data += n.encode() + b"\0"
        heap_data += bytes(_pad8(len(heap_data)) - len(heap_data))
    heap_data += bytes(16)
    data_addr = out.put(heap_data)
    heap = out.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), len(heap_data) - 16, data_addr))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", offs[n], headers[n], 0, 0)
    snod += bytes(8 + 8 * 40 - len(snod))
    if len(names) > 8:
        raise NotImplementedError("the fixture writer keeps all objects in one symbol node (at most 8)")
    snod_addr = out.put(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])
    tree += bytes(8 + 16 + 2 * 16 * 16 + 8 - len(tree))
    tree_addr = out.put(tree)
    root = out.put(_object_header([(0x11, struct.pack("<QQ", tree_addr, heap))]))
    eof = _pad8(out.f.tell())
    out.f.write(bytes(eof - out.f.tell()))
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap)
    assert len(sb) == 96
    out.f.seek(0)
    out.f.write(sb)
    out.f.close()
