"""Host-resident rasters: chunked host->device feed overlapped with the temporal kernels.

The reference reads its raster lazily, chunk by chunk, inside ``dask.compute``
(aggfly/dataset/dataset.py:636-740 opens with ``chunks={"time": 24, ...}``; the single compute is at
aggfly/aggregate/spatial.py:125).  Here a raster that lives in host memory (numpy array, CPU torch
tensor, pinned or pageable) is cut into row chunks; each chunk goes to the device with one
``cudaMemcpyAsync`` on a copy stream, and as soon as the rows of a time stripe are resident the
stripe's temporal kernel is launched on the compute stream (``agf_temporal_run`` takes a stripe
range), so the copy engine and the SMs work at the same time and the call costs
``max(copy, compute)`` instead of ``copy + compute``.

* pinned source: chunks are copied straight from the caller's buffer;
* pageable source: worker threads memcpy chunks into a small ring of pinned staging buffers
  (numpy releases the GIL), the main thread issues the async copies in order.

Device footprint.  A raster up to ``OPTIONS["device_raster_budget_bytes"]`` is held whole on the device (one year
of global 0.25deg hourly data is 36.4 GB of the 180 GB HBM) and its buffer is kept for the next call.  A longer
record -- the reference streams any length by time chunk (aggfly/aggregate/spatial.py:189-199) and its CLI loops
years (aggfly/cli/pipeline.py:138-150) -- goes through a RING of ``ring_slots`` device windows of about
``ring_slot_bytes``: a window ends where every raster-reading program has a stripe end, its kernels are launched
with the window as their raster (``row0`` of agf_temporal_run / agf_temporal_regional_run), and its slot is reused
once they have run.  Device raster memory is then independent of the length of the time axis.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np

# Measured on a B200 box (16 cores, tools/feed_sweep.py, global year = 36.4 GB of pageable NumPy data):
# 4 threads x 4 slots x 256 MB 949 ms; 8 x 8 x 128 MB 715 ms; 8 x 8 x 64 MB 710 ms; 16 x 16 x 64 MB 694 ms
# -- the last three are PCIe-bound like the pinned path (655-700 ms).
OPTIONS = {
    "chunk_bytes": 256 << 20,          # bytes per host->device copy, pinned sources
    "staging_chunk_bytes": 64 << 20,   # ... pageable sources (also the size of a staging slot)
    "staging_slots": 8,                # pinned ring depth for pageable sources
    "staging_threads": 8,              # host threads filling the ring
    "keep_device_raster": True,        # keep the device copy's buffer between calls (see _device_raster)
    # rasters larger than this are streamed through a ring of device windows instead of being held whole
    "device_raster_budget_bytes": int(float(__import__("os").environ.get("AGF_RASTER_BUDGET_GB", "48")) * (1 << 30)),
    "ring_slot_bytes": 16 << 30,        # target size of one device window (grown to the largest window the cuts allow)
    "ring_slots": 3,
    "chunked_ring_bytes": 2 << 30,     # pinned (and device) staging for chunked stores: slots x decoded chunk size
    "device_decompress": True,         # Blosc-LZ4 chunks: inflate on the GPU's decompression engine when it has one
    # The per-chunk tables of the device decode (expected stream lengths, raw-segment table) are uploaded from
    # pageable memory, which makes the driver synchronise the compute stream twice per chunk.  True: the feed threads append them to the staged slot so that they ride the chunk's own copy.
    # Validated on a B200 in round 2 (GPU suite green both ways; the 24-hour-chunk Blosc store feeds in 156 ms either
    # way, so the uploads were not its bound) and kept on: one copy per chunk instead of three.
    "inline_chunk_tables": True,
    # Blosc packs its streams back to back at arbitrary byte offsets.  0: hand them to the engine where they are;
    # n > 1: first move every stream to an n-byte aligned offset of a second device buffer (one segment-copy launch)
    "device_decompress_align": int(__import__("os").environ.get("AGF_DE_ALIGN", "0")),
}


def _host_source(values):
    """(pinned torch tensor | None, numpy view | None): a pinned contiguous tensor is copied from in place;
    everything else -- numpy arrays (any strides: memory maps, clipped views), pageable tensors -- goes
    through the pinned staging ring chunk by chunk, so no whole-raster host copy is ever made."""
    import torch
    if getattr(values, "lazy_rows", False) or getattr(values, "is_packed_raster", False):
        # dataset.TimeConcat (several files along time): row slices are NumPy views / lazy windows of the parts, turned
        # into arrays by np.copyto inside the staging threads -- the whole raster is never materialised on the host
        return None, values
    if isinstance(values, np.ndarray) or not type(values).__module__.startswith("torch"):
        arr = np.asarray(values)
        if arr.dtype not in (np.float32, np.float64) or not arr.dtype.isnative:
            arr = arr.astype(np.float64 if arr.dtype.itemsize > 4 or arr.dtype.kind != "f" else np.float32)
        t = _pinned_piece(torch, arr, torch.float64 if arr.dtype == np.float64 else torch.float32)
        if t is not None:                      # a NumPy view of page-locked memory (tensor.numpy() of a pinned tensor)
            return t, None
        return None, arr
    t = values
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if t.is_pinned() and t.is_contiguous():
        return t, None
    return None, t.detach().numpy()


def chunk_rows(n_rows: int, row_bytes: int, chunk_bytes: int):
    """[(r0, r1)] covering [0, n_rows) with about ``chunk_bytes`` per chunk (at least one row)."""
    step = max(1, int(chunk_bytes // max(1, row_bytes)))
    return [(r, min(n_rows, r + step)) for r in range(0, n_rows, step)]


def _drop(cache: dict) -> None:
    """Empty a cache of pinned / device staging buffers.  Copies and kernels of an earlier feed may still be using them
    (the feeds are asynchronous; ``aggregate_dataset`` synchronises before it returns, a caller that drives ``feed_*`` on
    its own stream need not have), so the device is synchronised first."""
    if cache:
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        cache.clear()


_RINGS = {}          # (dtype, slot_elems, n_slots) -> pinned slots, kept across calls: pinning 512 MB costs ~100 ms


class _Staging:
    """Ring of pinned buffers filled by worker threads (pageable sources only)."""

    def __init__(self, torch, dtype, slot_elems: int, n_slots: int, n_threads: int):
        self.torch = torch
        key = (str(dtype), int(slot_elems), int(n_slots))
        if key not in _RINGS:
            _drop(_RINGS)                              # one ring at a time: it is pinned memory
            _RINGS[key] = [torch.empty(slot_elems, dtype=dtype, pin_memory=True) for _ in range(n_slots)]
        self.slots = _RINGS[key]
        self.events = [None] * n_slots                 # copy-done event of the slot's last use
        self.pool = ThreadPoolExecutor(max_workers=max(1, n_threads))

    def fill(self, slot: int, chunk: np.ndarray):
        ev = self.events[slot]
        if ev is not None:
            ev.synchronize()                           # the previous async copy out of this slot
        dst = self.slots[slot][: chunk.size]
        np.copyto(dst.numpy().reshape(chunk.shape), chunk)   # (strided) memcpy, GIL released
        return dst

    def close(self):
        self.pool.shutdown(wait=True)


def plan_windows(cuts, n_rows: int, slot_rows: int):
    """[(r0, r1)] covering [0, n_rows): every window ends at one of ``cuts`` (ascending rows, the last one is
    n_rows) and holds at most ``slot_rows`` rows where the cuts allow it (else the shortest window that reaches the
    next cut).  Greedy: the farthest cut inside the slot."""
    cuts = np.unique(np.asarray(cuts, dtype=np.int64))
    cuts = cuts[(cuts > 0) & (cuts <= n_rows)]
    if len(cuts) == 0 or cuts[-1] != n_rows:
        cuts = np.append(cuts, n_rows)
    out, r0 = [], 0
    while r0 < n_rows:
        j = int(np.searchsorted(cuts, r0 + slot_rows, side="right")) - 1
        r1 = int(cuts[j]) if j >= 0 and cuts[j] > r0 else int(cuts[int(np.searchsorted(cuts, r0, side="right"))])
        out.append((r0, r1))
        r0 = r1
    return out


def _source_breaks(values, n_rows: int):
    """Rows where a lazily concatenated source changes part (chunks are cut there, so that every chunk is one
    part's own slice)."""
    starts = getattr(values, "starts", None)
    if starts is None:
        return np.zeros(0, dtype=np.int64)
    b = np.asarray(starts, dtype=np.int64)
    return b[(b > 0) & (b < n_rows)]


def _chunks_of(r0: int, r1: int, row_bytes: int, chunk_bytes: int, breaks):
    """Row chunks of [r0, r1) of about chunk_bytes, never straddling a source break."""
    edges = np.unique(np.concatenate([[r0, r1], breaks[(breaks > r0) & (breaks < r1)]])).astype(np.int64)
    out = []
    for a, b in zip(edges[:-1], edges[1:]):
        out += [(int(a) + x, int(a) + y) for x, y in chunk_rows(int(b - a), row_bytes, chunk_bytes)]
    return out


_RING_RASTERS = {}       # (device, dtype, slot elements, slots) -> device windows, kept across calls like _DEVICE_RASTERS


def _pinned_piece(torch, piece, tdtype):
    """A pinned torch view of a host chunk that can be copied from in place, else None."""
    if type(piece).__module__.startswith("torch"):
        t = piece
    elif isinstance(piece, np.ndarray) and piece.flags.c_contiguous and piece.dtype.isnative and piece.flags.writeable:
        t = torch.from_numpy(piece)
    else:
        return None
    if t.dtype != tdtype or not t.is_contiguous():
        return None
    try:
        return t if t.is_pinned() else None
    except Exception:
        return None


def _feed_ring(torch, runner, values, host, host_np, T: int, n_cells: int, tdtype, cuts, comp, copy, k1_events, chunk_bytes):
    """feed_and_run for a record longer than the device budget: windows of the time axis go through a ring of
    device slots (module docstring).  Returns (result, ring tensor, stats)."""
    dev = runner.device
    item = torch.empty(0, dtype=tdtype).element_size()
    row_bytes = n_cells * item
    slot_rows = max(1, int(OPTIONS["ring_slot_bytes"] // row_bytes))
    windows = plan_windows(cuts, T, slot_rows)
    slot_rows = max(r1 - r0 for r0, r1 in windows)
    n_slots = int(max(2, min(OPTIONS["ring_slots"], len(windows))))
    key = (dev.index, str(tdtype), slot_rows * n_cells, n_slots)
    ring = _RING_RASTERS.get(key)
    if ring is None:
        _drop(_RING_RASTERS)
        _drop(_DEVICE_RASTERS)
        ring = torch.empty((n_slots, slot_rows, n_cells), dtype=tdtype, device=dev)
        _RING_RASTERS[key] = ring
    pinned = host is not None
    cb = chunk_bytes or OPTIONS["chunk_bytes" if pinned else "staging_chunk_bytes"]
    breaks = _source_breaks(values, T)
    staging = None
    if not pinned:
        staging = _Staging(torch, tdtype, max(1, int(cb // row_bytes)) * n_cells, OPTIONS["staging_slots"], OPTIONS["staging_threads"])
    work = []                                                  # (window index, r0, r1, last chunk of its window)
    for k, (w0, w1) in enumerate(windows):
        cs = _chunks_of(w0, w1, row_bytes, cb, breaks)
        work += [(k, a, b, j == len(cs) - 1) for j, (a, b) in enumerate(cs)]
    consumed = [None] * n_slots                                # last kernel that read the slot
    copy.wait_stream(comp)
    ev_first, ev_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_first.record(copy)
    runner.begin_streamed(comp)
    launches = direct = 0
    pk_dev, pk_placed, n_packed, h2d_saved = [None] * 3, [None] * 3, 0, 0     # device staging of packed pieces
    try:
        futs = {}
        ahead = len(staging.slots) if staging else 0
        n_staged = 0

        def submit(i):
            nonlocal n_staged
            _, a, b, _last = work[i]
            piece = host_np[a:b]
            if getattr(piece, "is_packed_raster", False):
                # CF-packed integers: a pinned piece is copied as stored and unpacked into the window on the device
                st = piece.stored
                sdt = np.dtype(str(st.dtype).replace("torch.", ""))
                ts = _pinned_piece(torch, st, getattr(torch, sdt.name, None)) if sdt.str[1:] in _TILE_DTYPES else None
                if ts is not None:
                    futs[i] = ("packed", ts, piece, sdt)
                    return
            t = _pinned_piece(torch, piece, tdtype)
            if t is not None:
                futs[i] = ("direct", t)
            else:
                futs[i] = ("staged", staging.pool.submit(staging.fill, n_staged % ahead, piece), n_staged % ahead)
                n_staged += 1

        if staging:
            for i in range(min(ahead, len(work))):
                submit(i)
        for i, (k, a, b, last) in enumerate(work):
            slot = k % n_slots
            w0 = windows[k][0]
            sslot = None
            packed_piece = None
            if staging:
                got = futs.pop(i)
                if got[0] == "direct":
                    src = got[1].view(b - a, n_cells)
                    direct += 1
                elif got[0] == "packed":
                    src, packed_piece, sdt = got[1].reshape(-1), got[2], got[3]
                    direct += 1
                else:
                    src, sslot = got[1].result().view(b - a, n_cells), got[2]
            else:
                src = host[a:b]
            with torch.cuda.stream(copy):
                if packed_piece is not None:
                    j = n_packed % len(pk_placed)
                    if pk_dev[j] is None or pk_dev[j].dtype != src.dtype or pk_dev[j].numel() < src.numel():
                        if pk_placed[j] is not None:
                            pk_placed[j].synchronize()
                        pk_dev[j] = torch.empty(max(src.numel(), max(1, int(cb // row_bytes)) * n_cells), dtype=src.dtype, device=dev)
                    if pk_placed[j] is not None:
                        copy.wait_event(pk_placed[j])          # the unpack kernel that last read this staging buffer
                    pk_dev[j][: src.numel()].copy_(src, non_blocking=True)
                else:
                    if a == w0 and consumed[slot] is not None:
                        copy.wait_event(consumed[slot])        # the slot's previous window has been scanned
                    ring[slot, a - w0:b - w0].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            if packed_piece is not None:
                comp.wait_event(ev)                            # (comp runs behind the kernels of the slot's previous window)
                _, Yp, Xp = packed_piece.shape
                from . import _lib
                _lib.check(_lib.lib().agf_tile_place_run(
                    pk_dev[j].data_ptr(), _TILE_DTYPES[sdt.str[1:]], b - a, Yp, Xp, n_cells, Xp, 1, ring[slot].data_ptr(),
                    _lib.F64 if tdtype == torch.float64 else _lib.F32, n_cells, Xp, a - w0, 0, 0, 1, float(packed_piece.scale),
                    float(packed_piece.offset), int(packed_piece.fill is not None),
                    float(packed_piece.fill if packed_piece.fill is not None else 0.0), slot_rows, comp.cuda_stream))
                pe = torch.cuda.Event()
                pe.record(comp)
                pk_placed[j] = pe
                n_packed += 1
                h2d_saved += (b - a) * row_bytes - src.numel() * src.element_size()
            if staging:
                if sslot is not None:
                    staging.events[sslot] = ev
                if i + ahead < len(work):
                    submit(i + ahead)
            comp.wait_event(ev)
            launches += runner.feed(ring[slot], b, comp, k1_events, row0=w0, flush=last)
            if last:
                ce = torch.cuda.Event()
                ce.record(comp)
                consumed[slot] = ce
        ev_last.record(copy)
        res = runner.finish_streamed(None, comp)
    finally:
        if staging:
            staging.close()
    ring.record_stream(comp)
    stats = dict(chunks=len(work), pinned=bool(pinned), h2d_bytes=T * row_bytes - h2d_saved, k1_launches=launches,
                 copy_events=(ev_first, ev_last), ring=True, ring_slots=n_slots, ring_slot_rows=slot_rows,
                 ring_bytes=int(ring.numel() * item), windows=len(windows), direct_chunks=direct, unpacked_chunks=n_packed)
    return res, ring, stats


def feed_and_run(runner, values, n_cells: int, stream=None, k1_events: Optional[list] = None,
                 chunk_bytes: Optional[int] = None, stats: Optional[dict] = None):
    """Run ``runner`` (an ``engine.StageRunner``) over a HOST raster ``values[T, ...cells]``.

    Returns (StageResult, device raster [T, n_cells]).  Everything is asynchronous with respect to
    the host except the staging memcpys of a pageable source."""
    import torch
    if getattr(values, "is_chunked_raster", False):
        return feed_chunked(runner, values, n_cells, stream, k1_events, stats)
    if getattr(values, "is_packed_raster", False):
        over = int(np.prod(values.shape)) * values.dtype.itemsize > OPTIONS["device_raster_budget_bytes"]
        if not (over and hasattr(runner, "window_cuts") and runner.window_cuts() is not None):
            return feed_packed(runner, values, n_cells, stream, k1_events, stats)
    host, host_np = _host_source(values)
    pinned = host is not None
    T = int(host.shape[0] if pinned else host_np.shape[0])
    if pinned:
        host = host.reshape(T, n_cells)
        tdtype = host.dtype
    else:
        if int(np.prod(host_np.shape[1:])) != n_cells:
            raise ValueError(f"raster of shape {host_np.shape} does not have {n_cells} cells per step")
        tdtype = torch.float64 if host_np.dtype == np.float64 else torch.float32
    dev = runner.device
    comp = torch.cuda.current_stream(dev) if stream is None else stream
    copy = _copy_stream(dev)
    global LAST_STATS
    if T * n_cells * (8 if tdtype == torch.float64 else 4) > OPTIONS["device_raster_budget_bytes"]:
        cuts = runner.window_cuts() if hasattr(runner, "window_cuts") else None
        if cuts is not None:
            res, ring, LAST_STATS = _feed_ring(torch, runner, values, host, host_np, T, n_cells, tdtype, cuts, comp, copy,
                                               k1_events, chunk_bytes)
            if stats is not None:
                stats.update(LAST_STATS)
            return res, ring
    raster = _device_raster(torch, dev, tdtype, T, n_cells)
    row_bytes = n_cells * raster.element_size()
    chunks = chunk_rows(T, row_bytes, chunk_bytes or OPTIONS["chunk_bytes" if pinned else "staging_chunk_bytes"])
    staging = None
    if not pinned:
        slot_elems = max(r1 - r0 for r0, r1 in chunks) * n_cells
        staging = _Staging(torch, tdtype, slot_elems, OPTIONS["staging_slots"], OPTIONS["staging_threads"])

    copy.wait_stream(comp)                    # the raster buffer may be a recycled block still in use
    ev_first, ev_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_first.record(copy)
    runner.begin_streamed(comp)
    launches = 0
    try:
        futs = {}
        ahead = len(staging.slots) if staging else 0

        def submit(i):
            r0, r1 = chunks[i]
            futs[i] = staging.pool.submit(staging.fill, i % ahead, host_np[r0:r1])

        if staging:
            for i in range(min(ahead, len(chunks))):
                submit(i)
        for i, (r0, r1) in enumerate(chunks):
            if staging:
                src = futs.pop(i).result().view(r1 - r0, n_cells)
            else:
                src = host[r0:r1]
            with torch.cuda.stream(copy):
                raster[r0:r1].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            if staging:
                staging.events[i % ahead] = ev
                if i + ahead < len(chunks):
                    submit(i + ahead)
            comp.wait_event(ev)
            launches += runner.feed(raster, r1, comp, k1_events)
        ev_last.record(copy)
        res = runner.finish_streamed(raster, comp)
    finally:
        if staging:
            staging.close()
    raster.record_stream(comp)
    LAST_STATS = dict(chunks=len(chunks), pinned=bool(pinned), h2d_bytes=T * row_bytes, k1_launches=launches,
                      copy_events=(ev_first, ev_last))      # elapsed_time() once the caller has synchronised
    if stats is not None:
        stats.update(LAST_STATS)
    return res, raster


LAST_STATS: dict = {}          # what the most recent feed did (tests, bench bookkeeping)


# ---------------------------------------------------------------------------------------------
# chunked stores (zarr): decode on host threads -> pinned slot -> device slot -> placement kernel
# ---------------------------------------------------------------------------------------------
_CHUNK_RINGS = {}        # (device, slot_bytes, n_slots) -> (pinned slots, device slots), kept across calls

_TILE_DTYPES = {"f4": 0, "f8": 1, "i2": 2, "i4": 3, "u1": 4, "i1": 5, "u2": 6}      # AGF_F32 .. AGF_U16


def _chunk_ring(torch, dev, slot_bytes: int, n_slots: int, de: bool):
    """Staging of ``n_slots`` chunks: pinned host bytes + a device copy; with ``de`` (device decompression)
    the pinned / first device buffer hold the COMPRESSED file (a little headroom for incompressible
    chunks) and two more device buffers hold the inflated and the unshuffled chunk."""
    key = (dev.index, int(slot_bytes), int(n_slots), bool(de))
    if key not in _CHUNK_RINGS:
        _drop(_CHUNK_RINGS)
        cap = slot_bytes + (slot_bytes // 64 + 65536 if de else 0)
        cap = (cap + 15) // 16 * 16
        ring = {"cap": cap,
                "pinned": [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(n_slots)],
                "dev": [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(n_slots)]}
        if de:
            ring["inflated"] = [torch.empty(slot_bytes, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
            ring["plain"] = [torch.empty(slot_bytes, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
            ring["aligned_cap"] = cap + cap // 4 + (1 << 20)
            ring["aligned"] = None                     # allocated on first use (device_decompress_align)
        _CHUNK_RINGS[key] = ring
    return _CHUNK_RINGS[key]


_PENDING_CHECKS = []     # device booleans "every engine operation produced the bytes its frame promised"


def check_device_decompress() -> None:
    """Call after synchronising the stream a chunked feed ran on: raises if any decompression-engine
    operation of the feeds since the last call wrote a different number of bytes than its Blosc frame said."""
    pending, _PENDING_CHECKS[:] = list(_PENDING_CHECKS), []
    for ok, what in pending:
        if not bool(ok):
            raise IOError(f"{what}: the decompression engine produced a stream of unexpected length (corrupt chunk?)")


def device_decompress_caps():
    """(algorithm mask, bytes per operation) of the device's decompression engine (0, 0: none)."""
    import ctypes as C
    from . import _lib
    mask, mx = C.c_int32(0), C.c_int64(0)
    _lib.check(_lib.lib().agf_decompress_caps(C.byref(mask), C.byref(mx)))
    return int(mask.value), int(mx.value)


def feed_chunked(runner, src, n_cells: int, stream=None, k1_events: Optional[list] = None,
                 stats: Optional[dict] = None, device=None):
    """``feed_and_run`` for a ``zarrio.ChunkedRaster``: every storage chunk is brought into a pinned slot by
    a host thread (file read, decompression and copies all release the GIL), copied to the device as it is
    stored, and placed into the time-major raster by ``agf_tile_place_run`` (axis permutation, CF unpacking
    and fill -> NaN on the device).

    * Chunks that are runs of whole raster rows and need no decoding are copied straight into the raster.
    * Blosc-LZ4 chunks (zarr v2's default compressor) are NOT inflated on the host when the device has a
      decompression engine (``OPTIONS["device_decompress"]``): the compressed file crosses PCIe, its LZ4
      streams are inflated by ``agf_decompress_lz4_run`` and un-shuffled by ``agf_unshuffle_run``.
    * Tiles arrive time chunk by time chunk, so for a time-major store the temporal kernels of a stripe
      start as soon as its rows are complete; a time-contiguous store (chunks ``[s, s, T]``) is scanned once
      its last tile has landed.

    ``runner=None`` only builds the device raster (``engine.to_device``)."""
    import ctypes as C
    import torch
    from . import _lib
    L = _lib.lib()
    T, Y, X = (int(v) for v in src.shape)
    if Y * X != n_cells:
        raise ValueError(f"raster of shape {src.shape} does not have {n_cells} cells per step")
    dev = runner.device if runner is not None else (device or torch.device("cuda", torch.cuda.current_device()))
    comp = torch.cuda.current_stream(dev) if stream is None else stream
    copy = _copy_stream(dev)
    tdtype = torch.float64 if src.dtype == np.float64 else torch.float32
    dst_code = _lib.F64 if tdtype == torch.float64 else _lib.F32
    raster = _device_raster(torch, dev, tdtype, T, n_cells)
    sdt = src.array.dtype.newbyteorder("=")
    code = _TILE_DTYPES.get(sdt.str[1:])
    if code is None:
        raise NotImplementedError(f"no tile decoder for stored dtype {sdt}")
    tiles = src.tiles()
    slot_bytes = src.slot_elems * sdt.itemsize
    de, max_len = False, 0
    if OPTIONS.get("device_decompress", True) and src.array.blosc_only:
        mask, max_len = device_decompress_caps()
        de = bool(mask & 4) and max_len > 0
    n_slots = int(max(2, min(OPTIONS["staging_slots"], OPTIONS["chunked_ring_bytes"] // max(1, slot_bytes))))
    ring = _chunk_ring(torch, dev, slot_bytes, n_slots, de)
    pinned, dslots = ring["pinned"], ring["dev"]
    bytes_np = [p.numpy() for p in pinned]
    views = [b[:slot_bytes].view(sdt) for b in bytes_np]
    tviews = [p[:slot_bytes].view(tdtype) for p in pinned] if sdt == src.dtype else None
    h2d_done = [None] * n_slots          # last copy out of the pinned slot
    placed = [None] * n_slots            # last kernel reading the slot's device buffers
    fill_scalar = None

    def fill(i):
        slot = i % n_slots
        ev = h2d_done[slot]
        if ev is not None:
            ev.synchronize()
        if de:
            return src.load_stored(tiles[i], bytes_np[slot], max_len, bool(OPTIONS.get("inline_chunk_tables", True)))
        return ("host",) if src.load(tiles[i], views[slot]) else None

    def place(ptr, tile):
        nt, ny, nx = tile.extent
        _lib.check(L.agf_tile_place_run(
            ptr + tile.offset * sdt.itemsize, code, nt, ny, nx, tile.st, tile.sy, tile.sx, raster.data_ptr(), dst_code,
            n_cells, X, tile.t0, tile.y0, tile.x0, int(src.packed), float(src.scale), float(src.offset),
            int(src.fill is not None), float(src.fill if src.fill is not None else 0.0), T, comp.cuda_stream))

    i64p = C.POINTER(C.c_int64)
    align = int(OPTIONS.get("device_decompress_align", 0) or 0)
    copy.wait_stream(comp)
    ev_first, ev_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_first.record(copy)
    if runner is not None:
        runner.begin_streamed(comp)
    launches = places = direct = absent = inflated = engine_ops = raw_copies = 0
    h2d_bytes = 0
    pool = ThreadPoolExecutor(max_workers=max(1, int(OPTIONS["staging_threads"])))
    try:
        futs = {i: pool.submit(fill, i) for i in range(min(n_slots, len(tiles)))}
        for i, tile in enumerate(tiles):
            slot = i % n_slots
            got = futs.pop(i).result()
            nt, ny, nx = tile.extent
            if got is None:
                if fill_scalar is None:
                    fv = src.array.fill_value
                    fill_scalar = float(src.decode_host(np.full(1, 0 if fv is None else fv, sdt))[0])
                with torch.cuda.stream(comp):
                    raster.view(T, Y, X)[tile.t0:tile.t1, tile.y0:tile.y1, tile.x0:tile.x1] = fill_scalar
                absent += 1
            elif got[0] == "host" and tviews is not None and src.direct_rows(tile):
                with torch.cuda.stream(copy):
                    raster[tile.t0:tile.t1].copy_(tviews[slot][tile.offset: tile.offset + nt * n_cells].view(nt, n_cells),
                                                  non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                h2d_done[slot] = ev
                comp.wait_event(ev)
                h2d_bytes += nt * n_cells * sdt.itemsize
                direct += 1
            else:
                n_up = slot_bytes if got[0] == "host" else got[1]
                with torch.cuda.stream(copy):
                    if placed[slot] is not None:
                        copy.wait_event(placed[slot])
                    dslots[slot][:n_up].copy_(pinned[slot][:n_up], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                h2d_done[slot] = ev
                comp.wait_event(ev)
                h2d_bytes += n_up
                if got[0] == "host":
                    place(dslots[slot].data_ptr(), tile)
                else:
                    plan = got[2]
                    if plan.kind == "memcpy":
                        place(dslots[slot].data_ptr() + 16, tile)
                    else:
                        infl, plain = ring["inflated"][slot], ring["plain"][slot]
                        n_ops = len(plan.src_off)
                        with torch.cuda.stream(comp):
                            if n_ops:
                                src_ptr, src_off = dslots[slot].data_ptr(), plan.src_off
                                if align > 1:
                                    padded = (plan.src_len + (align - 1)) // align * align
                                    src_off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64))
                                    if int(src_off[-1] + padded[-1]) <= ring["aligned_cap"]:
                                        if ring["aligned"] is None:
                                            ring["aligned"] = [torch.empty(ring["aligned_cap"], dtype=torch.uint8, device=dev)
                                                               for _ in range(n_slots)]
                                        moved = torch.from_numpy(np.stack([plan.src_off, src_off, plan.src_len])).to(dev, non_blocking=True)
                                        _lib.check(L.agf_copy_segments_run(src_ptr, ring["aligned"][slot].data_ptr(), moved.data_ptr(),
                                                                           n_ops, comp.cuda_stream))
                                        src_ptr = ring["aligned"][slot].data_ptr()
                                    else:
                                        src_off = plan.src_off
                                actual = torch.empty(n_ops, dtype=torch.int32, device=dev)
                                _lib.check(L.agf_decompress_lz4_run(
                                    src_ptr, src_off.ctypes.data_as(i64p), plan.src_len.ctypes.data_as(i64p),
                                    infl.data_ptr(), plan.dst_off.ctypes.data_as(i64p), plan.dst_len.ctypes.data_as(i64p),
                                    n_ops, actual.data_ptr(), comp.cuda_stream))
                                if plan.inline is not None:
                                    want = dslots[slot][plan.inline[0]: plan.inline[0] + 4 * n_ops].view(torch.int32)
                                else:
                                    want = torch.from_numpy(plan.dst_len.astype(np.int32)).to(dev, non_blocking=True)
                                _PENDING_CHECKS.append(((actual == want).all(), src.array.chunk_path(tile.index)))
                            if plan.raw.shape[1]:                          # streams Blosc stored uncompressed
                                if plan.inline is not None:
                                    table_ptr = dslots[slot].data_ptr() + plan.inline[1]
                                else:
                                    table = torch.from_numpy(plan.raw).to(dev, non_blocking=True)
                                    table_ptr = table.data_ptr()
                                _lib.check(L.agf_copy_segments_run(dslots[slot].data_ptr(), infl.data_ptr(), table_ptr,
                                                                   plan.raw.shape[1], comp.cuda_stream))
                                raw_copies += int(plan.raw.shape[1])
                        if plan.shuffled:
                            _lib.check(L.agf_unshuffle_run(infl.data_ptr(), plain.data_ptr(), plan.nbytes, plan.typesize,
                                                           plan.blocksize, comp.cuda_stream))
                            place(plain.data_ptr(), tile)
                        else:
                            place(infl.data_ptr(), tile)
                        inflated += 1
                        engine_ops += n_ops
                pe = torch.cuda.Event()
                pe.record(comp)
                placed[slot] = pe
                places += 1
            if i + n_slots < len(tiles):
                futs[i + n_slots] = pool.submit(fill, i + n_slots)
            last_of_rows = i + 1 == len(tiles) or tiles[i + 1].t0 != tile.t0
            if runner is not None and last_of_rows:
                launches += runner.feed(raster, tile.t1, comp, k1_events)
        ev_last.record(copy)
        res = runner.finish_streamed(raster, comp) if runner is not None else None
    finally:
        pool.shutdown(wait=True)
    raster.record_stream(comp)
    global LAST_STATS
    LAST_STATS = dict(chunks=len(tiles), pinned=False, chunked=True, h2d_bytes=h2d_bytes, k1_launches=launches,
                      place_launches=places, direct_copies=direct, absent_chunks=absent, ring_slots=n_slots,
                      device_decompress=bool(de), inflated_on_device=inflated, engine_ops=engine_ops, raw_streams=raw_copies,
                      copy_events=(ev_first, ev_last))
    if stats is not None:
        stats.update(LAST_STATS)
    return res, raster


# ---------------------------------------------------------------------------------------------
# packed integers in host memory (dataset.PackedRaster): copied as stored, decoded on the device
# ---------------------------------------------------------------------------------------------
_PACKED_DEV = {}         # (device, dtype, slot elements, slots) -> device staging for stored chunks


def feed_packed(runner, src, n_cells: int, stream=None, k1_events: Optional[list] = None, stats: Optional[dict] = None,
                device=None):
    """``feed_and_run`` for a ``dataset.PackedRaster``: row chunks of the STORED integers go to a small device ring
    (pinned sources are copied from in place, pageable ones through the pinned staging ring), ``agf_tile_place_run``
    unpacks them into the float raster on the compute stream (CF scale / offset in double, fill -> NaN), and the
    temporal kernels of a stripe start as soon as its rows are decoded.  Half the PCIe bytes of a float32 raster for
    int16 data.  ``runner=None`` only builds the device raster."""
    import torch
    from . import _lib
    L = _lib.lib()
    T, Y, X = src.shape
    if Y * X != n_cells:
        raise ValueError(f"raster of shape {src.shape} does not have {n_cells} cells per step")
    dev = runner.device if runner is not None else (device or torch.device("cuda", torch.cuda.current_device()))
    comp = torch.cuda.current_stream(dev) if stream is None else stream
    copy = _copy_stream(dev)
    tdtype = torch.float64 if src.dtype == np.float64 else torch.float32
    dst_code = _lib.F64 if tdtype == torch.float64 else _lib.F32
    stored = src.stored
    is_t = type(stored).__module__.startswith("torch")
    sdt = np.dtype(str(stored.dtype).replace("torch.", "")) if is_t else np.dtype(stored.dtype).newbyteorder("=")
    code = _TILE_DTYPES.get(sdt.str[1:])
    if code is None or (not is_t and not np.dtype(stored.dtype).isnative):
        raise NotImplementedError(f"no tile decoder for stored dtype {stored.dtype}")
    st_tdtype = getattr(torch, sdt.name)
    pinned = bool(is_t and stored.is_pinned() and stored.is_contiguous())
    host = stored.reshape(T, n_cells) if pinned else None
    host_np = None if pinned else (stored.numpy() if is_t else stored).reshape(T, n_cells)
    raster = _device_raster(torch, dev, tdtype, T, n_cells)
    row_bytes = n_cells * sdt.itemsize
    chunks = chunk_rows(T, row_bytes, OPTIONS["chunk_bytes" if pinned else "staging_chunk_bytes"])
    slot_rows = max(r1 - r0 for r0, r1 in chunks)
    n_dev = 3
    key = (dev.index, sdt.str, slot_rows * n_cells, n_dev)
    if key not in _PACKED_DEV:
        _drop(_PACKED_DEV)
        _PACKED_DEV[key] = [torch.empty(slot_rows * n_cells, dtype=st_tdtype, device=dev) for _ in range(n_dev)]
    dslots = _PACKED_DEV[key]
    staging = None if pinned else _Staging(torch, st_tdtype, slot_rows * n_cells, OPTIONS["staging_slots"], OPTIONS["staging_threads"])
    placed = [None] * n_dev
    copy.wait_stream(comp)
    ev_first, ev_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_first.record(copy)
    if runner is not None:
        runner.begin_streamed(comp)
    launches = 0
    try:
        futs = {}
        ahead = len(staging.slots) if staging else 0

        def submit(i):
            r0, r1 = chunks[i]
            futs[i] = staging.pool.submit(staging.fill, i % ahead, host_np[r0:r1])

        if staging:
            for i in range(min(ahead, len(chunks))):
                submit(i)
        for i, (r0, r1) in enumerate(chunks):
            slot = i % n_dev
            n = (r1 - r0) * n_cells
            hsrc = futs.pop(i).result()[:n] if staging else host[r0:r1].reshape(-1)
            with torch.cuda.stream(copy):
                if placed[slot] is not None:
                    copy.wait_event(placed[slot])              # the unpack kernel that last read this device slot
                dslots[slot][:n].copy_(hsrc, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            if staging:
                staging.events[i % ahead] = ev
                if i + ahead < len(chunks):
                    submit(i + ahead)
            comp.wait_event(ev)
            _lib.check(L.agf_tile_place_run(
                dslots[slot].data_ptr(), code, r1 - r0, Y, X, n_cells, X, 1, raster.data_ptr(), dst_code, n_cells, X, r0, 0, 0,
                1, float(src.scale), float(src.offset), int(src.fill is not None),
                float(src.fill if src.fill is not None else 0.0), T, comp.cuda_stream))
            pe = torch.cuda.Event()
            pe.record(comp)
            placed[slot] = pe
            if runner is not None:
                launches += runner.feed(raster, r1, comp, k1_events)
        ev_last.record(copy)
        res = runner.finish_streamed(raster, comp) if runner is not None else None
    finally:
        if staging:
            staging.close()
    raster.record_stream(comp)
    global LAST_STATS
    LAST_STATS = dict(chunks=len(chunks), pinned=pinned, packed=True, h2d_bytes=T * row_bytes, k1_launches=launches,
                      place_launches=len(chunks), copy_events=(ev_first, ev_last))
    if stats is not None:
        stats.update(LAST_STATS)
    return res, raster


_DEVICE_RASTERS = {}     # (device, dtype, elements) -> the device copy of the last host raster of that shape


def _device_raster(torch, dev, tdtype, T: int, n_cells: int):
    """Device buffer for a host raster.  A yearly loop feeds same-shaped rasters again and again; taking the
    36 GB block from torch's caching allocator each time worked until a smaller buffer (partial records, X)
    was carved out of the cached block between two calls -- the next call then paid a fresh cudaMalloc /
    cudaFree of tens of GB (0.3-1.2 s, seen as sporadic slow calls in bench e2e).  The buffer of the most
    recent shape is therefore kept here; ``release_device_rasters()`` returns it."""
    if not OPTIONS.get("keep_device_raster", True):
        return torch.empty((T, n_cells), dtype=tdtype, device=dev)
    key = (dev.index, str(tdtype), T * n_cells)
    buf = _DEVICE_RASTERS.get(key)
    if buf is None:
        _drop(_DEVICE_RASTERS)                   # one shape at a time
        buf = torch.empty(T * n_cells, dtype=tdtype, device=dev)
        _DEVICE_RASTERS[key] = buf
    return buf.view(T, n_cells)


def release_device_rasters() -> None:
    """Give the cached device raster (and the pinned staging rings) back."""
    for cache in (_DEVICE_RASTERS, _RING_RASTERS, _RINGS, _CHUNK_RINGS, _PACKED_DEV):
        _drop(cache)


_COPY_STREAMS = {}


def _copy_stream(device):
    import torch
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def bind_host_to_device(device_index: int) -> dict:
    """Restrict this process to the CPUs that are local to GPU ``device_index`` (NVML's ideal CPU
    affinity), so that pinned host buffers allocated afterwards -- the caller's rasters, the staging
    ring -- land on the GPU's NUMA node and host->device copies do not cross the socket interconnect.
    On an 8-GPU box the un-bound ranks of a sharded run copied at 29 GB/s per GPU instead of 55 GB/s.
    Call it once per rank before allocating host memory.  Returns what it did (never raises)."""
    import os
    info = {"device": device_index, "bound": False}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_words = (os.cpu_count() + 63) // 64 + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        ideal = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        info.update(ideal=len(ideal), allowed=len(allowed))
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as exc:                                  # no NVML / restricted container: stay unbound
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info
