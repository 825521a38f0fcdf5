"""Launcher: runs planned stages (spec.py) and the CSR regional average through the C-ABI.

PyTorch is used for plumbing only -- device buffers, streams, host<->device copies; every
number is produced by the kernels in ``csrc/`` (no torch math on the data path, no CPU path).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib
from .spec import ProgramSpec, Stage, build_desc
from .weights import HostCSR


# Tunables (tests pin target_stripes to exercise the stripe/merge path; 0 = library heuristic).
OPTIONS = {
    "target_stripes": 0,
    # aggregate._panel_frame: select / gather / transpose the panel's rows on the device and join the region ids without
    # DataFrame.merge.  Frame-identical to the literal NumPy + merge route (tests/test_panel.py on CPU tensors; the whole GPU
    # suite ran green through it on a B200 in round 2, daily-panel call 3.97 s -> 1.84 s).  AGF_DEVICE_PANEL=0 selects the
    # literal route (kept as the cross-check the tests compare against).
    "device_panel_frame": bool(int(__import__("os").environ.get("AGF_DEVICE_PANEL", "1") or 0)),
    # Panels with many periods (hourly raster -> daily panel): temporal scan + regional average in ONE kernel
    # (agf_temporal_regional_run, csrc/agf_regional.cuh) instead of K1 -> X -> K2.  "auto": when the program has a regional
    # instantiation and the panel has at least ``regional_min_periods`` periods; True / False force it on / off.
    "regional": {"0": False, "1": True}.get(__import__("os").environ.get("AGF_REGIONAL", ""), "auto"),
    "regional_min_periods": 32,
}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("aggfly_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def _tdtype(np_dtype):
    torch = _torch()
    return torch.float64 if np.dtype(np_dtype) == np.float64 else torch.float32


def to_device(values, device=None):
    """Raster -> contiguous device tensor [T, cells...] (numpy, CPU tensor or CUDA tensor)."""
    torch = _torch()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    if getattr(values, "is_chunked_raster", False):                # zarr store: chunks decoded + placed on the device
        from . import stream as _stream
        T, Y, X = values.shape
        _, raster = _stream.feed_chunked(None, values, Y * X, device=device)
        torch.cuda.current_stream(device).synchronize()
        _stream.check_device_decompress()
        return raster.view(T, Y, X).clone()                         # the streamed buffer is recycled by the next feed
    if getattr(values, "is_packed_raster", False):                 # packed integers: copied as stored, decoded on the device
        from . import stream as _stream
        T, Y, X = values.shape
        _, raster = _stream.feed_packed(None, values, Y * X, device=device)
        torch.cuda.current_stream(device).synchronize()
        return raster.view(T, Y, X).clone()
    if getattr(values, "lazy_rows", False):
        values = np.asarray(values)
    if isinstance(values, np.ndarray):
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        values = torch.from_numpy(np.ascontiguousarray(values))
    if values.dtype not in (torch.float32, torch.float64):
        values = values.to(torch.float64)
    return values.to(device, non_blocking=True).contiguous()


class Program:
    """RAII wrapper of an ``agf_program_t``."""

    def __init__(self, spec: ProgramSpec, out_dtype, n_cells: int, target_stripes: int = 0):
        self.spec = spec
        desc, self._keep = build_desc(spec, out_dtype)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().agf_program_create(C.byref(self.handle), C.byref(desc), n_cells, target_stripes))
        self.info = _lib.ProgramInfo()
        _lib.check(_lib.lib().agf_program_info(self.handle, C.byref(self.info)))

    def stripe_rows(self, s: int) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().agf_program_stripe_rows(self.handle, s, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self.handle:
            _lib.lib().agf_program_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceCSR:
    """CSR arrays resident on the device + the ``agf_csr_t`` handle."""

    def __init__(self, host: HostCSR, device=None):
        torch = _torch()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.host = host
        self.row_ptr = torch.from_numpy(host.row_ptr).to(device)
        self.cell_idx = torch.from_numpy(host.cell_idx).to(device)
        self.w = torch.from_numpy(host.w).to(device)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().agf_csr_create(C.byref(self.handle), host.n_regions, host.n_cells, host.nnz,
                                             self.row_ptr.data_ptr(), self.cell_idx.data_ptr(), self.w.data_ptr()))

    def close(self):
        if self.handle:
            _lib.lib().agf_csr_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RegionalPlan:
    """``agf_rplan_t``: a CSR lowered onto the 8 x 32 cell tiles of the raster's lat x lon grid (kept on the DeviceCSR)."""

    def __init__(self, host: HostCSR, n_lat: int, n_lon: int):
        _torch()
        self.n_lat, self.n_lon = int(n_lat), int(n_lon)
        self.handle = C.c_void_p()
        rp, ci, w = (np.ascontiguousarray(host.row_ptr, dtype=np.int32), np.ascontiguousarray(host.cell_idx, dtype=np.int32),
                     np.ascontiguousarray(host.w, dtype=np.float64))
        _lib.check(_lib.lib().agf_rplan_create(C.byref(self.handle), host.n_regions, self.n_lat, self.n_lon, host.nnz,
                                               rp.ctypes.data, ci.ctypes.data, w.ctypes.data))
        self.info = _lib.RPlanInfo()
        _lib.check(_lib.lib().agf_rplan_info(self.handle, C.byref(self.info)))

    def close(self):
        if self.handle:
            _lib.lib().agf_rplan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def regional_plan_of(csr: "DeviceCSR", n_lat: int, n_lon: int) -> RegionalPlan:
    cache = csr.__dict__.setdefault("_rplans", {})
    key = (int(n_lat), int(n_lon))
    if key not in cache:
        if int(n_lat) * int(n_lon) != csr.host.n_cells:
            raise ValueError(f"grid {n_lat} x {n_lon} does not match the CSR's {csr.host.n_cells} cells")
        cache[key] = RegionalPlan(csr.host, n_lat, n_lon)
    return cache[key]


class PanelResult:
    """What a regional run leaves on the device: the panel itself (no per-cell X / V)."""

    def __init__(self, panel, den, labels, nodes):
        self.panel, self.den, self.labels, self.nodes = panel, den, labels, nodes


def regional_candidate(stage: Stage, n_lon: int) -> bool:
    """Cheap host-side test: one raster-reading single-level float32 program fills the whole stage."""
    opt = OPTIONS.get("regional", "auto")
    if opt is False or stage.inputs or stage.elementwise is not None or len(stage.programs) != 1:
        return False
    spec = stage.programs[0]
    if spec.two_level or getattr(spec, "_source", None) is not None or np.dtype(spec.in_dtype) != np.float32:
        return False
    b1 = np.asarray(spec.bounds1)
    if len(b1) < 2 or not np.all(np.diff(b1) == 24) or n_lon % 4 != 0:
        return False
    return opt is True or len(stage.labels) >= int(OPTIONS.get("regional_min_periods", 32))


class RegionalRunner:
    """One program + one regional plan -> the panel, without X.  Same streamed interface as ``StageRunner``
    (``begin_streamed`` / ``feed`` / ``finish_streamed``), so ``stream.feed_and_run`` drives either."""

    GROUP_QUANTUM = 16          # periods per launch of a streamed feed (a few period blocks: amortises the launch)

    def __init__(self, stage: Stage, csr: "DeviceCSR", n_lat: int, n_lon: int, device=None, want_den: bool = False):
        torch = _torch()
        self.stage, self.csr = stage, csr
        self.n_cells = int(n_lat) * int(n_lon)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.plan = regional_plan_of(csr, n_lat, n_lon)
        self.program = Program(stage.programs[0], stage.dtype, self.n_cells, 1)
        self.programs = [self.program]
        self.G, self.n_cols = len(stage.labels), len(stage.nodes)
        self.info = _lib.RegionalInfo()
        _lib.check(_lib.lib().agf_temporal_regional_plan(self.program.handle, self.plan.handle, self.G, C.byref(self.info)))
        self.supported = bool(self.info.supported)
        self.workspace = self.panel = self.den = None
        self.want_den = want_den
        self.b1 = np.asarray(stage.programs[0].bounds1, dtype=np.int64)
        self._cursor = 0

    def _buffers(self):
        torch = _torch()
        if self.workspace is None:
            self.workspace = torch.empty(int(self.info.workspace_bytes), dtype=torch.uint8, device=self.device)
        R = self.csr.host.n_regions
        self.panel = torch.empty((R, self.G, self.n_cols), dtype=torch.float64, device=self.device)
        self.den = torch.empty((R, self.G), dtype=torch.float64, device=self.device) if self.want_den else None

    def _launch(self, raster, g0: int, g1: int, stream, k1_events=None, row0: int = 0) -> None:
        torch = _torch()
        if raster.dtype != torch.float32:
            raise TypeError(f"raster dtype {raster.dtype} does not match the planned float32")
        ev = None
        if k1_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(stream)
        with torch.cuda.stream(stream):
            _lib.check(_lib.lib().agf_temporal_regional_run(
                self.program.handle, self.plan.handle, raster.data_ptr(), self.n_cells, int(row0), g0, g1,
                self.workspace.data_ptr(), self.workspace.numel(), self.panel.data_ptr(), self.G, self.n_cols,
                self.den.data_ptr() if self.den is not None else None, stream.cuda_stream))
        if ev is not None:
            ev[1].record(stream)
            k1_events.append(ev)

    @property
    def launches_per_run(self) -> int:
        return 1 + (1 if self.plan.info.n_partial_rows else 0) + (1 if self.plan.info.n_empty_regions else 0)

    def partial_bytes(self) -> int:
        """Scratch rows of the regions that straddle tiles: written by the scan kernel, read once by the merge kernel."""
        return int(self.plan.info.n_partial_rows) * self.G * int(self.info.lanes_per_slot) * 16

    def algorithmic_input_bytes(self) -> int:
        """Raster bytes the kernel must read: the tiles that hold at least one weighted cell (cells no region
        touches cannot change the panel and are never loaded)."""
        T = int(self.b1[-1] - self.b1[0])
        return T * int(self.plan.info.n_active_tiles) * 256 * 4

    def algorithmic_output_bytes(self) -> int:
        return self.csr.host.n_regions * self.G * self.n_cols * 8

    def run(self, raster, stream=None, k1_events=None) -> PanelResult:
        """Device-resident raster [T, n_cells]: one launch for the whole period range."""
        torch = _torch()
        st = torch.cuda.current_stream() if stream is None else stream
        self._buffers()
        self._launch(raster, 0, self.G, st, k1_events)
        return PanelResult(self.panel, self.den, self.stage.labels, self.stage.nodes)

    # ---- streamed execution (stream.feed_and_run) ----
    def begin_streamed(self, stream) -> None:
        self._buffers()
        self._cursor = 0

    def window_cuts(self) -> np.ndarray:
        """Rows at which a device window of the raster may end (stream.feed_and_run's ring mode): ends of whole
        launch quanta, and the end of the time axis."""
        q = np.arange(self.GROUP_QUANTUM, self.G, self.GROUP_QUANTUM)
        return np.unique(np.concatenate([self.b1[q], self.b1[-1:]])).astype(np.int64)

    def feed(self, raster, rows_ready: int, stream, k1_events=None, row0: int = 0, flush: bool = False) -> int:
        """``raster`` holds rows [row0, rows_ready) of the time axis (row0 = 0: the whole raster so far).  ``flush``:
        the window ends at rows_ready -- launch every complete period, not only whole quanta."""
        g_ready = int(np.searchsorted(self.b1, rows_ready, side="right")) - 1     # complete periods
        g_ready = min(g_ready, self.G)
        if g_ready < self.G and not flush:
            g_ready = self._cursor + (g_ready - self._cursor) // self.GROUP_QUANTUM * self.GROUP_QUANTUM
        if g_ready <= self._cursor:
            return 0
        if self.b1[self._cursor] < row0:
            raise RuntimeError("ring window starts past the first row of the next period")
        self._launch(raster, self._cursor, g_ready, stream, k1_events, row0)
        self._cursor = g_ready
        return 1

    def finish_streamed(self, raster, stream) -> PanelResult:
        if self._cursor != self.G:
            raise RuntimeError(f"streamed run ended with periods {self._cursor}..{self.G} not launched "
                               "(raster shorter than the time axis?)")
        return PanelResult(self.panel, self.den, self.stage.labels, self.stage.nodes)

    def close(self):
        self.program.close()


class StageResult:
    """X[G, cells, n_cols] + V[G, cells] on the device."""

    def __init__(self, X, V, dtype, labels, nodes):
        self.X, self.V, self.dtype, self.labels, self.nodes = X, V, np.dtype(dtype), labels, nodes

    @property
    def n_groups(self) -> int:
        return int(self.X.shape[0])

    @property
    def n_cols(self) -> int:
        return int(self.X.shape[2])


class StageRunner:
    """A planned stage made executable: program handles and device buffers are created once and
    reused by every ``run`` (the yearly loop of a pipeline, benchmark steps)."""

    def __init__(self, stage: Stage, n_cells: int, device=None, target_stripes: int = 0, _shared=None):
        torch = _torch()
        self.stage, self.n_cells = stage, n_cells
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        target_stripes = target_stripes or OPTIONS["target_stripes"]
        shared = {} if _shared is None else _shared            # sub-stages shared between consumers
        self.inputs: List[StageRunner] = []
        for sub in stage.inputs:
            if id(sub) not in shared:
                shared[id(sub)] = StageRunner(sub, n_cells, self.device, target_stripes, shared)
            self.inputs.append(shared[id(sub)])
        self._by_stage = {id(r.stage): r for r in self.inputs}
        G, n_cols = len(stage.labels), len(stage.nodes)
        self.X = torch.empty((G, n_cells, n_cols), dtype=_tdtype(stage.dtype), device=self.device)
        self.V = torch.empty((G, n_cells), dtype=torch.uint8, device=self.device)
        self.programs: List[Program] = []
        self.partials = []
        for spec in stage.programs:
            # a negative target ("at least") is meant for the programs that scan the streamed raster
            ts = target_stripes if getattr(spec, "_source", None) is None else max(target_stripes, 0)
            prog = Program(spec, stage.dtype, n_cells, ts)
            self.programs.append(prog)
            self.partials.append(torch.empty(prog.info.partial_bytes // 8, dtype=torch.float64, device=self.device)
                                 if prog.info.partial_bytes else None)
        self.other = None
        if stage.elementwise is not None and stage.elementwise.other is not None:
            o = stage.elementwise.other
            if int(np.prod(o.shape)) != G * n_cells:
                raise AssertionError(f"interaction array of shape {o.shape} does not match the series "
                                     f"({G} steps x {n_cells} cells)")                     # dataset.py:500
            self.other = to_device(np.ascontiguousarray(o).reshape(G, n_cells), self.device)
        self.result = StageResult(self.X, self.V, stage.dtype, stage.labels, stage.nodes)
        self._ran_token = None

    def _run_elementwise(self, raster, stream) -> None:
        """X = f(source series) for a materialised-transform stage (one launch)."""
        ew = self.stage.elementwise
        src = raster if ew.source is None else self._by_stage[id(ew.source)].result.X
        G = len(self.stage.labels)
        if ew.source is None and (raster.dtype != _tdtype(ew.in_dtype) or raster.shape[0] != G):
            raise TypeError("raster does not match the planned transform (dtype / length)")
        code = {"pow": _lib.XF_POWI if (float(ew.xparam).is_integer() and 0 <= ew.xparam <= 64) else _lib.XF_POW,
                "spline2": _lib.XF_SPLINE2, "inter": _lib.XF_NONE}[ew.xf]
        f = lambda dt: _lib.F64 if np.dtype(dt) == np.float64 else _lib.F32          # noqa: E731
        pre = (_lib.Pre * max(1, len(ew.pre)))()
        for i, (op, c) in enumerate(ew.pre):
            pre[i].op, pre[i].c = int(op), float(c)
        _lib.check(_lib.lib().agf_elementwise_run(
            src.data_ptr(), f(ew.in_dtype), self.X.data_ptr(), f(self.stage.dtype), G * self.n_cells, code,
            float(ew.xparam), self.other.data_ptr() if self.other is not None else None,
            f(np.float64 if (self.other is not None and self.other.dtype == _torch().float64) else np.float32),
            self.V.data_ptr(), len(ew.pre), pre, stream.cuda_stream))

    @property
    def launches_per_run(self) -> int:
        n = sum(r.launches_per_run for r in self.inputs)
        return (n + sum(2 if (p.spec.two_level and not p.info.direct_out) else 1 for p in self.programs)
                + (1 if self.stage.elementwise else 0))

    def algorithmic_input_bytes(self) -> int:
        """Bytes of raster the temporal kernels of this stage must read (each program reads its
        input series once)."""
        n = 0
        for p in self.programs:
            if getattr(p.spec, "_source", None) is None:
                n += int(p.spec.bounds1[-1] - p.spec.bounds1[0]) * self.n_cells * np.dtype(p.spec.in_dtype).itemsize
        return n

    def algorithmic_output_bytes(self) -> int:
        """Bytes the raster-reading temporal kernels must write: the columns + validity mask of single-level
        programs, the partial records of two-level ones (their merge is the finalize kernel's traffic)."""
        n = 0
        for p in self.programs:
            if getattr(p.spec, "_source", None) is None:
                if p.spec.two_level and not p.info.direct_out:
                    n += int(p.info.partial_bytes)
                else:
                    G = len(self.stage.labels)
                    n += G * self.n_cells * (len(p.spec.cols) * self.X.element_size() + 1)
        return n

    def run(self, raster, stream=None, token=None, k1_events=None) -> StageResult:
        """Launch everything on ``stream`` (asynchronous).  ``raster``: device tensor [T, n_cells].
        ``k1_events``: optional list that receives (start, end) CUDA event pairs bracketing each
        temporal-kernel launch that reads the raster."""
        torch = _torch()
        L = _lib.lib()
        token = object() if token is None else token
        if self._ran_token is token:
            return self.result
        for sub in self.inputs:
            sub.run(raster, stream, token, k1_events)
        st = torch.cuda.current_stream() if stream is None else stream
        sptr = st.cuda_stream
        n_cols = len(self.stage.nodes)
        first = True
        with torch.cuda.stream(st):
            if self.stage.elementwise is not None:
                self._run_elementwise(raster, st)
            for prog, partial in zip(self.programs, self.partials):
                spec = prog.spec
                src = getattr(spec, "_source", None)
                if src is None:
                    x_ptr, ld = raster.data_ptr(), self.n_cells
                    if raster.dtype != _tdtype(spec.in_dtype):
                        raise TypeError(f"raster dtype {raster.dtype} does not match the planned {spec.in_dtype}")
                else:
                    sub_res = self._by_stage[id(src[0])].result
                    assert sub_res.n_cols == 1 and src[1] == 0     # a materialised series is its own raster
                    ld, x_ptr = self.n_cells, sub_res.X.data_ptr()
                    assert sub_res.dtype == np.dtype(spec.in_dtype)
                pptr = partial.data_ptr() if partial is not None else None
                ev = None
                if k1_events is not None and src is None:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record(st)
                _lib.check(L.agf_temporal_run(prog.handle, x_ptr, ld, 0, 0, prog.info.n_stripes, pptr,
                                              self.X.data_ptr(), self.V.data_ptr(), n_cols, 0 if first else 1, sptr))
                if ev is not None:
                    ev[1].record(st)
                    k1_events.append(ev)
                _lib.check(L.agf_temporal_finalize(prog.handle, pptr, self.X.data_ptr(), self.V.data_ptr(),
                                                   n_cols, 0 if first else 1, sptr))
                first = False
        self._ran_token = token
        return self.result

    # ---- streamed execution (host-resident raster, see stream.py) -------------------------------
    def _walk(self, seen=None) -> List["StageRunner"]:
        """This runner and the sub-stage runners it reads from, dependencies first, each once."""
        seen = [] if seen is None else seen
        for sub in self.inputs:
            sub._walk(seen)
        if self not in seen:
            seen.append(self)
        return seen

    def begin_streamed(self, stream) -> None:
        """Reset the per-program stripe cursors.  The programs of a stage launch in chunk order, not
        program order, so the shared validity mask starts at 1 and every program ANDs into it."""
        torch = _torch()
        for r in self._walk():
            r._cursor = [0] * len(r.programs)
            if not hasattr(r, "_stripe_ends"):
                r._stripe_ends = [np.array([p.stripe_rows(s)[1] for s in range(p.info.n_stripes)], dtype=np.int64)
                                  for p in r.programs]
            with torch.cuda.stream(stream):
                r.V.fill_(1)
            r._ran_token = None

    def window_cuts(self) -> np.ndarray:
        """Rows at which a device window of the raster may end (stream.feed_and_run's ring mode): stripe ends shared
        by EVERY raster-reading program of the stage tree (the end of the time axis always is one).  None: a stage
        that needs the whole raster at once (an elementwise transform of the raw raster)."""
        cuts = None
        for r in self._walk():
            if r.stage.elementwise is not None and r.stage.elementwise.source is None:
                return None
            for p in r.programs:
                if getattr(p.spec, "_source", None) is not None:
                    continue
                ends = np.array([p.stripe_rows(s)[1] for s in range(p.info.n_stripes)], dtype=np.int64)
                cuts = ends if cuts is None else np.intersect1d(cuts, ends)
        return cuts

    def feed(self, raster, rows_ready: int, stream, k1_events=None, row0: int = 0, flush: bool = False) -> int:
        """Rows [row0, rows_ready) of the time axis are (stream-ordered) resident in ``raster`` (row0 = 0: the whole
        raster so far; otherwise a ring window whose first row is row0): launch every stripe of every raster-reading
        program that is now complete.  Returns the number of launches."""
        torch = _torch()
        L = _lib.lib()
        n = 0
        for r in self._walk():
            n_cols = len(r.stage.nodes)
            for i, (prog, partial) in enumerate(zip(r.programs, r.partials)):
                if getattr(prog.spec, "_source", None) is not None:
                    continue
                if raster.dtype != _tdtype(prog.spec.in_dtype):
                    raise TypeError(f"raster dtype {raster.dtype} does not match the planned {prog.spec.in_dtype}")
                s0 = r._cursor[i]
                # stripes are ordered in time; a stripe is complete when its last row is resident
                s1 = int(np.searchsorted(r._stripe_ends[i], rows_ready, side="right"))
                if s1 <= s0:
                    continue
                pptr = partial.data_ptr() if partial is not None else None
                ev = None
                if k1_events is not None:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record(stream)
                if row0 and prog.stripe_rows(s0)[0] < row0:
                    raise RuntimeError("ring window starts past the first row of the next stripe")
                _lib.check(L.agf_temporal_run(prog.handle, raster.data_ptr(), r.n_cells, int(row0), s0, s1, pptr,
                                              r.X.data_ptr(), r.V.data_ptr(), n_cols, 1, stream.cuda_stream))
                if ev is not None:
                    ev[1].record(stream)
                    k1_events.append(ev)
                r._cursor[i] = s1
                n += 1
        return n

    def finish_streamed(self, raster, stream) -> StageResult:
        """After the last ``feed``: merge partial records, then run the programs that read
        materialised sub-stage results (never the raster)."""
        L = _lib.lib()
        sptr = stream.cuda_stream
        for r in self._walk():
            n_cols = len(r.stage.nodes)
            if r.stage.elementwise is not None:
                with _torch().cuda.stream(stream):
                    r._run_elementwise(raster, stream)          # overwrites the V = 1 of begin_streamed
            for i, (prog, partial) in enumerate(zip(r.programs, r.partials)):
                pptr = partial.data_ptr() if partial is not None else None
                src = getattr(prog.spec, "_source", None)
                if src is None:
                    if r._cursor[i] != prog.info.n_stripes:
                        raise RuntimeError(f"streamed run ended with {prog.info.n_stripes - r._cursor[i]} stripes "
                                           "of a program not launched (raster shorter than the time axis?)")
                else:
                    sub_res = r._by_stage[id(src[0])].result
                    assert sub_res.n_cols == 1 and src[1] == 0
                    ld, x_ptr = r.n_cells, sub_res.X.data_ptr()
                    _lib.check(L.agf_temporal_run(prog.handle, x_ptr, ld, 0, 0, prog.info.n_stripes, pptr,
                                                  r.X.data_ptr(), r.V.data_ptr(), n_cols, 1, sptr))
                _lib.check(L.agf_temporal_finalize(prog.handle, pptr, r.X.data_ptr(), r.V.data_ptr(), n_cols, 1, sptr))
        return self.result

    def close(self):
        for p in self.programs:
            p.close()
        for r in self.inputs:
            r.close()


def run_stage(stage: Stage, raster, n_cells: int, stream=None, target_stripes: int = 0) -> StageResult:
    """Plan-and-run convenience: execute a stage (and the stages it reads from) once."""
    runner = StageRunner(stage, n_cells, raster.device, target_stripes)
    try:
        return runner.run(raster, stream)
    finally:
        # handles own only small tables; X / V / partials are torch tensors kept alive by the result
        if stream is None:
            _torch().cuda.current_stream().synchronize()
        else:
            stream.synchronize()
        runner.close()


def run_spmm(csr: DeviceCSR, res: StageResult, stream=None, want_den: bool = False):
    """panel[R, G, n_cols] (float64, device) = weighted regional average of a stage result."""
    torch = _torch()
    st = torch.cuda.current_stream() if stream is None else stream
    R, G, NC = csr.host.n_regions, res.n_groups, res.n_cols
    panel = torch.empty((R, G, NC), dtype=torch.float64, device=res.X.device)
    den = torch.empty((R, G), dtype=torch.float64, device=res.X.device) if want_den else None
    with torch.cuda.stream(st):
        _lib.check(_lib.lib().agf_spmm_run(csr.handle, res.X.data_ptr(),
                                           _lib.F64 if res.dtype == np.float64 else _lib.F32,
                                           res.V.data_ptr(), G, NC, panel.data_ptr(),
                                           den.data_ptr() if den is not None else None, st.cuda_stream))
    return (panel, den) if want_den else panel


def valid_mask(X, dtype, V, stream=None) -> None:
    """V[g, cell] = AND over columns of ~isnan(X[g, c, cell]) (library kernel, no torch math)."""
    torch = _torch()
    st = torch.cuda.current_stream() if stream is None else stream
    G, n_cells, NC = X.shape
    with torch.cuda.stream(st):
        _lib.check(_lib.lib().agf_valid_mask_run(X.data_ptr(), _lib.F64 if np.dtype(dtype) == np.float64 else _lib.F32,
                                                 G, NC, n_cells, V.data_ptr(), st.cuda_stream))
