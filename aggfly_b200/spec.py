"""Spec compiler: the ``('aggregate'|'transform', {...})`` DSL -> fused kernel programs.

The reference executes each output name independently, step by step, materialising one lazy
dask array per step (aggfly/aggregate/aggregate.py:101-162).  Here the whole call is first turned
into a DAG of symbolic series (``Node``; identical prefixes such as a shared ``mean/date`` step are
hash-consed, so the raster is read once for all names) and then lowered to *programs*: one
launch of the fused temporal kernel each (see include/aggfly_b200.h):

* single-level   ``agg(raw)``                        -> lanes + columns
* two-level      ``agg2(xform?(agg1(raw)))``         -> lanes + slots + columns (registers only)
* anything deeper / not fusable                      -> the inner series is materialised by a
  recursive sub-plan and the outer step runs as a single-level program over that array.

Key / dtype rules reproduced from the reference:
  - multi-ddargs fan-out keys ``f"{name}_{lo}_{hi}"``            aggregate.py:143-148, 299
  - power keys ``f"{name}_{exp}"``, ``exp`` taken from ``exp[0]``  aggregate.py:54-63
  - spline keys ``_spline1`` / ``_spline2``                        aggregate.py:70-73
  - a step's result has the dtype of its input                   nb_kernels.py:260, 266
  - ``np.power(array, exp)`` promotion (NumPy NEP 50): numpy-integer exponents turn a float32
    series into float64, python ints do not                      dataset.py:543
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .timeaxis import group_bounds, translate_groupby

FUSABLE_L2 = ("mean", "sum", "min", "max", "dd", "bins")
ALL_CALCS = ("mean", "nanmean", "sum", "min", "max", "dd", "bins", "sine_dd")


# ---------------------------------------------------------------------------------------------
# symbolic series
# ---------------------------------------------------------------------------------------------
class Node:
    """A per-cell time series: the raster, an aggregate of a series, or a transform of one."""
    __slots__ = ("kind", "src", "calc", "freq", "dd", "xf", "xparam", "dtype", "depth", "key",
                 "_axis", "other")

    def __init__(self, kind, src=None, calc=None, freq=None, dd=None, xf=None, xparam=None,
                 dtype=None, key=None, other=None):
        self.kind, self.src, self.calc, self.freq, self.dd = kind, src, calc, freq, dd
        self.xf, self.xparam, self.dtype, self.key, self.other = xf, xparam, np.dtype(dtype), key, other
        self.depth = 0 if src is None else src.depth + (1 if kind == "agg" else 0)
        self._axis = None

    def __repr__(self):
        return f"Node{self.key}"


class Graph:
    """Hash-consing factory for nodes over one raster (dtype + time axis)."""

    def __init__(self, raster_dtype, time_index, pre_ops=None):
        self.nodes: Dict[tuple, Node] = {}
        self.pre_ops = list(pre_ops or [])          # fused preprocess of the raster (preprocess.py)
        self.raw = Node("raw", dtype=raster_dtype, key=("raw",))
        self.raw._axis = (None, time_index)
        self.nodes[self.raw.key] = self.raw

    def _intern(self, node: Node) -> Node:
        return self.nodes.setdefault(node.key, node)

    def agg(self, src: Node, calc: str, freq: str, dd: Optional[Tuple[float, float, float]]) -> Node:
        if calc not in ALL_CALCS:
            raise ValueError(f"unsupported calc {calc!r}; expected one of {ALL_CALCS}")
        if calc in ("dd", "bins", "sine_dd"):
            if dd is None:
                raise ValueError(f"calc={calc!r} needs ddargs")
            dd = (float(dd[0]), float(dd[1]), float(dd[2]))
            if calc == "sine_dd" and dd[2] not in (0.0, 1.0):
                raise ValueError("Invalid ddargs[2] value")          # temporal.py:324-325
        else:
            dd = None
        key = ("agg", src.key, calc, freq, dd)
        return self._intern(Node("agg", src, calc=calc, freq=freq, dd=dd, dtype=src.dtype, key=key))

    def power(self, src: Node, exp) -> Node:
        dtype = np.result_type(src.dtype, exp)                        # NEP 50, like np.power itself
        if dtype.kind != "f" or dtype.itemsize < 4:
            dtype = np.result_type(dtype, np.float32)
        key = ("pow", src.key, float(exp), str(dtype))
        return self._intern(Node("xf", src, xf="pow", xparam=float(exp), dtype=dtype, key=key))

    def inter(self, src: Node, other) -> Node:
        """``Dataset.interact`` (aggfly/dataset/dataset.py:483-563): the series times another array of the
        same shape.  ``other``: a Dataset-like (``.values``) or an array [G, lat, lon]."""
        arr = getattr(other, "values", other)
        if type(arr).__module__.startswith("torch"):
            arr = arr.detach().cpu().numpy()
        arr = np.asarray(arr)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        n = len(self.labels(src))
        if arr.shape[0] != n:
            raise AssertionError(f"interaction array has {arr.shape[0]} time steps, the series has {n}")   # dataset.py:500
        dtype = np.result_type(src.dtype, arr.dtype)
        key = ("inter", src.key, id(arr))
        node = self._intern(Node("xf", src, xf="inter", dtype=dtype, key=key, other=arr))
        return node

    def spline2(self, src: Node) -> Node:
        key = ("spline2", src.key)
        return self._intern(Node("xf", src, xf="spline2", dtype=src.dtype, key=key))

    # time axis of a node: (bounds over the parent's axis | None, labels)
    def axis(self, node: Node):
        if node._axis is None:
            if node.kind == "xf":
                node._axis = (None, self.axis(node.src)[1])
            else:
                node._axis = group_bounds(self.axis(node.src)[1], node.freq)
        return node._axis

    def labels(self, node: Node):
        return self.axis(node)[1]


class TemporalAggregator:
    """One ``('aggregate', {...})`` step (aggfly/aggregate/temporal.py:19-163): same constructor,
    same attributes (``calc``, ``groupby`` as the resample frequency, ``ddargs``, ``multi_dd``)."""

    def __init__(self, calc: str, groupby: str, ddargs=None, pre_compute: bool = False, engine: str = "auto"):
        self.calc = calc
        self.groupby = translate_groupby(groupby)
        self.ddargs = self.get_ddargs(ddargs)
        self.pre_compute = pre_compute
        self.engine = engine
        self.kwargs = {"ddargs": self.ddargs} if calc in ("dd", "bins", "sine_dd") else {}
        if calc not in ALL_CALCS:
            # the reference fails later with an UnboundLocalError from assign_func
            raise ValueError(f"unsupported calc {calc!r}; expected one of {ALL_CALCS}")

    def get_ddargs(self, ddargs):
        if ddargs is None:
            self.multi_dd = False
            return None
        self.multi_dd = len(np.array(ddargs).shape) > 1
        return ddargs

    def execute(self, dataset, weights=None, update: bool = False, **kwargs):
        """Run this single step on a Dataset (temporal.py:165-263) through the CUDA engine."""
        from .aggregate import _execute_single_step
        return _execute_single_step(self, dataset)


def _exp_list(params: dict):
    exp = params["exp"]
    if not isinstance(exp, list):
        exp = [exp]
    exps = exp[0]                                            # aggregate.py:56-59 (the [0] quirk)
    if np.ndim(exps) == 0:
        raise TypeError("transform 'exp' must be an array of exponents (e.g. np.arange(1, 3)), "
                        "as in the reference (aggregate.py:59)")
    return list(exps)


def compile_spec(graph: Graph, aggregator_dict: Dict[str, Sequence]) -> Dict[str, Node]:
    """aggregate.py:101-162 on symbolic series -> {output name: Node}, insertion-ordered."""
    out: Dict[str, Node] = {}
    for key, steps in aggregator_dict.items():
        keys, data = [key], [graph.raw]
        for step in steps:
            kind, params = step
            if kind == "aggregate":
                agg = params if isinstance(params, TemporalAggregator) else TemporalAggregator(**params)
                if agg.multi_dd:
                    if len(data) > 1:
                        raise ValueError("Cannot aggregate multiple datasets with multiple ddargs, "
                                         "e.g., multiple polynomials for multiple bins")
                    rows = [tuple(r) for r in agg.ddargs]
                    if len(rows) < 2:
                        raise ValueError("2-D ddargs needs at least two rows (a single-row 2-D ddargs "
                                         "breaks the reference, temporal.py:250-253); pass a flat triple")
                    data = [graph.agg(data[0], agg.calc, agg.groupby, r) for r in rows]
                    keys = [f"{key}_{r[0]}_{r[1]}" for r in agg.ddargs]            # aggregate.py:299
                else:
                    data = [graph.agg(d, agg.calc, agg.groupby, agg.ddargs) for d in data]
            elif kind == "transform":
                nd, nk = [], []
                for d, k in zip(data, keys):
                    if "exp" in params:
                        exps = _exp_list(params)
                        nd.extend(graph.power(d, e) for e in exps)
                        nk.extend(f"{k}_{e}" for e in exps)
                    elif "inter" in params:
                        nd.append(graph.inter(d, params["inter"]))                     # aggregate.py:65-69
                        nk.append(k)
                    elif "spline" in params.get("transform", ""):
                        nd.extend([d, graph.spline2(d)])
                        nk.extend([f"{k}_spline1", f"{k}_spline2"])
                    else:
                        raise ValueError("No valid transform argument provided.")
                data, keys = nd, nk
            else:
                raise ValueError(f"unknown step type {kind!r} (expected 'aggregate' or 'transform')")
        out.update(dict(zip(keys, data)))
    return out


# ---------------------------------------------------------------------------------------------
# lowering to programs
# ---------------------------------------------------------------------------------------------
@dataclass
class LaneSpec:
    calc: str
    dd: Optional[Tuple[float, float, float]] = None


@dataclass
class SlotSpec:
    src: int
    xf: Optional[str]
    xparam: float
    x_f64: bool
    calc: str
    dd: Optional[Tuple[float, float, float]]


@dataclass
class ColSpec:
    src: int
    xf: Optional[str]
    xparam: float
    x_f64: bool
    node: Node = None
    out_col: int = -1            # column in the destination X


@dataclass
class ProgramSpec:
    input: Node                  # graph.raw or a materialised node
    in_dtype: np.dtype
    bounds1: np.ndarray
    bounds2: Optional[np.ndarray]
    lanes: List[LaneSpec] = field(default_factory=list)
    slots: List[SlotSpec] = field(default_factory=list)
    cols: List[ColSpec] = field(default_factory=list)
    freq1: str = ""
    freq2: Optional[str] = None
    n_time: int = 0
    pre: list = field(default_factory=list)   # (op, constant) chain applied to raster values

    @property
    def two_level(self) -> bool:
        return self.bounds2 is not None


@dataclass
class Stage:
    """Programs that together fill one X[G, cells, n_cols] (+ shared validity mask)."""
    programs: List[ProgramSpec]
    nodes: List[Node]            # node of each X column
    dtype: np.dtype              # dtype of X
    labels: Any
    inputs: List["Stage"] = field(default_factory=list)   # stages that must run first
    column_of: Dict[tuple, int] = field(default_factory=dict)
    elementwise: Optional["ElemSpec"] = None              # a materialised transform: X = f(source series)


@dataclass
class ElemSpec:
    """X[G, cells, 1] = xf(source): source is the raster (``source is None``) or column 0 of a stage."""
    source: Optional["Stage"]
    in_dtype: np.dtype
    xf: str
    xparam: float
    other: Optional[np.ndarray] = None
    pre: list = field(default_factory=list)   # preprocess chain when the source is the raster


def _peel(node: Node):
    """node = xf_k(... xf_1(base)) -> (base, [xf_1..xf_k])"""
    xfs = []
    while node.kind == "xf":
        xfs.append(node)
        node = node.src
    return node, xfs[::-1]


def _xf_fields(xfs: List[Node]):
    if not xfs:
        return None, 0.0, None
    x = xfs[0]
    return x.xf, (x.xparam if x.xf == "pow" else 0.0), x


MAX_TYPED_BINS = 28


def _lane_index(lanes: List[LaneSpec], calc, dd) -> int:
    for i, l in enumerate(lanes):
        if l.calc == calc and l.dd == dd:
            return i
    lanes.append(LaneSpec(calc, dd))
    return len(lanes) - 1


def _ensure_sine_helpers(lanes: List[LaneSpec]):
    if not lanes or lanes[0].calc != "_hidden_sum":
        lanes[:0] = [LaneSpec("_hidden_sum"), LaneSpec("_hidden_min"), LaneSpec("_hidden_max")]
        return 3
    return 0


class Planner:
    def __init__(self, graph: Graph):
        self.g = graph
        self._materialised: Dict[tuple, Tuple[Stage, int]] = {}   # node.key -> (stage, column)

    # ---- public ------------------------------------------------------------------------------
    def plan(self, nodes: List[Node], dtype=None) -> Stage:
        """A stage whose X columns are exactly ``nodes`` (same final time axis required)."""
        if dtype is None:
            dtype = np.result_type(*[n.dtype for n in nodes])
        stage = Stage(programs=[], nodes=list(nodes), dtype=np.dtype(dtype), labels=self.g.labels(nodes[0]))
        groups: Dict[tuple, List[Tuple[int, Node]]] = {}
        for col, node in enumerate(nodes):
            groups.setdefault(self._pattern(node), []).append((col, node))
        for pat, members in groups.items():
            self._lower_group(stage, pat, members)
        return stage

    # ---- classification ------------------------------------------------------------------------
    def _pattern(self, node: Node) -> tuple:
        a, tail = _peel(node)
        if a.kind == "raw":
            return ("identity",) if not tail else ("mat", node.key)   # transforms of the raster itself
        if len(tail) > 1 or any(x.xf == "inter" for x in tail):
            return ("mat", node.key)            # transforms the programs cannot carry: materialise, copy out
        s1, mid = _peel(a.src)
        if s1.kind == "raw":
            if mid:                             # aggregate of a transformed raster
                return ("outer", a.src.key, a.freq)
            return ("l1", a.freq)
        # s1 is an aggregate: can agg2(mid?(agg1(raw))) be fused?
        s2, pre = _peel(s1.src)
        fusable_mid = len(mid) == 0 or (len(mid) == 1 and mid[0].xf != "inter")
        if s2.kind == "raw" and not pre and not mid and self._collapsed_lane(s1, a) is not None:
            return ("l1c", s1.freq, a.freq)
        if (s2.kind == "raw" and not pre and fusable_mid and a.calc in FUSABLE_L2):
            return ("l2", s1.freq, a.freq)
        return ("outer", a.src.key, a.freq)     # single-level program over the materialised a.src

    def _collapsed_lane(self, s1: Node, a: Node):
        """``a(s1(raw))`` where every group of ``s1`` is exactly ONE row (daily data grouped by date), or a sum of bin
        counts over non-empty inner groups of any length:
        the inner step is the identity (mean / sum / min / max / nanmean of one value), a 0/1 bin
        indicator, or one degree-day term, so the pair is a single-level reduction of the raster
        over the composed bounds.  Returns the (calc, ddargs) of that lane, or None.

        Same bits as the two-step chain: a one-value mean is the value itself, counts are exact,
        and ``dd_r`` rounds every degree-day term to the raster dtype before adding it (the inner
        step stores its result in the input dtype, nb_kernels.py:260)."""
        b1, _ = self.g.axis(s1)
        n_rows = len(self.g.labels(self.g.raw))
        if s1.calc == "bins" and a.calc == "sum" and len(b1) > 1 and np.all(np.diff(b1) >= 1):
            # Counts of counts: ``bins / date -> sum / year`` (the chain of the reference's own golden matrix,
            # tests/test_aggregate.py:275-280) adds exact small integers, so it is the bin count over the composed
            # group whatever the inner group length -- as long as no inner group is EMPTY (an empty group's bins are
            # NaN, nb_kernels.py:197-199, and would poison the outer sum).  One single-level pass with typed bin lanes
            # instead of sixteen general lanes feeding sixteen slots (C3-size year: 67 ms -> 10 ms).
            return ("bins", s1.dd)
        if len(b1) - 1 != n_rows or (n_rows and not np.all(np.diff(b1) == 1)):
            return None
        if s1.calc in ("mean", "sum", "min", "max", "nanmean"):
            return (a.calc, a.dd)
        if s1.calc == "dd" and a.calc == "sum":
            return ("dd_r", s1.dd)
        if s1.calc == "bins" and a.calc == "sum":
            return ("bins", s1.dd)
        return None

    # ---- lowering ------------------------------------------------------------------------------
    def _lower_group(self, stage: Stage, pat: tuple, members: List[Tuple[int, Node]]):
        kind = pat[0]
        if kind == "identity":
            T = len(self.g.labels(self.g.raw))
            b1 = np.arange(T + 1, dtype=np.int64)
            self._emit_single_level(stage, self.g.raw, b1, "id", members, identity=True)
        elif kind == "mat":
            # the whole node is materialised by elementwise passes in its own stage, then copied into
            # this stage's column by an identity program (one row per group)
            for col, node in members:
                sub, _ = self._materialise(node)
                n = len(self.g.labels(node))
                self._emit_single_level(stage, node, np.arange(n + 1, dtype=np.int64), "id", [(col, node)],
                                        identity=True, source=(sub, 0))
        elif kind == "l1":
            a0, _ = _peel(members[0][1])
            b1, _ = self.g.axis(a0)
            self._emit_single_level(stage, self.g.raw, b1, pat[1], members)
        elif kind == "l1c":
            a0, _ = _peel(members[0][1])
            b1, _ = self.g.axis(a0.src)                       # one row per inner group
            b2, _ = self.g.axis(a0)
            composed = np.asarray(b1)[np.asarray(b2)]
            self._emit_single_level(stage, self.g.raw, composed, pat[2], members,
                                    lane_of=lambda a: self._collapsed_lane(_peel(a.src)[0], a))
        elif kind == "outer":
            a0, _ = _peel(members[0][1])
            inner = a0.src
            sub, col = self._materialise(inner)
            b1, _ = self.g.axis(a0)
            self._emit_single_level(stage, inner, b1, pat[2], members, source=(sub, col))
        elif kind == "l2":
            self._emit_two_level(stage, pat[1], pat[2], members)
        else:
            raise AssertionError(pat)

    def _materialise(self, node: Node) -> Tuple[Stage, int]:
        hit = self._materialised.get(node.key)
        if hit is None:
            if node.kind == "xf":
                # X = xf(source series): one elementwise pass over the materialised source (or the raster)
                src_stage = None if node.src.kind == "raw" else self._materialise(node.src)[0]
                xf, xparam = ("inter", 0.0) if node.xf == "inter" else (node.xf, node.xparam if node.xf == "pow" else 0.0)
                sub = Stage(programs=[], nodes=[node], dtype=np.dtype(node.dtype), labels=self.g.labels(node),
                            inputs=[] if src_stage is None else [src_stage],
                            elementwise=ElemSpec(src_stage, np.dtype(node.src.dtype), xf, float(xparam), node.other,
                                                 pre=self.g.pre_ops if src_stage is None else []))
            else:
                sub = self.plan([node], dtype=node.dtype)      # X dtype == node dtype: no extra rounding
            hit = (sub, 0)
            self._materialised[node.key] = hit
        return hit

    def _emit_single_level(self, stage, input_node, b1, freq, members, identity=False, source=None, lane_of=None):
        n_time = int(len(self.g.labels(input_node)))
        prog = None
        for col, node in members:
            a, tail = _peel(node)
            xf, xparam, xnode = (None, 0.0, None) if identity else _xf_fields(tail)
            calc, dd = ("mean", None) if identity else (lane_of(a) if lane_of is not None else (a.calc, a.dd))
            need = (1 if calc != "sine_dd" else 4)
            # the typed-lane kernels hold at most 28 bin counters (csrc/agf_k1_*: NB = 6 / 14 / 28): a 29th bin lane would
            # send the whole program to the general per-value dispatch
            bins_full = (calc == "bins" and prog is not None and (calc, dd) not in [(l.calc, l.dd) for l in prog.lanes]
                         and sum(l.calc == "bins" for l in prog.lanes) >= MAX_TYPED_BINS)
            if (prog is None or bins_full or len(prog.lanes) + need > _lib.MAX_LANES
                    or len(prog.cols) + 1 > _lib.MAX_COLS):
                prog = ProgramSpec(input=input_node, in_dtype=input_node.dtype, bounds1=b1, bounds2=None,
                                   freq1=freq, n_time=n_time, pre=self.g.pre_ops if source is None else [])
                prog._source = source
                stage.programs.append(prog)
                if source is not None and source[0] not in stage.inputs:
                    stage.inputs.append(source[0])
            shift = 0
            if calc == "sine_dd":
                shift = _ensure_sine_helpers(prog.lanes)
                for c in prog.cols:
                    c.src += shift
            lane = _lane_index(prog.lanes, calc, dd)
            prog.cols.append(ColSpec(lane, xf, xparam, node.dtype == np.float64, node=node, out_col=col))

    def _emit_two_level(self, stage, freq1, freq2, members):
        n_time = int(len(self.g.labels(self.g.raw)))
        prog = None
        for col, node in members:
            a, tail = _peel(node)
            s1, mid = _peel(a.src)
            txf, txparam, _ = _xf_fields(tail)
            mxf, mxparam, _ = _xf_fields(mid)
            placed = False
            for attempt in range(2):
                if prog is None or attempt == 1:
                    b1, _ = self.g.axis(s1)
                    b2, _ = self.g.axis(a)
                    prog = ProgramSpec(input=self.g.raw, in_dtype=self.g.raw.dtype, bounds1=b1, bounds2=b2,
                                       freq1=freq1, freq2=freq2, n_time=n_time, pre=self.g.pre_ops)
                    prog._source = None
                    stage.programs.append(prog)
                lanes = [LaneSpec(l.calc, l.dd) for l in prog.lanes]
                slots = list(prog.slots)
                shift = _ensure_sine_helpers(lanes) if s1.calc == "sine_dd" else 0
                if shift:
                    slots = [SlotSpec(s.src + shift, s.xf, s.xparam, s.x_f64, s.calc, s.dd) for s in slots]
                lane = _lane_index(lanes, s1.calc, s1.dd)
                want = SlotSpec(lane, mxf, mxparam, a.src.dtype == np.float64, a.calc, a.dd)
                slot = next((j for j, s in enumerate(slots) if s == want), None)
                if slot is None:
                    slots.append(want)
                    slot = len(slots) - 1
                if _fits_two_level(lanes, slots) and len(prog.cols) + 1 <= _lib.MAX_COLS:
                    prog.lanes, prog.slots = lanes, slots
                    prog.cols.append(ColSpec(slot, txf, txparam, node.dtype == np.float64, node=node, out_col=col))
                    placed = True
                    break
            if not placed:
                raise AssertionError("a single output must fit an empty program")


def _fits_two_level(lanes: List[LaneSpec], slots: List[SlotSpec]) -> bool:
    if len(lanes) <= 4 and len(slots) <= _lib.MAX_SLOTS:
        return True
    # Degree-day / mean / sum lanes have specialised kernels for up to four lanes (csrc/agf_k1_f32_tma_uni.cu: KIND_DD,
    # KIND_MIX_SD); the sixteen-lane diagonal form below runs on the general ragged-group kernel, which is slower than
    # several passes of those (six dd thresholds of a global hourly year: 154 ms in one pass, 11 ms in two)
    if all(l.calc in ("dd", "mean", "sum") for l in lanes):
        return False
    # diagonal form: lane j feeds slot j only
    if len(lanes) <= 16 and len(slots) == len(lanes):
        return all(s.src == j for j, s in enumerate(slots))
    return False


# ---------------------------------------------------------------------------------------------
# ProgramSpec -> C descriptor
# ---------------------------------------------------------------------------------------------
_XF_CODE = {None: _lib.XF_NONE, "spline2": _lib.XF_SPLINE2}


def _xf_code(xf: Optional[str], xparam: float) -> int:
    if xf == "pow":
        return _lib.XF_POWI if (float(xparam).is_integer() and 0 <= xparam <= 64) else _lib.XF_POW
    return _XF_CODE[xf]


def build_desc(prog: ProgramSpec, out_dtype) -> Tuple[_lib.ProgramDesc, list]:
    """Fill the ctypes descriptor; returns it plus the arrays it points into (keep them alive)."""
    d = _lib.ProgramDesc()
    d.in_dtype = _lib.F64 if prog.in_dtype == np.float64 else _lib.F32
    d.out_dtype = _lib.F64 if np.dtype(out_dtype) == np.float64 else _lib.F32
    d.n_lanes, d.n_slots, d.n_cols = len(prog.lanes), len(prog.slots), len(prog.cols)
    d.n_time = prog.n_time
    b1 = np.ascontiguousarray(prog.bounds1, dtype=np.int32)
    keep = [b1]
    d.n_groups1 = len(b1) - 1
    d.bounds1 = b1.ctypes.data_as(_lib.C.POINTER(_lib.C.c_int32))
    if prog.two_level:
        b2 = np.ascontiguousarray(prog.bounds2, dtype=np.int32)
        keep.append(b2)
        d.n_groups2 = len(b2) - 1
        d.bounds2 = b2.ctypes.data_as(_lib.C.POINTER(_lib.C.c_int32))
    for i, l in enumerate(prog.lanes):
        d.lanes[i].calc = _lib.CALC[l.calc]
        if l.dd is not None:
            d.lanes[i].t0, d.lanes[i].t1, d.lanes[i].flag = l.dd[0], l.dd[1], int(l.dd[2] != 0)
    for j, s in enumerate(prog.slots):
        S = d.slots[j]
        S.src, S.xform, S.xparam, S.x_f64 = s.src, _xf_code(s.xf, s.xparam), s.xparam, int(s.x_f64)
        S.calc = _lib.CALC[s.calc]
        if s.dd is not None:
            S.t0, S.t1, S.flag = s.dd[0], s.dd[1], int(s.dd[2] != 0)
    d.n_pre = len(prog.pre)
    for i, (op, c) in enumerate(prog.pre):
        d.pre[i].op, d.pre[i].c = int(op), float(c)
    for c, k in enumerate(prog.cols):
        Cc = d.cols[c]
        Cc.src, Cc.xform, Cc.xparam, Cc.x_f64 = k.src, _xf_code(k.xf, k.xparam), k.xparam, int(k.x_f64)
        Cc.dst = k.out_col
    return d, keep
