"""ctypes binding of ``libaggfly_b200.so`` (the C-ABI declared in ``include/aggfly_b200.h``).

The product has no CPU compute path: if the shared library is missing or a call fails, this
module raises -- it never falls back to anything else.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# AGF_B200_LIB: a developer switch -- tools/build_variant.sh builds the same sources with extra -D flags into
# csrc/variants/ so that one GPU visit can time several builds side by side.
LIB_PATH = os.environ.get("AGF_B200_LIB") or os.path.join(_HERE, "csrc", "libaggfly_b200.so")

ABI_VERSION = 5                      # AGF_ABI_VERSION of include/aggfly_b200.h
MAX_LANES, MAX_SLOTS, MAX_COLS = 32, 32, 64
E_INVALID, E_UNSUPPORTED, E_NOMEM, E_STATE = -1, -2, -3, -4

CALC = {"mean": 0, "sum": 1, "min": 2, "max": 3, "nanmean": 4, "dd": 5, "bins": 6, "sine_dd": 7,
        "_hidden_sum": 8, "_hidden_min": 9, "_hidden_max": 10, "dd_r": 11}
XF_NONE, XF_POWI, XF_POW, XF_SPLINE2 = 0, 1, 2, 3
F32, F64 = 0, 1
I16, I32, U8, I8, U16 = 2, 3, 4, 5, 6          # storage dtypes of chunked sources (agf_tile_place_run)


class Lane(C.Structure):
    _fields_ = [("calc", C.c_int32), ("flag", C.c_int32), ("t0", C.c_double), ("t1", C.c_double)]


class Slot(C.Structure):
    _fields_ = [("src", C.c_int32), ("xform", C.c_int32), ("xparam", C.c_double), ("x_f64", C.c_int32),
                ("calc", C.c_int32), ("flag", C.c_int32), ("pad_", C.c_int32),
                ("t0", C.c_double), ("t1", C.c_double)]


class Col(C.Structure):
    _fields_ = [("src", C.c_int32), ("xform", C.c_int32), ("xparam", C.c_double),
                ("x_f64", C.c_int32), ("dst", C.c_int32)]


MAX_PRE = 4


class Pre(C.Structure):
    _fields_ = [("op", C.c_int32), ("pad_", C.c_int32), ("c", C.c_double)]


class ProgramDesc(C.Structure):
    _fields_ = [("in_dtype", C.c_int32), ("out_dtype", C.c_int32), ("n_lanes", C.c_int32),
                ("n_slots", C.c_int32), ("n_cols", C.c_int32), ("pad_", C.c_int32),
                ("n_time", C.c_int64), ("n_groups1", C.c_int64), ("n_groups2", C.c_int64),
                ("bounds1", C.POINTER(C.c_int32)), ("bounds2", C.POINTER(C.c_int32)),
                ("lanes", Lane * MAX_LANES), ("slots", Slot * MAX_SLOTS), ("cols", Col * MAX_COLS),
                ("n_pre", C.c_int32), ("pad2_", C.c_int32), ("pre", Pre * MAX_PRE)]


class ProgramInfo(C.Structure):
    _fields_ = [("n_stripes", C.c_int32), ("n_recs", C.c_int32), ("n_cols", C.c_int32),
                ("out_dtype", C.c_int32), ("n_out_groups", C.c_int64), ("partial_bytes", C.c_int64),
                ("out_bytes", C.c_int64), ("valid_bytes", C.c_int64), ("kernel_lanes", C.c_int32),
                ("kernel_slots", C.c_int32), ("kernel_mode", C.c_int32), ("uses_tma", C.c_int32),
                ("kernel_kinds", C.c_int32), ("direct_out", C.c_int32)]


class RPlanInfo(C.Structure):
    _fields_ = [("n_tiles", C.c_int32), ("n_active_tiles", C.c_int32), ("n_slots", C.c_int32),
                ("max_slots_per_tile", C.c_int32), ("n_entries", C.c_int64), ("tile_lat", C.c_int32),
                ("tile_lon", C.c_int32), ("n_empty_regions", C.c_int32), ("n_partial_rows", C.c_int32), ("table_bytes", C.c_int64)]


class RegionalInfo(C.Structure):
    _fields_ = [("supported", C.c_int32), ("lanes_per_slot", C.c_int32), ("kernel_lanes", C.c_int32),
                ("smem_bytes", C.c_int32), ("ctas_per_sm", C.c_int32), ("pad_", C.c_int32), ("workspace_bytes", C.c_int64)]


class AgfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libaggfly_b200 error {code}: {msg}")
        self.code = code


class AgfUnsupported(AgfError):
    """The fused kernels do not cover this program; the host must split it."""


_lib: Optional[C.CDLL] = None

# every symbol include/aggfly_b200.h declares
SYMBOLS = ("agf_version", "agf_last_error", "agf_program_create", "agf_program_destroy",
           "agf_program_plan", "agf_program_info", "agf_program_stripe_rows", "agf_temporal_run",
           "agf_temporal_finalize", "agf_csr_create", "agf_csr_destroy", "agf_spmm_run",
           "agf_valid_mask_run", "agf_elementwise_run", "agf_tile_place_run", "agf_decompress_caps", "agf_decompress_lz4_run",
           "agf_unshuffle_run", "agf_copy_segments_run", "agf_overlap_create", "agf_overlap_fetch", "agf_overlap_destroy",
           "agf_rplan_create", "agf_rplan_destroy", "agf_rplan_info", "agf_rplan_tables", "agf_rplan_check_segments",
           "agf_temporal_regional_plan",
           "agf_temporal_regional_run")


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C aggfly_b200/csrc`). aggfly_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.agf_version.restype = C.c_int
    L.agf_last_error.restype = C.c_char_p
    L.agf_program_create.argtypes = [C.POINTER(vp), C.POINTER(ProgramDesc), i64, i32]
    L.agf_program_destroy.argtypes = [vp]
    L.agf_program_plan.argtypes = [C.POINTER(ProgramDesc), i64, i32, i32, C.POINTER(i32), i32,
                                   C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                   C.POINTER(i32)]
    L.agf_program_info.argtypes = [vp, C.POINTER(ProgramInfo)]
    L.agf_program_stripe_rows.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64)]
    L.agf_temporal_run.argtypes = [vp, vp, i64, i64, i32, i32, vp, vp, vp, i32, i32, u64]
    L.agf_temporal_finalize.argtypes = [vp, vp, vp, vp, i32, i32, u64]
    L.agf_csr_create.argtypes = [C.POINTER(vp), i32, i64, i64, vp, vp, vp]
    L.agf_csr_destroy.argtypes = [vp]
    L.agf_spmm_run.argtypes = [vp, vp, i32, vp, i64, i32, vp, vp, u64]
    L.agf_valid_mask_run.argtypes = [vp, i32, i64, i32, i64, vp, u64]
    L.agf_elementwise_run.argtypes = [vp, i32, vp, i32, i64, i32, C.c_double, vp, i32, vp, i32, C.POINTER(Pre), u64]
    L.agf_tile_place_run.argtypes = [vp, i32, i64, i64, i64, i64, i64, i64, vp, i32, i64, i64, i64, i64, i64,
                                     i32, C.c_double, C.c_double, i32, C.c_double, i64, u64]
    L.agf_decompress_caps.argtypes = [C.POINTER(i32), C.POINTER(i64)]
    L.agf_decompress_lz4_run.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), vp, C.POINTER(i64), C.POINTER(i64), i64, vp, u64]
    L.agf_unshuffle_run.argtypes = [vp, vp, i64, i32, i64, u64]
    L.agf_copy_segments_run.argtypes = [vp, vp, vp, i64, u64]
    dp, i64p, i32p = C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i32)
    L.agf_overlap_create.argtypes = [C.POINTER(vp), i32, i64p, i64p, dp, i32, dp, C.c_double, i32, dp, C.c_double, i64p]
    L.agf_overlap_fetch.argtypes = [vp, i32p, i64p, dp]
    L.agf_overlap_destroy.argtypes = [vp]
    L.agf_rplan_create.argtypes = [C.POINTER(vp), i32, i32, i32, i64, vp, vp, vp]
    L.agf_rplan_destroy.argtypes = [vp]
    L.agf_rplan_tables.argtypes = [i32, i32, i32, i64, vp, vp, vp, C.POINTER(RPlanInfo)] + [vp] * 9
    L.agf_rplan_info.argtypes = [vp, C.POINTER(RPlanInfo)]
    L.agf_rplan_check_segments.argtypes = [i32, i32, i32, i64, vp, vp, vp, i32, vp]
    L.agf_temporal_regional_plan.argtypes = [vp, vp, i64, C.POINTER(RegionalInfo)]
    L.agf_temporal_regional_run.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, i32, vp, u64]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("agf_version", "agf_last_error"):
            fn.restype = C.c_int
    if L.agf_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.agf_version()} != {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = (lib().agf_last_error() or b"").decode("utf-8", "replace")
    if rc == E_UNSUPPORTED:
        raise AgfUnsupported(rc, msg)
    raise AgfError(rc, msg)
