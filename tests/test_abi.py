"""The C-ABI library loads and exports every symbol include/aggfly_b200.h declares (no compute
calls: this runs without a GPU), and the host-only planner cuts stripes / records consistently."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from aggfly_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "aggfly_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*\*?(agf_\w+)\(", hdr, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.agf_version() == 5          # AGF_ABI_VERSION


def test_struct_layouts_match_the_header_as_gcc_sees_it(tmp_path):
    """sizeof / offsetof of every struct of the header, from a C program compiled here, against the
    ctypes mirror in _lib.py (a silent mismatch would shift every field after it)."""
    import subprocess
    fields = {"agf_lane_t": (_lib.Lane, ["calc", "flag", "t0", "t1"]),
              "agf_slot_t": (_lib.Slot, ["src", "xform", "xparam", "x_f64", "calc", "flag", "t0", "t1"]),
              "agf_col_t": (_lib.Col, ["src", "xform", "xparam", "x_f64", "dst"]),
              "agf_pre_t": (_lib.Pre, ["op", "c"]),
              "agf_program_desc_t": (_lib.ProgramDesc, ["in_dtype", "out_dtype", "n_lanes", "n_slots", "n_cols", "n_time",
                                                        "n_groups1", "n_groups2", "bounds1", "bounds2", "lanes", "slots",
                                                        "cols", "n_pre", "pre"]),
              "agf_program_info_t": (_lib.ProgramInfo, ["n_stripes", "n_recs", "n_cols", "out_dtype", "n_out_groups",
                                                        "partial_bytes", "out_bytes", "valid_bytes", "kernel_lanes",
                                                        "kernel_slots", "kernel_mode", "uses_tma", "kernel_kinds", "direct_out"]),
              "agf_rplan_info_t": (_lib.RPlanInfo, ["n_tiles", "n_active_tiles", "n_slots", "max_slots_per_tile", "n_entries",
                                                    "tile_lat", "tile_lon", "n_empty_regions", "n_partial_rows", "table_bytes"]),
              "agf_regional_info_t": (_lib.RegionalInfo, ["supported", "lanes_per_slot", "kernel_lanes", "smem_bytes",
                                                          "ctas_per_sm", "workspace_bytes"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "aggfly_b200.h"', "int main(void) {"]
    for st, (_, names) in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        lines += [f'printf("{st}.{n} %zu\\n", offsetof({st}, {n}));' for n in names]
    lines += ['printf("abi %d\\n", AGF_ABI_VERSION);', "return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st, (cls, names) in fields.items():
        assert int(got[st]) == C.sizeof(cls), st
        for n in names:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, (st, n)
    assert int(got["abi"]) == _lib.ABI_VERSION


def _desc(b1, b2=None, n_lanes=1, n_slots=0, n_cols=1):
    d = _lib.ProgramDesc()
    d.in_dtype, d.out_dtype = _lib.F32, _lib.F32
    d.n_lanes, d.n_slots, d.n_cols = n_lanes, n_slots, n_cols
    b1 = np.ascontiguousarray(b1, dtype=np.int32)
    d.n_time, d.n_groups1 = int(b1[-1]), len(b1) - 1
    d.bounds1 = b1.ctypes.data_as(C.POINTER(C.c_int32))
    keep = [b1]
    if b2 is not None:
        b2 = np.ascontiguousarray(b2, dtype=np.int32)
        d.n_groups2 = len(b2) - 1
        d.bounds2 = b2.ctypes.data_as(C.POINTER(C.c_int32))
        keep.append(b2)
    for c in range(n_cols):
        d.cols[c].dst = c
    return d, keep


def _plan(d, n_cells, target):
    buf = (C.c_int32 * (4 * 4096))()
    n_str, n_rec, kl, ks, kd = (C.c_int32() for _ in range(5))
    rc = _lib.lib().agf_program_plan(C.byref(d), n_cells, target, 148, buf, 4096, C.byref(n_str), C.byref(n_rec),
                                     C.byref(kl), C.byref(ks), C.byref(kd))
    _lib.check(rc)
    s = np.array(buf[:4 * n_str.value]).reshape(-1, 4)
    return s, n_rec.value, (kl.value, ks.value, kd.value)


def test_planner_stripes_cover_groups_and_records_count_intersections():
    b1 = np.arange(0, 24 * 365 + 1, 24)                       # 365 days
    month_days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    b2 = np.concatenate([[0], np.cumsum(month_days)])
    d, keep = _desc(b1, b2, n_slots=1)
    for target in (1, 7, 12, 50, 365, 1000):
        s, n_rec, kern = _plan(d, 24544, target)
        assert s[0, 0] == 0 and s[-1, 1] == 365 and np.array_equal(s[1:, 0], s[:-1, 1])
        want = 0
        for g0, g1, g2_first, rec0 in s:
            assert rec0 == want
            assert b2[g2_first] <= g0 < b2[g2_first + 1]
            want += len({int(np.searchsorted(b2, g, side="right") - 1) for g in range(g0, g1)})
        assert n_rec == want and kern == (1, 1, 0)
    s, n_rec, _ = _plan(d, 1038240, 0)                        # big grid: few stripes
    assert len(s) <= 2
    s, n_rec, _ = _plan(d, 24544, 0)                          # small grid: many stripes to fill 148 SMs
    assert len(s) >= 12


def test_planner_rejects_bad_descriptors():
    d, keep = _desc([0, 5, 3])
    with pytest.raises(_lib.AgfError, match="monotonic"):
        _plan(d, 100, 1)
    d, keep = _desc([0, 5, 9], [0, 1], n_slots=1)
    with pytest.raises(_lib.AgfError, match="cover"):
        _plan(d, 100, 1)
    d, keep = _desc([0, 5, 9], [0, 2], n_lanes=9, n_slots=3)
    with pytest.raises(_lib.AgfUnsupported):
        _plan(d, 100, 1)
