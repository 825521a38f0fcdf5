"""The ctypes stub printed in INTEGRATION.md ("Route B") is executed as it stands, so the documented binding cannot rot:
its structure layouts are compared with the header's (CPU), and its ``cuda_daily_mean`` runs on the GPU against the oracle."""
import os
import re

import numpy as np
import pytest

from aggfly_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source() -> str:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "def cuda_daily_mean" in b]
    assert len(stub) == 1, "INTEGRATION.md no longer holds the ctypes stub"
    assert 'C.CDLL("libaggfly_b200.so")' in stub[0]
    return stub[0].replace('C.CDLL("libaggfly_b200.so")', f"C.CDLL({_lib.LIB_PATH!r})")


def test_documented_struct_layouts_match_the_library():
    import ctypes as C
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)                      # loads the .so, defines the structs; no device call
    for doc, ours in (("Lane", _lib.Lane), ("Slot", _lib.Slot), ("Col", _lib.Col), ("Pre", _lib.Pre), ("Desc", _lib.ProgramDesc)):
        assert C.sizeof(ns[doc]) == C.sizeof(ours), doc
        assert [(n, getattr(ns[doc], n).offset) for n, _ in ns[doc]._fields_] == \
               [(n, getattr(ours, n).offset) for n, _ in ours._fields_], doc


@pytest.mark.gpu
def test_documented_stub_runs_on_the_device_and_matches_the_oracle():
    import torch
    from oracle import oracle as orc
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    rng = np.random.default_rng(4)
    T, Y, X = 24 * 6 + 5, 6, 10
    cube = rng.normal(10, 6, (T, Y, X)).astype(np.float32)
    cube[30:33, 2, 3] = np.nan
    bounds = np.array([0, 24, 48, 72, 96, 120, 144, T], dtype=np.int64)
    got, valid = ns["cuda_daily_mean"](torch.from_numpy(cube).cuda().reshape(T, Y * X), bounds)
    torch.cuda.synchronize()
    want = orc.block_stat(cube, bounds, "mean").reshape(len(bounds) - 1, Y * X)
    assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)
    assert np.array_equal(valid.cpu().numpy().astype(bool), ~np.isnan(want))
