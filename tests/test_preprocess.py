"""Fused preprocess chains (aggfly_b200/preprocess.py) -- host logic, no GPU.  The NumPy evaluation
of a chain must equal evaluating the same expression the way the reference does
(aggfly/cli/preprocess.py: operators dispatched to numpy on the float32 array)."""
import numpy as np
import pytest

from aggfly_b200 import preprocess as pp


def _ref_eval(expr, x):
    return eval(expr, {"__builtins__": {}}, {"x": x})          # what the reference's AST walker computes


@pytest.mark.parametrize("expr", ["x - 273.15", "x + 273.15", "x / 1000.0", "x * 1000.0", "(x - 32) * 5 / 9",
                                  "9 / 5 * x + 32", "-x + 1", "1 - x", "100 / x", "-(x - 2)", "x * (5 / 9)", "2 * (3 + x) - 1"])
def test_chain_equals_numpy_evaluation_bit_for_bit(expr):
    x = np.random.default_rng(1).normal(280, 20, 4096).astype(np.float32)
    chain = pp.compile_expression(expr)
    assert 1 <= len(chain.ops) <= pp.MAX_OPS
    got, want = chain(x), _ref_eval(expr, x)
    assert got.dtype == want.dtype == np.float32
    assert np.array_equal(got, want)
    x64 = x.astype(np.float64)
    assert np.array_equal(chain(x64), _ref_eval(expr, x64))


def test_builtins_and_identity():
    assert pp.resolve("kelvin_to_celsius").ops == [(pp.SUB, 273.15)]
    assert pp.resolve("celsius_to_kelvin").ops == [(pp.ADD, 273.15)]
    assert pp.resolve("pa_to_kpa").ops == [(pp.DIV, 1000.0)]
    assert pp.resolve("m_to_mm").ops == [(pp.MUL, 1000.0)]
    assert pp.resolve("identity") is None and pp.resolve(None) is None and pp.resolve("x") is None
    assert pp.constants_for([(pp.SUB, 273.15)], np.float32) == [(pp.SUB, float(np.float32(273.15)))]


@pytest.mark.parametrize("bad", ["x * x", "y + 1", "x ** 2", "x % 3", "abs(x)", "x.real", "3 + 4", "x +", "x // 2",
                                 "((((x + 1) * 2) - 3) / 4) + 5"])
def test_rejections(bad):
    with pytest.raises(pp.PreprocessError):
        pp.compile_expression(bad)


def test_dataset_keeps_the_raster_untouched_and_records_the_chain():
    import pandas as pd
    import aggfly_b200 as af
    arr = np.full((4, 2, 2), 300.0, np.float32)
    t = pd.date_range("2001-01-01", periods=4, freq="h")
    ds = af.Dataset.from_arrays(arr, t, [1.0, 0.0], [0.0, 1.0], preprocess="kelvin_to_celsius")
    assert ds.values is arr and ds.pre_ops == [(pp.SUB, float(np.float32(273.15)))]
    ds2 = af.Dataset(af.RasterArray(arr, ("time", "latitude", "longitude"), {"time": t, "latitude": [1.0, 0.0], "longitude": [0.0, 1.0]}),
                     preprocess=lambda v: v - 1.0)
    assert ds2.pre_ops == [] and float(ds2.values[0, 0, 0]) == 299.0          # arbitrary callables run eagerly
    with pytest.raises(TypeError):
        af.Dataset.from_arrays(arr, t, [1.0, 0.0], [0.0, 1.0], preprocess=lambda v: v)
