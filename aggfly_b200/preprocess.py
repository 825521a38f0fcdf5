"""``preprocess`` of a raster, fused into the temporal kernel.

The reference resolves a config's ``preprocess`` into a callable that is applied lazily to the
dask-backed DataArray (aggfly/cli/preprocess.py:24-30 named builtins, :33-113 safe arithmetic
expressions in the single variable ``x``; applied in ``Dataset.__init__``,
aggfly/dataset/dataset.py:47-118).  For float32 data NumPy computes ``x - 273.15`` in float32
with the Python scalar cast to float32 (NEP 50 weak scalars), one IEEE operation per node.

Here the same specs resolve to a ``FusedPreprocess``: a short chain of (op, constant) pairs that
the temporal kernel applies to every raster value in the raster's dtype as it leaves shared
memory -- no extra pass over the 36 GB, same bits as the NumPy evaluation (the library is built
with ``--fmad=false``, so ``x * a + b`` is two roundings exactly like NumPy's two ufuncs).
Calling the object on a NumPy array evaluates the chain with NumPy (what the oracle / tests do).
"""
from __future__ import annotations

import ast
from typing import List, Tuple

import numpy as np

# op codes == AGF_PRE_* in include/aggfly_b200.h
ADD, SUB, RSUB, MUL, DIV, RDIV, NEG = range(7)
MAX_OPS = 4

BUILTINS = {                                  # aggfly/cli/preprocess.py:24-30
    "identity": "x",
    "kelvin_to_celsius": "x - 273.15",
    "celsius_to_kelvin": "x + 273.15",
    "pa_to_kpa": "x / 1000.0",
    "m_to_mm": "x * 1000.0",
}


class PreprocessError(Exception):
    """Same name as the reference's (aggfly/cli/preprocess.py:33)."""


class FusedPreprocess:
    """A chain ``x -> op_k(... op_1(x))`` of at most ``MAX_OPS`` elementwise operations."""

    def __init__(self, ops: List[Tuple[int, float]], source: str = ""):
        if len(ops) > MAX_OPS:
            raise PreprocessError(f"expression {source!r} needs {len(ops)} operations; the fused preprocess "
                                  f"holds at most {MAX_OPS}")
        self.ops = [(int(o), float(c)) for o, c in ops]
        self.source = source

    def __call__(self, x):
        """NumPy evaluation with NumPy's own promotion rules (python-scalar constants are weak)."""
        for op, c in self.ops:
            if op == ADD:
                x = x + c
            elif op == SUB:
                x = x - c
            elif op == RSUB:
                x = c - x
            elif op == MUL:
                x = x * c
            elif op == DIV:
                x = x / c
            elif op == RDIV:
                x = c / x
            else:
                x = -x
        return x

    def __repr__(self):
        return f"FusedPreprocess({self.source!r}, ops={self.ops})"


def _lower(node):
    """-> ("const", value) | ("chain", [ops]) for the allow-listed arithmetic AST."""
    if isinstance(node, ast.Expression):
        return _lower(node.body)
    if isinstance(node, ast.Constant):
        if not isinstance(node.value, (int, float)) or isinstance(node.value, bool):
            raise PreprocessError(f"only numeric constants are allowed, got {node.value!r}")
        return ("const", node.value)
    if isinstance(node, ast.Name):
        if node.id != "x":
            raise PreprocessError(f"only the variable 'x' is allowed, got {node.id!r}")
        return ("chain", [])
    if isinstance(node, ast.UnaryOp):
        kind, val = _lower(node.operand)
        if isinstance(node.op, ast.UAdd):
            return (kind, val)
        if isinstance(node.op, ast.USub):
            return ("const", -val) if kind == "const" else ("chain", val + [(NEG, 0.0)])
        raise PreprocessError(f"unary {type(node.op).__name__} is not allowed")
    if isinstance(node, ast.BinOp):
        lk, lv = _lower(node.left)
        rk, rv = _lower(node.right)
        op = type(node.op)
        if lk == "const" and rk == "const":                 # folded by Python, in Python arithmetic
            import operator
            fold = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
                    ast.Pow: operator.pow, ast.Mod: operator.mod, ast.FloorDiv: operator.floordiv}
            if op not in fold:
                raise PreprocessError(f"operator {op.__name__} is not allowed")
            return ("const", fold[op](lv, rv))
        if lk == "chain" and rk == "chain":
            raise PreprocessError("expressions that combine x with x (x * x, x / x ...) are not fusable; "
                                  "use a ('transform', {'transform': 'power', ...}) step or preprocess the array")
        table = {ast.Add: (ADD, ADD), ast.Sub: (SUB, RSUB), ast.Mult: (MUL, MUL), ast.Div: (DIV, RDIV)}
        if op not in table:
            raise PreprocessError(f"operator {op.__name__} on x is not fusable (supported: + - * /)")
        if lk == "chain":
            return ("chain", lv + [(table[op][0], rv)])
        return ("chain", rv + [(table[op][1], lv)])
    raise PreprocessError(f"expression element {type(node).__name__} is not allowed "
                          "(only arithmetic on 'x' and numbers)")


def compile_expression(expr: str) -> FusedPreprocess:
    """A safe arithmetic-in-``x`` string (``"x - 273.15"``, ``"(x - 32) * 5 / 9"``) -> fused chain."""
    try:
        tree = ast.parse(expr, mode="eval")
    except SyntaxError as e:
        raise PreprocessError(f"could not parse expression {expr!r}: {e.msg}")
    kind, val = _lower(tree)
    if kind != "chain":
        raise PreprocessError(f"expression {expr!r} must use the variable 'x' (e.g. 'x - 273.15')")
    return FusedPreprocess(val, expr)


def resolve(spec) -> "FusedPreprocess | None":
    """``None`` | builtin name | expression string | FusedPreprocess -> FusedPreprocess (or None)."""
    if spec is None:
        return None
    if isinstance(spec, FusedPreprocess):
        return spec
    if isinstance(spec, str):
        pre = compile_expression(BUILTINS.get(spec.strip(), spec))
        return pre if pre.ops else None
    raise TypeError(f"preprocess must be a builtin name, an expression in x, or a FusedPreprocess; got {type(spec)}")


def constants_for(ops, dtype) -> List[Tuple[int, float]]:
    """Constants as the kernel applies them: rounded once to the raster dtype (NEP 50)."""
    dt = np.dtype(dtype).type
    return [(op, float(dt(c))) for op, c in ops]
