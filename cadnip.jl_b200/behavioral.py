"""Behavioural sources (B elements): ``BehavioralVoltageSource`` / ``BehavioralCurrentSource``
(src/mna/devices.jl:1003-1058, stamps :1079-1131).

The reference stores a Julia closure ``value_fn(get_voltage)`` and calls it on the host at every
rebuild with the current iterate; the value goes into ``b`` only (an explicit source: no Jacobian
entries for its dependence on the node voltages).  A closure cannot run on the device, so here the
closure is TRACED once on the host: ``get_voltage(name)`` returns a symbolic node voltage, the
arithmetic and numpy functions applied to it build an expression, and that expression becomes a
one-statement Verilog-A module that goes through the emitter like any other model (CUDA for the
kernels, C for the oracle).  ``$explicit(expr)`` -- value without partials -- keeps the reference's
explicit stamping: ``stamp_b!`` with the value at the current iterate, nothing in G.

Closures may use ``+ - * / **``, unary minus, ``abs``, comparisons inside ``np.where`` and the numpy
functions of `_UFUNCS`; Python floats become literals, lane arrays (swept parameters) become
per-lane module parameters.
"""
from __future__ import annotations

import hashlib
from typing import Any, Callable, Dict, List, Tuple

import numpy as np

from .verilog_a import VAError, VAModel

_UFUNCS = {"exp": "exp", "log": "ln", "log10": "log", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
           "tanh": "tanh", "sinh": "sinh", "cosh": "cosh", "arctan": "atan", "arcsin": "asin", "arccos": "acos",
           "absolute": "abs", "fabs": "abs", "floor": "floor", "ceil": "ceil"}
_BINARY = {"add": "+", "subtract": "-", "multiply": "*", "divide": "/", "true_divide": "/",
           "less": "<", "less_equal": "<=", "greater": ">", "greater_equal": ">="}
_BINARY_FN = {"power": "pow", "float_power": "pow", "minimum": "min", "maximum": "max", "arctan2": "atan2",
              "hypot": "hypot"}


class _Trace:
    """What one traced closure referenced: controlling nodes (in order of first use) and lane-array
    constants (module parameters)."""

    def __init__(self):
        self.nodes: List[Any] = []
        self.consts: List[np.ndarray] = []

    def node(self, name) -> "_Expr":
        if name not in self.nodes:
            self.nodes.append(name)
        return _Expr(self, f"V(c{self.nodes.index(name)})")

    def lift(self, x) -> "_Expr":
        if isinstance(x, _Expr):
            return x
        a = np.asarray(x, dtype=np.float64)
        if a.ndim == 0 or a.size == 1:
            return _Expr(self, _literal(float(a.reshape(-1)[0])))
        for i, c in enumerate(self.consts):
            if c is x or (c.shape == a.shape and np.array_equal(c, a)):
                return _Expr(self, f"k{i}")
        self.consts.append(a)
        return _Expr(self, f"k{len(self.consts) - 1}")


def _literal(v: float) -> str:
    if v != v or v in (float("inf"), float("-inf")):
        raise VAError("behavioural source: non-finite constant")
    r = repr(v)
    if "e" not in r and "." not in r:
        r += ".0"
    return f"({r})"


class _Expr:
    """A node of the traced expression; ``text`` is Verilog-A."""
    __array_priority__ = 1000.0
    __slots__ = ("tr", "text")

    def __init__(self, tr: _Trace, text: str):
        self.tr, self.text = tr, text

    def _bin(self, op, other, swap=False):
        o = self.tr.lift(other)
        a, b = (o, self) if swap else (self, o)
        return _Expr(self.tr, f"({a.text} {op} {b.text})")

    def _fn(self, fn, *args):
        return _Expr(self.tr, f"{fn}({', '.join(self.tr.lift(a).text for a in args)})")

    __add__ = lambda s, o: s._bin("+", o)
    __radd__ = lambda s, o: s._bin("+", o, True)
    __sub__ = lambda s, o: s._bin("-", o)
    __rsub__ = lambda s, o: s._bin("-", o, True)
    __mul__ = lambda s, o: s._bin("*", o)
    __rmul__ = lambda s, o: s._bin("*", o, True)
    __truediv__ = lambda s, o: s._bin("/", o)
    __rtruediv__ = lambda s, o: s._bin("/", o, True)
    __lt__ = lambda s, o: s._bin("<", o)
    __le__ = lambda s, o: s._bin("<=", o)
    __gt__ = lambda s, o: s._bin(">", o)
    __ge__ = lambda s, o: s._bin(">=", o)
    __neg__ = lambda s: _Expr(s.tr, f"(-{s.text})")
    __pos__ = lambda s: s
    __abs__ = lambda s: s._fn("abs", s)
    __pow__ = lambda s, o: s._fn("pow", s, o)
    __rpow__ = lambda s, o: s._fn("pow", o, s)

    def __bool__(self):
        raise VAError("behavioural source: the value function branches on a node voltage; use np.where(cond, a, b)")

    def __float__(self):
        raise VAError("behavioural source: the value function converts a node voltage to float (math.* functions); "
                      "use the numpy functions")

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if name in _UFUNCS and len(inputs) == 1:
            return self._fn(_UFUNCS[name], inputs[0])
        if name == "negative":
            return -self.tr.lift(inputs[0])
        if name == "square":
            x = self.tr.lift(inputs[0])
            return x * x
        if name in _BINARY and len(inputs) == 2:
            return self.tr.lift(inputs[0])._bin(_BINARY[name], inputs[1])
        if name in _BINARY_FN and len(inputs) == 2:
            return self._fn(_BINARY_FN[name], inputs[0], inputs[1])
        raise VAError(f"behavioural source: numpy function {name!r} has no device counterpart")

    def __array_function__(self, func, types, args, kwargs):
        if func is np.where and len(args) == 3:
            c, a, b = (self.tr.lift(x) for x in args)
            return _Expr(self.tr, f"(({c.text}) ? {a.text} : {b.text})")
        if func is np.clip and len(args) == 3:
            x, lo, hi = args
            return self._fn("min", self._fn("max", x, lo), hi)
        raise VAError(f"behavioural source: numpy function {func.__name__!r} has no device counterpart")


class _Behavioral:
    kind = "?"

    def __init__(self, value_fn: Callable, name: str = "B"):
        self.value_fn, self.name = value_fn, str(name)
        self._lowered = None

    def lower(self) -> Tuple[Any, List[Any]]:
        """(VA instance, controlling node names): traced once per source object."""
        if self._lowered is None:
            tr = _Trace()
            val = tr.lift(self.value_fn(tr.node))
            ports = ["p", "n"] + [f"c{i}" for i in range(len(tr.nodes))]
            pars = "".join(f"    parameter real k{i} = 0.0;\n" for i in range(len(tr.consts)))
            contrib = ("V(p, n) <+ $explicit(%s);" if self.kind == "V" else "I(n, p) <+ $explicit(%s);") % val.text
            uid = hashlib.sha256((self.kind + val.text).encode()).hexdigest()[:10]
            src = (f"module bsrc_{self.kind.lower()}_{uid}({', '.join(ports)});\n{pars}    inout {', '.join(ports)};\n"
                   f"    electrical {', '.join(ports)};\n    analog begin\n        {contrib}\n    end\nendmodule\n")
            model = _MODEL_CACHE.get(src)
            if model is None:
                model = _MODEL_CACHE[src] = VAModel(src)
            inst = model(name=self.name, **{f"k{i}": c for i, c in enumerate(tr.consts)})
            self._lowered = (inst, list(tr.nodes))
        return self._lowered


_MODEL_CACHE: Dict[str, VAModel] = {}


class BehavioralVoltageSource(_Behavioral):
    """``BehavioralVoltageSource(value_fn; name=:B)`` (devices.jl:1003-1029): V(p) - V(n) =
    value_fn(get_voltage); allocates the branch current ``I_<name>``."""
    kind = "V"


class BehavioralCurrentSource(_Behavioral):
    """``BehavioralCurrentSource(value_fn; name=:B)`` (devices.jl:1032-1058): the current
    value_fn(get_voltage) flows from n to p (into the positive terminal)."""
    kind = "I"
