"""MNASpec / MNACircuit / alter -- host mirror of src/mna/solve.jl:57-135, :1585-1732."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any, Callable, Dict, Iterable, Mapping, Optional

import numpy as np

MODES = {"dcop": 0, "tran": 1, "tranop": 2, "ac": 3}


@dataclass(frozen=True)
class MNASpec:
    """``MNASpec`` (solve.jl:57-70); defaults identical."""
    temp: float = 27.0
    mode: str = "tran"
    time: float = 0.0
    gmin: float = 1e-12
    gshunt: float = 0.0
    srcFact: float = 1.0
    tnom: float = 27.0
    abstol: float = 1e-12
    reltol: float = 1e-3
    vntol: float = 1e-6
    iabstol: float = 1e-12

    def __post_init__(self):
        if self.mode not in MODES:
            raise ValueError(f"unknown analysis mode {self.mode!r}")


def with_mode(obj, mode: str):
    """``with_mode(spec, mode)`` keeps every field (solve.jl:88-90);
    ``with_mode(circuit, mode)`` rebuilds the spec from ``temp`` and ``mode`` only,
    resetting gmin/gshunt/tolerances to defaults (solve.jl:1976-1979)."""
    if isinstance(obj, MNASpec):
        return replace(obj, mode=mode)
    if isinstance(obj, MNACircuit):
        return MNACircuit(obj.builder, obj.params, MNASpec(temp=obj.spec.temp, mode=mode))
    raise TypeError("with_mode expects an MNASpec or MNACircuit")


def with_temp(spec: MNASpec, temp: float) -> MNASpec:
    return replace(spec, temp=float(temp))


def with_gshunt(spec: MNASpec, gshunt: float) -> MNASpec:
    return replace(spec, gshunt=float(gshunt))


def with_srcfact(spec: MNASpec, srcFact: float) -> MNASpec:
    return replace(spec, srcFact=float(srcFact))


class Params(Mapping):
    """Immutable, attribute-accessible parameter tree: the NamedTuple the reference
    passes to builders.  Leaves are floats, or 1-D arrays (one value per sweep lane)
    while a CircuitSweep is being lowered."""

    def __init__(self, mapping: Optional[Mapping] = None, **kw):
        d: Dict[str, Any] = {}
        for src in (mapping or {}), kw:
            for k, v in src.items():
                d[str(k)] = Params(v) if isinstance(v, Mapping) and not isinstance(v, Params) else v
        object.__setattr__(self, "_d", d)

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise AttributeError(f"circuit has no parameter {k!r}") from None

    def __setattr__(self, k, v):
        raise AttributeError("Params is immutable; use alter()")

    def __getitem__(self, k):
        return self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __repr__(self):
        return "Params(" + ", ".join(f"{k}={v!r}" for k, v in self._d.items()) + ")"

    def get(self, k, default=None):
        return self._d.get(k, default)

    def set_path(self, parts, value) -> "Params":
        """``_set_param_path`` (solve.jl:1663-1680): dotted selector = path."""
        parts = list(parts)
        d = dict(self._d)
        if len(parts) == 1:
            d[parts[0]] = value
        else:
            sub = d.get(parts[0], Params())
            if not isinstance(sub, Params):
                sub = Params()
            d[parts[0]] = sub.set_path(parts[1:], value)
        return Params(d)


class MNACircuit:
    """``MNACircuit(builder, params, spec)`` / ``MNACircuit(builder; spec, kw...)``
    (solve.jl:1585-1600, :1602-1640).  ``builder(params, spec, t=0.0, x=ZERO_VECTOR,
    ctx=None) -> MNAContext``."""

    def __init__(self, builder: Callable, params: Optional[Mapping] = None,
                 spec: Optional[MNASpec] = None, **kwargs):
        if not callable(builder):
            raise TypeError("builder must be callable")
        p = params if isinstance(params, Params) else Params(params or {})
        for sel, val in kwargs.items():
            p = p.set_path(str(sel).split("."), val)
        self.builder = builder
        self.params = p
        self.spec = spec if spec is not None else MNASpec()


def alter(circuit: MNACircuit, spec: Optional[MNASpec] = None, **kwargs) -> MNACircuit:
    """``alter`` (solve.jl:1719-1732): dotted selectors are paths, ``None`` skipped."""
    p = circuit.params
    for sel, val in kwargs.items():
        if val is None:
            continue
        p = p.set_path(str(sel).split("."), val)
    return MNACircuit(circuit.builder, p, spec if spec is not None else circuit.spec)
