"""Sweep iterator algebra and CircuitSweep -- host mirror of src/sweeps.jl:58-424.

Iteration semantics follow the reference exactly (each point is a tuple of
``(selector, value)`` pairs sorted by selector, ``ProductSweep`` varies its first
axis fastest, ``SerialSweep`` yields ``None`` for variables the active sub-sweep
does not set).  In addition every sweep can report itself as *columns* -- one
numpy array per swept variable, lane index = iteration order -- which is how a
sweep becomes the per-lane parameter struct-of-arrays of the B200 path without a
Python loop over points.
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Set, Tuple

import numpy as np

from .circuit import MNACircuit, MNASpec, Params, alter

Point = Tuple[Tuple[str, Any], ...]


def _expand(x) -> Point:
    """``expand`` (sweeps.jl:137-145): flatten nesting, sort by selector."""
    rets: List[Tuple[str, Any]] = []

    def rec(v):
        if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], str):
            rets.append(v)
        else:
            for w in v:
                rec(w)

    rec(x)
    return tuple(sorted(rets, key=lambda kv: kv[0]))


class SweepLike:
    def sweepvars(self) -> Set[str]:
        raise NotImplementedError

    def __len__(self) -> int:
        raise NotImplementedError

    def size(self) -> Tuple[int, ...]:
        return (len(self),)

    def columns(self) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """(values, isset): per swept variable an array over lanes, plus a boolean
        mask that is False where the point leaves the variable at its default."""
        raise NotImplementedError

    def points(self) -> List[Dict[str, Any]]:
        return [{k: v for k, v in pt if v is not None} for pt in self]


class Sweep(SweepLike):
    """``Sweep(selector, values)`` / ``Sweep(R1=values)`` (sweeps.jl:181-202)."""

    def __init__(self, selector=None, values=None, **kwargs):
        if selector is None:
            if len(kwargs) != 1:
                raise ValueError("`Sweep` takes a single variable at a time!")
            (selector, values), = kwargs.items()
        elif isinstance(selector, tuple) and values is None:
            selector, values = selector
        elif kwargs:
            raise ValueError("`Sweep` takes a single variable at a time!")
        self.selector = str(selector)
        vals = np.asarray(list(values) if not isinstance(values, np.ndarray) else values)
        self.values = vals.reshape(-1)

    def __eq__(self, other):
        return (isinstance(other, Sweep) and self.selector == other.selector
                and np.array_equal(self.values, other.values))

    def __iter__(self) -> Iterator[Point]:
        for v in self.values:
            yield ((self.selector, _py(v)),)

    def __len__(self):
        return int(self.values.shape[0])

    def sweepvars(self):
        return {self.selector}

    def columns(self):
        return ({self.selector: self.values.astype(np.float64)},
                {self.selector: np.ones(len(self), dtype=bool)})

    def __repr__(self):
        if len(self) > 1:
            return (f"Sweep of {self.selector} with {len(self)} values over "
                    f"[{self.values.min()} .. {self.values.max()}]")
        return f"Sweep of {self.selector} set to {self.values[0]}"


def _py(v):
    return v.item() if isinstance(v, np.generic) else v


def _as_sweeps(args, kwargs) -> List[SweepLike]:
    out: List[SweepLike] = []
    for a in args:
        out.append(a if isinstance(a, SweepLike) else Sweep(a))
    for k, v in kwargs.items():
        out.append(Sweep(k, v))
    return out


class _Product(SweepLike):
    def __init__(self, its: List[SweepLike]):
        self.iterators = its

    def __iter__(self):
        # Iterators.product: FIRST axis fastest (sweeps.jl:272)
        for combo in itertools.product(*[list(it) for it in reversed(self.iterators)]):
            yield _expand(tuple(reversed(combo)))

    def __len__(self):
        n = 1
        for it in self.iterators:
            n *= len(it)
        return n

    def size(self):
        return tuple(len(it) for it in self.iterators)

    def sweepvars(self):
        return set().union(*[it.sweepvars() for it in self.iterators])

    def columns(self):
        vals: Dict[str, np.ndarray] = {}
        isset: Dict[str, np.ndarray] = {}
        inner = 1
        total = len(self)
        for it in self.iterators:
            m = len(it)
            v, s = it.columns()
            outer = total // (inner * m) if inner * m else 0
            for k in v:
                vals[k] = np.tile(np.repeat(v[k], inner), outer)
                isset[k] = np.tile(np.repeat(s[k], inner), outer)
            inner *= m
        return vals, isset


class _Tandem(SweepLike):
    def __init__(self, its: List[SweepLike]):
        lens = [len(i) for i in its]
        if any(l != lens[0] for l in lens):
            raise ValueError("TandemSweep requires all sweeps be of the same length!")
        self.iterators = its

    def __iter__(self):
        for combo in zip(*self.iterators):
            yield _expand(combo)

    def __len__(self):
        return len(self.iterators[0])

    def sweepvars(self):
        return set().union(*[it.sweepvars() for it in self.iterators])

    def columns(self):
        vals, isset = {}, {}
        for it in self.iterators:
            v, s = it.columns()
            vals.update(v); isset.update(s)
        return vals, isset


class _Serial(SweepLike):
    def __init__(self, its: List[SweepLike]):
        self.iterators = its
        self.vars = set().union(*[it.sweepvars() for it in its])

    def __iter__(self):
        # sweeps.jl:306-330: every variable appears, unset ones are `nothing`
        for it in self.iterators:
            for pt in it:
                m = {v: None for v in self.vars}
                m.update(dict(pt))
                yield tuple(sorted(m.items(), key=lambda kv: kv[0]))

    def __len__(self):
        return sum(len(it) for it in self.iterators)

    def sweepvars(self):
        return set(self.vars)

    def columns(self):
        total = len(self)
        vals = {v: np.zeros(total) for v in self.vars}
        isset = {v: np.zeros(total, dtype=bool) for v in self.vars}
        off = 0
        for it in self.iterators:
            m = len(it)
            v, s = it.columns()
            for k in v:
                vals[k][off:off + m] = v[k]
                isset[k][off:off + m] = s[k]
            off += m
        return vals, isset


def ProductSweep(*args, **kwargs) -> SweepLike:
    """sweeps.jl:267-273."""
    its = _as_sweeps(args, kwargs)
    return its[0] if len(its) == 1 else _Product(its)


def TandemSweep(*args, **kwargs) -> SweepLike:
    """sweeps.jl:284-295."""
    its = _as_sweeps(args, kwargs)
    return its[0] if len(its) == 1 else _Tandem(its)


def SerialSweep(*args, **kwargs) -> SweepLike:
    """sweeps.jl:306-320."""
    its = _as_sweeps(args, kwargs)
    return its[0] if len(its) == 1 else _Serial(its)


def sweepvars(*sweeps) -> Set[str]:
    out: Set[str] = set()
    for s in sweeps:
        out |= s.sweepvars()
    return out


def split_axes(ps: SweepLike, axes: Iterable[str]):
    """``split_axes`` (sweeps.jl:101-134): (outer, inner) of a ProductSweep."""
    if not isinstance(ps, _Product):
        raise ValueError("split_axes only works with ProductSweep objects!")
    idxs = []
    for ax in axes:
        for i, it in enumerate(ps.iterators):
            if isinstance(it, Sweep) and it.selector == str(ax):
                idxs.append(i)
                break
        else:
            raise ValueError(f"Unable to find product axis matching '{ax}'")
    inner = [ps.iterators[i] for i in idxs]
    outer = [it for i, it in enumerate(ps.iterators) if i not in idxs]
    mk = lambda its: its[0] if len(its) == 1 else _Product(its)
    return mk(outer), mk(inner)


def sweepify(x) -> SweepLike:
    """sweeps.jl:352-360."""
    if isinstance(x, SweepLike):
        return x
    if isinstance(x, dict):
        return ProductSweep(**x)
    if isinstance(x, (list, tuple)):
        return SerialSweep(*[sweepify(v) for v in x])
    return Sweep(x)


def find_param_ranges(params: SweepLike) -> Dict[str, Tuple[float, float, int]]:
    """sweeps.jl:709-749."""
    ranges: Dict[str, List[Tuple[float, float, int]]] = {}

    def rec(it):
        if isinstance(it, Sweep):
            ranges.setdefault(it.selector, []).append(
                (float(it.values.min()), float(it.values.max()), len(it)))
        else:
            for c in it.iterators:
                rec(c)

    rec(params)
    out = {}
    for k, rs in ranges.items():
        out[k] = (min(r[0] for r in rs), max(r[1] for r in rs), sum(r[2] for r in rs))
    return out


class CircuitSweep:
    """``CircuitSweep(builder, iterator; spec, default_params...)`` (sweeps.jl:387-424).
    Iterating yields the altered ``MNACircuit`` per point."""

    def __init__(self, builder, iterator: SweepLike, spec: Optional[MNASpec] = None,
                 circuit: Optional[MNACircuit] = None, **default_params):
        if not isinstance(iterator, SweepLike):
            raise ValueError("Must give some kind of iterator!")
        self.builder = builder
        self.iterator = iterator
        if circuit is None:
            base = MNACircuit(builder, spec=spec or MNASpec(), **default_params)
            first = next(iter(iterator), None)
            if first is not None:
                base = alter(base, **dict(first))
            circuit = base
        self.circuit = circuit

    def __len__(self):
        return len(self.iterator)

    def size(self):
        return self.iterator.size()

    def sweepvars(self):
        return self.iterator.sweepvars()

    def __iter__(self):
        for pt in self.iterator:
            yield alter(self.circuit, **dict(pt))

    def lane_params(self) -> Tuple[Params, int]:
        """The base circuit's parameter tree with every swept variable replaced by
        its lane column (unset SerialSweep entries keep the base value)."""
        vals, isset = self.iterator.columns()
        P = len(self.iterator)
        p = self.circuit.params
        for name, col in vals.items():
            col = np.asarray(col, dtype=np.float64)
            mask = isset[name]
            if not mask.all():
                base = _get_path(p, name.split("."))
                if base is None:
                    raise KeyError(f"sweep leaves {name!r} unset and the circuit has no default")
                col = np.where(mask, col, float(base))
            p = p.set_path(name.split("."), col)
        return p, P


def _get_path(p: Params, parts):
    cur: Any = p
    for part in parts:
        if not isinstance(cur, Params) or part not in cur:
            return None
        cur = cur[part]
    return cur


class SweepResult:
    """``SweepResult{P,S}`` (sweeps.jl:477-487): aligned points and solutions;
    iterates as ``(params, sol)`` pairs.  ``solutions`` may be a lazy sequence."""

    def __init__(self, points: Sequence, solutions: Sequence):
        self.points = points
        self.solutions = solutions

    def __len__(self):
        return len(self.solutions)

    def __iter__(self):
        for i in range(len(self)):
            yield self.points[i], self.solutions[i]

    def __getitem__(self, i):
        return self.points[i], self.solutions[i]
