"""Lowering: run the builder once against the recording MNAContext and flatten what
it recorded into the circuit IR of include/cadnip_b200.h (``cb200_desc`` + the
per-lane parameter SoA).

This is the host half of ``build_with_detection`` + ``compile_structure``
(src/mna/solve.jl:1793-1822, src/mna/precompile.jl:312-366): typed indices are
resolved once ``n_nodes`` is final, the COO coordinate lists are emitted in
program order, and every ``stamp`` call becomes a device-table row.  The CSC
pattern and the COO->nz maps themselves are built by the C++ host code inside
``cb200_create``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .circuit import MNACircuit, MNASpec, Params
from .mna import (MNAContext, ZERO_VECTOR, CurrentIndex, DeviceRow)
from . import verilog_a as _va


class StructuralSweepError(ValueError):
    """The builder could not be traced with lane-array parameters (it branches on a
    swept value), i.e. the sweep crosses a structural boundary."""


@dataclass
class LoweredCircuit:
    n_nodes: int
    n_currents: int
    n_charges: int
    n_limits: int
    node_names: List[str]
    current_names: List[str]
    charge_names: List[str]
    limit_names: List[str]
    G_I: np.ndarray
    G_J: np.ndarray
    C_I: np.ndarray
    C_J: np.ndarray
    b_I: np.ndarray
    dev_kind: np.ndarray
    dev_flags: np.ndarray
    dev_node_ptr: np.ndarray
    dev_nodes: np.ndarray
    dev_param_ptr: np.ndarray
    dev_params: np.ndarray
    dev_gbase: np.ndarray
    dev_cbase: np.ndarray
    dev_bbase: np.ndarray
    uniform: np.ndarray
    limit_init_ref: np.ndarray
    lane_soa: np.ndarray            # [n_lane_cols][P]
    P: int
    dev_names: List[str]
    dev_user_nodes: List[List[int]]
    breakpoints: List[Any]
    va_models: List[Any] = field(default_factory=list)   # VAModel per DEV_VA flags value
    n_user_nodes: int = -1          # nodes allocated by get_node! (the rest are internal nodes)
    dev_state_ptr: Optional[np.ndarray] = None   # [n_dev+1] private state slots per device (VA set-up values)
    va_cuda_header: str = ""        # emitted CUDA for the circuit's Verilog-A modules (cb200_load_va_models)
    va_c_source: str = ""           # the same modules as plain C: input of the CPU oracle, never run by the product
    # outcome of the voltage-dependent-charge detection per detect_or_cached! call of one builder pass
    # (VA devices in stamping order, their reactive branches in branch order).  "Outcome only" is what
    # SURVEY 8c asks to carry over from the host; a checker that wants the SAME structure for a bulk
    # netlist (per-parameter-set verdicts, verilog_a._BULK_AFTER) replays it instead of probing.
    va_detect_seq: Optional[np.ndarray] = None

    @property
    def n(self) -> int:
        return self.n_nodes + self.n_currents + self.n_charges + self.n_limits

    @property
    def n_lane_cols(self) -> int:
        return int(self.lane_soa.shape[0])

    def unknown_names(self) -> List[str]:
        return self.node_names + self.current_names + self.charge_names + self.limit_names

    def index_of(self, name) -> int:
        """1-based unknown index of a node / current / charge / limit name."""
        if isinstance(name, (int, np.integer)):
            return int(name)
        names = self.unknown_names()
        try:
            return names.index(str(name)) + 1
        except ValueError:
            raise KeyError(f"no unknown named {name!r}; have {names}") from None

    # ------------------------------------------------------------------ #
    # Reviewable on-disk form (gzip'd JSON: integer / float tables as lists, the emitted CUDA
    # and C text as strings).  Replaces pickled fixtures: loading executes nothing.
    _ARRAYS = {"G_I": np.int64, "G_J": np.int64, "C_I": np.int64, "C_J": np.int64, "b_I": np.int64,
               "dev_kind": np.int32, "dev_flags": np.int32, "dev_node_ptr": np.int32, "dev_nodes": np.int32,
               "dev_param_ptr": np.int32, "dev_params": np.int32, "dev_gbase": np.int64, "dev_cbase": np.int64,
               "dev_bbase": np.int64, "uniform": np.float64, "limit_init_ref": np.int32, "va_detect_seq": np.int32,
               "dev_state_ptr": np.int32}
    _PLAIN = ("n_nodes", "n_currents", "n_charges", "n_limits", "node_names", "current_names", "charge_names",
              "limit_names", "P", "dev_names", "dev_user_nodes", "n_user_nodes", "va_cuda_header", "va_c_source")

    def to_dict(self) -> dict:
        from .analysis import breakpoints as _bp
        d = {k: getattr(self, k) for k in self._PLAIN}
        for k in self._ARRAYS:
            v = getattr(self, k)
            d[k] = None if v is None else [float(x).hex() for x in v] if self._ARRAYS[k] is np.float64 else [int(x) for x in v]
        d["lane_soa"] = [[float(x).hex() for x in col] for col in self.lane_soa]
        bps = []
        for w in self.breakpoints:              # waveform objects -> (times, period, count)
            t = w if (w is None or isinstance(w, (tuple, list))) else _bp(w)
            bps.append(None if t is None else [[float(x) for x in t[0]], float(t[1]), int(t[2])])
        d["breakpoints"] = bps
        if getattr(self, "lane_exprs", None) is not None:
            d["lane_exprs"] = list(self.lane_exprs)
        d["format"] = "cadnip-b200 lowered circuit v1"
        return d

    def save(self, path: str) -> None:
        import gzip
        import json
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(self.to_dict(), sort_keys=True).encode())

    @classmethod
    def from_dict(cls, d: dict) -> "LoweredCircuit":
        if d.get("format") != "cadnip-b200 lowered circuit v1":
            raise ValueError("not a cadnip-b200 lowered-circuit file")
        kw = {k: d[k] for k in cls._PLAIN}
        for k, dt in cls._ARRAYS.items():
            v = d.get(k)
            if v is None:
                kw[k] = None
            elif dt is np.float64:
                kw[k] = np.asarray([float.fromhex(x) for x in v], dtype=np.float64)
            else:
                kw[k] = np.asarray(v, dtype=dt)
        cols = [[float.fromhex(x) for x in col] for col in d["lane_soa"]]
        kw["lane_soa"] = (np.ascontiguousarray(cols, dtype=np.float64) if cols
                          else np.zeros((0, int(d["P"])), dtype=np.float64))
        kw["breakpoints"] = [None if b is None else (list(b[0]), float(b[1]), int(b[2])) for b in d["breakpoints"]]
        lc = cls(**kw)
        if "lane_exprs" in d:
            lc.lane_exprs = list(d["lane_exprs"])
        return lc

    @classmethod
    def load(cls, path: str) -> "LoweredCircuit":
        import gzip
        import json
        with gzip.open(path, "rb") as f:
            return cls.from_dict(json.loads(f.read().decode()))

    def param_value(self, ref: int) -> np.ndarray:
        """Value of a parameter reference for every lane, shape (P,)."""
        if ref >= 0:
            return np.full(self.P, self.uniform[ref])
        return self.lane_soa[~ref]

    # ------------------------------------------------------------------ #
    # The same netlist as plain arrays of *values* (one row per lane): the input
    # format of the CPU oracle.  Lives here because it is pure data reshaping;
    # the product never evaluates it.
    def netlist_tables(self):
        kind = self.dev_kind.astype(np.int32)
        flags = self.dev_flags.astype(np.int32)
        node_ptr = [0]
        nodes: List[int] = []
        for un in self.dev_user_nodes:
            nodes += un
            node_ptr.append(len(nodes))
        par = np.empty((self.P, len(self.dev_params)), dtype=np.float64)
        for j, ref in enumerate(self.dev_params):
            par[:, j] = self.param_value(int(ref))
        return dict(kind=kind, flags=flags,
                    node_ptr=np.asarray(node_ptr, dtype=np.int32),
                    nodes=np.asarray(nodes, dtype=np.int32),
                    par_ptr=self.dev_param_ptr.astype(np.int32),
                    par=np.ascontiguousarray(par),
                    n_nodes=self.n_user_nodes if self.n_user_nodes >= 0 else self.n_nodes)


class _ParamPool:
    def __init__(self, P: int):
        self.P = P
        self.uniform: List[float] = []
        self._u_index: Dict[float, int] = {}
        self.cols: List[np.ndarray] = []
        self._c_index: Dict[bytes, int] = {}

    def ref(self, v) -> int:
        a = np.asarray(v, dtype=np.float64)
        if a.ndim == 0:
            f = float(a)
            key = f if f == f else "nan"
            i = self._u_index.get(key)
            if i is None:
                i = len(self.uniform)
                self.uniform.append(f)
                self._u_index[key] = i
            return i
        if a.ndim != 1 or a.shape[0] != self.P:
            raise StructuralSweepError(
                f"device parameter has shape {a.shape}; expected a scalar or ({self.P},) lane array")
        a = np.ascontiguousarray(a)
        key = a.tobytes()
        i = self._c_index.get(key)
        if i is None:
            i = len(self.cols)
            self.cols.append(a)
            self._c_index[key] = i
        return ~i


_N_DETECTION_PASSES = 5


def _probe_stream():
    """The detection probes x in [-1, 1]: the reference draws them from
    ``MersenneTwister(0xDEADBEEF)`` (solve.jl:999, :1799), a stream that is not reproducible
    outside Julia; this xorshift64* restates the oracle's so both sides probe the same points."""
    s = 0xDEADBEEF
    mask = (1 << 64) - 1
    while True:
        s ^= s >> 12
        s ^= (s << 25) & mask
        s ^= s >> 27
        r = (s * 2685821657736338717) & mask
        yield ((r >> 11) / 9007199254740992.0 - 0.5) * 2.0


def _call_builder(builder, params, spec, x, ctx):
    from . import mna as _mna
    _mna._SPEC_STACK.append(spec)
    _mna._X_STACK.append(x)
    try:
        out = builder(params, spec, 0.0, x=x, ctx=ctx)
    except StructuralSweepError:
        raise
    except ValueError as e:
        if "ambiguous" in str(e):
            raise StructuralSweepError(
                "builder branches on a swept parameter: a CircuitSweep must not cross a "
                "structural boundary (run such points as separate sweeps)") from e
        raise
    finally:
        _mna._SPEC_STACK.pop()
        _mna._X_STACK.pop()
    if not isinstance(out, MNAContext):
        raise TypeError("builder must return the MNAContext it stamped into")
    return out


def run_builder(builder, params: Params, spec: MNASpec) -> MNAContext:
    """``build_with_detection`` (solve.jl:1793-1822): pass 1 runs the builder at
    ``x = ZERO_VECTOR``; passes 2-5 re-stamp the SAME context at random operating points so
    that Verilog-A devices can compare their branch charges across passes
    (``detect_or_cached!``, contrib.jl:214-257) and settle ``charge_is_vdep``.  The system may
    grow between passes (charge states appear), so each probe is sized to the previous pass.
    Circuits without reactive Verilog-A branches need only pass 1: the primitive devices'
    structure is independent of x."""
    ctx = _call_builder(builder, params, spec, ZERO_VECTOR, None)
    if not ctx.charge_Q_values:
        return ctx
    probes = _probe_stream()
    for _ in range(2, _N_DETECTION_PASSES + 1):
        known = ctx.system_size()
        x = np.fromiter((next(probes) for _ in range(known)), dtype=np.float64, count=known)
        ctx.reset_for_restamping()
        out = _call_builder(builder, params, spec, x, ctx)
        if out is not ctx:
            raise TypeError("builder must stamp into the context it is given (ctx=...) when re-run "
                            "for charge detection")
    return ctx


def lower(builder, params: Params, spec: MNASpec, P: int = 1) -> LoweredCircuit:
    ctx = run_builder(builder, params, spec)
    pool = _ParamPool(P)
    res = ctx.resolve_index

    def coords(lst):
        return np.asarray([res(i) for i in lst], dtype=np.int64)

    kind, flags, node_ptr, nodes, par_ptr, pars = [], [], [0], [], [0], []
    gbase, cbase, bbase, names, user_nodes = [], [], [], [], []
    state_ptr = [0]
    va_models: List[Any] = []
    for d in ctx.devices:
        f = d.flags
        if d.model is not None:                        # DEV_VA: flags = index of the model
            uids = [m.uid for m in va_models]
            if d.model.uid not in uids:
                va_models.append(d.model)
                uids.append(d.model.uid)
            f = uids.index(d.model.uid)
        kind.append(d.kind); flags.append(f); names.append(d.name)
        nodes += [res(i) for i in d.nodes]
        node_ptr.append(len(nodes))
        pars += [pool.ref(v) for v in d.params]
        par_ptr.append(len(pars))
        gbase.append(d.gbase); cbase.append(d.cbase); bbase.append(d.bbase)
        state_ptr.append(state_ptr[-1] + (d.model.n_state if d.model is not None else 0))
        user_nodes.append(list(d.user_nodes))
    gbase.append(len(ctx.G_I)); cbase.append(len(ctx.C_I)); bbase.append(len(ctx.b_I))
    limit_init_ref = [pool.ref(v) for v in ctx.limit_init]
    soa = (np.stack(pool.cols) if pool.cols else np.zeros((0, P), dtype=np.float64))
    return LoweredCircuit(
        n_nodes=ctx.n_nodes, n_currents=ctx.n_currents, n_charges=ctx.n_charges,
        n_limits=ctx.n_limits,
        node_names=list(ctx.node_names), current_names=list(ctx.current_names),
        charge_names=list(ctx.charge_names), limit_names=list(ctx.limit_names),
        G_I=coords(ctx.G_I), G_J=coords(ctx.G_J), C_I=coords(ctx.C_I), C_J=coords(ctx.C_J),
        b_I=coords(ctx.b_I),
        dev_kind=np.asarray(kind, dtype=np.int32), dev_flags=np.asarray(flags, dtype=np.int32),
        dev_node_ptr=np.asarray(node_ptr, dtype=np.int32), dev_nodes=np.asarray(nodes, dtype=np.int32),
        dev_param_ptr=np.asarray(par_ptr, dtype=np.int32), dev_params=np.asarray(pars, dtype=np.int32),
        dev_gbase=np.asarray(gbase, dtype=np.int64), dev_cbase=np.asarray(cbase, dtype=np.int64),
        dev_bbase=np.asarray(bbase, dtype=np.int64),
        uniform=np.asarray(pool.uniform, dtype=np.float64),
        limit_init_ref=np.asarray(limit_init_ref, dtype=np.int32),
        lane_soa=np.ascontiguousarray(soa, dtype=np.float64), P=P,
        dev_names=names, dev_user_nodes=user_nodes, breakpoints=list(ctx.breakpoints),
        va_models=va_models, n_user_nodes=_user_node_count(ctx),
        va_detect_seq=np.asarray([int(d.model.vdep[bi]) for d in ctx.devices if d.model is not None
                                  for bi in range(len(d.model.branches)) if d.model.reactive[bi]], dtype=np.int32),
        dev_state_ptr=np.asarray(state_ptr, dtype=np.int32),
        va_cuda_header=_va.cuda_header(va_models) if va_models else "",
        va_c_source=_va.c_source(va_models) if va_models else "")


def _user_node_count(ctx: MNAContext) -> int:
    """Nodes the builder allocated itself; internal nodes (alloc_internal_node!) must all
    come after them -- the reference's builders allocate every external node before the
    first stamp! call (offset-stability discipline, test/mna/pcnr.jl:33-36)."""
    flags = ctx.internal_node_flags
    n_user = sum(1 for f in flags if not f)
    if any(flags[:n_user]):
        raise ValueError("allocate all circuit nodes with get_node before stamping devices that "
                         "create internal nodes")
    return n_user


def lower_circuit(circuit: MNACircuit, spec: Optional[MNASpec] = None) -> LoweredCircuit:
    """Single circuit = one lane, every parameter uniform."""
    return lower(circuit.builder, circuit.params, spec or circuit.spec, P=1)
