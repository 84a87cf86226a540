"""Host-side mirror of Cadnip's MNA stamping API (structure discovery only).

Mirrors ``src/mna/context.jl`` (MNAContext, typed indices, ``get_node!``,
``alloc_*!``, ``stamp_G!/C!/b!``) and the ``stamp!`` methods of
``src/mna/devices.jl`` with the ``!`` dropped from the names.  In the reference
the builder is run once per Newton iteration and the devices compute their own
values; here the builder is run ONCE on the host, against a *recording*
context, and what it records -- the COO coordinates in program order plus one
device-table row per ``stamp`` call -- is everything the CUDA side needs to
re-evaluate the devices for every sweep lane and Newton iteration.

Device parameters may be Python floats (uniform over the sweep) or 1-D numpy
arrays of length ``P`` (one value per lane): running the builder once with
lane-array parameters is how a ``CircuitSweep`` becomes a struct of arrays.
No device arithmetic happens on the host: stamp *values* are produced only by
the CUDA kernels (there is no CPU fallback).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

# device kind / wave ids: must match include/cadnip_b200.h
DEV_RESISTOR, DEV_CAPACITOR, DEV_INDUCTOR, DEV_VSOURCE, DEV_ISOURCE = 1, 2, 3, 4, 5
DEV_VCVS, DEV_VCCS, DEV_CCVS, DEV_CCCS, DEV_DIODE, DEV_DIODECAP, DEV_SIMPLEMOS = 6, 7, 8, 9, 10, 11, 12
DEV_VA = 13
WAVE_NONE, WAVE_PWL, WAVE_PULSE, WAVE_SIN = 0, 1, 2, 3


# --------------------------------------------------------------------------- #
# typed indices  (src/mna/context.jl:40-101)
# --------------------------------------------------------------------------- #
@dataclass(frozen=True)
class CurrentIndex:
    k: int


@dataclass(frozen=True)
class ChargeIndex:
    k: int


@dataclass(frozen=True)
class LimitIndex:
    k: int


Index = Union[int, CurrentIndex, ChargeIndex, LimitIndex]


def _iszero(i: Index) -> bool:
    """Ground check for stamps (context.jl:95-99)."""
    return isinstance(i, (int, np.integer)) and int(i) == 0


class ZeroVector:
    """Phantom state vector: 0.0 for any index (context.jl:107-125)."""

    def __len__(self):
        return 0

    def __getitem__(self, i):
        return 0.0


ZERO_VECTOR = ZeroVector()

_GROUND_NAMES = ("gnd", "0", "gnd!")


@dataclass
class DeviceRow:
    """One ``stamp!`` call: what the device-table export records."""
    kind: int
    flags: int
    name: str
    user_nodes: List[int]          # terminals exactly as passed (ints; for the
                                   # "existing current" forms the CurrentIndex.k)
    nodes: List[Index]             # terminals + indices allocated by the stamp
    params: List[Any]              # float or ndarray(P,)
    gbase: int
    cbase: int
    bbase: int
    model: Any = None              # VAModel for DEV_VA rows


class MNAContext:
    """COO-recording context (src/mna/context.jl:248-372)."""

    def __init__(self):
        self.node_names: List[str] = []
        self.node_to_idx: Dict[str, int] = {}
        self.internal_node_flags: List[bool] = []
        self.n_nodes = 0
        self.current_names: List[str] = []
        self.n_currents = 0
        self.charge_names: List[str] = []
        self.charge_branches: List[Tuple[int, int]] = []
        self.n_charges = 0
        self.limit_names: List[str] = []
        self.limit_branches: List[Tuple[int, int]] = []
        self.limit_init: List[Any] = []
        self.n_limits = 0
        self.G_I: List[Index] = []
        self.G_J: List[Index] = []
        self.G_V: List[Any] = []
        self.C_I: List[Index] = []
        self.C_J: List[Index] = []
        self.C_V: List[Any] = []
        self.b_I: List[Index] = []
        self.b_V: List[Any] = []
        self.devices: List[DeviceRow] = []
        self.breakpoints: List[Any] = []
        # voltage-dependent-charge detection cache (context.jl:338-341): survives restamping
        if not hasattr(self, "charge_is_vdep"):
            self.charge_is_vdep: List[bool] = []
            self.charge_Q_values: List[float] = []
            self.charge_V_values: List[float] = []
        self.charge_detection_pos = 0

    def detect_or_cached(self, V: float, Q: float) -> bool:
        """``detect_or_cached!`` (src/mna/contrib.jl:214-257): positional, sticky."""
        pos = self.charge_detection_pos
        self.charge_detection_pos = pos + 1
        if pos >= len(self.charge_Q_values):              # first run: assume linear
            self.charge_is_vdep.append(False)
            self.charge_Q_values.append(float(Q))
            self.charge_V_values.append(float(V))
            return False
        Vs, Qs = self.charge_V_values[pos], self.charge_Q_values[pos]
        if abs(V) > 1e-6 and abs(Vs) > 1e-6:
            Cc, Cs = Q / V, Qs / Vs
            diff, maxC = abs(Cc - Cs), max(abs(Cc), abs(Cs))
            if diff > 1e-15 and (maxC < 1e-30 or diff / maxC > 1e-6):
                self.charge_is_vdep[pos] = True
        self.charge_Q_values[pos] = float(Q)
        self.charge_V_values[pos] = float(V)
        return self.charge_is_vdep[pos]

    # -- sizes ------------------------------------------------------------- #
    def system_size(self) -> int:
        """context.jl:438."""
        return self.n_nodes + self.n_currents + self.n_charges + self.n_limits

    # -- allocation -------------------------------------------------------- #
    def get_node(self, name) -> int:
        """``get_node!`` (context.jl:467-486); ints pass through (:500)."""
        if isinstance(name, (int, np.integer)):
            return int(name)
        name = str(name)
        if name in _GROUND_NAMES:
            return 0
        idx = self.node_to_idx.get(name, 0)
        if idx != 0:
            return idx
        self.n_nodes += 1
        self.node_names.append(name)
        self.node_to_idx[name] = self.n_nodes
        self.internal_node_flags.append(False)
        return self.n_nodes

    def alloc_internal_node(self, name: str, instance_name: str = "") -> int:
        """``alloc_internal_node!`` (context.jl:654-679)."""
        full = name if instance_name == "" else f"{instance_name}_{name}"
        idx = self.get_node(full)
        self.internal_node_flags[idx - 1] = True
        return idx

    def alloc_current(self, *name_parts: str) -> CurrentIndex:
        """``alloc_current!`` (context.jl:523-527, :548-562): parts concatenated."""
        self.n_currents += 1
        self.current_names.append("".join(str(p) for p in name_parts))
        return CurrentIndex(self.n_currents)

    def get_current_idx(self, name: str) -> CurrentIndex:
        """context.jl:591-595."""
        try:
            return CurrentIndex(self.current_names.index(str(name)) + 1)
        except ValueError:
            raise KeyError(f"Current variable {name} not found in MNA context")

    def alloc_charge(self, name: str, p: int, n: int) -> ChargeIndex:
        """``alloc_charge!`` (context.jl:741-746)."""
        self.n_charges += 1
        self.charge_names.append(str(name))
        self.charge_branches.append((p, n))
        return ChargeIndex(self.n_charges)

    def alloc_limit(self, base_name: str, instance_name: str, p: int, n: int, init=0.0) -> LimitIndex:
        """``alloc_limit!`` (context.jl:826-850)."""
        name = base_name if instance_name == "" else f"{instance_name}_{base_name}"
        self.n_limits += 1
        self.limit_names.append(name)
        self.limit_branches.append((p, n))
        self.limit_init.append(init)
        return LimitIndex(self.n_limits)

    def resolve_index(self, idx: Index) -> int:
        """context.jl:577-581."""
        if isinstance(idx, CurrentIndex):
            return self.n_nodes + idx.k
        if isinstance(idx, ChargeIndex):
            return self.n_nodes + self.n_currents + idx.k
        if isinstance(idx, LimitIndex):
            return self.n_nodes + self.n_currents + self.n_charges + idx.k
        return int(idx)

    # -- stamping primitives (context.jl:945-999) --------------------------- #
    def stamp_G(self, i: Index, j: Index, val) -> None:
        if _iszero(i) or _iszero(j):
            return
        self.G_I.append(i); self.G_J.append(j); self.G_V.append(val)

    def stamp_C(self, i: Index, j: Index, val) -> None:
        if _iszero(i) or _iszero(j):
            return
        self.C_I.append(i); self.C_J.append(j); self.C_V.append(val)

    def stamp_b(self, i: Index, val) -> None:
        if _iszero(i):
            return
        self.b_I.append(i); self.b_V.append(val)

    def stamp_conductance(self, p: int, n: int, G) -> None:
        """context.jl:1362-1368 (call order matters: it is the COO order)."""
        self.stamp_G(p, p, G); self.stamp_G(p, n, _neg(G))
        self.stamp_G(n, p, _neg(G)); self.stamp_G(n, n, G)

    def stamp_capacitance(self, p: int, n: int, C) -> None:
        """context.jl:1379-1385."""
        self.stamp_C(p, p, C); self.stamp_C(p, n, _neg(C))
        self.stamp_C(n, p, _neg(C)); self.stamp_C(n, n, C)

    def reset_for_restamping(self) -> None:
        """context.jl:1528-1596."""
        self.__init__()

    # -- device-table recording -------------------------------------------- #
    def _begin(self) -> Tuple[int, int, int]:
        return len(self.G_I), len(self.C_I), len(self.b_I)

    def _record(self, kind, flags, name, user_nodes, nodes, params, base) -> None:
        self.devices.append(DeviceRow(kind, flags, str(name), [int(u) for u in user_nodes],
                                      list(nodes), list(params), base[0], base[1], base[2]))


# value marker for stamps whose value depends on the operating point: the host
# records structure only (values come from the CUDA device evaluation).
STATE_DEPENDENT = float("nan")


def _neg(v):
    return -v if not (isinstance(v, float) and math.isnan(v)) else v


# module-level aliases mirroring the reference's free functions
def get_node(ctx: MNAContext, name) -> int:
    return ctx.get_node(name)


def alloc_current(ctx: MNAContext, *parts) -> CurrentIndex:
    return ctx.alloc_current(*parts)


def alloc_internal_node(ctx: MNAContext, name, instance_name="") -> int:
    return ctx.alloc_internal_node(name, instance_name)


def alloc_limit(ctx: MNAContext, base_name, instance_name, p, n, init=0.0) -> LimitIndex:
    return ctx.alloc_limit(base_name, instance_name, p, n, init)


def resolve_index(ctx: MNAContext, idx) -> int:
    return ctx.resolve_index(idx)


def stamp_G(ctx, i, j, v):
    ctx.stamp_G(i, j, v)


def stamp_C(ctx, i, j, v):
    ctx.stamp_C(i, j, v)


def stamp_b(ctx, i, v):
    ctx.stamp_b(i, v)


def system_size(ctx: MNAContext) -> int:
    return ctx.system_size()


def reset_for_restamping(ctx: MNAContext) -> None:
    ctx.reset_for_restamping()


# --------------------------------------------------------------------------- #
# waveforms  (src/mna/devices.jl:122-216) -- data only; evaluated on the device
# --------------------------------------------------------------------------- #
class Wave:
    wave_kind = WAVE_NONE

    def wave_params(self) -> List[Any]:
        return []


class PWLWave(Wave):
    """``PWLWave(ts, ys)`` (devices.jl:130-135)."""
    wave_kind = WAVE_PWL

    def __init__(self, ts: Sequence, ys: Sequence):
        if len(ts) != len(ys) or len(ts) == 0:
            raise ValueError("PWLWave needs equally many times and values")
        self.ts, self.ys = list(ts), list(ys)

    def wave_params(self):
        out: List[Any] = []
        for t, y in zip(self.ts, self.ys):
            out += [t, y]
        return out


class PulseWave(Wave):
    """``PulseWave(v1, v2, td, tr, tf, pw, per)`` (devices.jl:189-203)."""
    wave_kind = WAVE_PULSE

    def __init__(self, v1, v2, td, tr, tf, pw, per):
        self.p = [v1, v2, td, tr, tf, pw, per]

    def wave_params(self):
        return list(self.p)


class SinWave(Wave):
    """``SinWave(vo, va, freq, td=0, theta=0, phase=0)`` (devices.jl:155-174)."""
    wave_kind = WAVE_SIN

    def __init__(self, vo, va, freq, td=0.0, theta=0.0, phase=0.0):
        self.p = [vo, va, freq, td, theta, phase]

    def wave_params(self):
        return list(self.p)


# --------------------------------------------------------------------------- #
# devices  (src/mna/devices.jl:264-460, :1304-1323, :1473-1492, :1637-1655)
# --------------------------------------------------------------------------- #
@dataclass
class Resistor:
    r: Any
    name: str = "R"


@dataclass
class Capacitor:
    c: Any
    name: str = "C"


@dataclass
class Inductor:
    l: Any
    name: str = "L"


class _Source:
    def __init__(self, dc, ac=0.0, tran: Optional[Wave] = None, name: str = "V"):
        if tran is not None and not isinstance(tran, Wave):
            raise TypeError("tran= must be a Wave struct (PWLWave/PulseWave/SinWave): arbitrary "
                            "closures cannot be lowered to the device (devices.jl:111-120)")
        self.dc, self.ac, self.tran, self.name = dc, ac, tran, name

    def _params(self):
        return [self.dc] + (self.tran.wave_params() if self.tran is not None else [])

    def _wave(self):
        return self.tran.wave_kind if self.tran is not None else WAVE_NONE


class VoltageSource(_Source):
    """devices.jl:329-339."""
    def __init__(self, v, ac=0.0, tran=None, name="V"):
        super().__init__(v, ac, tran, name)


class CurrentSource(_Source):
    """devices.jl:382-392."""
    def __init__(self, i, ac=0.0, tran=None, name="I"):
        super().__init__(i, ac, tran, name)


@dataclass
class VCVS:
    gain: Any
    name: str = "E"


@dataclass
class VCCS:
    gm: Any
    name: str = "G"


@dataclass
class CCVS:
    rm: Any
    name: str = "H"


@dataclass
class CCCS:
    gain: Any
    name: str = "F"


class Diode:
    """``Diode(; Is, Vt, n, limit, name)`` (devices.jl:1304-1323).

    ``vcrit = nVt*log(nVt/(sqrt(2)*Is))`` is precomputed by the constructor, as
    in the reference, and handed to the device evaluation as a parameter.
    """

    def __init__(self, Is=1e-14, Vt=0.026, n=1.0, KF=0.0, AF=1.0, FFE=1.0, limit=True, name="D"):
        self.Is, self.Vt, self.n, self.limit, self.name = Is, Vt, n, bool(limit), name
        nVt = np.asarray(n, dtype=np.float64) * np.asarray(Vt, dtype=np.float64)
        vcrit = nVt * np.log(nVt / (math.sqrt(2.0) * np.asarray(Is, dtype=np.float64)))
        self.vcrit = float(vcrit) if np.ndim(vcrit) == 0 else vcrit


@dataclass
class DiodeWithCap:
    """devices.jl:1473-1492."""
    Is: Any = 1e-14
    Vt: Any = 0.026
    n: Any = 1.0
    Cj0: Any = 1e-12
    Vj: Any = 0.7
    m: Any = 0.5
    name: str = "D"


@dataclass
class SimpleMOSFET:
    """devices.jl:1637-1655."""
    Vth: Any = 0.5
    K: Any = 1e-3
    lambda_: Any = 0.0
    Cgd: Any = 1e-15
    Cgs: Any = 1e-15
    name: str = "M"


# --------------------------------------------------------------------------- #
# stamp!  -- one method per device, COO order identical to the reference
# --------------------------------------------------------------------------- #
def stamp(dev, ctx: MNAContext, *ports, t=0.0, mode="tran", x=ZERO_VECTOR):
    """``stamp!(device, ctx, ports...)`` (devices.jl:475).

    ``ports`` may be node names or indices (devices.jl:981-996).  Time/mode/x
    keywords are accepted for signature compatibility; the host records
    structure only, so they do not influence what is recorded.
    """
    rest: Tuple = ()
    # trailing (t, mode) positional form of the source methods (devices.jl:643, :719)
    if isinstance(dev, _Source) and len(ports) == 4 and isinstance(ports[3], str) \
            and ports[3] in ("dcop", "tran", "tranop", "ac"):
        ports = ports[:2]
    from .behavioral import _Behavioral
    if isinstance(dev, _Behavioral):                   # devices.jl:1079-1131 (traced closure -> emitted module)
        inst, ctrl = dev.lower()
        before = ctx.n_currents
        stamp(inst, ctx, *ports, *ctrl)
        if dev.kind == "V" and ctx.n_currents == before + 1:
            ctx.current_names[-1] = "I_" + dev.name    # alloc_current!(ctx, :I_, B.name)
            return CurrentIndex(ctx.n_currents)
        return None
    pr: List[Index] = []
    for p in ports:
        if isinstance(p, (CurrentIndex, ChargeIndex, LimitIndex)):
            pr.append(p)
        else:
            pr.append(ctx.get_node(p))
    base = ctx._begin()

    if isinstance(dev, Resistor):                      # devices.jl:498-510
        p, n = pr
        ctx.stamp_conductance(p, n, _recip(dev.r))
        ctx._record(DEV_RESISTOR, 0, dev.name, [p, n], [p, n], [dev.r], base)
        return None
    if isinstance(dev, Capacitor):                     # devices.jl:531-534
        p, n = pr
        ctx.stamp_capacitance(p, n, dev.c)
        ctx._record(DEV_CAPACITOR, 0, dev.name, [p, n], [p, n], [dev.c], base)
        return None
    if isinstance(dev, Inductor):                      # devices.jl:569-586
        p, n = pr
        I = ctx.alloc_current("I_", dev.name)
        ctx.stamp_G(p, I, 1.0); ctx.stamp_G(n, I, -1.0)
        ctx.stamp_G(I, p, 1.0); ctx.stamp_G(I, n, -1.0)
        ctx.stamp_C(I, I, _neg(dev.l))
        ctx._record(DEV_INDUCTOR, 0, dev.name, [p, n], [p, n, I], [dev.l], base)
        return I
    if isinstance(dev, VoltageSource):                 # devices.jl:619-663
        p, n = pr
        I = ctx.alloc_current("I_", dev.name)
        if dev.tran is not None:
            ctx.breakpoints.append(dev.tran)           # register_breakpoints! :647
        ctx.stamp_G(p, I, 1.0); ctx.stamp_G(n, I, -1.0)
        ctx.stamp_G(I, p, 1.0); ctx.stamp_G(I, n, -1.0)
        ctx.stamp_b(I, dev.dc if dev.tran is None else STATE_DEPENDENT)
        ctx._record(DEV_VSOURCE, dev._wave(), dev.name, [p, n], [p, n, I], dev._params(), base)
        return I
    if isinstance(dev, CurrentSource):                 # devices.jl:698-737
        p, n = pr
        if dev.tran is not None:
            ctx.breakpoints.append(dev.tran)
        v = dev.dc if dev.tran is None else STATE_DEPENDENT
        ctx.stamp_b(p, v); ctx.stamp_b(n, _neg(v))
        ctx._record(DEV_ISOURCE, dev._wave(), dev.name, [p, n], [p, n], dev._params(), base)
        return None
    if isinstance(dev, VCVS):                          # devices.jl:760-775
        op, on, ip, in_ = pr
        I = ctx.alloc_current("I_", dev.name)
        ctx.stamp_G(op, I, 1.0); ctx.stamp_G(on, I, -1.0)
        ctx.stamp_G(I, op, 1.0); ctx.stamp_G(I, on, -1.0)
        ctx.stamp_G(I, ip, _neg(dev.gain)); ctx.stamp_G(I, in_, dev.gain)
        ctx._record(DEV_VCVS, 0, dev.name, [op, on, ip, in_], [op, on, ip, in_, I], [dev.gain], base)
        return I
    if isinstance(dev, VCCS):                          # devices.jl:797-808
        op, on, ip, in_ = pr
        gm = dev.gm
        ctx.stamp_G(op, ip, _neg(gm)); ctx.stamp_G(op, in_, gm)
        ctx.stamp_G(on, ip, gm); ctx.stamp_G(on, in_, _neg(gm))
        ctx._record(DEV_VCCS, 0, dev.name, [op, on, ip, in_], [op, on, ip, in_], [gm], base)
        return None
    if isinstance(dev, CCVS):
        if len(pr) == 4:                               # devices.jl:824-849
            op, on, ip, in_ = pr
            Iin = ctx.alloc_current("I_", dev.name, "_in")
            Iout = ctx.alloc_current("I_", dev.name, "_out")
            ctx.stamp_G(ip, Iin, 1.0); ctx.stamp_G(in_, Iin, -1.0)
            ctx.stamp_G(Iin, ip, 1.0); ctx.stamp_G(Iin, in_, -1.0)
            ctx.stamp_G(op, Iout, 1.0); ctx.stamp_G(on, Iout, -1.0)
            ctx.stamp_G(Iout, op, 1.0); ctx.stamp_G(Iout, on, -1.0)
            ctx.stamp_G(Iout, Iin, _neg(dev.rm))
            ctx._record(DEV_CCVS, 0, dev.name, [op, on, ip, in_], [op, on, ip, in_, Iin, Iout],
                        [dev.rm], base)
            return (Iout, Iin)
        op, on, Iin = pr                               # devices.jl:898-913, :941-956
        Iin = _as_current(ctx, Iin)
        Iout = ctx.alloc_current("I_", dev.name)
        ctx.stamp_G(op, Iout, 1.0); ctx.stamp_G(on, Iout, -1.0)
        ctx.stamp_G(Iout, op, 1.0); ctx.stamp_G(Iout, on, -1.0)
        ctx.stamp_G(Iout, Iin, _neg(dev.rm))
        ctx._record(DEV_CCVS, 1, dev.name, [op, on, Iin.k], [op, on, Iin, Iout], [dev.rm], base)
        return Iout
    if isinstance(dev, CCCS):
        if len(pr) == 4:                               # devices.jl:865-881
            op, on, ip, in_ = pr
            Iin = ctx.alloc_current("I_", dev.name, "_in")
            ctx.stamp_G(ip, Iin, 1.0); ctx.stamp_G(in_, Iin, -1.0)
            ctx.stamp_G(Iin, ip, 1.0); ctx.stamp_G(Iin, in_, -1.0)
            ctx.stamp_G(op, Iin, _neg(dev.gain)); ctx.stamp_G(on, Iin, dev.gain)
            ctx._record(DEV_CCCS, 0, dev.name, [op, on, ip, in_], [op, on, ip, in_, Iin],
                        [dev.gain], base)
            return Iin
        op, on, Iin = pr                               # devices.jl:924-931, :963-970
        Iin = _as_current(ctx, Iin)
        ctx.stamp_G(op, Iin, _neg(dev.gain)); ctx.stamp_G(on, Iin, dev.gain)
        ctx._record(DEV_CCCS, 1, dev.name, [op, on, Iin.k], [op, on, Iin], [dev.gain], base)
        return None
    if isinstance(dev, Diode):                         # devices.jl:1370-1428
        p, n = pr
        params = [dev.Is, dev.Vt, dev.n, dev.vcrit]
        if dev.limit:
            # limit! (devices.jl:1209-1234): alloc, tracking row, then companion at w
            l = ctx.alloc_limit("vdlim", dev.name, p, n, init=dev.vcrit)
            ctx.stamp_G(l, l, 1.0); ctx.stamp_G(l, p, -1.0); ctx.stamp_G(l, n, 1.0)
            ctx.stamp_conductance(p, n, STATE_DEPENDENT)
            ctx.stamp_b(p, STATE_DEPENDENT); ctx.stamp_b(n, STATE_DEPENDENT)
            ctx._record(DEV_DIODE, 1, dev.name, [p, n], [p, n, l], params, base)
        else:
            ctx.stamp_conductance(p, n, STATE_DEPENDENT)
            ctx.stamp_b(p, STATE_DEPENDENT); ctx.stamp_b(n, STATE_DEPENDENT)
            ctx._record(DEV_DIODE, 0, dev.name, [p, n], [p, n], params, base)
        return None
    if isinstance(dev, DiodeWithCap):                  # devices.jl:1558-1602
        p, n = pr
        ctx.stamp_conductance(p, n, STATE_DEPENDENT)
        ctx.stamp_b(p, STATE_DEPENDENT); ctx.stamp_b(n, STATE_DEPENDENT)
        ctx.stamp_capacitance(p, n, STATE_DEPENDENT)
        ctx._record(DEV_DIODECAP, 0, dev.name, [p, n], [p, n],
                    [dev.Is, dev.Vt, dev.n, dev.Cj0, dev.Vj, dev.m], base)
        return None
    if isinstance(dev, SimpleMOSFET):                  # devices.jl:1667-1761
        d, g, s = pr[:3]                               # 4-terminal form ignores the body
        S = STATE_DEPENDENT
        ctx.stamp_G(d, d, S); ctx.stamp_G(d, g, S); ctx.stamp_G(d, s, S)
        ctx.stamp_G(s, d, S); ctx.stamp_G(s, g, S); ctx.stamp_G(s, s, S)
        ctx.stamp_b(d, S); ctx.stamp_b(s, S)
        ctx.stamp_capacitance(g, s, dev.Cgs)
        ctx.stamp_capacitance(g, d, dev.Cgd)
        ctx._record(DEV_SIMPLEMOS, 0, dev.name, [d, g, s], [d, g, s],
                    [dev.Vth, dev.K, dev.lambda_, dev.Cgd, dev.Cgs], base)
        return None
    from .verilog_a import VAInstance
    if isinstance(dev, VAInstance):                    # generated VA stamp!, src/vasim.jl:2993-3985
        m = dev.model
        if len(pr) != len(m.ports):
            raise TypeError(f"{m.name} has {len(m.ports)} ports, got {len(pr)}")
        spec, xvec = current_spec(), current_x()
        collapsed = dev.collapsed(spec)                # V(int, ext) <+ 0 aliasing, vasim.jl:3532-3564
        loc: List[Index] = list(pr)
        for node in m.internal:                        # alloc_internal_node! in declaration order
            if node in collapsed:
                loc.append(pr[m.ports.index(collapsed[node])])
            else:
                loc.append(ctx.alloc_internal_node(f"{m.name}_{node}", dev.name))
        nodes_now = [int(x) for x in loc]
        var = m.variant(dev.given, dev.detect_vdep(spec, ctx, xvec, nodes_now), dev.vsites(spec, nodes_now))
        aliased = dev.aliased_sites(nodes_now)
        S = STATE_DEPENDENT
        for item in var.stamp_plan():
            if item[0] == "I":                         # alloc_current!(ctx, name, instance), vasim.jl:3256-3278, :2366
                _, slot, iname, site = item
                assert slot == len(loc)
                if site is not None and aliased[site]:     # `if p_node != n_node`: nothing allocated, nothing stamped
                    loc.append(0)
                else:
                    loc.append(ctx.alloc_current(f"{m.name}_{iname}", dev.name))
            elif item[0] == "L":                         # alloc_limit!(ctx, name, instance, p, n; init=0.0)
                _, slot, pi, ni, lname = item
                assert slot == len(loc)
                loc.append(ctx.alloc_limit(f"{m.name}_{lname}", dev.name, loc[pi],
                                           0 if ni is None else loc[ni], init=0.0))
            elif item[0] == "Q":                       # alloc_charge!(ctx, name, instance, p, n)
                _, slot, pi, ni, qname = item
                assert slot == len(loc)
                loc.append(ctx.alloc_charge(f"{dev.name}_{m.name}_{qname}", loc[pi], 0 if ni is None else loc[ni]))
            elif item[0] == "G":
                ctx.stamp_G(loc[item[1]], loc[item[2]], S)
            elif item[0] == "C":
                ctx.stamp_C(loc[item[1]], loc[item[2]], S)
            else:
                ctx.stamp_b(loc[item[1]], S)
        params = list(dev.params) + [dev.mfactor] + dev.extra_values(var, spec)
        ctx._record(DEV_VA, 0, dev.name, list(pr), loc, params, base)
        ctx.devices[-1].model = var
        return None
    raise TypeError(f"no stamp method for device of type {type(dev).__name__}")


_SPEC_STACK: List[Any] = []
_X_STACK: List[Any] = []


def current_x():
    """The operating point ``x`` of the builder call in progress (``_mna_x_``)."""
    return _X_STACK[-1] if _X_STACK else ZERO_VECTOR


def current_spec():
    """The MNASpec of the builder call in progress (the reference passes it to every
    generated stamp! as ``_mna_spec_``); defaults when stamping outside a builder."""
    if _SPEC_STACK:
        return _SPEC_STACK[-1]
    from .circuit import MNASpec
    return MNASpec()


def _recip(r):
    if isinstance(r, np.ndarray):
        return 1.0 / r
    return 1.0 / float(r)


def _as_current(ctx: MNAContext, i) -> CurrentIndex:
    if isinstance(i, CurrentIndex):
        return i
    k = int(i) - ctx.n_nodes        # plain Int form: an already-resolved system index
    if k < 1:
        raise ValueError("current-controlled source needs the index of a current variable")
    return CurrentIndex(k)
