"""dc / tran entry points -- host mirror of ``dc!`` / ``tran!`` (src/sweeps.jl:450-532,
:588-707) on top of the C ABI.  A CircuitSweep is lowered once (one builder run with
lane-array parameters) and solved for all lanes by the CUDA kernels.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import backend
from .circuit import MNACircuit, MNASpec, Params, with_mode
from .lowering import LoweredCircuit, lower
from .mna import PWLWave, PulseWave, SinWave, Wave
from .sweeps import CircuitSweep, SweepResult

_RETCODES = {0: "Success", 1: "MaxIters", 2: "Unstable", 3: "Unstable", 4: "DtLessThanMin"}


class CedarTranOp:
    """``CedarTranOp`` (src/mna/dcop.jl:160-203): DC solve in ``:tranop`` mode from zeros with the
    fallback chain -- the default initialisation of ``tran``."""

    def __init__(self, abstol: float = 1e-9, maxiters: int = 500):
        self.abstol, self.maxiters = abstol, maxiters


class CedarUICOp:
    """``CedarUICOp`` (src/mna/dcop.jl:145-151, :311-411): no DC solve; ``warmup_steps`` fixed
    backward-Euler steps of ``dt`` relax the algebraic constraints from ``u0`` (zeros when none is
    given).  For oscillators and circuits without a stable DC operating point."""

    def __init__(self, warmup_steps: int = 10, dt: float = 1e-12, use_shampine: bool = False):
        self.warmup_steps, self.dt, self.use_shampine = int(warmup_steps), float(dt), bool(use_shampine)


def _init_fields(initializealg, u0, init_abstol, init_maxiters):
    """(init, init_abstol, init_maxiters, uic_steps, uic_dt) of cb200_tran_opts."""
    if isinstance(initializealg, CedarUICOp):
        return 2, init_abstol, init_maxiters, initializealg.warmup_steps, initializealg.dt
    if isinstance(initializealg, CedarTranOp):
        return 0, initializealg.abstol, initializealg.maxiters, 10, 1e-12
    if initializealg is not None:
        raise TypeError("initializealg must be CedarTranOp(...) or CedarUICOp(...)")
    return (0 if u0 is None else 1), init_abstol, init_maxiters, 10, 1e-12


def state_abstol(lc: LoweredCircuit, vntol: float = 1e-6, iabstol: float = 1e-12,
                 chgtol: float = 1e-14) -> np.ndarray:
    """``state_abstol(sys; vntol, iabstol, chgtol)`` (src/mna/build.jl:276-283): per-class absolute
    tolerance vector over ``[nodes | currents | charges | limits]``; limit unknowns are branch
    voltages and share ``vntol``."""
    tol = np.empty(lc.n)
    a, b, c = lc.n_nodes, lc.n_nodes + lc.n_currents, lc.n_nodes + lc.n_currents + lc.n_charges
    tol[:a], tol[a:b], tol[b:c], tol[c:] = vntol, iabstol, chgtol, vntol
    return tol


# --------------------------------------------------------------------------- #
# breakpoints (src/mna/breakpoints.jl, src/mna/solve.jl:1847-1918)
# --------------------------------------------------------------------------- #
def breakpoints(wave) -> Optional[Tuple[List[float], float, int]]:
    """``breakpoints(wave)`` -> (times, period, count) or None (devices.jl:145,180,211)."""
    if isinstance(wave, PWLWave):
        return ([float(t) for t in wave.ts], 0.0, -1)
    if isinstance(wave, SinWave):
        td = float(wave.p[3])
        return ([td], 0.0, -1) if td > 0 else None
    if isinstance(wave, PulseWave):
        td, tr, tf, pw, per = [float(x) for x in wave.p[2:]]      # the levels v1, v2 may be swept per lane
        edges = [td, td + tr, td + tr + pw, td + tr + pw + tf]
        return (edges, per, -1) if per > 0 else (edges, 0.0, -1)
    return None


def expand_breakpoints(specs: Sequence, tspan: Tuple[float, float], max_points: int = 100_000) -> List[float]:
    """``expand_breakpoints`` (solve.jl:1847-1918): strictly inside tspan, sorted,
    truncated to ``max_points``, dedup tolerance ``4*max(eps(a), eps(b))``."""
    t0, t1 = float(tspan[0]), float(tspan[1])
    out: List[float] = []
    for spec in specs:
        if spec is None:
            continue
        if isinstance(spec, Wave):
            spec = breakpoints(spec)
            if spec is None:
                continue
        times, period, count = spec
        if not times:
            continue
        if period <= 0:
            out += [t for t in times if t0 < t < t1]
        else:
            tmin, tmax = min(times), max(times)
            k_start = int(min(max(math.floor((t0 - tmax) / period), 0.0), 1e15))
            k_end = int(min(max(math.ceil((t1 - tmin) / period), -1.0), 1e15))
            if count >= 0:
                k_end = min(k_end, count - 1)
            if k_end < k_start:
                continue
            if k_end - k_start + 1 > max_points:
                k_end = k_start + max_points - 1
            for k in range(k_start, k_end + 1):
                base = k * period
                out += [t + base for t in times if t0 < t + base < t1]
    if not out:
        return out
    out.sort()
    out = out[:max_points]
    dedup = [out[0]]
    for x in out[1:]:
        tol = 4 * max(np.spacing(abs(dedup[-1])), np.spacing(abs(x)))
        if x - dedup[-1] > tol:
            dedup.append(x)
    return dedup


# --------------------------------------------------------------------------- #
# solutions
# --------------------------------------------------------------------------- #
class DCSolution:
    """``DCSolution`` (solve.jl:156-166): ``x``, name vectors, ``converged``;
    ``sol["out"]`` looks a node / current / charge / limit variable up by name."""

    def __init__(self, lc: LoweredCircuit, x: np.ndarray, converged: bool, iters: int = 0):
        self.x = x
        self.node_names = lc.node_names
        self.current_names = lc.current_names
        self.charge_names = lc.charge_names
        self.limit_names = lc.limit_names
        self.n_nodes = lc.n_nodes
        self.converged = bool(converged)
        self.iters = int(iters)
        self._lc = lc

    def __getitem__(self, name):
        return float(self.x[self._lc.index_of(name) - 1])

    def terminal_currents(self) -> dict:
        """Operating-point info: the current flowing INTO every terminal of every device that reports
        one, ``{"i_<device>_<terminal>": amperes}`` (``terminal_currents(ctx, x)``,
        src/mna/context.jl:1190-1270; registrations devices.jl:706-707, :733-734, :1424-1425, :1589-1590,
        :1737-1739, :1759, resistor: ohmic).  Evaluated on the host at the converged ``x`` from the
        lowered device table -- a handful of flops per device, not part of the batched hot path."""
        return terminal_currents(self._lc, self.x, getattr(self, "_lane", 0))

    def __repr__(self):
        return f"DCSolution(n={len(self.x)}, converged={self.converged})"


def terminal_currents(lc: LoweredCircuit, x: np.ndarray, lane: int = 0) -> dict:
    """``terminal_currents(ctx, x)`` (src/mna/context.jl:1251-1270) for one lane of a lowered circuit:
    entries in registration (device) order, repeated names summed.  Resistors report G (V_p - V_n)
    (``register_ohmic_terminal_current!``); current sources their DC value; the native diodes the
    junction current at the solution (the PCNR limit voltage equals V at convergence); SimpleMOSFET the
    square-law drain current.  (Verilog-A instances are not reported yet.)"""
    from . import mna as M
    x = np.asarray(x, dtype=np.float64)
    out: dict = {}

    def add(dev, term, val):
        key = f"i_{dev}_{term}"
        out[key] = out.get(key, 0.0) + float(val)

    def v(i):
        return 0.0 if i == 0 else float(x[i - 1])

    def par(d, k):
        r = int(lc.dev_params[lc.dev_param_ptr[d] + k])
        return float(lc.uniform[r]) if r >= 0 else float(lc.lane_soa[~r][lane])

    for d, kind in enumerate(lc.dev_kind):
        nb = int(lc.dev_node_ptr[d])
        nodes = [int(q) for q in lc.dev_nodes[nb:int(lc.dev_node_ptr[d + 1])]]
        name = lc.dev_names[d]
        if kind == M.DEV_RESISTOR:
            g = 1.0 / par(d, 0)
            add(name, "p", g * (v(nodes[0]) - v(nodes[1])))
            add(name, "n", g * (v(nodes[1]) - v(nodes[0])))
        elif kind == M.DEV_ISOURCE:
            i = par(d, 0)
            add(name, "p", -i); add(name, "n", i)
        elif kind in (M.DEV_DIODE, M.DEV_DIODECAP):
            Is, nVt = par(d, 0), par(d, 1) * par(d, 2)
            a = (v(nodes[0]) - v(nodes[1])) / nVt
            i0 = Is * (math.exp(80.0) * (1.0 + (a - 80.0)) - 1.0) if a > 80.0 else Is * (math.exp(a) - 1.0)
            add(name, "p", i0); add(name, "n", -i0)
        elif kind == M.DEV_SIMPLEMOS:
            vgs, vds = v(nodes[1]) - v(nodes[2]), v(nodes[0]) - v(nodes[2])
            vth, K, lam = par(d, 0), par(d, 1), par(d, 2)
            if vgs <= vth:
                ids = 0.0
            elif vds <= vgs - vth:
                ids = K * ((vgs - vth) * vds - vds * vds / 2)
            else:
                ids = K / 2 * (vgs - vth) ** 2 * (1 + lam * vds)
            add(name, "d", ids); add(name, "g", 0.0); add(name, "s", -ids)
    return out


class TranSolution:
    """One lane of a transient sweep: the subset of the SciML solution interface the
    reference's callers use (``sol.t``, ``sol.u``, ``sol[name]``, ``sol(t)``,
    ``sol.retcode``, ``sol.stats``)."""

    def __init__(self, lc: LoweredCircuit, save_idx: Sequence[int], t: np.ndarray, u: np.ndarray,
                 status: int, newton_iters: int):
        self._lc = lc
        self._save = list(save_idx)
        self.t = t
        self.u = u                      # [T][n_save] view
        self.status = int(status)
        self.retcode = _RETCODES.get(int(status), "Failure")
        self.stats = dict(nnonliniter=int(newton_iters), naccept=len(t) - 1)

    def __getitem__(self, name):
        idx = self._lc.index_of(name)
        try:
            col = self._save.index(idx)
        except ValueError:
            raise KeyError(f"{name!r} was not saved (save_idxs)") from None
        return self.u[:, col]

    def __call__(self, t: float) -> np.ndarray:
        """Linear interpolation between saved points."""
        return np.array([np.interp(t, self.t, self.u[:, q]) for q in range(self.u.shape[1])])


class _LazyTranSolutions:
    def __init__(self, lc, save_idx, t, u, count, status, iters, adaptive, stats=None, retried=0):
        self.lc, self.save_idx, self.t, self.u = lc, save_idx, t, u
        self.count, self.status, self.iters, self.adaptive = count, status, iters, adaptive
        self.stats = dict(stats or {})       # sweep-level counters (cb200_stats) of the solve
        self.stats["lanes_repivoted"] = int(retried)

    def __len__(self):
        return self.u.shape[2]

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        T = int(self.count[i])
        t = self.t[:T, i] if self.adaptive else self.t[:T]
        return TranSolution(self.lc, self.save_idx, t, self.u[:, :T, i].T, self.status[i], self.iters[i])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class _LazyDCSolutions:
    def __init__(self, lc, x, status, iters):
        self.lc, self.x, self.status, self.iters = lc, x, status, iters

    def __len__(self):
        return self.x.shape[1]

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        sol = DCSolution(self.lc, self.x[:, i].copy(), self.status[i] == 0, self.iters[i])
        sol._lane = i
        return sol

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


# --------------------------------------------------------------------------- #
# compiled sweeps
# --------------------------------------------------------------------------- #
class CompiledSweep:
    """A lowered circuit bound to one GPU: pattern, maps and device table built once,
    lane parameters resident in HBM.  ``lanes`` restricts the handle to a contiguous
    block of the sweep (multi-GPU sharding)."""

    def __init__(self, lc: LoweredCircuit, spec: MNASpec, device: int = 0,
                 lanes: Optional[slice] = None):
        self.lc = lc
        self.spec = spec
        self.device = device
        self.handle = backend.Handle(lc, device)
        soa = lc.lane_soa if lanes is None else np.ascontiguousarray(lc.lane_soa[:, lanes])
        self.P = int(soa.shape[1])
        self.lane_slice = lanes
        self._soa = soa
        self.handle.set_lanes(soa, self.P)

    def upload_lanes(self):
        """Re-send the lane parameters (the per-step H2D copy of the e2e measurement)."""
        self.handle.set_lanes(self._soa, self.P)

    def save_indices(self, save_idxs) -> List[int]:
        if save_idxs is None:
            return list(range(1, self.lc.n + 1))
        return [self.lc.index_of(s) for s in save_idxs]

    def dc(self, u0=None, abstol=1e-10, maxiters=100, use_stepping=True, mode="dcop"):
        return self.handle.dc(self.spec, u0, abstol, maxiters, use_stepping, mode)

    def specialize(self, dt, method="be", limit=False, fixed_only=False):
        """Compile and load kernels specialised for this circuit and step size."""
        self.handle.specialize(self.spec, method, dt, limit=limit, fixed_only=fixed_only)

    def tran(self, tspan, dt, method="be", save_idxs=None, save_every=1, abstol=1e-10,
             max_nl_iters=10, u0=None, init_abstol=1e-9, init_maxiters=500,
             specialize=False, limit=False, initializealg=None) -> backend.Wave:
        """limit=True applies the PCNR corrector inside the transient Newton loop too
        (CB200_TRAN_LIMIT): the models' $limit functions then damp the iteration.
        initializealg: None (CedarTranOp, or ``u0`` as given), CedarTranOp(...) or CedarUICOp(...)."""
        if specialize:
            self.specialize(dt, method, limit=limit, fixed_only=True)
        init, ia, im, us, ud = _init_fields(initializealg, u0, init_abstol, init_maxiters)
        opts = backend.make_tran_opts(method=method, adaptive=False, dt=dt, abstol=abstol,
                                      max_nl_iters=max_nl_iters, save_every=save_every,
                                      init=init, init_abstol=ia, init_maxiters=im, limit=limit,
                                      uic_steps=us, uic_dt=ud)
        return self.handle.tran(self.spec, tspan[0], tspan[1], opts, self.save_indices(save_idxs), u0)

    def tran_fetch(self, tspan, dt, out_u, method="be", save_idxs=None, save_every=1, abstol=1e-10,
                   max_nl_iters=10, n_segments=4, init_abstol=1e-9, init_maxiters=500, limit=False):
        """Fixed-step transient delivered into the (pinned) host array ``out_u``
        [save][T][P]; D2H of each time segment overlaps the next segment's compute."""
        opts = backend.make_tran_opts(method=method, adaptive=False, dt=dt, abstol=abstol,
                                      max_nl_iters=max_nl_iters, save_every=save_every,
                                      init_abstol=init_abstol, init_maxiters=init_maxiters, limit=limit)
        return self.handle.tran_fetch(self.spec, tspan[0], tspan[1], opts,
                                      self.save_indices(save_idxs), out_u, None, n_segments)

    def tran_adaptive(self, tspan, dt0=None, method="trap", save_idxs=None, abstol=1e-10,
                      reltol=1e-6, lte_abstol=1e-9, max_points=4096, dtmin=0.0, dtmax=0.0,
                      max_nl_iters=10, tstops=None, u0=None, specialize=False, limit=False,
                      initializealg=None, class_abstol=None) -> backend.Wave:
        """LTE-controlled stepping, one time axis per lane.  ``tstops`` defaults to the
        source breakpoints (``auto_tstops``, src/sweeps.jl:620-627).  ``class_abstol`` =
        (vntol, iabstol, chgtol): per-class absolute tolerances of the error test
        (``state_abstol``, src/mna/build.jl:276-283) instead of the one ``lte_abstol``."""
        if tstops is None:
            try:
                tstops = expand_breakpoints(self.lc.breakpoints, tspan)
            except TypeError as e:
                raise ValueError("source breakpoints depend on a swept parameter; pass tstops=") from e
        self.handle.set_tstops(tstops)
        dt0 = float(dt0) if dt0 else (tspan[1] - tspan[0]) * 1e-4
        if specialize:
            self.specialize(dt0, method, limit=limit)
        init, ia, im, us, ud = _init_fields(initializealg, u0, 1e-9, 500)
        opts = backend.make_tran_opts(method=method, adaptive=True, dt=dt0, abstol=abstol,
                                      reltol=reltol, lte_abstol=lte_abstol, dtmin=dtmin, dtmax=dtmax,
                                      max_nl_iters=max_nl_iters, max_points=max_points,
                                      init=init, init_abstol=ia, init_maxiters=im, limit=limit,
                                      uic_steps=us, uic_dt=ud, class_abstol=class_abstol)
        return self.handle.tran(self.spec, tspan[0], tspan[1], opts, self.save_indices(save_idxs), u0)

    def close(self):
        self.handle.close()


def compile_sweep(cs: Union[CircuitSweep, MNACircuit], spec: Optional[MNASpec] = None,
                  device: int = 0, lanes: Optional[slice] = None) -> CompiledSweep:
    if isinstance(cs, MNACircuit):
        sp = spec or cs.spec
        return CompiledSweep(lower(cs.builder, cs.params, sp, P=1), sp, device)
    params, P = cs.lane_params()
    sp = spec or cs.circuit.spec
    return CompiledSweep(lower(cs.builder, params, sp, P=P), sp, device, lanes)


# --------------------------------------------------------------------------- #
# dc! / tran!
# --------------------------------------------------------------------------- #
def dc(obj, u0=None, continuation: bool = True, abstol: float = 1e-10, maxiters: int = 100,
       device: int = 0):
    """``dc!(circuit; u0)`` (sweeps.jl:450-454) / ``dc!(cs::CircuitSweep; continuation)``
    (sweeps.jl:511-532).  ``dc!`` goes through ``with_mode(circuit, :dcop)``, which
    rebuilds the spec from temp+mode only (solve.jl:1976-1979) -- preserved.

    For a sweep all lanes are solved concurrently from a cold start; ``continuation``
    is accepted for API compatibility (the reference's tests assert it changes the
    path Newton takes, not where it lands: test/sweep.jl:340-345)."""
    if isinstance(obj, MNACircuit):
        c = with_mode(obj, "dcop")
        comp = compile_sweep(c, device=device)
        try:
            uu = None if u0 is None or len(u0) != comp.lc.n else np.asarray(u0, float).reshape(-1, 1)
            x, st, it = comp.dc(uu, abstol, maxiters)
        finally:
            comp.close()
        return DCSolution(comp.lc, x[:, 0].copy(), st[0] == 0, it[0])
    if isinstance(obj, CircuitSweep):
        spec = MNASpec(temp=obj.circuit.spec.temp, mode="dcop")
        comp = compile_sweep(obj, spec=spec, device=device)
        try:
            x, st, it = comp.dc(None, abstol, maxiters)
            r = dict(status=st, newton_iters=it, count=np.zeros_like(st), u=np.zeros((0, 0, len(st))),
                     t=np.zeros(0), x=x, weak=comp.handle.weak_pivot_lanes())

            def rerun(c):
                x2, st2, it2 = c.dc(None, abstol, maxiters)
                return dict(status=st2, newton_iters=it2, count=np.zeros_like(st2),
                            u=np.zeros((0, 0, len(st2))), t=np.zeros(0), x=x2)
            _retry_singular_lanes(comp, r, rerun)      # static-pivot safeguard: re-pivot on the failed lanes
        finally:
            comp.close()
        return SweepResult(obj.iterator.points(), _LazyDCSolutions(comp.lc, x, st, it))
    raise TypeError("dc expects an MNACircuit or CircuitSweep")


def _retry_singular_lanes(comp: "CompiledSweep", r: dict, run) -> int:
    """Static-pivot safeguard.  The pivot order is chosen once from magnitudes probed
    on 16 sample lanes and reused for every lane, step and gamma; KLU in the reference re-pivots
    when a refactor meets a vanishing pivot.  Lanes that end ``CB200_LANE_SINGULAR`` or were marked
    by the kernels' multiplier check (``cb200_weak_pivot_lanes``: |l| > 1e8) are therefore solved again on a handle that holds ONLY those lanes, so the probe -- and the pivot order --
    comes from their own parameter values; the results are merged back.  Returns the number of
    lanes retried."""
    bad = np.flatnonzero((r["status"] == backend.LANE_SINGULAR) | r.get("weak", False))
    if bad.size == 0 or bad.size == comp.P:
        return 0
    lc = comp.lc
    base = comp.lane_slice.start if comp.lane_slice is not None else 0
    sub = LoweredCircuit(**{f: getattr(lc, f) for f in lc.__dataclass_fields__})
    sub.lane_soa = np.ascontiguousarray(lc.lane_soa[:, base + bad])
    sub.P = int(bad.size)
    c2 = CompiledSweep(sub, comp.spec, comp.device)
    try:
        r2 = run(c2)
    finally:
        c2.close()
    for key in ("status", "newton_iters", "count"):
        r[key][bad] = r2[key]
    r["u"][:, :, bad] = r2["u"]
    if r["t"].ndim == 2:
        r["t"][:, bad] = r2["t"]
    if "x" in r:
        r["x"][:, bad] = r2["x"]
    return int(bad.size)


def tran(obj, tspan: Tuple[float, float], solver: Optional[str] = None, abstol: float = 1e-10,
         reltol: float = 1e-8, dt: Optional[float] = None, adaptive: Optional[bool] = None,
         saveat: Optional[float] = None, save_idxs=None, max_nl_iters: int = 10, device: int = 0,
         max_points: int = 4096, initializealg=None, u0=None):
    """``tran!(circuit, tspan; solver, abstol, reltol, kw...)`` (sweeps.jl:588-601) and
    ``tran!(cs::CircuitSweep, tspan; kw...)`` (sweeps.jl:692-707).

    ``solver``: "ImplicitEuler" | "Trapezoid" | "gear2", or -- adaptive mode only -- "IDA" (also "FBDF",
    "QNDF", "bdf"): the variable-order (1..5) variable-step BDF controller, the family of the reference's
    default ``Sundials.IDA`` (DESIGN.md s. 5; the default here stays the trapezoidal controller).  Fixed-step
    mode is the reference's ``solver=ImplicitEuler()/Trapezoid(), adaptive=false, dt=h``.
    ``initializealg``: ``CedarTranOp()`` (default) or ``CedarUICOp(warmup_steps, dt)``
    (src/mna/dcop.jl); ``u0`` ([n] or [n][P]) is the start state of the latter.
    ``abstol``: a number, or -- adaptive mode -- a mapping with any of ``vntol`` / ``iabstol`` /
    ``chgtol`` (defaults 1e-6 / 1e-12 / 1e-14): the reference's NamedTuple form, resolved per class
    of unknown by ``state_abstol`` (sweeps.jl:556, 615-618; build.jl:276-283)."""
    class_abstol = None
    if isinstance(abstol, dict):
        unknown = set(abstol) - {"vntol", "iabstol", "chgtol"}
        if unknown:
            raise ValueError(f"abstol: unknown tolerance class {sorted(unknown)}")
        class_abstol = (float(abstol.get("vntol", 1e-6)), float(abstol.get("iabstol", 1e-12)),
                        float(abstol.get("chgtol", 1e-14)))
        abstol = 1e-10                                    # Newton residual test keeps the scalar default
    method = solver or "Trapezoid"
    single = isinstance(obj, MNACircuit)
    if not single and not isinstance(obj, CircuitSweep):
        raise TypeError("tran expects an MNACircuit or CircuitSweep")
    is_adaptive = bool(adaptive or (adaptive is None and dt is None))
    save_every = 1
    if saveat is not None and not is_adaptive:
        save_every = max(1, int(round(saveat / dt)))
    comp = compile_sweep(obj, device=device)
    save = comp.save_indices(save_idxs)

    def run(c: CompiledSweep) -> dict:
        uu = None
        if u0 is not None:
            uu = np.asarray(u0, dtype=np.float64)
            if uu.ndim == 1:
                uu = np.repeat(uu[:, None], c.P, axis=1)
            elif uu.shape[1] != c.P:                      # a retry handle holds a subset of the lanes
                uu = uu[:, :c.P]
            uu = np.ascontiguousarray(uu)
        if is_adaptive:
            # the reference's default is a variable-step integrator (IDA, sweeps.jl:599-601)
            wave = c.tran_adaptive(tspan, dt0=dt, method=method, save_idxs=save, abstol=abstol,
                                   reltol=reltol, lte_abstol=max(abstol, 1e-12), u0=uu,
                                   max_nl_iters=max_nl_iters, max_points=max_points,
                                   initializealg=initializealg, class_abstol=class_abstol)
        else:
            wave = c.tran(tspan, dt, method=method, save_idxs=save, save_every=save_every,
                          abstol=abstol, max_nl_iters=max_nl_iters, u0=uu, initializealg=initializealg)
        out = wave.fetch()
        out["x"] = wave.final_state()
        out["stats"] = c.handle.stats()
        out["weak"] = c.handle.weak_pivot_lanes()
        wave.free()
        return out

    try:
        r = run(comp)
        retried = _retry_singular_lanes(comp, r, run) if u0 is None or np.ndim(u0) == 1 else 0
    finally:
        comp.close()
    sols = _LazyTranSolutions(comp.lc, save, r["t"], r["u"], r["count"], r["status"],
                              r["newton_iters"], is_adaptive, stats=r["stats"], retried=retried)
    return sols[0] if single else SweepResult(obj.iterator.points(), sols)
