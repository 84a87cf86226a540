"""Hand-written builders shared by the CPU and GPU tests.  They restate the builders
of the reference's own tests (cited per function) with the Python host API."""
import numpy as np

import cadnip_b200 as cb
from cadnip_b200 import (MNAContext, ZERO_VECTOR, get_node, stamp, Resistor, Capacitor, Inductor,
                         VoltageSource, CurrentSource, VCVS, VCCS, CCVS, CCCS, Diode, DiodeWithCap,
                         SimpleMOSFET, SinWave, PWLWave, PulseWave)


def _ctx(ctx):
    return MNAContext() if ctx is None else ctx


def rectifier(limit=True, V=5.0):
    """test/mna/pcnr.jl:39-66 (rectifier_lim / rectifier_nolim)."""
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = _ctx(ctx)
        vin = get_node(ctx, "vin"); out = get_node(ctx, "out")
        stamp(VoltageSource(V, name="V1"), ctx, vin, 0)
        stamp(Resistor(1000.0), ctx, vin, out)
        stamp(Diode(Is=1e-14, Vt=0.026, limit=limit, name="D1"), ctx, out, 0)
        return ctx
    return build


def chain(limit=True, V=50.0):
    """test/mna/pcnr.jl:69-104 (chain_lim / chain_nolim)."""
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = _ctx(ctx)
        vin = get_node(ctx, "vin"); n1 = get_node(ctx, "n1"); n2 = get_node(ctx, "n2"); n3 = get_node(ctx, "n3")
        stamp(VoltageSource(getattr(params, "vsrc", V) if "vsrc" in params else V, name="V1"), ctx, vin, 0)
        stamp(Resistor(1000.0), ctx, vin, n1)
        stamp(Diode(Is=1e-14, Vt=0.026, limit=limit, name="D1"), ctx, n1, n2)
        stamp(Diode(Is=1e-14, Vt=0.026, limit=limit, name="D2"), ctx, n2, n3)
        stamp(Diode(Is=1e-14, Vt=0.026, limit=limit, name="D3"), ctx, n3, 0)
        return ctx
    return build


def divider(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """test/sweep.jl build_two_resistor: V=1 across R1 + R2."""
    ctx = _ctx(ctx)
    vcc = get_node(ctx, "vcc"); out = get_node(ctx, "out")
    stamp(VoltageSource(1.0, name="V"), ctx, vcc, 0)
    stamp(Resistor(params.R1), ctx, vcc, out)
    stamp(Resistor(params.R2), ctx, out, 0)
    return ctx


def rc_charge(Vcc=5.0, R=1000.0, C=1e-6):
    """test/mna/core.jl:785-800."""
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = _ctx(ctx)
        vcc = get_node(ctx, "vcc"); out = get_node(ctx, "out")
        stamp(VoltageSource(Vcc, name="V1"), ctx, vcc, 0)
        stamp(Resistor(R), ctx, vcc, out)
        stamp(Capacitor(C), ctx, out, 0)
        return ctx
    return build


def rl_circuit(V=5.0, R=100.0, L=1e-3):
    """test/mna/core.jl:914-935: V - R - L to ground."""
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = _ctx(ctx)
        vcc = get_node(ctx, "vcc"); mid = get_node(ctx, "mid")
        stamp(VoltageSource(V, name="V1"), ctx, vcc, 0)
        stamp(Resistor(R), ctx, vcc, mid)
        stamp(Inductor(L, name="L1"), ctx, mid, 0)
        return ctx
    return build


def controlled_sources(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """One circuit exercising every linear primitive (E, G, H, F, L, I, both CC forms)."""
    ctx = _ctx(ctx)
    a = get_node(ctx, "a"); b = get_node(ctx, "b"); c = get_node(ctx, "c"); d = get_node(ctx, "d")
    e = get_node(ctx, "e"); f = get_node(ctx, "f"); g = get_node(ctx, "g")
    Iv = stamp(VoltageSource(2.0, tran=SinWave(2.0, 1.0, 1e6), name="V1"), ctx, a, 0)
    stamp(Resistor(1e3), ctx, a, b)
    stamp(Resistor(2e3), ctx, b, 0)
    stamp(VCVS(3.0, name="E1"), ctx, c, 0, b, 0)
    stamp(Resistor(500.0), ctx, c, d)
    stamp(VCCS(0.002, name="G1"), ctx, d, 0, b, 0)
    stamp(Resistor(1e3), ctx, d, 0)
    stamp(CCVS(100.0, name="H1"), ctx, e, 0, Iv)          # SPICE-style: senses I(V1)
    stamp(Resistor(1e3), ctx, e, 0)
    stamp(CCCS(2.0, name="F1"), ctx, f, 0, Iv)
    stamp(Resistor(1e3), ctx, f, 0)
    stamp(CurrentSource(1e-3, name="I1"), ctx, g, 0)
    stamp(Resistor(1e3), ctx, g, 0)
    hh = get_node(ctx, "h")
    stamp(Inductor(1e-3, name="L1"), ctx, g, hh)
    stamp(Resistor(2e3), ctx, hh, 0)
    stamp(Capacitor(1e-9), ctx, d, 0)
    k = get_node(ctx, "k"); m = get_node(ctx, "m"); q = get_node(ctx, "q")
    stamp(Resistor(1e3), ctx, a, k)                       # a -[1k]- k -(H2 sense)- m -[1k]- gnd
    stamp(Resistor(1e3), ctx, m, 0)
    stamp(CCVS(50.0, name="H2"), ctx, q, 0, k, m)         # 4-terminal form (own sensing branch)
    stamp(Resistor(1e3), ctx, q, 0)
    r = get_node(ctx, "r"); s2 = get_node(ctx, "s"); u = get_node(ctx, "u")
    stamp(Resistor(2e3), ctx, a, r)
    stamp(Resistor(1e3), ctx, s2, 0)
    stamp(CCCS(0.5, name="F2"), ctx, u, 0, r, s2)         # 4-terminal form
    stamp(Resistor(1e3), ctx, u, 0)
    return ctx


def mos_amp(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """Common-source stage with the native SimpleMOSFET and a DiodeWithCap clamp."""
    ctx = _ctx(ctx)
    vdd = get_node(ctx, "vdd"); g = get_node(ctx, "g"); d = get_node(ctx, "d")
    stamp(VoltageSource(3.3, name="VDD"), ctx, vdd, 0)
    stamp(VoltageSource(params.vg, tran=PulseWave(0.0, 2.0, 1e-9, 1e-9, 1e-9, 5e-9, 14e-9), name="VG"), ctx, g, 0)
    stamp(Resistor(params.rd), ctx, vdd, d)
    stamp(SimpleMOSFET(Vth=0.6, K=2e-3, lambda_=0.0, Cgd=2e-15, Cgs=5e-15, name="M1"), ctx, d, g, 0)
    stamp(DiodeWithCap(Is=1e-14, Vt=0.026, n=1.0, Cj0=2e-13, Vj=0.7, m=0.5, name="DC1"), ctx, 0, d)
    stamp(Capacitor(1e-13), ctx, d, 0)
    return ctx


def clipper(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    from cadnip_b200.workloads import clipper_builder
    return clipper_builder(params, spec, t, x, ctx)


def rectifier_v(limit=False):
    """The rectifier of test/mna/pcnr.jl:39-66 with the source voltage as a sweep parameter.  Without
    limiting, PCNR is skipped and plain Newton from zeros overshoots the exponential for V >= 4:
    the lanes fall through to gshunt / source stepping (solve.jl:871-929)."""
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = _ctx(ctx)
        vin = get_node(ctx, "vin"); out = get_node(ctx, "out")
        stamp(VoltageSource(params.vsrc, name="V1"), ctx, vin, 0)
        stamp(Resistor(1000.0), ctx, vin, out)
        stamp(Diode(Is=1e-14, Vt=0.026, limit=limit, name="D1"), ctx, out, 0)
        return ctx
    return build


def cancelling_vccs(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """a -[R]- b -[R]- c -[R]- gnd, R from b to gnd, and a VCCS b->gnd controlled by V(b): the (b, b)
    entry is 3/R - gm.  With R = 1024 (1/R exact in binary) and gm = 3/1024 the entry is EXACTLY zero,
    so a pivot order chosen on lanes with another gm meets a vanishing pivot on that lane; the matrix
    itself stays regular (b couples to a and c)."""
    ctx = _ctx(ctx)
    a = get_node(ctx, "a"); b = get_node(ctx, "b"); c = get_node(ctx, "c")
    stamp(VoltageSource(1.0, name="V1"), ctx, a, 0)
    stamp(Resistor(1024.0), ctx, a, b)
    stamp(Resistor(1024.0), ctx, b, 0)
    stamp(Resistor(1024.0), ctx, b, c)
    stamp(Resistor(1024.0), ctx, c, 0)
    stamp(VCCS(params.gm, name="G1"), ctx, b, 0, b, 0)
    stamp(Capacitor(1e-9), ctx, c, 0)
    return ctx
