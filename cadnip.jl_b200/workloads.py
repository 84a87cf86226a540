"""Builders and sweeps of the BASELINE.json configurations, written the
way the reference's hand-written test builders are (explicit node order, e.g.
test/mna/pcnr.jl:39-52)."""
from __future__ import annotations

import numpy as np

from .circuit import MNASpec
from .mna import (MNAContext, ZERO_VECTOR, Capacitor, Diode, Resistor, SimpleMOSFET, SinWave,
                  PWLWave, VoltageSource, get_node, stamp)
from .sweeps import CircuitSweep, ProductSweep, Sweep


# ---- C2: RC / diode clipper -------------------------------------------------
def clipper_builder(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """``V1 in 0 SIN(0 5 1k)``; ``R1 in out R``; ``D1 out 0`` native
    ``Diode(Is=1e-14, Vt=0.026, n=1, limit=true)``; ``C1 out 0 C``.
    Node order in, out; unknowns [in, out, I_V1, D1_vdlim] (n = 4, nnz = 8)."""
    ctx = MNAContext() if ctx is None else ctx
    vin = get_node(ctx, "in")
    out = get_node(ctx, "out")
    stamp(VoltageSource(0.0, tran=SinWave(0.0, 5.0, 1e3), name="V1"), ctx, vin, 0)
    stamp(Resistor(params.R, name="R1"), ctx, vin, out)
    stamp(Diode(Is=1e-14, Vt=0.026, n=1.0, limit=True, name="D1"), ctx, out, 0)
    stamp(Capacitor(params.C, name="C1"), ctx, out, 0)
    return ctx


def clipper_sweep(nR: int = 256, nC: int = 256) -> CircuitSweep:
    """P = nR x nC: R log-spaced [1e2, 1e4] (fastest) x C log-spaced [1e-10, 1e-8]."""
    R = np.logspace(2, 4, nR)
    C = np.logspace(-10, -8, nC)
    return CircuitSweep(clipper_builder, ProductSweep(R=R, C=C), spec=MNASpec())


CLIPPER_TSPAN = (0.0, 2e-3)
CLIPPER_DT = 1e-6


# ---- C1/C3 stand-in with the native square-law MOSFET ------------------------
def inverter_builder(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    """CMOS inverter of benchmarks/benchmark_common.jl:82-106 restated with the
    in-tree native ``SimpleMOSFET`` (devices.jl:1637): NMOS pull-down plus a PMOS
    modelled as the complementary square-law device is not expressible with
    SimpleMOSFET (n-type only), so the pull-up is a resistive load (lambda = 0: with
    lambda != 0 the reference's square-law model is discontinuous at Vds = Vgs - Vth); the `sp_mos1`
    Verilog-A tier replaces this builder once the VA emitter lands.
    Nodes vdd, in, out."""
    ctx = MNAContext() if ctx is None else ctx
    vdd = get_node(ctx, "vdd")
    vin = get_node(ctx, "in")
    out = get_node(ctx, "out")
    stamp(VoltageSource(params.Vdd, name="VDD"), ctx, vdd, 0)
    pwl = PWLWave([0.0, 100e-9, 110e-9, 200e-9, 210e-9, 300e-9, 310e-9, 400e-9],
                  [0.0, 0.0, params.Vdd, params.Vdd, 0.0, 0.0, params.Vdd, params.Vdd])
    stamp(VoltageSource(0.0, tran=pwl, name="VIN"), ctx, vin, 0)
    stamp(Resistor(params.Rload, name="RL"), ctx, vdd, out)
    stamp(SimpleMOSFET(Vth=0.7, K=params.K, lambda_=0.0, Cgd=1e-15, Cgs=2e-15, name="MN"),
          ctx, out, vin, 0)
    stamp(Capacitor(params.CL, name="CL"), ctx, out, 0)
    return ctx


def inverter_sweep(nW: int = 50, nV: int = 50, nC: int = 40) -> CircuitSweep:
    """P = nW x nV x nC (first axis fastest): K ~ W/L lin, Vdd lin [1.8, 5], C_L log [1, 100] fF."""
    K = 100e-6 * np.linspace(0.36, 3.6, nW) / 0.18
    Vdd = np.linspace(1.8, 5.0, nV)
    CL = np.logspace(-15, -13, nC)
    return CircuitSweep(inverter_builder, ProductSweep(K=K, Vdd=Vdd, CL=CL), spec=MNASpec(),
                        Rload=20e3)
