"""Builders and sweeps of the BASELINE.json configurations, written the
way the reference's hand-written test builders are (explicit node order, e.g.
test/mna/pcnr.jl:39-52)."""
from __future__ import annotations

import numpy as np

import os

from .circuit import MNASpec
from .mna import (MNAContext, ZERO_VECTOR, Capacitor, Diode, Resistor, SimpleMOSFET, SinWave,
                  PWLWave, PulseWave, VoltageSource, VCVS, get_node, stamp)
from .sweeps import CircuitSweep, ProductSweep, Sweep, TandemSweep


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


# ---- circuits on the reference's VADistiller models (sp_mos1, sp_diode) ----------------------
# The builders take the parsed models as arguments.  The .va sources live in the reference tree
# (models/VADistillerModels.jl/va/), which exists where the fixtures are generated but not on a
# GPU box: `load_workload` lowers from the .va source when the tree is present and otherwise
# loads the committed lowered circuit (tests/golden/va_<name>.json.gz: tables + emitted text,
# written by tests/golden/make_va_fixtures.py; tests/test_va_models.py asserts both agree).
VA_DIR = "/root/reference/models/VADistillerModels.jl/va/"


def _B(f):
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        f(ctx, params)
        return ctx
    return build


def mos1_corner(sp_mos1):
    """test/params.jl:377-385 corner_ckt: square-law MOS1 in saturation, W/L = 20, RD = 10k,
    VGS = 1.2; swept over the model card (vto, kp)."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); g = get_node(ctx, "g"); drain = get_node(ctx, "drain")
        stamp(VoltageSource(5.0, name="Vdd"), ctx, vdd, 0)
        stamp(VoltageSource(1.2, name="Vg"), ctx, g, 0)
        stamp(sp_mos1(type=1, vto=p.vt0, kp=p.kpn, w=20e-6, l=1e-6, name="M1"), ctx, drain, g, 0, 0)
        stamp(Resistor(10e3, name="Rd"), ctx, vdd, drain)
    return CircuitSweep(_B(f), ProductSweep(vt0=[0.6, 0.7, 0.9], kpn=[100e-6, 50e-6]))


def diode_chain(sp_diode):
    """test/sweep.jl:322-332 diode_chain: three series junctions behind 1k, vsrc 0.5:0.5:20."""
    def f(ctx, p):
        vin = get_node(ctx, "in"); n1 = get_node(ctx, "n1"); n2 = get_node(ctx, "n2"); n3 = get_node(ctx, "n3")
        stamp(VoltageSource(p.vsrc, name="v1"), ctx, vin, 0)
        stamp(Resistor(1e3, name="r1"), ctx, vin, n1)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d1"}), ctx, n1, n2)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d2"}), ctx, n2, n3)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d3"}), ctx, n3, 0)
    return CircuitSweep(_B(f), Sweep(vsrc=np.arange(0.5, 20.01, 0.5)))


def diode_rs_cap(sp_diode):
    """One sp_diode with series resistance (internal node kept), junction capacitance and
    transit time behind a resistor, driven by a pulse: exercises the charge-state rows."""
    def f(ctx, p):
        vin = get_node(ctx, "in"); a = get_node(ctx, "a")
        stamp(VoltageSource(0.0, tran=PWLWave([0.0, 1e-9, 2e-9, 6e-9, 7e-9], [-2.0, -2.0, 1.0, 1.0, -2.0]),
                            name="v1"), ctx, vin, 0)
        stamp(Resistor(p.r, name="r1"), ctx, vin, a)
        stamp(sp_diode(**{"is": 1e-14, "rs": 5.0, "cjo": 2e-12, "tt": 1e-10, "vj": 0.8, "m": 0.4,
                          "name": "d1"}), ctx, a, 0)
    return CircuitSweep(_B(f), Sweep(r=[200.0, 1e3]))


_NMOS = dict(type=1, vto=0.7, kp=100e-6)       # test/mna/oscillator_test.jl:43-44
_PMOS = dict(type=-1, vto=-0.7, kp=50e-6)


def mos1_inverter(sp_mos1):
    """CMOS inverter of benchmarks/benchmark_common.jl:82-106 with sp_mos1 cards:
    W_n x Vdd x C_L sweep, W_p = 1.375 W_n; the input ramp is scaled to Vdd by a VCVS."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); d = get_node(ctx, "d"); q = get_node(ctx, "q"); ramp = get_node(ctx, "ramp")
        stamp(VoltageSource(p.vdd, name="VDD"), ctx, vdd, 0)
        stamp(VoltageSource(0.0, tran=PWLWave([0.0, 1e-9, 2e-9, 6e-9, 7e-9, 10e-9], [0.0, 0.0, 1.0, 1.0, 0.0, 0.0]),
                            name="VR"), ctx, ramp, 0)
        stamp(VCVS(p.vdd, name="EIN"), ctx, d, 0, ramp, 0)
        stamp(sp_mos1(w=1.375 * p.wn, l=1e-6, name="MP", **_PMOS), ctx, q, d, vdd, vdd)
        stamp(sp_mos1(w=p.wn, l=1e-6, name="MN", **_NMOS), ctx, q, d, 0, 0)
        stamp(Capacitor(p.cl, name="CL"), ctx, q, 0)
    return CircuitSweep(_B(f), ProductSweep(wn=[0.36e-6, 3.6e-6], vdd=[1.8, 5.0], cl=[1e-15, 100e-15]))


C3_PWL_T = [0.0, 100e-9, 110e-9, 200e-9, 210e-9, 300e-9, 310e-9, 400e-9]
C3_PWL_Y = [0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 1.0]


def mos1_c3(sp_mos1):
    """SURVEY 8d config C3 (BASELINE configs[2]): the inverter deck of
    benchmarks/benchmark_common.jl:82-106 (Xneg W=3.6e-7 L=6e-7, Xpos W=4.95e-7 L=5e-7 = 1.375 W_n,
    VSS source, CQ D 0 1e-15, PWL 0/100n/110n/200n/210n/300n/310n/400n) with the PDK FETs
    replaced by sp_mos1 cards (vto=+-0.7, kp=100u/50u), a load C_L on Q, and the input ramp scaled
    to Vdd by a VCVS; swept over W_n x Vdd x C_L.  The fixture holds a 2x2x2 corner grid; bench.py
    regenerates the 50x50x40 = 100 000-lane columns from `lane_exprs`."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); vss = get_node(ctx, "vss"); d = get_node(ctx, "d"); q = get_node(ctx, "q")
        ramp = get_node(ctx, "ramp")
        stamp(sp_mos1(w=p.wn, l=6e-7, name="Xneg", **_NMOS), ctx, q, d, vss, vss)
        stamp(sp_mos1(w=1.375 * p.wn, l=5e-7, name="Xpos", **_PMOS), ctx, q, d, vdd, vdd)
        stamp(VoltageSource(p.vdd, name="VVDD"), ctx, vdd, 0)
        stamp(VoltageSource(0.0, name="VVSS"), ctx, vss, 0)
        stamp(Capacitor(1e-15, name="CQ"), ctx, d, 0)
        stamp(VoltageSource(0.0, tran=PWLWave(C3_PWL_T, C3_PWL_Y), name="VR"), ctx, ramp, 0)
        stamp(VCVS(p.vdd, name="VD"), ctx, d, 0, ramp, 0)
        stamp(Capacitor(p.cl, name="CL"), ctx, q, 0)
    return CircuitSweep(_B(f), ProductSweep(wn=[0.36e-6, 3.6e-6], vdd=[1.8, 5.0], cl=[1e-15, 100e-15]))


def c3_lane_exprs(lc, cs):
    """Which swept quantity each lane column of the lowered C3 circuit holds."""
    params, P = cs.lane_params()
    cand = {"wn": params.wn, "1.375*wn": 1.375 * params.wn, "vdd": params.vdd, "cl": params.cl}
    out = []
    for col in lc.lane_soa:
        hit = [k for k, v in cand.items() if np.array_equal(col, np.asarray(v, dtype=np.float64))]
        if len(hit) != 1:
            raise ValueError("unrecognised lane column in the C3 circuit")
        out.append(hit[0])
    return out


DFF_DIR = "/root/reference/test/DFF/"


def _spice_lines(path):
    """Logical lines of a SPICE deck: comments dropped, `+` continuations joined."""
    out = []
    for raw in open(path).read().split("\n"):
        ln = raw.strip()
        if not ln or ln.startswith("*"):
            continue
        if ln.startswith("+"):
            out[-1] += " " + ln[1:].strip()
        else:
            out.append(ln)
    return out


def mos1_dff(sp_mos1, n_lanes=4):
    """the gf180 D flip-flop of test/DFF/DFF_cap_all.cir +
    gf180mcu_fd_sc_mcu7t5v0__dffnq_4.ngspice (30 FETs, CQ = 1.7205e-13, CLKN / D PWL stimuli, the
    0 V probe sources VQ / VNW / VPW).  The PDK cards are not in the reference tree, so the FETs are
    sp_mos1 with a synthetic 5 V card (vto = +-0.7, kp = 100u / 50u: the FALLBACK tier SURVEY names) and
    every internal net carries a synthetic 5 fF parasitic to ground.  Lanes: process corners x
    Monte-Carlo draws on (vto_n, vto_p, kp_n, kp_p)."""
    fets, caps, srcs = [], [], []
    for ln in _spice_lines(DFF_DIR + "gf180mcu_fd_sc_mcu7t5v0__dffnq_4.ngspice"):
        f = ln.split()
        if f[0][0] in "xX":
            kw = dict(x.split("=") for x in f[6:])
            fets.append((f[0], f[1], f[2], f[3], f[4], f[5], float(kw["W"]), float(kw["L"])))
    for ln in _spice_lines(DFF_DIR + "DFF_cap_all.cir"):
        f = ln.replace("(", " ").replace(")", " ").split()
        if f[0][0] in "cC":
            caps.append((f[0], f[1], f[2], float(f[3])))
        elif f[0][0] in "vV":
            if len(f) > 3 and f[3].upper() == "PWL":
                vals = [float(x) for x in f[4:]]
                srcs.append((f[0], f[1], f[2], 0.0, (vals[0::2], vals[1::2])))
            else:
                srcs.append((f[0], f[1], f[2], float(f[3]), None))
    assert len(fets) == 30 and len(srcs) == 7
    nets = []
    for _, d, g, s_, b, *_ in fets:
        for n in (d, g, s_, b):
            if n not in nets:
                nets.append(n)
    for _, p, n, *_ in srcs + caps:
        for x in (p, n):
            if x != "0" and x not in nets:
                nets.append(x)
    driven = {p for _, p, n, *_ in srcs}

    def f(ctx, p):
        node = {"0": 0}
        for n in nets:
            node[n] = get_node(ctx, n)
        for name, a, b, dc, pwl in srcs:
            stamp(VoltageSource(dc if pwl is None else pwl[1][0], tran=None if pwl is None else PWLWave(*pwl),
                                name=name), ctx, node[a], node[b])
        for name, a, b, c in caps:
            stamp(Capacitor(c, name=name), ctx, node[a], node[b])
        for n in nets:
            if n not in driven and n not in ("VDD", "VSS"):
                stamp(Capacitor(5e-15, name="Cpar_" + n), ctx, node[n], 0)
        for name, d, g, s_, b, model, w, l in fets:
            if model.startswith("nfet"):
                card = dict(type=1, vto=p.vton, kp=p.kpn)
            else:
                card = dict(type=-1, vto=p.vtop, kp=p.kpp)
            stamp(sp_mos1(w=w, l=l, name=name, **card), ctx, node[d], node[g], node[s_], node[b])

    rng = np.random.default_rng(20261018)
    per = max(1, n_lanes // 4)
    corner = [(+1, +1), (+1, -1), (-1, +1), (-1, -1)]                   # (vto, kp) +-10 %
    vton, vtop, kpn, kpp = [], [], [], []
    for cv, ck in corner:
        dv = rng.normal(0.0, 15e-3, (per, 2))
        dk = rng.normal(0.0, 0.02, (per, 2))
        vton += list(0.7 * (1 + 0.1 * cv) + dv[:, 0]); vtop += list(-0.7 * (1 + 0.1 * cv) - dv[:, 1])
        kpn += list(100e-6 * (1 + 0.1 * ck) * (1 + dk[:, 0])); kpp += list(50e-6 * (1 + 0.1 * ck) * (1 + dk[:, 1]))
    return CircuitSweep(_B(f), TandemSweep(vton=vton, vtop=vtop, kpn=kpn, kpp=kpp))


def mos1_ring(sp_mos1, caps=False):
    """3-stage ring oscillator of test/mna/oscillator_test.jl:38-68; caps=True adds the
    device's own overlap / junction / Meyer capacitances (voltage-dependent charges)."""
    extra = dict(cgso=3e-10, cgdo=3e-10, cbd=2e-15, cbs=2e-15, tox=2e-8) if caps else {}

    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); in1 = get_node(ctx, "in1"); out1 = get_node(ctx, "out1"); out2 = get_node(ctx, "out2")
        stamp(VoltageSource(3.3, name="Vdd"), ctx, vdd, 0)
        for k, (i, o) in enumerate(((in1, out1), (out1, out2), (out2, in1)), 1):
            stamp(sp_mos1(w=2e-6, l=1e-6, name=f"MP{k}", **_PMOS, **extra), ctx, o, i, vdd, vdd)
            stamp(sp_mos1(w=1e-6, l=1e-6, name=f"MN{k}", **_NMOS, **extra), ctx, o, i, 0, 0)
        stamp(Capacitor(p.c, name="C1"), ctx, out1, 0)
        stamp(Capacitor(p.c, name="C2"), ctx, out2, 0)
        stamp(Capacitor(p.c, name="C3"), ctx, in1, 0)
    return CircuitSweep(_B(f), Sweep(c=[10e-15, 20e-15] if not caps else [10e-15]))


def bjt_ce(sp_bjt):
    """The sp_bjt circuit of test/mna/vadistiller_integration.jl:392-406 (Vcc 5 V, Vb through 10 k,
    Rc 1 k, bf = 100, is = 1e-15) with the base drive and bf swept, a collector load capacitor and a
    PULSE on the base source for the transient.  sp_bjt: 7 internal nodes, of which cx_int, b_int and
    e_int alias their port, c_int / sub_con are shorted to other INTERNAL nodes and xf1 / xf2 to ground by
    executed potential contributions (bjt.va:739-753, :1009-1013) -> branch-current unknowns."""
    def f(ctx, p):
        vcc = get_node(ctx, "vcc"); vb = get_node(ctx, "vb"); c = get_node(ctx, "collector"); b = get_node(ctx, "base")
        stamp(VoltageSource(5.0, name="V1"), ctx, vcc, 0)
        stamp(VoltageSource(p.vb, tran=PulseWave(p.vb, p.vb + 0.1, 2e-7, 1e-7, 1e-7, 6e-7, 2e-6), name="V2"), ctx, vb, 0)
        stamp(Resistor(10000.0, name="Rb"), ctx, vb, b)
        stamp(Resistor(1000.0, name="Rc"), ctx, vcc, c)
        stamp(Capacitor(1e-11, name="Cc"), ctx, c, 0)
        stamp(sp_bjt(bf=p.bf, name="Q1", **{"is": 1e-15}), ctx, c, b, 0, 0)
    return CircuitSweep(_B(f), ProductSweep(vb=[0.6, 0.7, 0.8], bf=[50.0, 100.0]))


def jfet2_cs(sp_jfet2):
    """test/mna/vadistiller_integration.jl:498-519: sp_jfet2 common-source stage, gate bias swept."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); d = get_node(ctx, "drain"); g = get_node(ctx, "gate")
        stamp(VoltageSource(10.0, name="Vdd"), ctx, vdd, 0)
        stamp(VoltageSource(p.vg, name="Vg"), ctx, g, 0)
        stamp(Resistor(1000.0, name="Rd"), ctx, vdd, d)
        stamp(sp_jfet2(name="J1"), ctx, d, g, 0)
    return CircuitSweep(_B(f), Sweep(vg=[-1.0, -0.5, 0.0]))


def vdmos_cs(sp_vdmos):
    """test/mna/vadistiller_integration.jl:695-711: sp_vdmos (vto 2, kp 0.5) with a 100 ohm drain load;
    named branch `tbr` (vdmos.va:105, :1431) carries a potential contribution on the thermal node."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); d = get_node(ctx, "drain"); g = get_node(ctx, "gate")
        stamp(VoltageSource(10.0, name="Vdd"), ctx, vdd, 0)
        stamp(VoltageSource(p.vg, name="Vg"), ctx, g, 0)
        stamp(Resistor(100.0, name="Rd"), ctx, vdd, d)
        stamp(sp_vdmos(vto=2.0, kp=0.5, name="M1"), ctx, d, g, 0, 0, 0)
    return CircuitSweep(_B(f), Sweep(vg=[1.0, 3.0, 5.0]))


def inductor_rl(sp_inductor):
    """test/mna/vadistiller_integration.jl:240-259: V1 5 V - 1 k - sp_inductor(1 mH) to ground; at DC
    the inductor is a short (V(mid) = 0)."""
    def f(ctx, p):
        vcc = get_node(ctx, "vcc"); mid = get_node(ctx, "mid")
        stamp(VoltageSource(5.0, name="V1"), ctx, vcc, 0)
        stamp(Resistor(p.r, name="R1"), ctx, vcc, mid)
        stamp(sp_inductor(inductance=1e-3, name="L1"), ctx, mid, 0)
    return CircuitSweep(_B(f), Sweep(r=[500.0, 1000.0]))


def bsim4_stage(sp_bsim4v8):
    """test/mna/vadistiller_integration.jl:758-775: sp_bsim4v8 (l = 100 nm, w = 1 um) with a 1 k drain
    load on 1 V, gate at 0.5 V: the reference expects 0.9 < V(drain) < 1.0.  Gate bias swept."""
    def f(ctx, p):
        vdd = get_node(ctx, "vdd"); d = get_node(ctx, "drain"); g = get_node(ctx, "gate")
        stamp(VoltageSource(1.0, name="Vdd"), ctx, vdd, 0)
        stamp(VoltageSource(p.vg, name="Vg"), ctx, g, 0)
        stamp(Resistor(1000.0, name="Rd"), ctx, vdd, d)
        stamp(sp_bsim4v8(l=100e-9, w=1e-6, name="M1"), ctx, d, g, 0, 0)
    return CircuitSweep(_B(f), Sweep(vg=[0.3, 0.5, 0.8, 1.0]))


C6288_DIR = "/root/reference/benchmarks/vacask/c6288/cedarsim/"


def c6288_netlist(path=C6288_DIR):
    """Flat device list of the ISCAS c6288 16x16 multiplier deck (SURVEY 8d config C5:
    benchmarks/vacask/c6288/cedarsim/{runme.sp, multiplier.inc}): (fets, drivers) with
    fets = [(name, d, g, s, b, type +1/-1, w, l)] -- the not / nor / and gate sub-circuits expanded,
    `.global vdd vss` honoured -- and drivers = [node] for the 32 `v01` pulse sources (0 -> 1.2 V,
    td = tr = 0.1 ns, 1 ohm series resistor)."""
    subckts, cur = {}, None
    top = []
    for fname in ("multiplier.inc", "runme.sp"):
        for ln in _spice_lines(path + fname)[(1 if fname == "runme.sp" else 0):]:
            f = ln.split()
            key = f[0].lower()
            if key == ".subckt":
                cur = dict(name=f[1].lower(), ports=[x for x in f[2:] if "=" not in x], body=[])
                continue
            if key == ".ends":
                subckts[cur["name"]] = cur
                cur = None
                continue
            if key.startswith("."):
                continue
            (cur["body"] if cur is not None else top).append(f)
    glob = {"vdd", "vss", "0"}
    fets, drivers = [], []

    def expand(inst, conn, prefix):
        sc = subckts[inst]
        env = dict(zip(sc["ports"], conn))
        for f in sc["body"]:
            name = prefix + f[0]
            args = [x for x in f[1:] if "=" not in x]
            kw = dict(x.split("=") for x in f[1:] if "=" in x)
            if f[0][0] in "xX":
                sub = args[-1].lower()
                nodes = [env.get(a, a if a.lower() in glob else prefix + a) for a in args[:-1]]
                if sub in ("nmos", "pmos"):
                    def num(v):
                        v = v.lower()
                        return float(v[:-1]) * 1e-6 if v.endswith("u") else float(v)
                    fets.append((name, nodes[0], nodes[1], nodes[2], nodes[3], 1 if sub == "nmos" else -1,
                                 num(kw.get("w", "1u")), num(kw.get("l", "0.2u"))))
                else:
                    expand(sub, nodes, name + ".")
    for f in top:
        if f[0][0] in "xX":
            args = [x for x in f[1:] if "=" not in x]
            sub = args[-1].lower()
            if sub == "v01":
                drivers.append(args[0])
            else:
                expand(sub, args[:-1], f[0] + ".")
    assert len(drivers) == 32, len(drivers)
    return fets, drivers


def mos1_c6288(sp_mos1, n_fets=None):
    """C5 on the FALLBACK tier: the c6288 multiplier with its 10 112 PSP103 FETs replaced by sp_mos1
    cards (vto = +-0.4 V for the 1.2 V supply, kp = 200u / 100u, 1 fF per net to ground standing in
    for the device capacitances PSP103 would bring) -- the emitter does not read PSP103 yet, so this
    exercises the LARGE-circuit path (sparse symbolic analysis, one lane on many threads) on the
    reference's own topology, not its device physics.  n_fets: keep only the first gates (tests)."""
    fets, drivers = c6288_netlist()
    if n_fets:
        fets = fets[:n_fets]
    nets = []
    seen = set()
    for _, d, g, s_, b, *_ in fets:
        for x in (d, g, s_, b):
            if x not in seen and x != "0":
                seen.add(x); nets.append(x)
    used_drivers = [x for x in drivers if x in seen]

    def f(ctx, p):
        node = {"0": 0}
        for x in nets:
            node[x] = get_node(ctx, x)
        stamp(VoltageSource(1.2, name="vdd"), ctx, node["vdd"], 0)
        stamp(VoltageSource(0.0, name="vss"), ctx, node["vss"], 0)
        for x in used_drivers:
            n_int = get_node(ctx, x + "_drv")
            stamp(VoltageSource(0.0, tran=PulseWave(0.0, 1.2, 1e-10, 1e-10, 1e-10, 1.0, 2.0), name="vdrv_" + x), ctx, n_int, 0)
            stamp(Resistor(1.0, name="rdrv_" + x), ctx, n_int, node[x])
        for x in nets:
            if x not in ("vdd", "vss"):
                stamp(Capacitor(1e-15, name="cpar_" + x), ctx, node[x], 0)
        for name, d, g, s_, b, typ, w, l in fets:
            card = dict(type=typ, vto=0.4 * typ, kp=200e-6 if typ > 0 else 100e-6)
            stamp(sp_mos1(w=w, l=l, name=name, **card), ctx, node[d], node[g], node[s_], node[b])
    return CircuitSweep(_B(f), Sweep(dummy=[0.0]), dummy=0.0)


PSP_DIR = "/root/reference/models/PSPModels.jl/va/"
RING_DIR = "/root/reference/benchmarks/vacask/ring/cedarsim/"


def spice_model_cards(path):
    """`.model name type key=value ...` cards of a SPICE include file -> {name: (type, {key: float})}."""
    cards = {}
    for ln in _spice_lines(path):
        f = ln.split()
        if f[0].lower() != ".model":
            continue
        kv = {}
        for tok in f[3:]:
            k, v = tok.split("=")
            kv[k.lower()] = float(v)
        cards[f[1].lower()] = (f[2].lower(), kv)
    return cards


def psp_ring(psp103, stages=9, lanes=None):
    """The 9-stage PSP103 ring oscillator of benchmarks/vacask/ring/cedarsim/{runme.sp, models.inc}: the
    one small circuit the reference publishes a per-Newton-iteration time for (1376 us, VACASK 27.8 us;
    doc/ring_oscillator_investigation.md:297-313).  Inverters of pmos w = 20u / nmos w = 10u, l = 1u with
    the deck's psp103p / psp103n cards (ad / as / pd / ps from the nmos / pmos wrappers: ld = ls = 0.5u),
    vdd 1.2 V, a 10 uA / 1 ns current pulse into node 1 to start it.  Lanes: supply voltage."""
    cards = spice_model_cards(RING_DIR + "models.inc")
    card_n, card_p = cards["psp103n"][1], cards["psp103p"][1]
    vdds = [1.2] if lanes is None else list(lanes)

    def fet(card, name, w, l):
        ld = 0.5e-6
        return psp103(name=name, w=w, l=l, ad=w * ld, **{"as": w * ld}, pd=2 * (w + ld), ps=2 * (w + ld), **card)

    def f(ctx, p):
        vdd = get_node(ctx, "vdd")
        nodes = [get_node(ctx, str(k + 1)) for k in range(stages)]
        from .mna import CurrentSource
        stamp(CurrentSource(0.0, tran=PulseWave(0.0, 10e-6, 1e-9, 1e-9, 1e-9, 1e-9, 1.0), name="i0"), ctx, 0, nodes[0])
        for k in range(stages):
            i, o = nodes[k], nodes[(k + 1) % stages]
            stamp(fet(card_p, f"xu{k + 1}.xmp", 20e-6, 1e-6), ctx, o, i, vdd, vdd)
            stamp(fet(card_n, f"xu{k + 1}.xmn", 10e-6, 1e-6), ctx, o, i, 0, 0)
        stamp(VoltageSource(p.vdd, name="vdd"), ctx, vdd, 0)
    return CircuitSweep(_B(f), Sweep(vdd=vdds))


def psp_c6288(psp103, n_fets=None):
    """SURVEY 8d config C5 as the reference runs it: the c6288 16x16 multiplier with its 10 112 PSP103VA
    FETs and the deck's model cards (benchmarks/vacask/c6288/cedarsim/{runme.sp, multiplier.inc,
    models.inc}; nmos w = 0.5u, pmos w = 1u, l = 0.2u, ad / as / pd / ps from ld = ls = 0.5u), vdd 1.2 V,
    32 pulse drivers 0 -> 1.2 V (td = tr = 0.1 ns) behind 1 ohm.  The reference reports 212 228 unknowns
    for it (90 850 nodes + 70 818 currents + 50 560 charges)."""
    fets, drivers = c6288_netlist()
    if n_fets:
        fets = fets[:n_fets]
    cards = spice_model_cards(C6288_DIR + "models.inc")
    card = {1: cards["psp103n"][1], -1: cards["psp103p"][1]}
    nets, seen = [], set()
    for _, d, g, s_, b, *_ in fets:
        for x in (d, g, s_, b):
            if x not in seen and x != "0":
                seen.add(x); nets.append(x)
    used_drivers = [x for x in drivers if x in seen]

    def f(ctx, p):
        node = {"0": 0}
        for x in nets:
            node[x] = get_node(ctx, x)
        stamp(VoltageSource(1.2, name="vdd"), ctx, node["vdd"], 0)
        stamp(VoltageSource(0.0, name="vss"), ctx, node["vss"], 0)
        for x in used_drivers:
            n_int = get_node(ctx, x + "_drv")
            stamp(VoltageSource(0.0, tran=PulseWave(0.0, 1.2, 1e-10, 1e-10, 1e-10, 1.0, 2.0), name="vdrv_" + x), ctx, n_int, 0)
            stamp(Resistor(1.0, name="rdrv_" + x), ctx, n_int, node[x])
        ld = 0.5e-6
        for name, d, g, s_, b, typ, w, l in fets:
            stamp(psp103(name=name, w=w, l=l, ad=w * ld, **{"as": w * ld}, pd=2 * (w + ld), ps=2 * (w + ld), **card[typ]),
                  ctx, node[d], node[g], node[s_], node[b])
    return CircuitSweep(_B(f), Sweep(dummy=[0.0]), dummy=0.0)


FIXTURES = {
    "mos1_corner": ("mos1", mos1_corner),
    "diode_chain": ("diode", diode_chain),
    "diode_rs_cap": ("diode", diode_rs_cap),
    "mos1_inverter": ("mos1", mos1_inverter),
    "mos1_c3": ("mos1", mos1_c3),
    "mos1_dff": ("mos1", mos1_dff),
    "mos1_ring": ("mos1", mos1_ring),
    "mos1_ring_caps": ("mos1", lambda m: mos1_ring(m, caps=True)),
    "bjt_ce": ("bjt", bjt_ce),
    "jfet2_cs": ("jfet2", jfet2_cs),
    "vdmos_cs": ("vdmos", vdmos_cs),
    "inductor_rl": ("inductor", inductor_rl),
    "bsim4_stage": ("bsim4v8", bsim4_stage),
    "mos1_c6288": ("mos1", mos1_c6288),
    "mos1_c6288_slice": ("mos1", lambda m: mos1_c6288(m, 420)),
    "psp_ring": ("psp103", lambda m: psp_ring(m, lanes=[1.1, 1.2])),
    "psp_c6288": ("psp103", lambda m: psp_c6288(m)),
}


def lower_fixture(name, models=None):
    """Lower one of the circuits above from the reference's .va sources."""
    from . import verilog_a
    model_file, make = FIXTURES[name]
    models = {} if models is None else models
    if model_file not in models:
        models[model_file] = verilog_a.load_va((PSP_DIR if model_file.startswith("psp") else VA_DIR) + model_file + ".va")
    cs = make(models[model_file])
    from .lowering import lower
    params, P = cs.lane_params()
    # 10 112 instances of two parameter sets: the voltage-dependent-charge verdict is taken per SET from the
    # first instance on, so that the circuit carries one emitted variant per set (verilog_a._BULK_AFTER)
    bulk, verilog_a._BULK_AFTER = verilog_a._BULK_AFTER, (0 if "c6288" in name else verilog_a._BULK_AFTER)
    try:
        lc = lower(cs.builder, params, MNASpec(mode="tran"), P=P)
    finally:
        verilog_a._BULK_AFTER = bulk
    if name == "mos1_c3":
        lc.lane_exprs = c3_lane_exprs(lc, cs)
    if name == "mos1_dff":
        cand = {k: np.asarray(getattr(params, k), dtype=np.float64) for k in ("vton", "vtop", "kpn", "kpp")}
        lc.lane_exprs = [next(k for k, v in cand.items() if np.array_equal(col, v)) for col in lc.lane_soa]
    return lc


FIXTURE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
# circuits whose kernel sets __graft_entry__.build() prebuilds (the GPU parity tests and bench.py)
GPU_VA_FIXTURES = ["mos1_corner", "diode_chain", "diode_rs_cap", "mos1_inverter", "mos1_c3", "mos1_dff",
                   "mos1_ring", "mos1_ring_caps", "bjt_ce", "jfet2_cs", "vdmos_cs", "inductor_rl", "bsim4_stage",
                   "mos1_c6288", "psp_ring"]


def fixture_path(name: str) -> str:
    return os.path.join(FIXTURE_DIR, f"va_{name}.json.gz")


def load_fixture(name: str):
    from .lowering import LoweredCircuit
    return LoweredCircuit.load(fixture_path(name))


def load_workload(name: str, from_source=None):
    """Lowered circuit ``name`` (a key of FIXTURES): va source -> emitter -> lower when the
    reference's .va files are present (from_source=None: auto), else the committed fixture."""
    model_file = FIXTURES[name][0]
    have = os.path.exists((PSP_DIR if model_file.startswith("psp") else VA_DIR) + model_file + ".va") and \
        (name != "mos1_dff" or os.path.isdir(DFF_DIR))
    if from_source is None:
        from_source = have
    if from_source:
        lc = lower_fixture(name)
        lc.va_models = []
        return lc
    return load_fixture(name)


def c3_lanes(lc, n_lanes=0, shape=None):
    """The C3 grid: W_n lin [0.36, 3.6] um x Vdd lin [1.8, 5.0] V x C_L log [1, 100] fF,
    first axis fastest (Iterators.product order, src/sweeps.jl:272); W_p = 1.375 W_n.
    Returns (lane SoA [col][P], P)."""
    if shape is None:
        if n_lanes:
            side = max(2, int(round(n_lanes ** (1.0 / 3.0))))
            shape = (side, side, side)
        else:
            shape = (50, 50, 40)
    wn = np.linspace(0.36e-6, 3.6e-6, shape[0])
    vdd = np.linspace(1.8, 5.0, shape[1])
    cl = np.logspace(-15, -13, shape[2])
    k, j, i = np.meshgrid(np.arange(shape[2]), np.arange(shape[1]), np.arange(shape[0]), indexing="ij")
    col = {"wn": wn[i.ravel()], "1.375*wn": 1.375 * wn[i.ravel()], "vdd": vdd[j.ravel()], "cl": cl[k.ravel()]}
    return np.ascontiguousarray(np.stack([col[e] for e in lc.lane_exprs])), int(i.size)


def c4_lanes(lc, n_lanes=0):
    """4 process corners (vto, kp +-10 %) x Monte-Carlo draws (dvto ~ N(0, 15 mV), dkp/kp ~ N(0, 2 %)),
    numpy.random.default_rng(20261018) -- the same construction as mos1_dff."""
    n_lanes = n_lanes or 16384
    rng = np.random.default_rng(20261018)
    per = max(1, n_lanes // 4)
    col = {"vton": [], "vtop": [], "kpn": [], "kpp": []}
    for cv, ck in ((+1, +1), (+1, -1), (-1, +1), (-1, -1)):
        dv = rng.normal(0.0, 15e-3, (per, 2))
        dk = rng.normal(0.0, 0.02, (per, 2))
        col["vton"] += list(0.7 * (1 + 0.1 * cv) + dv[:, 0]); col["vtop"] += list(-0.7 * (1 + 0.1 * cv) - dv[:, 1])
        col["kpn"] += list(100e-6 * (1 + 0.1 * ck) * (1 + dk[:, 0])); col["kpp"] += list(50e-6 * (1 + 0.1 * ck) * (1 + dk[:, 1]))
    return np.ascontiguousarray(np.stack([np.asarray(col[e]) for e in lc.lane_exprs])), 4 * per
