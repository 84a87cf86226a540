"""Circuits built from the reference's VADistiller models (sp_mos1, sp_diode), shared by the
fixture generator (tests/golden/make_va_fixtures.py) and the CPU tests.  The builders take
the parsed models as arguments: the GPU box has no /root/reference, so the GPU tests load the
lowered circuits from the committed fixtures instead of parsing the .va files."""
import numpy as np

import cadnip_b200 as cb
from cadnip_b200 import (MNAContext, ZERO_VECTOR, get_node, stamp, VoltageSource, Resistor, Capacitor,
                         VCVS, PWLWave)

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
    return cb.CircuitSweep(_B(f), cb.ProductSweep(vt0=[0.6, 0.7, 0.9], kpn=[100e-6, 50e-6]))


def diode_chain(sp_diode):
    """test/sweep.jl:322-332 diode_chain: three series junctions behind 1k, vsrc 0.5:0.5:20."""
    def f(ctx, p):
        vin = get_node(ctx, "in"); n1 = get_node(ctx, "n1"); n2 = get_node(ctx, "n2"); n3 = get_node(ctx, "n3")
        stamp(VoltageSource(p.vsrc, name="v1"), ctx, vin, 0)
        stamp(Resistor(1e3, name="r1"), ctx, vin, n1)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d1"}), ctx, n1, n2)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d2"}), ctx, n2, n3)
        stamp(sp_diode(**{"is": 1e-14, "n": 1.0, "name": "d3"}), ctx, n3, 0)
    return cb.CircuitSweep(_B(f), cb.Sweep(vsrc=np.arange(0.5, 20.01, 0.5)))


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
    return cb.CircuitSweep(_B(f), cb.Sweep(r=[200.0, 1e3]))


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
    return cb.CircuitSweep(_B(f), cb.ProductSweep(wn=[0.36e-6, 3.6e-6], vdd=[1.8, 5.0], cl=[1e-15, 100e-15]))


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
    return cb.CircuitSweep(_B(f), cb.ProductSweep(wn=[0.36e-6, 3.6e-6], vdd=[1.8, 5.0], cl=[1e-15, 100e-15]))


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
    return cb.CircuitSweep(_B(f), cb.TandemSweep(vton=vton, vtop=vtop, kpn=kpn, kpp=kpp))


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
    return cb.CircuitSweep(_B(f), cb.Sweep(c=[10e-15, 20e-15] if not caps else [10e-15]))


FIXTURES = {
    "mos1_corner": ("mos1", mos1_corner),
    "diode_chain": ("diode", diode_chain),
    "diode_rs_cap": ("diode", diode_rs_cap),
    "mos1_inverter": ("mos1", mos1_inverter),
    "mos1_c3": ("mos1", mos1_c3),
    "mos1_dff": ("mos1", mos1_dff),
    "mos1_ring": ("mos1", mos1_ring),
    "mos1_ring_caps": ("mos1", lambda m: mos1_ring(m, caps=True)),
}


def lower_fixture(name, models=None):
    """Lower one of the circuits above from the reference's .va sources."""
    from cadnip_b200 import verilog_a
    model_file, make = FIXTURES[name]
    models = {} if models is None else models
    if model_file not in models:
        models[model_file] = verilog_a.load_va(VA_DIR + model_file + ".va")
    cs = make(models[model_file])
    params, P = cs.lane_params()
    lc = cb.lower(cs.builder, params, cb.MNASpec(mode="tran"), P=P)
    if name == "mos1_c3":
        lc.lane_exprs = c3_lane_exprs(lc, cs)
    if name == "mos1_dff":
        cand = {k: np.asarray(getattr(params, k), dtype=np.float64) for k in ("vton", "vtop", "kpn", "kpp")}
        lc.lane_exprs = [next(k for k, v in cand.items() if np.array_equal(col, v)) for col in lc.lane_soa]
    return lc
