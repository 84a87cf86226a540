"""Verilog-A emitter (tier 0): the reference's inline VA test models, checked through the
oracle against the reference's known answers (CPU) and GPU-vs-oracle parity (-m gpu).
Model sources are the reference's test fixtures (test/mna/vadistiller.jl, va_mosfet.jl)."""
import math

import numpy as np
import pytest

import cadnip_b200 as cb
import cadnip_oracle as ora
from cadnip_b200 import MNAContext, ZERO_VECTOR, get_node, stamp, VoltageSource, Resistor

VADDiode = cb.va("""
module VADDiode(a, c);
    parameter real Is = 1e-14;
    parameter real N = 1.0;
    inout a, c;
    electrical a, c;
    analog begin
        I(a,c) <+ Is*(exp(V(a,c)/(N*0.02585)) - 1.0);
    end
endmodule""")                                   # test/mna/vadistiller.jl:141-151

VADDiodeRs = cb.va("""
module VADDiodeRs(a, c);
    parameter real Is = 1e-14;
    parameter real N = 1.0;
    parameter real Rs = 10.0;
    inout a, c;
    electrical a, c, a_int;
    analog begin
        I(a, a_int) <+ V(a, a_int) / Rs;
        I(a_int, c) <+ Is*(exp(V(a_int,c)/(N*0.02585)) - 1.0);
    end
endmodule""")                                   # test/mna/vadistiller.jl:184-196

ChainDiodeRs = cb.va("""
module ChainDiodeRs(a, c);
    parameter real Is = 76.9e-12;
    parameter real N = 1.45;
    parameter real Rs = 0.042;
    inout a, c;
    electrical a, c, a_int;
    analog begin
        I(a, a_int) <+ V(a, a_int) / Rs;
        I(a_int, c) <+ Is*(limexp(V(a_int,c)/(N*0.02585)) - 1.0);
    end
endmodule""")                                   # test/mna/vadistiller.jl:245-257

_MOS_BODY = """
    real Vgs, Vds, Vov, Ids;
    analog begin
        Vgs = V(g, s);
        Vds = V(d, s);
        Vov = Vgs - Vth;
        if (Vov <= 0) begin
            Ids = 0;
        end else if (Vds < Vov) begin
            Ids = K * (Vov * Vds - Vds * Vds / 2);
        end else begin
            Ids = K / 2 * Vov * Vov;
        end
        I(d, s) <+ Ids;
"""
SimpleMOS = cb.va("module SimpleMOS(d, g, s);\n parameter real K = 1e-3;\n parameter real Vth = 0.5;\n"
                  " inout d, g, s;\n electrical d, g, s;\n" + _MOS_BODY + " end\nendmodule")   # va_mosfet.jl:71-95
CapMOS = cb.va("module CapMOS(d, g, s);\n parameter real K = 1e-3;\n parameter real Vth = 0.5;\n"
               " parameter real Cgs = 10e-15;\n parameter real Cgd = 5e-15;\n inout d, g, s;\n electrical d, g, s;\n"
               + _MOS_BODY + "  I(g, s) <+ Cgs * ddt(V(g, s));\n  I(g, d) <+ Cgd * ddt(V(g, d));\n end\nendmodule")
PMOS = cb.va("""
module PMOS(d, g, s);
    parameter real K = 1e-3;
    parameter real Vth = 0.5;
    inout d, g, s;
    electrical d, g, s;
    real Vsg, Vsd, Vov, Ids;
    analog begin
        Vsg = V(s, g);
        Vsd = V(s, d);
        Vov = Vsg - Vth;
        if (Vov <= 0) begin
            Ids = 0;
        end else if (Vsd < Vov) begin
            Ids = K * (Vov * Vsd - Vsd * Vsd / 2);
        end else begin
            Ids = K / 2 * Vov * Vov;
        end
        I(s, d) <+ Ids;
    end
endmodule""")                                   # va_mosfet.jl:357-381
JunctionCap = cb.va("""
module JunctionCap(p, n);
    parameter real Cj0 = 1e-12;
    parameter real phi = 0.8;
    parameter real m = 0.5;
    inout p, n;
    electrical p, n;
    real V, C;
    analog begin
        V = V(p, n);
        if (V < 0.5 * phi) begin
            C = Cj0 / $pow(1 - V/phi, m);
        end else begin
            C = Cj0 / $pow(0.5, m);
        end
        I(p, n) <+ C * ddt(V(p, n));
    end
endmodule""")                                   # va_mosfet.jl:465-484


# ---- potential contributions, named branches, current probes (vasim.jl:2311-2395, :3225-3279,
# :3630-3745).  RShort: a resistor chain whose middle section is either a resistor or -- rm == 0 -- a
# zero-volt potential contribution between two INTERNAL nodes, which is not the alias idiom
# (detect_short_circuits only aliases internal <-> port) and therefore carries a branch current.
RShort = cb.va("""
module RShort(a, c);
    parameter real r1 = 1e3;
    parameter real rm = 0.0;
    parameter real r2 = 1e3;
    inout a, c;
    electrical a, c, m1, m2;
    analog begin
        if (rm == 0) begin
            V(m1, m2) <+ 0;
        end else begin
            I(m1, m2) <+ V(m1, m2) / rm;
        end
        I(a, m1) <+ V(a, m1) / r1;
        I(m2, c) <+ V(m2, c) / r2;
    end
endmodule""")
# VRef: a bias-dependent potential contribution to ground, V(o) = gain * V(i) + vofs
VRef = cb.va("""
module VRef(i, o);
    parameter real gain = 2.0;
    parameter real vofs = 0.25;
    inout i, o;
    electrical i, o;
    analog begin
        V(o) <+ gain * V(i) * V(i) + vofs;
    end
endmodule""")
# LBranch: potential contribution on a named branch with a current probe and ddt (the inductor idiom
# of models/VADistillerModels.jl/va/inductor.va:61, :266-272) plus a series resistance
LBranch = cb.va("""
module LBranch(p, n);
    parameter real L = 1e-3;
    parameter real rs = 5.0;
    inout p, n;
    electrical p, n;
    branch (p, n) br;
    real flux;
    analog begin
        flux = L * I(br);
        V(br) <+ rs * I(br) + ddt(flux);
    end
endmodule""")


def B(f):
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        f(ctx, params)
        return ctx
    return build


def mos_bias(model, vg, vd, **kw):
    def f(ctx, params):
        d = get_node(ctx, "d"); g = get_node(ctx, "g")
        stamp(VoltageSource(vg, name="Vg"), ctx, g, 0)
        stamp(VoltageSource(vd, name="Vd"), ctx, d, 0)
        stamp(model(K=1e-3, Vth=0.5, **kw), ctx, d, g, 0)
    return B(f)


def inverter(vin):
    def f(ctx, params):
        vdd = get_node(ctx, "vdd"); out = get_node(ctx, "out"); inp = get_node(ctx, "inp")
        stamp(VoltageSource(3.0, name="Vdd"), ctx, vdd, 0)
        stamp(VoltageSource(params.vin if "vin" in params else vin, name="Vin"), ctx, inp, 0)
        stamp(PMOS(K=1e-3, Vth=0.5, name="MP"), ctx, out, inp, vdd)
        stamp(SimpleMOS(K=params.kn if "kn" in params else 1e-3, Vth=0.5, name="MN"), ctx, out, inp, 0)
        stamp(Resistor(1e6, name="Rload"), ctx, out, 0)
    return B(f)


def chain(ctx, params):
    n1, na, nb, nc = [get_node(ctx, s) for s in ("n1", "na", "nb", "nc")]
    stamp(VoltageSource(50.0, name="V1"), ctx, n1, 0)
    stamp(ChainDiodeRs(name="d1"), ctx, n1, na)
    stamp(ChainDiodeRs(name="d2"), ctx, na, nb)
    stamp(ChainDiodeRs(name="d3"), ctx, nb, nc)


def oracle_dc(builder, mode="dcop", **params):
    lc = cb.lower_circuit(cb.MNACircuit(builder, **params))
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    x, ok, it = ora.solve_dc(nl, ora.make_spec(mode=mode))
    return lc, nl, x, ok


# ---- emitter structure ---------------------------------------------------------
def test_parser_and_structure():
    assert SimpleMOS.ports == ["d", "g", "s"] and SimpleMOS.internal == [] and SimpleMOS.branches == [("d", "s")]
    assert ChainDiodeRs.internal == ["a_int"] and ChainDiodeRs.branches == [("a", "a_int"), ("a_int", "c")]
    assert CapMOS.reactive == [False, True, True] and CapMOS.vdep == [False, False, False]
    assert JunctionCap.reactive == [True] and JunctionCap.vdep == [True] and JunctionCap.n_charges == 1
    # per branch: 2 G stamps per module node, then 2 b stamps (vasim.jl:3382-3392, :3490-3518)
    plan = SimpleMOS.stamp_plan()
    assert [p[0] for p in plan] == ["G"] * 6 + ["b"] * 2
    assert SimpleMOS(K=2e-3).params == [2e-3, 0.5] and SimpleMOS(k=2e-3, vth=0.7).params == [2e-3, 0.7]
    with pytest.raises(cb.VAError):
        SimpleMOS(W=1.0)
    for bad in ("module m(a,b); electrical a,b; analog I(a,b) <+ $limit(V(a,b), 1, 2); endmodule",):
        with pytest.raises(cb.VAError):
            cb.va(bad)
    # a potential contribution carries a branch-current unknown and the constraint row
    # (vasim.jl:2362-2393): alloc_current!, G[p,I], G[I,p], G[I,k] for every module node, b[I]
    vs = cb.va("module vsrc(a); electrical a; analog V(a) <+ 1.0; endmodule")
    assert [p[0] for p in vs.stamp_plan()] == ["I", "G", "G", "G", "b"]
    src = SimpleMOS.emit_cuda()
    assert "va_stamp_SimpleMOS_" in src and "VA_G(0, 0, dI0);" in src and "VA_B(2, Ieq);" in src


# ---- known answers through the oracle -------------------------------------------
def test_vaddiode_current_known_answer():
    def f(ctx, params):
        a = get_node(ctx, "anode")
        stamp(VoltageSource(0.6, name="V1"), ctx, a, 0)
        stamp(VADDiode(Is=1e-14, N=1.0), ctx, a, 0)
    lc, nl, x, ok = oracle_dc(B(f))
    expected = 1e-14 * (math.exp(0.6 / 0.02585) - 1)
    assert ok and -x[lc.index_of("I_V1") - 1] == pytest.approx(expected, rel=0.01)   # vadistiller.jl:172-176

    def g(ctx, params):
        a = get_node(ctx, "anode")
        stamp(VoltageSource(0.7, name="V1"), ctx, a, 0)
        stamp(VADDiodeRs(Is=1e-14, N=1.0, Rs=10.0), ctx, a, 0)
    lc, nl, x, ok = oracle_dc(B(g))
    i = -x[lc.index_of("I_V1") - 1]
    assert ok and 0 < i < 0.1                                                          # vadistiller.jl:217-219
    assert lc.node_names == ["anode", "VADDiodeRs_VADDiodeRs_a_int"] and lc.n == 3


def test_chain_internal_nodes_carry_no_anchor_current():
    lc = cb.lower_circuit(cb.MNACircuit(B(chain)))
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    S = ora.Structure(nl, ora.make_spec(mode="dcop"))
    assert lc.node_names[4:] == ["d1_ChainDiodeRs_a_int", "d2_ChainDiodeRs_a_int", "d3_ChainDiodeRs_a_int"]
    coo = S.coo()
    assert np.array_equal(coo["G_I"], lc.G_I) and np.array_equal(coo["G_J"], lc.G_J) and np.array_equal(coo["b_I"], lc.b_I)
    u = np.zeros(S.n); u[:S.n_nodes] = 50.0
    G, C, b, _ = S.rebuild(u)
    F = S.dense(G) @ u - b
    for i, name in enumerate(lc.node_names):
        if name != "n1":
            # no gmin anchor to ground: only rounding of the 1190 A-scale cancellation remains
            # (the old anchor injected 5e-11 A; vadistiller.jl:286-296)
            assert abs(F[i]) < 1e-12
    x, ok, it = ora.solve_dc(nl, ora.make_spec(mode="dcop"))
    for name in ("na", "nb", "nc"):
        assert x[lc.index_of(name) - 1] == pytest.approx(50.0, abs=2e-2)               # vadistiller.jl:302-306


@pytest.mark.parametrize("vg,vd,expect,atol", [(1.5, 2.0, -0.0005, 1e-5), (1.5, 0.3, -0.000255, 1e-6),
                                               (0.3, 2.0, 0.0, 1e-8)])
def test_simple_mos_regions(vg, vd, expect, atol):
    lc, nl, x, ok = oracle_dc(mos_bias(SimpleMOS, vg, vd))
    assert ok and x[lc.index_of("I_Vd") - 1] == pytest.approx(expect, abs=atol)        # va_mosfet.jl:113-160


def test_cmos_inverter_dc():
    lc, nl, x, ok = oracle_dc(inverter(0.0))
    assert ok and x[lc.index_of("out") - 1] == pytest.approx(3.0, abs=0.01)            # va_mosfet.jl:411-417
    lc, nl, x, ok = oracle_dc(inverter(3.0))
    assert ok and x[lc.index_of("out") - 1] == pytest.approx(0.0, abs=0.01)            # va_mosfet.jl:452-458


def test_capmos_dc_and_gate_charging():
    lc, nl, x, ok = oracle_dc(mos_bias(CapMOS, 1.5, 2.0, Cgs=10e-15, Cgd=5e-15))
    assert ok and x[lc.index_of("I_Vd") - 1] == pytest.approx(-0.0005, abs=1e-5)       # va_mosfet.jl:211-214
    S = ora.Structure(nl, ora.make_spec(mode="tran"))
    G, C, b, _ = S.rebuild(x)
    assert np.any(C != 0.0) and S.n_charges == 0                                       # linear caps: plain C stamps

    Rg, Cgs, Cgd = 1e3, 10e-12, 5e-12
    def f(ctx, params):
        d = get_node(ctx, "d"); g = get_node(ctx, "g"); vin = get_node(ctx, "vin")
        stamp(VoltageSource(2.0, name="Vin"), ctx, vin, 0)
        stamp(Resistor(Rg, name="Rg"), ctx, vin, g)
        stamp(VoltageSource(3.0, name="Vd"), ctx, d, 0)
        stamp(CapMOS(K=1e-3, Vth=0.5, Cgs=Cgs, Cgd=Cgd), ctx, d, g, 0)
    lc = cb.lower_circuit(cb.MNACircuit(B(f)))
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    tau = Rg * (Cgs + Cgd)
    u0 = np.zeros(lc.n); u0[lc.index_of("vin") - 1] = 2.0; u0[lc.index_of("d") - 1] = 3.0   # gate starts at 0 V
    o = ora.make_tran_opts(method=1, dt=tau / 200, init=1)
    r = ora.tran(nl, ora.make_spec(mode="tran"), 0.0, 5 * tau, o, [lc.index_of("g")], u0=u0)
    assert r["status"] == 0
    assert r["u"][-1, 0] == pytest.approx(2.0 * (1 - math.exp(-5)), rel=0.05)          # va_mosfet.jl:258-262


def test_voltage_dependent_charge_uses_charge_state_formulation():
    """charge_formulation.jl:151-220: C[p,q] = 1/CHARGE_SCALE, G[q,q] = 1,
    G[q,p] = -CHARGE_SCALE*dQ/dVp, C[q,q] = 0 (the charge row is algebraic)."""
    def f(ctx, params):
        p = get_node(ctx, "p")
        stamp(VoltageSource(0.4, name="V1"), ctx, p, 0)
        stamp(JunctionCap(Cj0=1e-12, phi=0.8, m=0.5, name="J1"), ctx, p, 0)
    lc = cb.lower_circuit(cb.MNACircuit(B(f)))
    assert lc.n_charges == 1 and lc.charge_names == ["J1_JunctionCap_Q_p_n"] and lc.n == 3
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    S = ora.Structure(nl, ora.make_spec(mode="tran"))
    assert S.n_charges == 1
    p, q = 0, 2                       # unknowns [p | I_V1 | q]
    CS = 1e12
    # V = 0.4 is not < 0.5*phi: the guard branch holds C constant, so dQ/dV = C = Cj0/sqrt(0.5)
    # (the reference's own check, va_mosfet.jl:500-504)
    G, C, b, _ = S.rebuild(np.array([0.4, 0.0, 0.0]))
    Gd, Cd = S.dense(G), S.dense(C)
    assert Cd[p, q] == 1.0 / CS and Cd[q, q] == 0.0 and Gd[q, q] == 1.0
    assert Gd[q, p] == pytest.approx(-CS * 1e-12 / math.sqrt(0.5), rel=1e-12)
    assert abs(b[q]) < 1e-12
    # below the guard the s-dual charge of C(V)*ddt(V) is Q = C(V)*V  =>  dQ/dV = C + V*dC/dV
    v = 0.2
    G, C, b, _ = S.rebuild(np.array([v, 0.0, 0.0]))
    Gd = S.dense(G)
    Cj = 1e-12 / math.sqrt(1 - v / 0.8)
    dC = 1e-12 * 0.5 / 0.8 * (1 - v / 0.8) ** -1.5
    assert Gd[q, p] == pytest.approx(-CS * (Cj + v * dC), rel=1e-12)
    assert b[q] == pytest.approx(CS * (Cj * v - (Cj + v * dC) * v), rel=1e-9)


# ---- potential contributions: structure and known answers ------------------------------
def _rshort(rm):
    def f(ctx, params):
        vin = get_node(ctx, "vin"); out = get_node(ctx, "out")
        stamp(VoltageSource(3.0, name="V1"), ctx, vin, 0)
        stamp(RShort(r1=1e3, rm=params.rm if "rm" in params else rm, r2=2e3, name="X1"), ctx, vin, out)
        stamp(Resistor(3e3, name="RL"), ctx, out, 0)
    return B(f)


def _vref(ctx, params):
    a = get_node(ctx, "a"); o = get_node(ctx, "o")
    stamp(VoltageSource(params.va if "va" in params else 1.5, name="V1"), ctx, a, 0)
    stamp(VRef(gain=2.0, vofs=0.25, name="X1"), ctx, a, o)
    stamp(Resistor(1e3, name="RL"), ctx, o, 0)


def _lbranch(ctx, params):
    vin = get_node(ctx, "vin"); mid = get_node(ctx, "mid")
    stamp(VoltageSource(5.0, name="V1"), ctx, vin, 0)
    stamp(Resistor(params.r if "r" in params else 95.0, name="R1"), ctx, vin, mid)
    stamp(LBranch(L=1e-3, rs=5.0, name="L1"), ctx, mid, 0)


def test_potential_contributions_structure_and_dc():
    # executed internal-to-internal short: one branch current, 2 + 2 + N stamps and a b stamp
    lc0, nl0, x0, ok0 = oracle_dc(_rshort(0.0))
    lc1, nl1, x1, ok1 = oracle_dc(_rshort(500.0))
    assert ok0 and ok1
    assert (lc0.n_nodes, lc0.n_currents) == (4, 2) and (lc1.n_nodes, lc1.n_currents) == (4, 1)
    assert x0[lc0.index_of("out") - 1] == pytest.approx(3.0 * 3e3 / 6e3, rel=1e-12)
    assert x1[lc1.index_of("out") - 1] == pytest.approx(3.0 * 3e3 / 6.5e3, rel=1e-12)
    i_br = x0[lc0.n_nodes + 1]                            # the site's branch current carries the chain current
    assert i_br == pytest.approx(3.0 / 6e3, rel=1e-12)
    for lc, nl in ((lc0, nl0), (lc1, nl1)):               # host structure == the oracle's own run of the builder
        coo = ora.Structure(nl, ora.make_spec(mode="dcop")).coo()
        assert np.array_equal(coo["G_I"], lc.G_I) and np.array_equal(coo["G_J"], lc.G_J)
        assert np.array_equal(coo["b_I"], lc.b_I)
    # bias-dependent potential contribution: Newton on the linearised constraint row
    lc, nl, x, ok = oracle_dc(B(_vref))
    assert ok and x[lc.index_of("o") - 1] == pytest.approx(2.0 * 1.5 ** 2 + 0.25, rel=1e-12)
    # named branch: DC short through rs (I = 5 / (95 + 5)), current probe = the branch unknown
    lc, nl, x, ok = oracle_dc(B(_lbranch))
    assert ok and lc.n_currents == 2
    assert x[lc.index_of("mid") - 1] == pytest.approx(5.0 * 5.0 / 100.0, rel=1e-9)
    assert [p[0] for p in LBranch.stamp_plan()] == ["I", "G", "G", "G", "G", "b", "C"]
    assert RShort(rm=0.0).vsites(cb.MNASpec(), [1, 2, 3, 4]) == (True,)
    assert RShort(rm=1.0).vsites(cb.MNASpec(), [1, 2, 3, 4]) == (False,)      # not executed
    assert RShort(rm=0.0).aliased_sites([1, 2, 3, 3]) == (True,)              # same circuit node: nothing to stamp
    assert RShort(rm=0.0).aliased_sites([1, 2, 3, 4]) == (False,)


def test_behavioral_sources_trace_and_explicit_stamping():
    """BehavioralVoltageSource / BehavioralCurrentSource (devices.jl:1003-1131): the closure is traced into a
    one-statement module; the value lands in b only (no Jacobian entries), the current of the voltage form is
    named I_<name>, current flows from n to p."""
    cs = GPU_SWEEPS["bsrc"]()
    params, P = cs.lane_params()
    lc = cb.lower(cs.builder, params, cb.MNASpec(mode="dcop"), P=P)
    assert lc.current_names == ["I_V1", "I_Vb", "I_B1"]
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    x, st, it = ora.sweep_dc(nl, ora.make_spec(mode="dcop", ), lc.n)
    assert (st == 0).all()
    gain = np.asarray(params.gain)
    # DC: the pulse is at 0 -> mid = 0; out = gain tanh(-0.2) + 0.1; sq = 0; fb: v = -1e3 (1e-3 + gfb v) -> v = -1 / 1.25
    assert np.allclose(x[:, lc.index_of("out") - 1], gain * np.tanh(-0.2) + 0.1, rtol=1e-12)
    assert np.allclose(x[:, lc.index_of("sq") - 1], 0.0, atol=1e-15)
    # (I1 pulls 1 mA out of fb; the fixed-point iteration stops at the Newton tolerance, not at rounding)
    assert np.allclose(x[:, lc.index_of("fb") - 1], -0.8, rtol=1e-6)
    assert np.allclose(x[:, lc.index_of("I_B1") - 1], -(gain * np.tanh(-0.2) + 0.1) / 2e3, rtol=1e-12)
    # explicit stamping: v_{k+1} = 1 - 0.25 v_k contracts by 4 per iteration -- a Newton iteration with the
    # partial in G would need two
    assert np.all(it > 8)
    with pytest.raises(cb.VAError):
        cb.BehavioralCurrentSource(lambda V: math.exp(V("a"))).lower()
    with pytest.raises(cb.VAError):
        cb.BehavioralCurrentSource(lambda V: 1.0 if V("a") > 0 else 0.0).lower()


# ---- GPU parity -----------------------------------------------------------------
def _sweep_vs_oracle(cs, tran=None):
    params, P = cs.lane_params()
    lc = cb.lower(cs.builder, params, cb.MNASpec(mode="tran"), P=P)
    ora.load_va_models(lc.va_c_source)
    nl = ora.OracleNetlist(lc.netlist_tables())
    comp = cb.CompiledSweep(lc, cb.MNASpec(mode="tran"))
    out = {}
    try:
        x, st, it = comp.dc()
        xo, sto, ito = ora.sweep_dc(nl, ora.make_spec(mode="dcop"), lc.n)
        out["dc"] = (x.T, xo, st, sto)
        if tran:
            tspan, dt, method, spec = tran
            save = list(range(1, lc.n + 1))
            wave = comp.tran(tspan, dt, method=method, save_idxs=save, specialize=spec)
            r = wave.fetch(); wave.free()
            o = ora.make_tran_opts(method={"be": 0, "trap": 1, "gear2": 2}[method], dt=dt)
            ro = ora.sweep_tran(nl, ora.make_spec(mode="tran"), tspan[0], tspan[1], o, save)
            out["tran"] = (np.transpose(r["u"], (2, 1, 0)), ro["u"], r, ro, comp.handle.is_specialized())
    finally:
        comp.close()
    return lc, out


def _close(a, b, rtol=1e-9, atol=1e-12):
    return np.all(np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b)))


def _capmos_stage(ctx, params):
    d = get_node(ctx, "d"); g = get_node(ctx, "g"); vin = get_node(ctx, "vin"); vdd = get_node(ctx, "vdd")
    stamp(VoltageSource(0.0, tran=cb.PulseWave(0.0, 2.0, 1e-9, 1e-9, 1e-9, 20e-9, 60e-9), name="Vin"), ctx, vin, 0)
    stamp(VoltageSource(3.0, name="Vdd"), ctx, vdd, 0)
    stamp(Resistor(params.rg, name="Rg"), ctx, vin, g)
    stamp(Resistor(2e3, name="Rd"), ctx, vdd, d)
    stamp(CapMOS(K=params.k, Vth=0.5, Cgs=1e-12, Cgd=0.5e-12), ctx, d, g, 0)
    stamp(JunctionCap(Cj0=2e-13, phi=0.8, m=0.5, name="Jdb"), ctx, 0, d)     # reverse-biased drain junction


def _bsrc(ctx, params):
    """Behavioural sources (devices.jl:1003-1131): a voltage B-source driven by a filtered pulse, a current
    B-source with np.where, and a current B-source that senses its own node (explicit stamping -> the
    Newton loop is a fixed-point iteration with loop gain gfb * 1 kOhm)."""
    stamp(VoltageSource(0.0, tran=cb.PulseWave(0.0, params.va, 2e-6, 1e-6, 1e-6, 6e-6, 20e-6), name="V1"), ctx, "in", 0)
    stamp(Resistor(1e3, name="R1"), ctx, "in", "mid")
    stamp(cb.Capacitor(1e-9, name="C1"), ctx, "mid", 0)
    stamp(Resistor(1e3, name="R2"), ctx, "mid", 0)
    stamp(VoltageSource(1.0, name="Vb"), ctx, "bias", 0)
    stamp(cb.BehavioralVoltageSource(lambda V: params.gain * np.tanh(V("mid") - 0.2) + 0.1 * V("bias"), name="B1"),
          ctx, "out", 0)
    stamp(Resistor(2e3, name="RL"), ctx, "out", 0)
    stamp(cb.BehavioralCurrentSource(lambda V: np.where(V("mid") > 0.3, 1e-3 * (V("mid") - 0.3) ** 2, 0.0), name="B2"),
          ctx, "sq", 0)
    stamp(Resistor(1e3, name="Rsq"), ctx, "sq", 0)
    stamp(cb.CurrentSource(1e-3, name="I1"), ctx, 0, "fb")
    stamp(Resistor(1e3, name="Rfb"), ctx, "fb", 0)
    stamp(cb.BehavioralCurrentSource(lambda V: -params.gfb * V("fb"), name="B3"), ctx, "fb", 0)


GPU_SWEEPS = {
    "bsrc": lambda: cb.CircuitSweep(B(_bsrc), cb.ProductSweep(va=[0.5, 1.0, 2.0], gain=[1.0, 2.5]), gfb=0.25e-3),
    "inverter": lambda: cb.CircuitSweep(inverter(0.0), cb.ProductSweep(vin=np.linspace(0.0, 3.0, 13),
                                                                        kn=[0.5e-3, 1e-3, 2e-3])),
    "capmos": lambda: cb.CircuitSweep(B(_capmos_stage), cb.ProductSweep(rg=[500.0, 1e3, 2e3], k=[0.5e-3, 1e-3])),
    "chain": lambda: cb.CircuitSweep(B(chain), cb.Sweep(dummy=[0.0, 1.0]), dummy=0.0),
    "rshort": lambda: cb.CircuitSweep(_rshort(0.0), cb.Sweep(dummy=[0.0, 1.0, 2.0]), dummy=0.0),
    "vref": lambda: cb.CircuitSweep(B(_vref), cb.Sweep(va=np.linspace(-1.0, 2.0, 7))),
    "lbranch": lambda: cb.CircuitSweep(B(_lbranch), cb.Sweep(r=[45.0, 95.0, 195.0])),
}


def gpu_va_headers():
    """Emitted CUDA headers of the GPU tests' circuits: __graft_entry__.build() pre-builds their
    kernel sets so the GPU box finds them in the in-tree cache."""
    out = []
    for make in GPU_SWEEPS.values():
        cs = make()
        params, P = cs.lane_params()
        out.append(cb.lower(cs.builder, params, cb.MNASpec(mode="tran"), P=P).va_cuda_header)
    return out


@pytest.mark.gpu
def test_gpu_va_inverter_sweep_dc_and_transient():
    cs = GPU_SWEEPS["inverter"]()
    lc, out = _sweep_vs_oracle(cs)
    x, xo, st, sto = out["dc"]
    assert np.array_equal(st, sto) and (st == 0).all()
    assert _close(x, xo)
    vout = x[:, lc.index_of("out") - 1].reshape(3, 13)
    assert np.all(np.diff(vout, axis=1) <= 1e-9)          # transfer curve is monotone
    assert vout[1, 0] == pytest.approx(3.0, abs=0.01) and vout[1, -1] == pytest.approx(0.0, abs=0.01)


@pytest.mark.gpu
@pytest.mark.parametrize("spec", [False, True], ids=["table", "specialised"])
def test_gpu_va_capmos_transient(spec):
    cs = GPU_SWEEPS["capmos"]()
    lc, out = _sweep_vs_oracle(cs, tran=((0.0, 4e-8), 2e-10, "trap", spec))
    assert lc.n_charges == 1
    gpu, ref, r, ro, is_spec = out["tran"]
    assert is_spec == spec
    assert np.array_equal(r["status"], ro["status"]) and (r["status"] == 0).all()
    assert _close(gpu, ref[:, :gpu.shape[1], :]), float(np.max(np.abs(gpu - ref[:, :gpu.shape[1], :])))
    assert np.array_equal(r["newton_iters"], ro["newton_iters"])
    d = gpu[:, :, lc.index_of("d") - 1]
    assert d.min() < 2.0 and d.max() > 2.9                 # the stage actually switches


@pytest.mark.gpu
def test_gpu_va_chain_dc():
    cs = GPU_SWEEPS["chain"]()
    lc, out = _sweep_vs_oracle(cs)
    x, xo, st, sto = out["dc"]
    # a floating chain end behind pA junctions: the operating point is determined only to
    # ~abstol/Is-slope, so pivot order shows at 1e-6 relative; the reference itself checks
    # this circuit to 2e-2 V (vadistiller.jl:302-306)
    assert np.array_equal(st, sto) and (st == 0).all()
    assert np.allclose(x, xo, rtol=0.0, atol=1e-3)
    for name in ("na", "nb", "nc"):
        assert np.allclose(x[:, lc.index_of(name) - 1], 50.0, atol=2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rshort", "vref", "lbranch"])
def test_gpu_va_potential_contributions(name):
    """Branch-current unknowns of potential contributions (internal-to-internal short, bias-dependent
    V(o) <+ f(V(i)), named branch with current probe and ddt): DC and a BE transient vs the oracle."""
    cs = GPU_SWEEPS[name]()
    lc, out = _sweep_vs_oracle(cs, tran=((0.0, 2e-5), 2e-7, "be", False))
    x, xo, st, sto = out["dc"]
    assert np.array_equal(st, sto) and (st == 0).all()
    assert _close(x, xo), float(np.max(np.abs(x - xo)))
    gpu, ref, r, ro, _ = out["tran"]
    assert np.array_equal(r["status"], ro["status"]) and (r["status"] == 0).all()
    assert _close(gpu, ref[:, :gpu.shape[1], :]), float(np.max(np.abs(gpu - ref[:, :gpu.shape[1], :])))
    assert np.array_equal(r["newton_iters"], ro["newton_iters"])
    if name == "vref":
        va = np.linspace(-1.0, 2.0, 7)
        assert np.allclose(x[:, lc.index_of("o") - 1], 2.0 * va ** 2 + 0.25, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("spec", [False, True], ids=["table", "specialised"])
def test_gpu_behavioral_sources(spec):
    """B-sources (traced closures, explicit stamping) on the device: DC and a BE transient vs the oracle."""
    cs = GPU_SWEEPS["bsrc"]()
    lc, out = _sweep_vs_oracle(cs, tran=((0.0, 2e-5), 1e-7, "be", spec))
    x, xo, st, sto = out["dc"]
    assert np.array_equal(st, sto) and (st == 0).all()
    assert _close(x, xo), float(np.max(np.abs(x - xo)))
    gpu, ref, r, ro, is_spec = out["tran"]
    assert is_spec == spec
    assert np.array_equal(r["status"], ro["status"]) and (r["status"] == 0).all()
    assert _close(gpu, ref[:, :gpu.shape[1], :]), float(np.max(np.abs(gpu - ref[:, :gpu.shape[1], :])))
    assert np.array_equal(r["newton_iters"], ro["newton_iters"])
    sq = gpu[:, :, lc.index_of("sq") - 1]
    assert sq.max() > 0.05                                 # the squarer switched on during the pulse


@pytest.mark.gpu
def test_gpu_va_requires_models():
    from cadnip_b200 import backend
    lc = cb.lower_circuit(cb.MNACircuit(mos_bias(SimpleMOS, 1.5, 2.0)))
    models = lc.va_cuda_header
    lc.va_cuda_header = ""               # handle created without loading the emitted models
    h = backend.Handle(lc)
    try:
        h.set_lanes(lc.lane_soa, 1)
        with pytest.raises(backend.CB200Error):
            h.dc(cb.MNASpec(mode="dcop"))
    finally:
        h.close()
        lc.va_cuda_header = models
