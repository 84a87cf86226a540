"""Pins the CPU oracle against the reference's own known-answer tests.
Golden constants live in tests/golden/reference_kat.json (hand-transcribed from the
reference's test files, each entry cites file:line)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import cadnip_b200 as cb
import cadnip_oracle as ora
import circuits
from cadnip_b200 import (MNAContext, ZERO_VECTOR, get_node, stamp, Resistor, Capacitor, Inductor,
                         VoltageSource, CurrentSource, VCVS, VCCS, CCVS, CCCS)

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))


def lower_one(builder, **params):
    return cb.lower_circuit(cb.MNACircuit(builder, **params))


def oracle_structure(builder, mode="dcop", **params):
    lc = lower_one(builder, **params)
    nl = ora.OracleNetlist(lc.netlist_tables())
    return lc, nl, ora.Structure(nl, ora.make_spec(mode=mode))


# ---- pure functions -------------------------------------------------------
def test_pnjlim_known_answers():
    L = ora.lib()
    for k in KAT["pnjlim"]:
        lim = C.c_int()
        v = L.ora_pnjlim(*k["args"], C.byref(lim))
        if k.get("exact"):
            assert v == k["vlim"], k["src"]
        else:
            assert v == pytest.approx(k["vlim"], rel=1e-12, abs=1e-15), k["src"]
        assert bool(lim.value) == k["limited"], k["src"]
    vnew, vold, vt, vcrit = KAT["pnjlim_compression"]["args"]
    lim = C.c_int()
    v = L.ora_pnjlim(vnew, vold, vt, vcrit, C.byref(lim))
    assert lim.value == 1 and vold < v < vnew and v < 1.0


def test_diode_iv_linear_extension():
    L = ora.lib()
    d = KAT["diode_iv"]
    Is, nVt = d["Is"], d["nVt"]

    def iv(v):
        i, g = C.c_double(), C.c_double()
        L.ora_diode_iv(Is, nVt, v, C.byref(i), C.byref(g))
        return i.value, g.value

    I_lo, G_lo = iv(80.0 * nVt - 1e-9)
    I_hi, G_hi = iv(80.0 * nVt + 1e-9)
    assert I_lo == pytest.approx(I_hi, rel=1e-6) and G_lo == pytest.approx(G_hi, rel=1e-6)
    I1, G1 = iv(10.0)
    I2, G2 = iv(11.0)
    assert math.isfinite(I1) and math.isfinite(I2) and G1 == G2
    assert I2 - I1 == pytest.approx(G1 * 1.0, rel=1e-12)
    I3, G3 = iv(d["exact_at"])
    assert I3 == pytest.approx(d["I_exact"], rel=1e-14)
    assert G3 == pytest.approx(d["G_exact"], rel=1e-14)


def test_vcrit_matches_constructor():
    # Diode constructor (devices.jl:1319-1320): host (numpy) and oracle (libm) agree
    d = cb.Diode(Is=1e-14, Vt=0.026, n=1.0)
    assert d.vcrit == pytest.approx(ora.lib().ora_diode_vcrit(1e-14, 0.026, 1.0), rel=1e-15)


def test_coo_to_csc_mapping():
    L = ora.lib()
    for k in KAT["coo_to_csc"]:
        I = np.array(k["I"], np.int64); J = np.array(k["J"], np.int64); V = np.array(k["V"])
        n = k["n"]
        colptr = np.zeros(n + 1, np.int64); rowval = np.zeros(len(I), np.int64)
        nz = np.zeros(len(I)); mp = np.zeros(len(I), np.int64)
        nnz = L.ora_sparse_coo(n, len(I), ora._lp(I), ora._lp(J), ora._dp(V), ora._lp(colptr),
                               ora._lp(rowval), ora._dp(nz), ora._lp(mp))
        # rebuild through the mapping reproduces sparse(I,J,V)  (precompile.jl:28-40)
        nz2 = np.zeros(nnz)
        for q in range(len(I)):
            assert mp[q] > 0
            nz2[mp[q] - 1] += V[q]
        assert np.array_equal(nz2, nz[:nnz])
        dense = np.zeros((n, n))
        for j in range(n):
            rows = rowval[colptr[j] - 1:colptr[j + 1] - 1]
            assert list(rows) == sorted(rows)            # rows ascending inside a column
            for idx in range(colptr[j], colptr[j + 1]):
                dense[rowval[idx - 1] - 1, j] = nz[idx - 1]
        if "dup" in k:
            a, b = k["dup"]
            assert mp[a] == mp[b]
            assert np.allclose(dense, np.array(k["dense"]))
        ref = np.zeros((n, n))
        for i, j, v in zip(I, J, V):
            ref[i - 1, j - 1] += v
        assert np.array_equal(dense, ref)


def test_waveforms():
    L = ora.lib()
    p = KAT["pulse"]
    for t, v in p["points"]:
        got = L.ora_pulse_at_time(p["v1"], p["v2"], p["td"], p["tr"], p["tf"], p["pw"], p["per"], t)
        assert got == pytest.approx(v, abs=1e-9), (t, got, v)
    # PWL: hold ends, linear inside, equal-time vertices average (devices.jl:47-71)
    ts = np.array([0.0, 1e-3, 1e-3, 2e-3]); ys = np.array([0.0, 1.0, 3.0, 3.0])
    f = lambda t: L.ora_pwl_at_time(ora._dp(ts), ora._dp(ys), 4, t)
    assert f(-1.0) == 0.0 and f(5.0) == 3.0
    assert f(0.5e-3) == pytest.approx(0.5)
    assert f(1e-3) == 2.0          # infinitely steep segment: (ys[i-1] + ys[i]) / 2
    assert f(1.5e-3) == 3.0
    # SIN uses sind of degrees (devices.jl:168-174)
    assert L.ora_sin_wave(0.0, 5.0, 1e3, 0.0, 0.0, 0.0, 0.25e-3) == pytest.approx(5.0, rel=1e-15)
    assert L.ora_sin_wave(0.0, 5.0, 1e3, 0.0, 0.0, 0.0, 0.5e-3) == 0.0      # sind(180) == 0 exactly
    assert L.ora_sin_wave(1.0, 2.0, 1e3, 1e-3, 0.0, 30.0, 0.0) == pytest.approx(2.0, rel=1e-15)
    for deg in (0.0, 30.0, 45.0, 90.0, 135.0, 180.0, 200.0, 270.0, 300.0, 359.0, 720.5, -30.0, -200.0):
        assert L.ora_sind(deg) == pytest.approx(math.sin(math.radians(deg % 360 if deg >= 0 else deg)), abs=2e-16)


def test_breakpoints():
    b = KAT["breakpoints_pulse"]
    w = cb.PulseWave(*b["wave"])
    times, period, count = cb.breakpoints(w)
    assert times == pytest.approx(b["edges"], rel=1e-15) and period == b["period"] and count == -1
    host = cb.expand_breakpoints([w], (0.0, 5e-3))
    L = ora.lib()
    wp = np.array(b["wave"]); out = np.zeros(64)
    m = L.ora_expand_breakpoints(ora.ORA_WAVE_PULSE if hasattr(ora, "ORA_WAVE_PULSE") else 2,
                                 ora._dp(wp), 7, 0.0, 5e-3, ora._dp(out), 64)
    assert list(out[:m]) == host
    assert all(0.0 < t < 5e-3 for t in host) and host == sorted(host)
    # tr = tf = 0: coincident edges are deduplicated by expand_breakpoints
    wz = cb.PulseWave(0.0, 1.0, 1e-6, 0.0, 0.0, 1e-3, 2e-3)
    hz = cb.expand_breakpoints([wz], (0.0, 3e-3))
    assert hz == pytest.approx([1e-6, 1.001e-3, 2.001e-3], rel=1e-12)
    assert cb.breakpoints(cb.SinWave(0.0, 1.0, 1e3)) is None
    assert cb.breakpoints(cb.SinWave(0.0, 1.0, 1e3, 1e-3))[0] == [1e-3]
    assert cb.breakpoints(cb.PWLWave([0.0, 1e-3, 2e-3], [0.0, 1.0, 0.0]))[0] == [0.0, 1e-3, 2e-3]


# ---- primitive stamps: exact matrix entries (test/mna/core.jl:265-507) -----
def _mats(builder):
    lc, nl, S = oracle_structure(builder)
    G, Cm, b, _ = S.rebuild(None)
    return lc, S, S.dense(G), S.dense(Cm), b


def _two_node(dev):
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        n1 = get_node(ctx, "n1"); n2 = get_node(ctx, "n2")
        stamp(dev, ctx, n1, n2)
        return ctx
    return build


def _four_node(dev):
    def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        nodes = [get_node(ctx, s) for s in ("out_p", "out_n", "in_p", "in_n")]
        stamp(dev, ctx, *nodes)
        return ctx
    return build


def test_primitive_stamps_exact():
    _, _, G, Cm, b = _mats(_two_node(Resistor(1000.0)))
    g = 1.0 / 1000.0
    assert np.array_equal(G, np.array([[g, -g], [-g, g]])) and not Cm.any()
    _, _, G, Cm, b = _mats(_two_node(Capacitor(1e-6)))
    assert not G.any() and np.array_equal(Cm, np.array([[1e-6, -1e-6], [-1e-6, 1e-6]]))
    lc, S, G, Cm, b = _mats(_two_node(Inductor(1e-3, name="L1")))
    assert S.n == 3 and G[0, 2] == 1.0 and G[1, 2] == -1.0 and G[2, 0] == 1.0 and G[2, 1] == -1.0
    assert Cm[2, 2] == -1e-3
    lc, S, G, Cm, b = _mats(_four_node(VCCS(0.01)))
    assert (G[0, 2], G[0, 3], G[1, 2], G[1, 3]) == (-0.01, 0.01, 0.01, -0.01)
    lc, S, G, Cm, b = _mats(_four_node(VCVS(10.0, name="E1")))
    assert S.n == 5 and (G[0, 4], G[1, 4], G[4, 0], G[4, 1], G[4, 2], G[4, 3]) == (1.0, -1.0, 1.0, -1.0, -10.0, 10.0)
    lc, S, G, Cm, b = _mats(_four_node(CCVS(1000.0, name="H1")))
    assert S.n == 6
    assert (G[2, 4], G[3, 4], G[4, 2], G[4, 3], G[0, 5], G[1, 5], G[5, 0], G[5, 1], G[5, 4]) == \
        (1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0, -1.0, -1000.0)
    assert lc.current_names == ["I_H1_in", "I_H1_out"]
    lc, S, G, Cm, b = _mats(_four_node(CCCS(2.0, name="F1")))
    assert S.n == 5 and (G[2, 4], G[3, 4], G[4, 2], G[4, 3], G[0, 4], G[1, 4]) == (1.0, -1.0, 1.0, -1.0, -2.0, 2.0)

    def vsrc(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        vcc = get_node(ctx, "vcc")
        I = stamp(VoltageSource(5.0, name="V1"), ctx, vcc, 0)
        assert isinstance(I, cb.CurrentIndex) and cb.resolve_index(ctx, I) == 2
        return ctx
    lc, S, G, Cm, b = _mats(vsrc)
    assert G.shape == (2, 2) and G[0, 1] == 1.0 and G[1, 0] == 1.0 and b[1] == 5.0

    def isrc(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        n1 = get_node(ctx, "n1")
        stamp(CurrentSource(0.001), ctx, n1, 0)
        return ctx
    lc, S, G, Cm, b = _mats(isrc)
    assert S.nnz == 0 and b[0] == 0.001


# ---- DC linear circuits (test/mna/core.jl:509-637) --------------------------
def _dc_builders():
    def B(f):
        def build(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
            ctx = MNAContext() if ctx is None else ctx
            f(ctx)
            return ctx
        return build

    def divider(ctx, r1=1e3):
        vcc = get_node(ctx, "vcc"); out = get_node(ctx, "out")
        stamp(VoltageSource(5.0, name="V1"), ctx, vcc, 0)
        stamp(Resistor(r1), ctx, vcc, out); stamp(Resistor(1e3), ctx, out, 0)

    def isrc(ctx):
        n1 = get_node(ctx, "n1")
        stamp(CurrentSource(1e-3, name="I1"), ctx, n1, 0)   # SPICE `I1 0 n1` (codegen.jl:2550 swaps)
        stamp(Resistor(1e3), ctx, n1, 0)

    def two_v(ctx):
        vcc = get_node(ctx, "vcc"); mid = get_node(ctx, "mid")
        stamp(VoltageSource(5.0, name="V1"), ctx, vcc, 0); stamp(VoltageSource(3.0, name="V2"), ctx, mid, 0)
        stamp(Resistor(1e3), ctx, vcc, mid); stamp(Resistor(1e3), ctx, mid, 0)

    def vccs(ctx):
        inp = get_node(ctx, "inp"); out = get_node(ctx, "out")
        stamp(VoltageSource(1.0, name="V1"), ctx, inp, 0)
        stamp(VCCS(0.01, name="G1"), ctx, out, 0, inp, 0)
        stamp(Resistor(1e3), ctx, out, 0)

    def vcvs(ctx):
        inp = get_node(ctx, "inp"); out = get_node(ctx, "out")
        stamp(VoltageSource(0.5, name="V1"), ctx, inp, 0)
        stamp(VCVS(-10.0, name="E1"), ctx, out, 0, inp, 0)

    def ccvs(ctx):
        inp = get_node(ctx, "inp"); out = get_node(ctx, "out")
        I = stamp(VoltageSource(0.0, name="V_sense"), ctx, inp, 0)
        stamp(CurrentSource(1e-3, name="I1"), ctx, inp, 0)
        stamp(CCVS(1000.0, name="H1"), ctx, out, 0, I)
        stamp(Resistor(1e6), ctx, out, 0)

    def cccs(ctx):
        inp = get_node(ctx, "inp"); out = get_node(ctx, "out")
        I = stamp(VoltageSource(0.0, name="V_sense"), ctx, inp, 0)
        stamp(CurrentSource(1e-3, name="I1"), ctx, inp, 0)
        stamp(CCCS(2.0, name="F1"), ctx, out, 0, I)
        stamp(Resistor(1e3), ctx, out, 0)

    def multi(ctx):
        n1, n2, n3, c = [get_node(ctx, s) for s in ("n1", "n2", "n3", "center")]
        stamp(VoltageSource(3.0, name="V1"), ctx, n1, 0); stamp(VoltageSource(6.0, name="V2"), ctx, n2, 0)
        stamp(VoltageSource(9.0, name="V3"), ctx, n3, 0)
        stamp(Resistor(1e3), ctx, n1, c); stamp(Resistor(2e3), ctx, n2, c); stamp(Resistor(3e3), ctx, n3, c)

    return {"divider": B(divider), "divider_unequal": B(lambda c: divider(c, 2e3)), "isrc_resistor": B(isrc),
            "two_vsources": B(two_v), "vccs_amp": B(vccs), "vcvs_inv": B(vcvs), "ccvs_transres": B(ccvs),
            "cccs_mirror": B(cccs), "multinode": B(multi)}


DC_BUILDERS = _dc_builders()


@pytest.mark.parametrize("case", KAT["dc_linear"], ids=lambda c: c["name"])
def test_dc_linear_known_answers(case):
    lc = lower_one(DC_BUILDERS[case["name"]])
    nl = ora.OracleNetlist(lc.netlist_tables())
    x, ok, _ = ora.solve_dc(nl, ora.make_spec(mode="dcop"))
    assert ok, case["src"]
    for name, v in case["expect"].items():
        sign = 1.0
        if case["name"] in ("ccvs_transres", "cccs_mirror") and name == "out":
            # I1 drives +1 mA into `inp`; V_sense (inp -> 0, 0 V) carries it to ground, so
            # I(V_sense) = +1 mA and out = rm*I = +1 V / gain*I*R = +2 V as in the reference.
            sign = 1.0
        assert x[lc.index_of(name) - 1] == pytest.approx(sign * v, abs=case["atol"]), case["src"]


# ---- PCNR (test/mna/pcnr.jl:268-357) ----------------------------------------
def test_rectifier_fixed_point_and_iterations():
    k = KAT["rectifier"]
    lc, nl, S = oracle_structure(circuits.rectifier(True))
    assert S.n == k["n"] and S.n_limits == k["n_limits"]
    assert lc.limit_names == ["D1_vdlim"]
    u, ok, iters = S.pcnr()
    assert ok and iters <= k["pcnr_iters_max"]
    out = u[lc.index_of("out") - 1]
    assert k["out_range"][0] < out < k["out_range"][1]
    lc2, nl2, S2 = oracle_structure(circuits.rectifier(False))
    u2, ok2, _ = S2.dc()
    assert ok2 and abs(out - u2[lc2.index_of("out") - 1]) < k["lim_vs_nolim_atol"]
    I_R = (5.0 - out) / 1000.0
    I_D = 1e-14 * (math.exp(out / 0.026) - 1.0)
    assert I_R == pytest.approx(I_D, rel=1e-2)
    assert u[lc.index_of("D1_vdlim") - 1] == pytest.approx(out, abs=1e-6)
    # warm start strictly fewer iterations (test/mna/pcnr.jl:365-385)
    _, okw, itw = S.pcnr(u0=u)
    assert okw and itw < iters
    # matrix structure at the ZERO_VECTOR build point (test/mna/pcnr.jl:242-262)
    G, _, _, _ = S.rebuild(None)
    D = S.dense(G)
    li, oi = S.n - 1, lc.index_of("out") - 1
    assert D[li, li] == 1.0 and D[li, oi] == -1.0 and D[oi, li] == 0.0
    assert D[oi, oi] == pytest.approx(1 / 1000.0, rel=1e-6)


def test_chain_stiff_series():
    k = KAT["chain"]
    lc, nl, S = oracle_structure(circuits.chain(True))
    assert S.n_limits == k["n_limits"]
    u, ok, iters = S.pcnr()
    assert ok and iters <= k["pcnr_iters_max"]
    n1, n2, n3 = [u[lc.index_of(s) - 1] for s in ("n1", "n2", "n3")]
    vd = [n1 - n2, n2 - n3, n3]
    assert vd[0] == pytest.approx(vd[1], rel=k["equal_rtol"]) and vd[1] == pytest.approx(vd[2], rel=k["equal_rtol"])
    assert k["vd_range"][0] < vd[2] < k["vd_range"][1]
    assert (50.0 - n1) / 1000.0 == pytest.approx(1e-14 * (math.exp(vd[2] / 0.026) - 1.0), rel=1e-2)


def test_structure_invariant_under_state():
    # nonlinear structure invariance (test/mna/precompile.jl:205-242): same pattern at any x
    lc, nl, S = oracle_structure(circuits.rectifier(True))
    a = S.arrays()
    rng = np.random.default_rng(0)
    L = ora.lib()
    for _ in range(3):
        x = rng.uniform(-1, 1, S.n)
        c2 = L.ora_build(C.byref(nl.c), C.byref(ora.make_spec(mode="dcop")), 0.0, ora._dp(x), S.n)
        sz = [C.c_int64() for _ in range(7)]
        L.ora_ctx_sizes(c2, *[C.byref(s) for s in sz])
        assert [s.value for s in sz] == [S.n_nodes, S.n_currents, S.n_charges, S.n_limits, S.nG, S.nC, S.nb]
        L.ora_ctx_free(c2)
    assert a["colptr"][-1] - 1 == S.nnz == 8


# ---- sweeps and transients ---------------------------------------------------
def test_sweep_divider_grid():
    k = KAT["sweep_divider"]
    r = np.arange(k["R1"][0], k["R1"][2] + 1, k["R1"][1])
    cs = cb.CircuitSweep(circuits.divider, cb.ProductSweep(R1=r, R2=r))
    params, P = cs.lane_params()
    lc = cb.lower(cs.builder, params, cb.MNASpec(mode="dcop"), P=P)
    nl = ora.OracleNetlist(lc.netlist_tables())
    x, st, it = ora.sweep_dc(nl, ora.make_spec(mode="dcop"), lc.n)
    iv = lc.index_of("I_V") - 1
    pts = cs.iterator.points()
    assert len(pts) == P == 400
    for p in (0, 1, 19, 20, 399):
        assert x[p, iv] == pytest.approx(-1.0 / (pts[p]["R1"] + pts[p]["R2"]), abs=k["atol"])
    assert (st == 0).all()
    # continuation == cold (test/sweep.jl:322-353)
    xw, stw, itw = ora.sweep_dc(nl, ora.make_spec(mode="dcop"), lc.n, continuation=True)
    assert np.allclose(xw, x, atol=1e-8)


def test_chain_continuation_matches_cold():
    v = np.arange(0.5, 20.01, 0.5)
    cs = cb.CircuitSweep(circuits.chain(True), cb.Sweep(vsrc=v))
    params, P = cs.lane_params()
    lc = cb.lower(cs.builder, params, cb.MNASpec(mode="dcop"), P=P)
    nl = ora.OracleNetlist(lc.netlist_tables())
    cold, st, itc = ora.sweep_dc(nl, ora.make_spec(mode="dcop"), lc.n)
    warm, stw, itw = ora.sweep_dc(nl, ora.make_spec(mode="dcop"), lc.n, continuation=True)
    assert P == 40 and (st == 0).all() and (stw == 0).all()
    n1 = lc.index_of("n1") - 1
    iv = lc.index_of("I_V1") - 1
    assert np.allclose(warm[:, n1], cold[:, n1], atol=1e-8)
    assert np.allclose(warm[:, iv], cold[:, iv], rtol=1e-6)
    assert (np.diff(warm[:, n1]) > 0).all() and 1.5 < warm[-1, n1] < 2.5
    assert itw.sum() < itc.sum()           # src/sweeps.jl:499-502: continuation saves iterations


@pytest.mark.parametrize("method", ["be", "trap", "gear2"])
def test_rc_charging_closed_form(method):
    k = KAT["rc_charge"]
    tau = k["R"] * k["C"]
    lc = lower_one(circuits.rc_charge(k["Vcc"], k["R"], k["C"]))
    nl = ora.OracleNetlist(lc.netlist_tables())
    # start from u0 = [Vcc, 0, 0] as the reference test does (core.jl:803-806)
    u0 = np.array([k["Vcc"], 0.0, 0.0])
    m = {"be": 0, "trap": 1, "gear2": 2}[method]
    dt = tau / (2000 if method == "be" else 200)
    o = ora.make_tran_opts(method=m, dt=dt, init=1)
    r = ora.tran(nl, ora.make_spec(mode="tran"), 0.0, 5 * tau, o, [lc.index_of("out")], u0=u0)
    assert r["status"] == 0
    for mult in k["times_tau"]:
        t = mult * tau
        v = np.interp(t, r["t"], r["u"][:, 0])
        exact = k["Vcc"] * (1 - math.exp(-mult))
        assert v == pytest.approx(exact, rel=k["rtol"], abs=1e-10)


def test_rl_current_closed_form():
    # test/mna/core.jl:914-981: i(t) = V/R (1 - exp(-t R/L)), rtol 1e-3
    lc = lower_one(circuits.rl_circuit(5.0, 100.0, 1e-3))
    nl = ora.OracleNetlist(lc.netlist_tables())
    tau = 1e-3 / 100.0
    u0 = np.zeros(lc.n); u0[0] = 5.0; u0[1] = 5.0
    o = ora.make_tran_opts(method=1, dt=tau / 400, init=1)
    r = ora.tran(nl, ora.make_spec(mode="tran"), 0.0, 5 * tau, o, [lc.index_of("I_L1")], u0=u0)
    for mult in (0.5, 1.0, 2.0, 5.0):
        exact = 5.0 / 100.0 * (1 - math.exp(-mult))
        assert np.interp(mult * tau, r["t"], r["u"][:, 0]) == pytest.approx(exact, rel=1e-3)


def test_tran_dc_initialised_and_tranop_mode():
    # CedarTranOp (dcop.jl:160-203): sources take their waveform value at t0 in :tranop mode
    lc = lower_one(circuits.clipper, R=1e3, C=1e-9)
    nl = ora.OracleNetlist(lc.netlist_tables())
    o = ora.make_tran_opts(method=0, dt=1e-6)
    r = ora.tran(nl, ora.make_spec(mode="tran"), 0.0, 2e-3, o, [1, 2])
    assert r["status"] == 0 and len(r["t"]) == 2001
    vin, vout = r["u"][:, 0], r["u"][:, 1]
    assert vin[250] == pytest.approx(5.0, rel=1e-12)         # sin peak at t = 0.25 ms
    assert 0.55 < vout.max() < 0.80                           # clipped at a diode drop
    assert vout.min() < -4.0                                  # negative half-wave passes
    assert r["newton_iters"] > 2000


# ---- variable-order BDF (the IDA family, sweeps.jl:599-601) --------------------------------------
def _rc_step(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
    ctx = MNAContext() if ctx is None else ctx
    stamp(VoltageSource(0.0, tran=cb.PulseWave(0.0, 5.0, 1e-4, 1e-9, 1e-9, 1.0, 2.0), name="V1"), ctx, "in", 0)
    stamp(Resistor(1e3, name="R1"), ctx, "in", "out")
    stamp(Capacitor(1e-6, name="C1"), ctx, "out", 0)
    return ctx


def test_bdf_controller_known_answer_and_order_raising():
    """Oracle's variable-order BDF (orders 1..5, fixed-leading-coefficient form): RC step response against
    the closed form, breakpoints hit exactly, and -- what makes it variable ORDER -- far fewer steps than the
    second-order trapezoidal controller at a tight tolerance, with a smaller error."""
    lc = lower_one(_rc_step)
    nl = ora.OracleNetlist(lc.netlist_tables())
    save = [lc.index_of("out")]
    res = {}
    for method in (1, 3):
        o = ora.make_tran_opts(method=method, adaptive=1, dt=1e-7, reltol=1e-7, lte_abstol=1e-10, max_points=200000)
        r = ora.sweep_tran(nl, ora.make_spec(mode="tran"), 0.0, 5e-3, o, save)
        assert r["status"][0] == 0
        T = int(r["T"][0])
        t, u = r["t"][0, :T], r["u"][0, :T, 0]
        assert t[0] == 0.0 and t[-1] == 5e-3 and np.all(np.diff(t) > 0)
        for stop in (1e-4, 1e-4 + 1e-9):                      # PULSE corners (solve.jl:1847-1918)
            assert np.min(np.abs(t - stop)) == 0.0
        after = t >= 1e-4 + 1e-9
        exact = 5.0 * (1.0 - np.exp(-(t[after] - (1e-4 + 0.5e-9)) / 1e-3))   # ramp of 1 ns ~ a step at its midpoint
        res[method] = (T, float(np.max(np.abs(u[after] - exact))))
    (T_trap, e_trap), (T_bdf, e_bdf) = res[1], res[3]
    print(f"reltol 1e-7: trapezoid {T_trap} points, max error {e_trap:.2e}; BDF(1..5) {T_bdf} points, max error {e_bdf:.2e}")
    assert e_bdf < 5e-6 and e_trap < 5e-5
    assert T_bdf < 0.6 * T_trap
    # loose tolerance: still converges to the same curve
    o = ora.make_tran_opts(method=3, adaptive=1, dt=1e-7, reltol=1e-3, lte_abstol=1e-6, max_points=20000)
    r = ora.sweep_tran(nl, ora.make_spec(mode="tran"), 0.0, 5e-3, o, save)
    T = int(r["T"][0])
    assert r["status"][0] == 0 and T < T_bdf
    t, u = r["t"][0, :T], r["u"][0, :T, 0]
    after = t >= 1e-4 + 1e-9
    assert np.max(np.abs(u[after] - 5.0 * (1.0 - np.exp(-(t[after] - (1e-4 + 0.5e-9)) / 1e-3)))) < 2e-2


def test_adaptive_negative_start_time_with_breakpoint_at_t0():
    """A transient that starts at a NEGATIVE time with a source breakpoint exactly at t0: the breakpoint must be
    skipped (4 eps(t) dedup, either sign of t), not turned into a zero-length step (h = 0 -> gamma = inf -> NaN)."""
    def b(params, spec, t=0.0, x=ZERO_VECTOR, ctx=None):
        ctx = MNAContext() if ctx is None else ctx
        stamp(VoltageSource(0.0, tran=cb.PulseWave(0.0, 2.0, -1e-6, 1e-7, 1e-7, 4e-7, 1e-6), name="V1"), ctx, "in", 0)
        stamp(Resistor(1e3, name="R1"), ctx, "in", "out")
        stamp(Capacitor(1e-10, name="C1"), ctx, "out", 0)
        return ctx
    lc = lower_one(b)
    nl = ora.OracleNetlist(lc.netlist_tables())
    for method in (1, 3):                                     # trapezoidal + LTE, variable-order BDF
        o = ora.make_tran_opts(method=method, adaptive=1, dt=1e-9, reltol=1e-5, lte_abstol=1e-8, max_points=20000)
        r = ora.sweep_tran(nl, ora.make_spec(mode="tran"), -1e-6, 1e-6, o, [lc.index_of("in"), lc.index_of("out")])
        T = int(r["T"][0])
        t, u = r["t"][0, :T], r["u"][0, :T]
        assert r["status"][0] == 0 and np.all(np.isfinite(u))
        assert t[0] == -1e-6 and t[-1] == 1e-6 and np.all(np.diff(t) > 0)
        for stop in (-1e-6 + 1e-7, -1e-6 + 5e-7, -1e-6 + 6e-7, 0.0, 1e-7):     # corners, negative and positive
            assert np.min(np.abs(t - stop)) <= 4e-16 * max(abs(stop), 1e-6)
        assert u[:, 0].max() == pytest.approx(2.0, abs=1e-12) and 1.5 < u[:, 1].max() <= 2.0 + 1e-9
