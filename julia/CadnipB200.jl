# CadnipB200.jl -- the Julia side of the drop-in boundary (SURVEY.md 8(b), 8(f)1).
#
# A maintainer of NyanCAD/Cadnip.jl adds this file next to src/sweeps.jl and `include`s it from
# src/Cadnip.jl.  It gives `dc!` / `tran!` a `backend = B200(...)` keyword; everything below the
# `ccall`s is include/cadnip_b200.h (libcadnip_b200.so).  Julia is not available in the image this
# repository is built in, so this file has NOT been executed here: the Python ctypes binding
# (cadnip.jl_b200/backend.py) drives exactly the same entry points with exactly the same struct
# layouts in every test, and the layouts below are checked against the header by
# tests/test_host.py::test_julia_shim_matches_header.
#
# Three parts:
#   1. mirrors of the C structs + thin ccall wrappers;
#   2. `export_device_table`: the one thing the reference does not already hold -- a device-table
#      row per `stamp!` call -- obtained by running the builder against a RECORDING context;
#   3. the `dc!` / `tran!` methods (src/sweeps.jl:450-454, :511-532, :588-601, :692-707).
module CadnipB200

using ..Cadnip
using ..Cadnip.MNA
using ..Cadnip.MNA: MNAContext, MNASpec, MNACircuit, resolve_index, system_size, ZERO_VECTOR
using ..Cadnip: CircuitSweep, SweepResult

const LIB = Ref{String}("libcadnip_b200")            # set to an absolute path if not on the loader path

# ----------------------------------------------------------------------------------------------
# 1. C ABI mirrors (field order and padding as in include/cadnip_b200.h)
# ----------------------------------------------------------------------------------------------
struct Spec                                           # cb200_spec
    temp::Cdouble; mode::Int32; _pad::Int32
    gmin::Cdouble; gshunt::Cdouble; srcFact::Cdouble; tnom::Cdouble
    abstol::Cdouble; reltol::Cdouble; vntol::Cdouble; iabstol::Cdouble
end
const MODES = Dict(:dcop => 0, :tran => 1, :tranop => 2, :ac => 3)
Spec(s::MNASpec; mode = s.mode) = Spec(s.temp, MODES[mode], 0, s.gmin, s.gshunt, s.srcFact, s.tnom,
                                       s.abstol, s.reltol, s.vntol, s.iabstol)

struct DcOpts                                         # cb200_dc_opts
    abstol::Cdouble; maxiters::Int32; use_stepping::Int32
end

struct TranOpts                                       # cb200_tran_opts
    method::Int32; adaptive::Int32
    dt::Cdouble; abstol::Cdouble; reltol::Cdouble; lte_abstol::Cdouble; dtmin::Cdouble; dtmax::Cdouble
    max_nl_iters::Int32; save_every::Int32; max_points::Int32; init::Int32
    init_abstol::Cdouble; init_maxiters::Int32; flags::Int32
    uic_steps::Int32; _pad::Int32; uic_dt::Cdouble
    vntol::Cdouble; iabstol::Cdouble; chgtol::Cdouble
end
const TRAN_LIMIT = Int32(1)
const TRAN_CLASS_ABSTOL = Int32(4)

struct Desc                                           # cb200_desc
    n_nodes::Int32; n_currents::Int32; n_charges::Int32; n_limits::Int32
    nG::Int64; nC::Int64; nb::Int64
    G_I::Ptr{Int64}; G_J::Ptr{Int64}; C_I::Ptr{Int64}; C_J::Ptr{Int64}; b_I::Ptr{Int64}
    n_devices::Int32; n_uniform::Int32
    dev_kind::Ptr{Int32}; dev_flags::Ptr{Int32}; dev_node_ptr::Ptr{Int32}; dev_nodes::Ptr{Int32}
    dev_param_ptr::Ptr{Int32}; dev_params::Ptr{Int32}
    dev_gbase::Ptr{Int64}; dev_cbase::Ptr{Int64}; dev_bbase::Ptr{Int64}
    uniform::Ptr{Cdouble}; limit_init_ref::Ptr{Int32}
    n_lane_cols::Int32; _pad::Int32
    dev_state_ptr::Ptr{Int32}
end

# CB200_DEV_* (include/cadnip_b200.h)
const DEV = (resistor = 1, capacitor = 2, inductor = 3, vsource = 4, isource = 5, vcvs = 6, vccs = 7,
             ccvs = 8, cccs = 9, diode = 10, diodecap = 11, simplemos = 12, va = 13)

last_error(h) = unsafe_string(ccall((:cb200_last_error, LIB[]), Cstring, (Ptr{Cvoid},), h))
check(h, rc) = rc == 0 ? nothing : error("cadnip-b200: " * last_error(h))      # usage errors throw

mutable struct Handle
    ptr::Ptr{Cvoid}
    n::Int
    P::Int
    keep::Any                                          # arrays the Desc pointed at during create
    function Handle(ptr, n, keep)
        h = new(ptr, n, 0, keep)
        finalizer(x -> (x.ptr == C_NULL || ccall((:cb200_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
    end
end

function set_lanes!(h::Handle, soa::Matrix{Float64})  # soa[lane, column]: column-major == [col][lane]
    GC.@preserve soa check(h.ptr, ccall((:cb200_set_lanes, LIB[]), Cint,
        (Ptr{Cvoid}, Int64, Int32, Ptr{Cdouble}), h.ptr, size(soa, 1), size(soa, 2), soa))
    h.P = size(soa, 1)
    return h
end

function load_va_models!(h::Handle, header::String, csrc_dir::String, cache_dir::String)
    check(h.ptr, ccall((:cb200_load_va_models, LIB[]), Cint, (Ptr{Cvoid}, Cstring, Cstring, Cstring),
                       h.ptr, header, csrc_dir, cache_dir))
end

function dc(h::Handle, spec::MNASpec; abstol = 1e-10, maxiters = 100, u0 = nothing, mode = :dcop)
    x = Matrix{Float64}(undef, h.P, h.n); status = Vector{Int32}(undef, h.P); iters = similar(status)
    s = Ref(Spec(spec; mode)); o = Ref(DcOpts(abstol, maxiters, 1))
    GC.@preserve x status iters u0 check(h.ptr, ccall((:cb200_dc, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Spec}, Ptr{DcOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int32}),
        h.ptr, s, o, u0 === nothing ? C_NULL : pointer(u0), x, status, iters))
    return x, status, iters                            # x[lane, unknown]
end

# lanes the static pivot order served badly (cb200_weak_pivot_lanes): re-solve them on a handle of their own
function weak_pivot_lanes(h::Handle)
    f = Vector{Int32}(undef, h.P)
    GC.@preserve f check(h.ptr, ccall((:cb200_weak_pivot_lanes, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}), h.ptr, f, C_NULL))
    return findall(!iszero, f)
end

function tran(h::Handle, spec::MNASpec, tspan, o::TranOpts, save_idx::Vector{Int64}; u0 = nothing)
    w = Ref{Ptr{Cvoid}}(C_NULL); s = Ref(Spec(spec; mode = :tran)); oo = Ref(o)
    GC.@preserve save_idx u0 check(h.ptr, ccall((:cb200_tran, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Spec}, Cdouble, Cdouble, Ptr{TranOpts}, Ptr{Int64}, Int32, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}),
        h.ptr, s, tspan[1], tspan[2], oo, save_idx, length(save_idx), u0 === nothing ? C_NULL : pointer(u0), w))
    T = Ref{Int64}(0); P = Ref{Int64}(0); ns = Ref{Int32}(0); ad = Ref{Int32}(0)
    ccall((:cb200_wave_info, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}), w[], T, P, ns, ad)
    adaptive = ad[] != 0
    t = adaptive ? Matrix{Float64}(undef, P[], T[]) : Vector{Float64}(undef, T[])      # [lane, point] when ragged
    u = Array{Float64,3}(undef, P[], T[], ns[])                                        # u[lane, point, save]
    count = Vector{Int32}(undef, P[]); status = similar(count); iters = similar(count)
    GC.@preserve t u count status iters check(h.ptr, ccall((:cb200_wave_fetch, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), w[], t, u, count, status, iters))
    ccall((:cb200_wave_free, LIB[]), Cvoid, (Ptr{Cvoid},), w[])
    return (; t, u, count, status, iters, adaptive)
end

# ----------------------------------------------------------------------------------------------
# 2. export_device_table: builder -> cb200_desc + lane parameter matrix
#
# The structure (COO coordinates, unknown counts, limit_init) is what build_with_detection already
# leaves in an MNAContext (src/mna/solve.jl:1793-1822).  The device table is recorded by wrapping the
# context: `RecordingContext` forwards every MNAContext primitive and, around each `stamp!` of a device
# it knows, notes (kind, flags, node indices, parameter references, first G / C / b stamp slot).  It
# needs `RecordingContext` added to the `AnyMNAContext` union (src/mna/value_only.jl:714) and -- for
# generated Verilog-A `stamp!` methods, which are `invoke`d with a concrete context type
# (src/spc/codegen.jl:856-863) -- the same widening in `vasim.jl:3966`.  The alternative with no change
# to dispatch is to emit the row in `cg_mna_instance!` (src/spc/codegen.jl:917-2527) next to the call.
#
# A parameter is either uniform over the sweep (reference >= 0 into `uniform`) or a lane column
# (reference ~col): `lane_of(value)` tells which, by looking the value up among the swept quantities.
# ----------------------------------------------------------------------------------------------
mutable struct DeviceTable
    kind::Vector{Int32}; flags::Vector{Int32}
    node_ptr::Vector{Int32}; nodes::Vector{Int32}
    param_ptr::Vector{Int32}; params::Vector{Int32}
    gbase::Vector{Int64}; cbase::Vector{Int64}; bbase::Vector{Int64}
    uniform::Vector{Float64}
    lane_cols::Vector{Vector{Float64}}                 # one vector of length P per swept quantity in use
    limit_init_ref::Vector{Int32}
end
DeviceTable() = DeviceTable(Int32[], Int32[], Int32[0], Int32[], Int32[0], Int32[], Int64[], Int64[], Int64[],
                            Float64[], Vector{Float64}[], Int32[])

"parameter reference of a value: a lane column when it is one of the swept vectors, else a uniform slot"
function param_ref!(tab::DeviceTable, v)
    if v isa AbstractVector                            # the builder was run with lane-array parameters
        col = findfirst(c -> c === v || c == v, tab.lane_cols)
        col === nothing && (push!(tab.lane_cols, collect(Float64, v)); col = length(tab.lane_cols))
        return Int32(~(col - 1))
    end
    push!(tab.uniform, Float64(v))
    return Int32(length(tab.uniform) - 1)
end

"record one device row; call immediately BEFORE forwarding the stamp! to the wrapped context"
function record!(tab::DeviceTable, ctx::MNAContext, kind, flags, nodes, params)
    push!(tab.kind, kind); push!(tab.flags, flags)
    append!(tab.nodes, Int32.(nodes)); push!(tab.node_ptr, length(tab.nodes))
    append!(tab.params, [param_ref!(tab, p) for p in params]); push!(tab.param_ptr, length(tab.params))
    push!(tab.gbase, length(ctx.G_I)); push!(tab.cbase, length(ctx.C_I)); push!(tab.bbase, length(ctx.b_I))
    return nothing
end

# Per-device recorders for the primitives of src/mna/devices.jl (node indices are the resolved unknown
# indices, 0 = ground; a current / limit / charge unknown allocated by the stamp is appended after the
# stamp has run -- see `finish!`).  Parameter order = what csrc/lane_kernels.cuh: eval_device reads.
recorder(d::MNA.Resistor) = (DEV.resistor, 0, (d.r,))
recorder(d::MNA.Capacitor) = (DEV.capacitor, 0, (d.c,))
recorder(d::MNA.Inductor) = (DEV.inductor, 0, (d.l,))
recorder(d::MNA.VCVS) = (DEV.vcvs, 0, (d.gain,))
recorder(d::MNA.VCCS) = (DEV.vccs, 0, (d.gm,))
recorder(d::MNA.Diode) = (DEV.diode, d.limit ? 1 : 0, (d.Is, d.Vt, d.n, MNA.diode_vcrit(d)))
recorder(d::MNA.DiodeWithCap) = (DEV.diodecap, 0, (d.Is, d.Vt, d.n, d.Cj0, d.Vj, d.m))
recorder(d::MNA.SimpleMOSFET) = (DEV.simplemos, 0, (d.Vth, d.K, d.lambda, d.Cgd, d.Cgs))
# sources: flags = waveform kind (0 dc, 1 PWL, 2 PULSE, 3 SIN), params = dc value then the waveform's
recorder(d::MNA.VoltageSource) = (DEV.vsource, wave_kind(d.tran), (d.dc, wave_params(d.tran)...))
recorder(d::MNA.CurrentSource) = (DEV.isource, wave_kind(d.tran), (d.dc, wave_params(d.tran)...))
wave_kind(::Nothing) = 0; wave_params(::Nothing) = ()
wave_kind(::MNA.PWLWave) = 1; wave_params(w::MNA.PWLWave) = (w.ts..., w.ys...)
wave_kind(::MNA.PulseWave) = 2; wave_params(w::MNA.PulseWave) = (w.v1, w.v2, w.td, w.tr, w.tf, w.pw, w.per)
wave_kind(::MNA.SinWave) = 3; wave_params(w::MNA.SinWave) = (w.vo, w.va, w.freq, w.td, w.theta, w.phase)

# Behavioural sources (src/mna/devices.jl:1003-1131).  The closure `value_fn(get_voltage)` cannot be called from
# the device, so it is TRACED once: `get_voltage(name)` returns a `SymV` whose arithmetic builds Verilog-A
# text; the text becomes the one-statement module `V(p,n) <+ $explicit(expr)` / `I(n,p) <+ $explicit(expr)` that
# goes to the emitter like any other model (INTEGRATION.md section 5b; cadnip_b200/behavioral.py is the Python
# statement of the same tracer).  `$explicit` = value without partials: the reference's b-only stamping.
struct SymV
    text::String
end
const BehaviouralTrace = Vector{Symbol}                 # controlling nodes in order of first use
sym(x::SymV) = x
sym(x::Real) = SymV("(" * repr(Float64(x)) * ")")
for (op, va) in ((:+, "+"), (:-, "-"), (:*, "*"), (:/, "/"))
    @eval Base.$op(a::SymV, b::SymV) = SymV("(" * a.text * " " * $va * " " * b.text * ")")
    @eval Base.$op(a::SymV, b::Real) = $op(a, sym(b))
    @eval Base.$op(a::Real, b::SymV) = $op(sym(a), b)
end
Base.:-(a::SymV) = SymV("(-" * a.text * ")")
Base.:^(a::SymV, b::Real) = SymV("pow(" * a.text * ", " * sym(b).text * ")")
Base.:^(a::SymV, b::SymV) = SymV("pow(" * a.text * ", " * b.text * ")")
for (fn, va) in ((:exp, "exp"), (:log, "ln"), (:log10, "log"), (:sqrt, "sqrt"), (:sin, "sin"), (:cos, "cos"),
                 (:tan, "tan"), (:tanh, "tanh"), (:sinh, "sinh"), (:cosh, "cosh"), (:atan, "atan"), (:abs, "abs"))
    @eval Base.$fn(a::SymV) = SymV($va * "(" * a.text * ")")
end
Base.min(a::SymV, b) = SymV("min(" * a.text * ", " * sym(b).text * ")")
Base.max(a::SymV, b) = SymV("max(" * a.text * ", " * sym(b).text * ")")
Base.:<(a::SymV, b) = SymV("(" * a.text * " < " * sym(b).text * ")")
Base.:>(a::SymV, b) = SymV("(" * a.text * " > " * sym(b).text * ")")
Base.ifelse(c::SymV, a, b) = SymV("((" * c.text * ") ? " * sym(a).text * " : " * sym(b).text * ")")

"""
    trace_behavioral(B) -> (va_source, controlling_nodes)

The module text for a `BehavioralVoltageSource` / `BehavioralCurrentSource` and the nodes its value reads
(they become the extra ports c0, c1, ... of the module, after p and n).
"""
function trace_behavioral(B::Union{MNA.BehavioralVoltageSource, MNA.BehavioralCurrentSource})
    nodes = BehaviouralTrace()
    function get_voltage(name)
        i = findfirst(==(Symbol(name)), nodes)
        i === nothing && (push!(nodes, Symbol(name)); i = length(nodes))
        SymV("V(c$(i - 1))")
    end
    expr = sym(B.value_fn(get_voltage)).text
    ports = join(vcat(["p", "n"], ["c$(k - 1)" for k in 1:length(nodes)]), ", ")
    contrib = B isa MNA.BehavioralVoltageSource ? "V(p, n) <+ \$explicit($expr);" : "I(n, p) <+ \$explicit($expr);"
    name = "bsrc_" * string(hash(expr), base = 16)
    src = "module $name($ports);\n    inout $ports;\n    electrical $ports;\n    analog begin\n        $contrib\n    end\nendmodule\n"
    return src, nodes
end

"""
    export_device_table(circuit_or_sweep) -> (Desc, keepalive, lane_matrix, names)

Runs `build_with_detection` once with the swept parameters bound to lane VECTORS (the structure must not
depend on them: `StructuralSweepError` otherwise, as in lowering.py), records the device rows, resolves the
typed indices and packs everything into a `cb200_desc`.
"""
function export_device_table(cs::CircuitSweep)
    P = length(cs)
    lane_params = sweep_columns(cs)                     # NamedTuple of Vector{Float64}, product order (sweeps.jl:272)
    tab = DeviceTable()
    ctx = MNA.build_with_detection(cs.circuit.builder, merge(cs.circuit.params, lane_params), cs.circuit.spec;
                                   on_stamp = (dev, ctx, nodes) -> begin
                                       kind, flags, params = recorder(dev)
                                       record!(tab, ctx, kind, flags, nodes, params)
                                   end)
    res(v) = Int64[resolve_index(ctx, i) for i in v]
    keep = (G_I = res(ctx.G_I), G_J = res(ctx.G_J), C_I = res(ctx.C_I), C_J = res(ctx.C_J), b_I = res(ctx.b_I),
            tab = tab, limit_ref = Int32[param_ref!(tab, v) for v in ctx.limit_init])
    push!(tab.gbase, length(ctx.G_I)); push!(tab.cbase, length(ctx.C_I)); push!(tab.bbase, length(ctx.b_I))
    desc = Desc(ctx.n_nodes, ctx.n_currents, ctx.n_charges, ctx.n_limits,
                length(keep.G_I), length(keep.C_I), length(keep.b_I),
                pointer(keep.G_I), pointer(keep.G_J), pointer(keep.C_I), pointer(keep.C_J), pointer(keep.b_I),
                length(tab.kind), length(tab.uniform),
                pointer(tab.kind), pointer(tab.flags), pointer(tab.node_ptr), pointer(tab.nodes),
                pointer(tab.param_ptr), pointer(tab.params), pointer(tab.gbase), pointer(tab.cbase), pointer(tab.bbase),
                pointer(tab.uniform), pointer(keep.limit_ref), length(tab.lane_cols), 0, C_NULL)
    lanes = isempty(tab.lane_cols) ? zeros(P, 0) : reduce(hcat, tab.lane_cols)        # [lane, column]
    names = (nodes = ctx.node_names, currents = ctx.current_names, charges = ctx.charge_names, limits = ctx.limit_names)
    return desc, keep, lanes, names
end

"swept quantities as lane vectors in `Iterators.product` order (first axis fastest, src/sweeps.jl:272)"
function sweep_columns(cs::CircuitSweep)
    pts = collect(cs.iterator)
    ks = keys(first(pts))
    return NamedTuple{ks}(Tuple([Float64(p[k]) for p in pts] for k in ks))
end

function create(cs::CircuitSweep; device::Integer = 0)
    desc, keep, lanes, names = export_device_table(cs)
    ptr = Ref{Ptr{Cvoid}}(C_NULL); d = Ref(desc)
    rc = GC.@preserve keep ccall((:cb200_create, LIB[]), Cint, (Ptr{Desc}, Int32, Ptr{Ptr{Cvoid}}), d, device, ptr)
    rc == 0 || error("cadnip-b200: " * last_error(C_NULL))
    h = Handle(ptr[], desc.n_nodes + desc.n_currents + desc.n_charges + desc.n_limits, keep)
    set_lanes!(h, lanes)
    # the library rebuilt pattern and maps itself: a host that wants to assert equality with its own
    # CompiledStructure reads them back with cb200_get_pattern / cb200_get_maps (1-based, Julia layout)
    return h, names
end

# ----------------------------------------------------------------------------------------------
# 3. drop-in methods
# ----------------------------------------------------------------------------------------------
struct B200
    device::Int
end
B200() = B200(0)

const RETCODES = Dict(0 => :Success, 1 => :MaxIters, 2 => :Unstable, 3 => :Unstable, 4 => :DtLessThanMin)
# 0 backward Euler, 1 trapezoidal, 2 Gear-2; 3 = the variable-order (1..5) variable-step BDF controller for the
# IDA / FBDF / QNDF family (adaptive only; `nothing` = tran!'s default Sundials.IDA, src/sweeps.jl:599-601)
method_id(solver) = solver === nothing ? Int32(3) :
                    solver isa Cadnip.ImplicitEuler ? Int32(0) : solver isa Cadnip.Trapezoid ? Int32(1) :
                    (occursin("IDA", string(typeof(solver))) || occursin("BDF", string(typeof(solver))) ||
                     occursin("QNDF", string(typeof(solver)))) ? Int32(3) : Int32(2)

"`dc!(cs::CircuitSweep; backend = B200())`: all lanes at once, cold start (sweeps.jl:511-532)"
function Cadnip.dc!(cs::CircuitSweep; backend::B200, abstol = 1e-10, maxiters = 100, continuation = true)
    h, names = create(cs; device = backend.device)
    x, status, iters = dc(h, cs.circuit.spec; abstol, maxiters)
    sols = [MNA.DCSolution(vec(x[p, :]), names.nodes, names.currents, names.charges, names.limits, status[p] == 0)
            for p in 1:length(cs)]
    return SweepResult(collect(cs.iterator), sols)
end

"`tran!(cs::CircuitSweep, tspan; backend = B200(), ...)` (sweeps.jl:692-707)"
function Cadnip.tran!(cs::CircuitSweep, tspan::Tuple{<:Real,<:Real}; backend::B200,
                      solver = Cadnip.Trapezoid(), abstol = 1e-10, reltol = 1e-8, dt = nothing, adaptive = dt === nothing,
                      saveat = nothing, save_idxs = nothing, max_points = 4096, initializealg = nothing)
    h, names = create(cs; device = backend.device)
    n = h.n
    save = save_idxs === nothing ? collect(Int64, 1:n) : Int64[unknown_index(names, s) for s in save_idxs]
    cls = abstol isa NamedTuple                        # (vntol = ..., iabstol = ..., chgtol = ...), sweeps.jl:615-618
    vn, ia, cq = cls ? (get(abstol, :vntol, 1e-6), get(abstol, :iabstol, 1e-12), get(abstol, :chgtol, 1e-14)) : (0.0, 0.0, 0.0)
    uic = initializealg isa Cadnip.MNA.CedarUICOp
    o = TranOpts(method_id(solver), adaptive ? 1 : 0, dt === nothing ? (tspan[2] - tspan[1]) * 1e-4 : dt,
                 cls ? 1e-10 : abstol, reltol, cls ? 1e-10 : max(abstol, 1e-12), 0.0, 0.0,
                 10, saveat === nothing || adaptive ? 1 : max(1, round(Int, saveat / dt)), max_points, uic ? 2 : 0,
                 1e-9, 500, cls ? TRAN_CLASS_ABSTOL : Int32(0),
                 uic ? initializealg.warmup_steps : 10, 0, uic ? initializealg.dt : 1e-12, vn, ia, cq)
    if adaptive                                         # _with_auto_tstops (solve.jl:1948-1960)
        ts = MNA.expand_breakpoints(cs.circuit, tspan)
        GC.@preserve ts check(h.ptr, ccall((:cb200_set_tstops, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int32), h.ptr, ts, length(ts)))
    end
    r = tran(h, cs.circuit.spec, tspan, o, save)
    sols = map(1:length(cs)) do p
        T = r.adaptive ? Int(r.count[p]) : size(r.u, 2)
        tt = r.adaptive ? r.t[p, 1:T] : r.t
        B200Solution(tt, r.u[p, 1:T, :], save, names, RETCODES[Int(r.status[p])], Int(r.iters[p]))
    end
    return SweepResult(collect(cs.iterator), sols)
end

# single circuits go through a one-point sweep
Cadnip.tran!(c::MNACircuit, tspan::Tuple{<:Real,<:Real}; backend::B200, kw...) =
    last(first(Cadnip.tran!(CircuitSweep(c.builder, Cadnip.Sweep(_b200_dummy = [0.0]); c.params...), tspan; backend, kw...)))

"time-series solution with the accessors the reference's results offer: sol[:name], sol(t), sol.t, sol.u, sol.retcode"
struct B200Solution
    t::Vector{Float64}
    u::Matrix{Float64}                                  # [point, saved unknown]
    save::Vector{Int64}
    names::Any
    retcode::Symbol
    newton_iters::Int
end
unknown_index(names, s::Symbol) = something(findfirst(==(s), vcat(names.nodes, names.currents, names.charges, names.limits)))
Base.getindex(s::B200Solution, name::Symbol) = s.u[:, something(findfirst(==(unknown_index(s.names, name)), s.save))]
function (s::B200Solution)(t::Real)                     # linear interpolation between saved points
    k = clamp(searchsortedlast(s.t, t), 1, length(s.t) - 1)
    a = (t - s.t[k]) / (s.t[k + 1] - s.t[k])
    return (1 - a) .* s.u[k, :] .+ a .* s.u[k + 1, :]
end

end # module
