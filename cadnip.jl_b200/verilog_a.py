"""Verilog-A device-model emitter (tier 0).

north_star: "a new emitter lowers each device model's stamp function ... to
hand-written-style sm_100a CUDA C".  This module reads a Verilog-A module (the subset
the reference's inline test models use: parameters, internal nodes, real locals,
assignments, if/else, ``I(a,b) <+ expr`` contributions with ``ddt``), and emits the
module's ``stamp!`` method as straight-line C with ANALYTIC partial derivatives
(forward-mode differentiation at emit time; no dual numbers at run time) -- once as a
CUDA device function for ``lane_kernels.cuh`` and once as plain C for the CPU oracle,
from the same IR so both share one arithmetic order.

The stamping protocol restates ``generate_mna_stamp_method_nterm``
(src/vasim.jl:2993-3985):
  * internal nodes are allocated in declaration order (``alloc_internal_node!``);
  * per current branch, in order of first contribution: ``G[p,k] += dI/dV_k``,
    ``G[n,k] -= dI/dV_k`` for EVERY node k of the module (ports then internals),
    each guarded by node != 0 (vasim.jl:3382-3392);
  * a branch with a ``ddt`` term is reactive BY TYPE (vasim.jl:3307-3308): constant
    capacitance -> plain C stamps (vasim.jl:3474-3485); voltage-dependent charge ->
    charge-state formulation with ``CHARGE_SCALE = 1e12`` (vasim.jl:3433-3472,
    src/mna/contrib.jl:39): the reference decides by probing Q/V numerically
    (contrib.jl:214-257), here the same outcome is derived statically from whether the
    charge's partials depend on a node voltage;
  * ``Ieq = I0 - sum_k dI/dV_k V_k``; ``b[p] -= Ieq``, ``b[n] += Ieq`` (vasim.jl:3490-3518).
``ddt`` follows the s-dual algebra of src/mna/contrib.jl:356-375: ``c * ddt(q)`` has
charge ``c*q`` with the full product rule on the partials.

Not supported in this tier (rejected loudly): ``$limit``, voltage contributions,
node collapse, analog functions, ``$vt`` / ``$temperature``, noise, strings.
"""
from __future__ import annotations

import hashlib
import math
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

CHARGE_SCALE = 1e12          # src/mna/contrib.jl:39


class VAError(ValueError):
    pass


# --------------------------------------------------------------------------- #
# lexer
# --------------------------------------------------------------------------- #
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/) |
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[TGMKkmunpfa]?) |
    (?P<id>[$A-Za-z_`][A-Za-z0-9_$]*) |
    (?P<str>"[^"]*") |
    (?P<op><\+|\*\*|<=|>=|==|!=|&&|\|\||[-+*/%<>=!?:;,()\[\]{}@^~&|#.])
""", re.X | re.S)

_SCALE = {"T": 1e12, "G": 1e9, "M": 1e6, "K": 1e3, "k": 1e3, "m": 1e-3, "u": 1e-6, "n": 1e-9,
          "p": 1e-12, "f": 1e-15, "a": 1e-18}


def _lex(src: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise VAError(f"Verilog-A: cannot tokenise at {src[pos:pos + 20]!r}")
        pos = m.end()
        if m.lastgroup == "ws":
            continue
        out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(("eof", ""))
    return out


def _number(tok: str) -> float:
    if tok[-1] in _SCALE and not tok[-1].isdigit():
        return float(tok[:-1]) * _SCALE[tok[-1]]
    return float(tok)


# --------------------------------------------------------------------------- #
# parser -> tiny AST (tuples)
# --------------------------------------------------------------------------- #
class _Parser:
    def __init__(self, src: str):
        self.t = _lex(src)
        self.i = 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, val):
        if self.t[self.i][1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise VAError(f"Verilog-A: expected {val!r}, got {self.t[self.i][1]!r}")

    def ident(self):
        k, v = self.next()
        if k != "id":
            raise VAError(f"Verilog-A: expected identifier, got {v!r}")
        return v

    # module ------------------------------------------------------------- #
    def module(self):
        self.expect("module")
        name = self.ident()
        ports = []
        if self.accept("("):
            while not self.accept(")"):
                ports.append(self.ident())
                self.accept(",")
        self.expect(";")
        params, electrical, reals, body = [], [], [], None
        while not self.accept("endmodule"):
            k, v = self.peek()
            if v == "parameter":
                self.next()
                if self.peek()[1] in ("real", "integer"):
                    self.next()
                pname = self.ident()
                self.expect("=")
                default = self.expr()
                while not self.accept(";"):          # skip `from [..]` / `exclude` ranges
                    self.next()
                params.append((pname, default))
            elif v in ("inout", "input", "output"):
                self.next()
                while not self.accept(";"):
                    self.next()
            elif v == "electrical":
                self.next()
                while True:
                    electrical.append(self.ident())
                    if self.accept(";"):
                        break
                    self.expect(",")
            elif v in ("real", "integer"):
                self.next()
                while True:
                    reals.append(self.ident())
                    if self.accept("="):
                        raise VAError("Verilog-A: variable initialisers are not supported")
                    if self.accept(";"):
                        break
                    self.expect(",")
            elif v == "analog":
                self.next()
                if self.peek()[1] == "function":
                    raise VAError("Verilog-A: analog functions are not supported in this tier")
                body = self.stmt()
            elif v in ("branch", "ground", "string", "genvar", "aliasparam", "localparam"):
                raise VAError(f"Verilog-A: `{v}` is not supported in this tier")
            elif k == "eof":
                raise VAError("Verilog-A: missing endmodule")
            else:
                raise VAError(f"Verilog-A: unexpected {v!r} in module body")
        if body is None:
            raise VAError("Verilog-A: module has no analog block")
        return dict(name=name, ports=ports, params=params, electrical=electrical, reals=reals, body=body)

    # statements ----------------------------------------------------------- #
    def stmt(self):
        k, v = self.peek()
        if v == "begin":
            self.next()
            if self.accept(":"):
                self.ident()
            stmts = []
            while not self.accept("end"):
                stmts.append(self.stmt())
            return ("block", stmts)
        if v == "if":
            self.next()
            self.expect("(")
            cond = self.expr()
            self.expect(")")
            then = self.stmt()
            other = self.stmt() if self.accept("else") else ("block", [])
            return ("if", cond, then, other)
        if v == ";":
            self.next()
            return ("block", [])
        if v in ("I", "V") and self.t[self.i + 1][1] == "(":
            kind = self.next()[1]
            self.expect("(")
            a = self.ident()
            b = self.ident() if self.accept(",") else None
            self.expect(")")
            self.expect("<+")
            e = self.expr()
            self.expect(";")
            if kind == "V":
                raise VAError("Verilog-A: voltage contributions V(..) <+ are not supported in this tier")
            return ("contrib", a, b, e)
        if k == "id" and v.startswith("$"):
            # system tasks ($strobe, $warning, ...) are no-ops on this path (vasim.jl:1181-1256)
            self.next()
            if self.accept("("):
                depth = 1
                while depth:
                    t = self.next()[1]
                    depth += (t == "(") - (t == ")")
            self.expect(";")
            return ("block", [])
        name = self.ident()
        self.expect("=")
        e = self.expr()
        self.expect(";")
        return ("assign", name, e)

    # expressions ---------------------------------------------------------- #
    def expr(self):
        c = self.lor()
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("?", c, a, b)
        return c

    def _binary(self, sub, ops):
        left = sub()
        while self.peek()[1] in ops:
            op = self.next()[1]
            left = ("bin", op, left, sub())
        return left

    def lor(self):
        return self._binary(self.land, ("||",))

    def land(self):
        return self._binary(self.cmp, ("&&",))

    def cmp(self):
        return self._binary(self.add, ("==", "!=", "<", "<=", ">", ">="))

    def add(self):
        return self._binary(self.mul, ("+", "-"))

    def mul(self):
        return self._binary(self.unary, ("*", "/", "%"))

    def unary(self):
        if self.accept("-"):
            return ("neg", self.unary())
        if self.accept("+"):
            return self.unary()
        if self.accept("!"):
            return ("not", self.unary())
        return self.power()

    def power(self):
        base = self.primary()
        if self.accept("**"):
            return ("call", "pow", [base, self.unary()])
        return base

    def primary(self):
        k, v = self.next()
        if k == "num":
            return ("num", _number(v))
        if v == "(":
            e = self.expr()
            self.expect(")")
            return e
        if k == "id":
            if self.peek()[1] == "(":
                self.next()
                if v in ("V", "I"):
                    a = self.ident()
                    b = self.ident() if self.accept(",") else None
                    self.expect(")")
                    if v == "I":
                        raise VAError("Verilog-A: current probes I(..) are not supported in this tier")
                    return ("V", a, b)
                args = []
                while not self.accept(")"):
                    if self.peek()[0] == "str":
                        args.append(("str", self.next()[1].strip('"')))
                    else:
                        args.append(self.expr())
                    self.accept(",")
                return ("call", v, args)
            return ("var", v)
        raise VAError(f"Verilog-A: unexpected token {v!r} in expression")


# --------------------------------------------------------------------------- #
# forward-mode differentiation at emit time
# --------------------------------------------------------------------------- #
@dataclass
class _D:
    """value + partials w.r.t. the module's node voltages; strings are C expressions,
    ``None`` partial = structurally zero; ``const`` = Python float when known."""
    v: str
    d: List[Optional[str]]
    const: Optional[float] = None

    def is_const(self):
        return self.const is not None


def _lit(x: float) -> str:
    if x != x:
        return "(0.0/0.0)"
    if x in (float("inf"), float("-inf")):
        return "(1.0/0.0)" if x > 0 else "(-1.0/0.0)"
    r = repr(float(x))
    if "e" not in r and "." not in r:
        r += ".0"
    return "(" + r + ")" if x < 0 else r


class _Emitter:
    def __init__(self, mod: dict, params: List[str], nodes: List[str]):
        self.mod, self.params, self.nodes = mod, params, nodes
        self.N = len(nodes)
        self.lines: List[str] = []
        self.tmp = 0
        self.indent = 1
        self.vars = set(mod["reals"])
        self.assigned: set = set()
        # branch accumulators, in order of first contribution
        self.branches: List[Tuple[str, Optional[str]]] = []
        self.reactive: Dict[int, bool] = {}
        self.vdep: Dict[int, bool] = {}

    # -- helpers ------------------------------------------------------------ #
    def emit(self, s: str):
        self.lines.append("    " * self.indent + s)

    def temp(self, expr: str) -> str:
        self.tmp += 1
        name = f"t{self.tmp}"
        self.emit(f"const double {name} = {expr};")
        return name

    def const(self, x: float) -> _D:
        return _D(_lit(x), [None] * self.N, float(x))

    def zero(self):
        return self.const(0.0)

    def _padd(self, a, b, sign="+"):
        if a is None and b is None:
            return None
        if b is None:
            return a
        if a is None:
            return b if sign == "+" else self.temp(f"-{b}")
        return self.temp(f"{a} {sign} {b}")

    def _pscale(self, p, s: str):
        return None if p is None else self.temp(f"{p} * {s}")

    # -- arithmetic on duals -------------------------------------------------- #
    def add(self, a: _D, b: _D, sign="+") -> _D:
        if a.is_const() and b.is_const():
            return self.const(a.const + b.const if sign == "+" else a.const - b.const)
        if b.is_const() and b.const == 0.0:
            return a
        if a.is_const() and a.const == 0.0 and sign == "+":
            return b
        v = self.temp(f"{a.v} {sign} {b.v}")
        return _D(v, [self._padd(x, y, sign) for x, y in zip(a.d, b.d)])

    def neg(self, a: _D) -> _D:
        if a.is_const():
            return self.const(-a.const)
        return _D(self.temp(f"-{a.v}"), [None if p is None else self.temp(f"-{p}") for p in a.d])

    def mul(self, a: _D, b: _D) -> _D:
        if a.is_const() and b.is_const():
            return self.const(a.const * b.const)
        for x, y in ((a, b), (b, a)):
            if x.is_const() and x.const == 1.0:
                return y
            if x.is_const() and x.const == 0.0:
                return self.zero()
        v = self.temp(f"{a.v} * {b.v}")
        d = []
        for pa, pb in zip(a.d, b.d):
            t1 = None if pa is None else f"{pa} * {b.v}"
            t2 = None if pb is None else f"{a.v} * {pb}"
            if t1 is None and t2 is None:
                d.append(None)
            elif t2 is None:
                d.append(self.temp(t1))
            elif t1 is None:
                d.append(self.temp(t2))
            else:
                d.append(self.temp(f"{t1} + {t2}"))
        return _D(v, d)

    def div(self, a: _D, b: _D) -> _D:
        if a.is_const() and b.is_const() and b.const != 0.0:
            return self.const(a.const / b.const)
        v = self.temp(f"{a.v} / {b.v}")
        if all(p is None for p in b.d):
            return _D(v, [None if p is None else self.temp(f"{p} / {b.v}") for p in a.d])
        d = []
        for pa, pb in zip(a.d, b.d):        # (a/b)' = (a' - (a/b) b') / b
            if pa is None and pb is None:
                d.append(None)
            elif pb is None:
                d.append(self.temp(f"{pa} / {b.v}"))
            elif pa is None:
                d.append(self.temp(f"-({v} * {pb}) / {b.v}"))
            else:
                d.append(self.temp(f"({pa} - {v} * {pb}) / {b.v}"))
        return _D(v, d)

    def unary_fn(self, a: _D, val_expr: str, dfac_expr) -> _D:
        """f(a): value val_expr, derivative factor dfac_expr(valname) (a C expression)."""
        v = self.temp(val_expr)
        if all(p is None for p in a.d):
            return _D(v, [None] * self.N)
        fac = self.temp(dfac_expr(v))
        return _D(v, [self._pscale(p, fac) for p in a.d])

    def select(self, c: str, a: _D, b: _D) -> _D:
        v = self.temp(f"({c}) ? {a.v} : {b.v}")
        d = []
        for pa, pb in zip(a.d, b.d):
            if pa is None and pb is None:
                d.append(None)
            else:
                d.append(self.temp(f"({c}) ? {pa or '0.0'} : {pb or '0.0'}"))
        return _D(v, d)

    # -- expressions: returns (resistive, reactive-or-None) ------------------- #
    def ev(self, e) -> Tuple[_D, Optional[_D]]:
        kind = e[0]
        if kind == "num":
            return self.const(e[1]), None
        if kind == "var":
            name = e[1]
            if name in self.params:
                return _D(f"p_{name}", [None] * self.N), None
            if name in self.vars:
                if name not in self.assigned:
                    return self.zero(), None           # VA reals start at 0
                return _D(f"v_{name}", [f"v_{name}_d{k}" for k in range(self.N)]), None
            if name in ("$mfactor",):
                return self.const(1.0), None
            if name == "$abstime":
                return _D("t", [None] * self.N), None
            raise VAError(f"Verilog-A: unknown identifier {name!r}")
        if kind == "V":
            a, b = e[1], e[2]
            d: List[Optional[str]] = [None] * self.N
            ia = self._node(a)
            if b is None:
                d[ia] = "1.0"
                return _D(f"V{ia}", d), None
            ib = self._node(b)
            if ia == ib:
                return self.zero(), None
            d[ia] = "1.0"
            d[ib] = "(-1.0)"
            return _D(self.temp(f"V{ia} - V{ib}"), d), None
        if kind == "neg":
            r, q = self.ev(e[1])
            return self.neg(r), (None if q is None else self.neg(q))
        if kind == "not":
            r, q = self.ev(e[1])
            self._no_react(q, "!")
            return _D(self.temp(f"({r.v} == 0.0) ? 1.0 : 0.0"), [None] * self.N), None
        if kind == "?":
            c, cq = self.ev(e[1])
            self._no_react(cq, "?:")
            ar, aq = self.ev(e[2])
            br, bq = self.ev(e[3])
            cond = f"{c.v} != 0.0"
            q = None
            if aq is not None or bq is not None:
                q = self.select(cond, aq or self.zero(), bq or self.zero())
            return self.select(cond, ar, br), q
        if kind == "bin":
            op = e[1]
            ar, aq = self.ev(e[2])
            br, bq = self.ev(e[3])
            if op in ("+", "-"):
                q = None
                if aq is not None or bq is not None:
                    q = self.add(aq or self.zero(), bq or self.zero(), op)
                return self.add(ar, br, op), q
            if op == "*":
                if aq is not None and bq is not None:
                    raise VAError("Verilog-A: product of two ddt() terms")
                q = None
                if aq is not None:
                    q = self.mul(aq, br)
                elif bq is not None:
                    q = self.mul(ar, bq)
                return self.mul(ar, br), q
            if op == "/":
                self._no_react(bq, "/")
                q = None if aq is None else self.div(aq, br)
                return self.div(ar, br), q
            self._no_react(aq, op)
            self._no_react(bq, op)
            if op == "%":
                return _D(self.temp(f"fmod({ar.v}, {br.v})"), list(ar.d)), None
            cop = {"==": "==", "!=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}.get(op)
            if cop:
                return _D(self.temp(f"({ar.v} {cop} {br.v}) ? 1.0 : 0.0"), [None] * self.N), None
            if op == "&&":
                return _D(self.temp(f"({ar.v} != 0.0 && {br.v} != 0.0) ? 1.0 : 0.0"), [None] * self.N), None
            if op == "||":
                return _D(self.temp(f"({ar.v} != 0.0 || {br.v} != 0.0) ? 1.0 : 0.0"), [None] * self.N), None
            raise VAError(f"Verilog-A: operator {op!r} not supported")
        if kind == "call":
            return self.call(e[1], e[2])
        raise VAError(f"Verilog-A: cannot lower {kind!r}")

    def _no_react(self, q, what):
        if q is not None:
            raise VAError(f"Verilog-A: ddt() inside {what} is not supported")

    def _node(self, name: str) -> int:
        try:
            return self.nodes.index(name)
        except ValueError:
            raise VAError(f"Verilog-A: {name!r} is not a node of the module") from None

    def call(self, fn: str, args) -> Tuple[_D, Optional[_D]]:
        if fn == "ddt":
            r, q = self.ev(args[0])
            self._no_react(q, "ddt")                 # second derivatives are out of scope
            return self.zero(), r                      # va_ddt: s-dual (contrib.jl:356-375)
        if fn in ("white_noise", "flicker_noise"):
            return self.zero(), None                   # 0 on the value path (vasim.jl)
        if fn == "$simparam":
            if len(args) > 1:
                r, _ = self.ev(args[1])
                return r, None
            raise VAError("Verilog-A: $simparam without default is not supported")
        if fn in ("$limit", "$vt", "$temperature", "$param_given", "analysis"):
            raise VAError(f"Verilog-A: {fn} is not supported in this tier")
        if fn.startswith("$") and fn[1:] in ("pow", "exp", "ln", "log", "sqrt", "abs", "sin", "cos", "tan",
                                             "tanh", "sinh", "cosh", "atan", "min", "max", "hypot",
                                             "floor", "ceil", "limexp"):
            fn = fn[1:]                              # $pow(...) etc.: system-function spelling
        ev = [self.ev(a) for a in args]
        for _, q in ev:
            self._no_react(q, fn)
        a = ev[0][0]
        b = ev[1][0] if len(ev) > 1 else None
        one = lambda f: (f, None)
        if fn == "exp":
            return one(self.unary_fn(a, f"CB_EXP({a.v})", lambda v: v))
        if fn == "limexp":          # exp(clamp(x,-80,80)), zero slope outside (va_env.jl)
            c = self.temp(f"fmin(fmax({a.v}, -80.0), 80.0)")
            return one(self.unary_fn(a, f"CB_EXP({c})",
                                     lambda v: f"({a.v} >= -80.0 && {a.v} <= 80.0) ? {v} : 0.0"))
        if fn == "ln":
            return one(self.unary_fn(a, f"log({a.v})", lambda v: f"1.0 / {a.v}"))
        if fn == "log":
            return one(self.unary_fn(a, f"log10({a.v})", lambda v: f"1.0 / ({a.v} * 2.302585092994046)"))
        if fn == "sqrt":
            return one(self.unary_fn(a, f"sqrt({a.v})", lambda v: f"0.5 / {v}"))
        if fn == "abs":
            return one(self.unary_fn(a, f"fabs({a.v})", lambda v: f"({a.v} < 0.0) ? -1.0 : 1.0"))
        if fn == "sin":
            return one(self.unary_fn(a, f"sin({a.v})", lambda v: f"cos({a.v})"))
        if fn == "cos":
            return one(self.unary_fn(a, f"cos({a.v})", lambda v: f"-sin({a.v})"))
        if fn == "tan":
            return one(self.unary_fn(a, f"tan({a.v})", lambda v: f"1.0 + {v} * {v}"))
        if fn == "tanh":
            return one(self.unary_fn(a, f"tanh({a.v})", lambda v: f"1.0 - {v} * {v}"))
        if fn == "sinh":
            return one(self.unary_fn(a, f"sinh({a.v})", lambda v: f"cosh({a.v})"))
        if fn == "cosh":
            return one(self.unary_fn(a, f"cosh({a.v})", lambda v: f"sinh({a.v})"))
        if fn == "atan":
            return one(self.unary_fn(a, f"atan({a.v})", lambda v: f"1.0 / (1.0 + {a.v} * {a.v})"))
        if fn in ("floor", "ceil"):
            return one(_D(self.temp(f"{fn}({a.v})"), [None] * self.N))
        if fn == "pow":
            v = self.temp(f"pow({a.v}, {b.v})")
            d = []
            fa = fb = None
            for pa, pb in zip(a.d, b.d):
                terms = []
                if pa is not None:
                    fa = fa or self.temp(f"{b.v} * pow({a.v}, {b.v} - 1.0)")
                    terms.append(f"{fa} * {pa}")
                if pb is not None:
                    fb = fb or self.temp(f"{v} * log({a.v})")
                    terms.append(f"{fb} * {pb}")
                d.append(self.temp(" + ".join(terms)) if terms else None)
            return one(_D(v, d))
        if fn in ("min", "max"):
            cmp = "<" if fn == "min" else ">"
            return one(self.select(f"{a.v} {cmp} {b.v}", a, b))
        if fn == "hypot":
            h = self.temp(f"sqrt({a.v} * {a.v} + {b.v} * {b.v})")
            d = []
            for pa, pb in zip(a.d, b.d):
                terms = [f"{a.v} * {pa}" if pa else None, f"{b.v} * {pb}" if pb else None]
                terms = [x for x in terms if x]
                d.append(self.temp("(" + " + ".join(terms) + f") / {h}") if terms else None)
            return one(_D(h, d))
        raise VAError(f"Verilog-A: function {fn!r} is not supported in this tier")

    # -- statements ------------------------------------------------------------ #
    def stmt(self, s):
        kind = s[0]
        if kind == "block":
            for x in s[1]:
                self.stmt(x)
        elif kind == "assign":
            name = s[1]
            if name not in self.vars:
                raise VAError(f"Verilog-A: assignment to undeclared variable {name!r}")
            r, q = self.ev(s[2])
            self._no_react(q, "assignment")
            self.emit(f"v_{name} = {r.v};")
            for k in range(self.N):
                self.emit(f"v_{name}_d{k} = {r.d[k] or '0.0'};")
            self.assigned.add(name)
        elif kind == "if":
            c, cq = self.ev(s[1])
            self._no_react(cq, "if")
            # a variable first assigned inside a branch must read as declared afterwards
            before = set(self.assigned)
            self.emit(f"if ({c.v} != 0.0) {{")
            self.indent += 1
            self.stmt(s[2])
            self.indent -= 1
            a1 = set(self.assigned)
            self.assigned = set(before)
            self.emit("} else {")
            self.indent += 1
            self.stmt(s[3])
            self.indent -= 1
            self.emit("}")
            self.assigned = a1 | self.assigned
        elif kind == "contrib":
            key = (s[1], s[2])
            if key not in self.branches:
                self.branches.append(key)
            bi = self.branches.index(key)
            r, q = self.ev(s[3])
            self.emit(f"I{bi} += {r.v};")
            for k in range(self.N):
                if r.d[k] is not None:
                    self.emit(f"I{bi}_d{k} += {r.d[k]};")
            if q is not None:
                self.reactive[bi] = True
                self.emit(f"Q{bi} += {q.v};")
                for k in range(self.N):
                    if q.d[k] is not None:
                        self.emit(f"Q{bi}_d{k} += {q.d[k]};")
                # voltage-dependent iff a partial of the charge is not a parameter-only constant
                if any(p is not None and not self._param_only(p) for p in q.d):
                    self.vdep[bi] = True
        else:
            raise VAError(f"Verilog-A: statement {kind!r} not supported")

    def _param_only(self, expr: str, _seen=None) -> bool:
        """True when the C expression depends on parameters / literals only."""
        _seen = _seen or set()
        for name in re.findall(r"[A-Za-z_][A-Za-z0-9_]*", expr):
            if name in ("fmin", "fmax", "pow", "log", "exp", "sqrt", "CB_EXP", "fabs", "e", "E"):
                continue
            if name.startswith("p_"):
                continue
            if re.fullmatch(r"t\d+", name):
                if name in _seen:
                    continue
                _seen.add(name)
                rhs = self._temp_rhs(name)
                if rhs is None or not self._param_only(rhs, _seen):
                    return False
                continue
            return False             # V*, v_* (state-dependent) or anything unknown
        return True

    def _temp_rhs(self, name: str) -> Optional[str]:
        pat = f"const double {name} = "
        for ln in self.lines:
            s = ln.strip()
            if s.startswith(pat):
                return s[len(pat):-1]
        return None


# --------------------------------------------------------------------------- #
# model object
# --------------------------------------------------------------------------- #
class VAModel:
    """A parsed Verilog-A module: knows its structure (ports, internal nodes, branches)
    and emits its stamp function for the CUDA kernels and for the C oracle."""

    def __init__(self, source: str):
        self.source = source
        self.mod = _Parser(source).module()
        m = self.mod
        self.name = m["name"]
        self.ports = list(m["ports"])
        self.internal = [n for n in m["electrical"] if n not in self.ports]
        self.nodes = self.ports + self.internal           # "all nodes": ports, then internals
        self.param_names = [p[0] for p in m["params"]]
        self.param_defaults = [p[1] for p in m["params"]]
        em = _Emitter(m, self.param_names, self.nodes)
        em.stmt(m["body"])
        self._em = em
        self.branches = list(em.branches)
        self.reactive = [bool(em.reactive.get(i)) for i in range(len(self.branches))]
        self.vdep = [bool(em.vdep.get(i)) for i in range(len(self.branches))]
        self.n_charges = sum(1 for i in range(len(self.branches)) if self.reactive[i] and self.vdep[i])
        self.uid = hashlib.sha256(source.encode()).hexdigest()[:12]
        self.cname = re.sub(r"\W", "_", self.name) + "_" + self.uid

    # ---- instances ----------------------------------------------------------- #
    def __call__(self, **kwargs):
        return VAInstance(self, kwargs)

    def default_params(self, given: Dict[str, Any]) -> List[Any]:
        lower = {k.lower(): v for k, v in given.items() if k != "name"}
        unknown = set(lower) - {p.lower() for p in self.param_names}
        if unknown:
            raise VAError(f"{self.name}: unknown parameter(s) {sorted(unknown)}")
        vals: Dict[str, Any] = {}
        for name, default in zip(self.param_names, self.param_defaults):
            if name.lower() in lower:
                vals[name] = lower[name.lower()]
            else:
                vals[name] = _const_eval(default, vals)
        return [vals[n] for n in self.param_names]

    # ---- structure (the COO order the emitted code produces) ------------------ #
    def stamp_plan(self):
        """[(kind, args)] in emission order; kinds: G(i,j) C(i,j) b(i) where i, j are
        local node slots (0..N-1 = module nodes, N.. = charge unknowns) ."""
        N = len(self.nodes)
        plan = []
        q = 0
        for bi, (a, b) in enumerate(self.branches):
            p = self.nodes.index(a)
            n = self.nodes.index(b) if b is not None else None
            for k in range(N):
                plan.append(("G", p, k))
                if n is not None:
                    plan.append(("G", n, k))
            if self.reactive[bi]:
                if self.vdep[bi]:
                    qi = N + q
                    q += 1
                    plan.append(("Q", p, n, f"Q_{a}_{b if b is not None else '0'}"))   # alloc_charge!
                    plan.append(("C", p, qi))
                    if n is not None:
                        plan.append(("C", n, qi))
                    plan.append(("G", qi, qi))
                    for k in range(N):
                        plan.append(("G", qi, k))
                    plan.append(("b", qi))
                else:
                    for k in range(N):
                        plan.append(("C", p, k))
                        if n is not None:
                            plan.append(("C", n, k))
            plan.append(("b", p))
            if n is not None:
                plan.append(("b", n))
        return plan

    # ---- code emission --------------------------------------------------------- #
    def _body(self, target: str) -> str:
        """target: 'cuda' or 'c'.  The two differ only in the prologue / stamping macros."""
        em, N = self._em, len(self.nodes)
        L: List[str] = []
        for k in range(N):
            L.append(f"    const double V{k} = VA_V({k});")
        for i, name in enumerate(self.param_names):
            L.append(f"    const double p_{name} = VA_PARAM({i}); (void)p_{name};")
        for name in self.mod["reals"]:
            L.append(f"    double v_{name} = 0.0; (void)v_{name};")
            L.append("    double " + ", ".join(f"v_{name}_d{k} = 0.0" for k in range(N)) + ";")
            L.append("    " + " ".join(f"(void)v_{name}_d{k};" for k in range(N)))
        for bi in range(len(self.branches)):
            L.append(f"    double I{bi} = 0.0, " + ", ".join(f"I{bi}_d{k} = 0.0" for k in range(N)) + ";")
            if self.reactive[bi]:
                L.append(f"    double Q{bi} = 0.0, " + ", ".join(f"Q{bi}_d{k} = 0.0" for k in range(N)) + ";")
        L += em.lines
        q = 0
        for bi, (a, b) in enumerate(self.branches):
            p = self.nodes.index(a)
            n = self.nodes.index(b) if b is not None else None
            L.append(f"    /* branch {bi}: I({a}{',' + b if b else ''}) */")
            for k in range(N):
                L.append(f"    VA_G({p}, {k}, I{bi}_d{k});")
                if n is not None:
                    L.append(f"    VA_G({n}, {k}, -I{bi}_d{k});")
            if self.reactive[bi]:
                if self.vdep[bi]:
                    qi = N + q
                    q += 1
                    L.append(f"    VA_C({p}, {qi}, 1.0 / {CHARGE_SCALE!r});")
                    if n is not None:
                        L.append(f"    VA_C({n}, {qi}, -1.0 / {CHARGE_SCALE!r});")
                    L.append(f"    VA_G({qi}, {qi}, 1.0);")
                    for k in range(N):
                        L.append(f"    VA_G({qi}, {k}, -{CHARGE_SCALE!r} * Q{bi}_d{k});")
                    L.append(f"    {{ double bq = Q{bi};")
                    for k in range(N):
                        L.append(f"      bq -= Q{bi}_d{k} * V{k};")
                    L.append(f"      VA_B({qi}, {CHARGE_SCALE!r} * bq); }}")
                else:
                    for k in range(N):
                        L.append(f"    VA_C({p}, {k}, Q{bi}_d{k});")
                        if n is not None:
                            L.append(f"    VA_C({n}, {k}, -Q{bi}_d{k});")
            L.append(f"    {{ double Ieq = I{bi};")
            for k in range(N):
                L.append(f"      Ieq += -I{bi}_d{k} * V{k};")
            L.append(f"      VA_B({p}, -Ieq);")
            if n is not None:
                L.append(f"      VA_B({n}, Ieq);")
            L.append("    }")
        return "\n".join(L)

    def emit_cuda(self) -> str:
        """``template <int PASS, typename PG, typename W> va_stamp_<cname>(...)`` for
        lane_kernels.cuh: node slot i of the module = ``pg.dev_node(nb + i)``."""
        nloc = len(self.nodes) + self.n_charges
        head = [f"// Verilog-A module {self.name} (emitted by cadnip_b200.verilog_a)",
                "template <int PASS, typename PG, typename W>",
                f"__device__ __forceinline__ void va_stamp_{self.cname}(const PG &pg, W &w, int d, double t)",
                "{",
                "    (void)t;",
                "    if (PASS == 2) return;                      // nothing depends on time alone",
                "    const int nb = pg.dev_node_ptr(d), pb = pg.dev_param_ptr(d);",
                "    int g = pg.off_SG() + pg.dev_gbase(d), c = pg.off_SC() + pg.dev_cbase(d);",
                "    int b = pg.off_SB() + pg.dev_bbase(d);",
                "    (void)c;"]
        for i in range(nloc):
            head.append(f"    const int n{i} = pg.dev_node(nb + {i});")
        return "\n".join(head) + "\n" + self._body("cuda") + "\n}\n"

    def emit_c(self) -> str:
        """Plain C for the oracle: ``void ora_va_<cname>(const ora_va_api *A, void *ctx, ...)``."""
        nloc = len(self.nodes) + self.n_charges
        N = len(self.nodes)
        L = [f"/* Verilog-A module {self.name} (emitted by cadnip_b200.verilog_a) */",
             f"void ora_va_{self.cname}(const ora_va_api *A, void *ctx, const int *ports, const double *par,",
             "                          const double *x, long nx, double t)",
             "{",
             "    (void)t;",
             f"    long n[{nloc}];"]
        for i in range(len(self.ports)):
            L.append(f"    n[{i}] = ports[{i}];")
        for j, name in enumerate(self.internal):
            L.append(f"    n[{len(self.ports) + j}] = A->alloc_internal_node(ctx);   /* {name} */")
        L.append(f"    double Vn[{N}];")
        L.append(f"    for (int k = 0; k < {N}; k++) Vn[k] = A->xval(ctx, n[k], x, nx);")
        # charge unknowns are allocated where the reference allocates them: inside the branch loop
        body = self._body("c")
        q = 0
        for bi in range(len(self.branches)):
            if self.reactive[bi] and self.vdep[bi]:
                a, b = self.branches[bi]
                p = self.nodes.index(a)
                nn = self.nodes.index(b) if b is not None else -1
                marker = f"    VA_C({p}, {N + q}, 1.0 / {CHARGE_SCALE!r});"
                alloc = (f"    n[{N + q}] = A->alloc_charge(ctx, n[{p}], "
                         f"{'n[%d]' % nn if nn >= 0 else '0'});\n")
                body = body.replace(marker, alloc + marker, 1)
                q += 1
        return "\n".join(L) + "\n" + body + "\n}\n"


def _const_eval(e, env: Dict[str, Any]):
    """Evaluate a parameter default (may reference earlier parameters)."""
    k = e[0]
    if k == "num":
        return e[1]
    if k == "var":
        if e[1] in env:
            return env[e[1]]
        raise VAError(f"Verilog-A: parameter default references unknown {e[1]!r}")
    if k == "neg":
        return -_const_eval(e[1], env)
    if k == "bin":
        a, b = _const_eval(e[2], env), _const_eval(e[3], env)
        return {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b}[e[1]]()
    if k == "call" and e[1] in ("exp", "ln", "sqrt", "pow"):
        import numpy as np
        args = [_const_eval(a, env) for a in e[2]]
        return {"exp": np.exp, "ln": np.log, "sqrt": np.sqrt, "pow": np.power}[e[1]](*args)
    raise VAError("Verilog-A: unsupported expression in a parameter default")


class VAInstance:
    """``NMOS(K=1e-3, Vth=0.5)``: a module with its parameter values bound
    (``spicecall`` with case-insensitive names, src/spectre.jl:474-513)."""

    def __init__(self, model: VAModel, kwargs: Dict[str, Any]):
        self.model = model
        self.name = kwargs.pop("name", model.name)
        self.params = model.default_params(kwargs)


def va(source: str) -> VAModel:
    """``va\"\"\"...\"\"\"`` (src/vasim.jl:4050 make_mna_module): parse one module."""
    return VAModel(source)


# --------------------------------------------------------------------------- #
# per-circuit headers
# --------------------------------------------------------------------------- #
_CUDA_PRELUDE = """// generated by cadnip_b200.verilog_a -- Verilog-A device models for lane_kernels.cuh
#define CB_EXP(x) d_exp(x)
#define VA_V(k) xval(pg, w, n##k)
#define VA_PARAM(i) param(pg, w, pb, i)
#define VA_G(i, j, v) do { if (n##i != 0 && n##j != 0) { w(g) = (v); g++; } } while (0)
#define VA_C(i, j, v) do { if (n##i != 0 && n##j != 0) { w(c) = (v); c++; } } while (0)
#define VA_B(i, v) do { if (n##i != 0) { w(b) = (v); b++; } } while (0)
/* included from lane_kernels.cuh, inside namespace cb200 */
"""

_CUDA_EPILOGUE = """
#undef CB_EXP
#undef VA_V
#undef VA_PARAM
#undef VA_G
#undef VA_C
#undef VA_B
"""

_C_PRELUDE = """/* generated by cadnip_b200.verilog_a -- Verilog-A device models for the CPU oracle */
#include <math.h>
typedef struct ora_va_api {
    long (*alloc_internal_node)(void *ctx);
    long (*alloc_charge)(void *ctx, long p, long n);
    double (*xval)(void *ctx, long node, const double *x, long nx);
    void (*stamp_G)(void *ctx, long i, long j, double v);
    void (*stamp_C)(void *ctx, long i, long j, double v);
    void (*stamp_b)(void *ctx, long i, double v);
} ora_va_api;
#define CB_EXP(x) exp(x)
#define VA_V(k) Vn[k]
#define VA_PARAM(i) par[i]
#define VA_G(i, j, v) A->stamp_G(ctx, n[i], n[j], (v))
#define VA_C(i, j, v) A->stamp_C(ctx, n[i], n[j], (v))
#define VA_B(i, v) A->stamp_b(ctx, n[i], (v))
"""


def cuda_header(models: Sequence[VAModel]) -> str:
    """All models of a circuit + the dispatcher ``va_dispatch<PASS>(pg, w, d, model, t)``."""
    out = [_CUDA_PRELUDE]
    for m in models:
        out.append(m.emit_cuda())
    out.append("template <int PASS, typename PG, typename W>\n"
               "__device__ __forceinline__ void va_dispatch(const PG &pg, W &w, int d, int model, double t)\n{\n"
               "    switch (model) {")
    for i, m in enumerate(models):
        out.append(f"    case {i}: va_stamp_{m.cname}<PASS>(pg, w, d, t); break;")
    out.append("    default: break;\n    }\n}")
    out.append(_CUDA_EPILOGUE)
    return "\n".join(out)


def c_source(models: Sequence[VAModel]) -> str:
    out = [_C_PRELUDE]
    for m in models:
        out.append(m.emit_c())
    out.append("typedef void (*ora_va_fn)(const ora_va_api *, void *, const int *, const double *, const double *, long, double);")
    out.append(f"int ora_va_count(void) {{ return {len(models)}; }}")
    out.append("ora_va_fn ora_va_table(int i)\n{\n    switch (i) {")
    for i, m in enumerate(models):
        out.append(f"    case {i}: return ora_va_{m.cname};")
    out.append("    default: return 0;\n    }\n}")
    return "\n".join(out)
