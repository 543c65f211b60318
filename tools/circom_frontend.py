"""Circom-subset frontend: .circom sources -> circom-witnesscalc graph (.bin).

TOOLING, not product: it replaces the reference's offline `build-circuit` binary
(/root/reference/src/bin/build-circuit.rs, which needs the un-vendored circom
compiler crates and a Rust toolchain, neither present) so that the test circuits
of the reference can be turned into `wtns.graph.001` files for the evaluator.

It is an independent implementation: a hand-written lexer / recursive-descent
parser for the circom 2.1 subset used by circomlib and the iden3 authV2
circuits, and a symbolic executor that runs templates at "compile time" and emits
one graph node per signal-dependent operation, like build-circuit does
(build-circuit.rs:512 process_instruction, :1843 calc_expression).  The passes
of graph::optimize (/root/reference/src/graph.rs:358-365: tree_shake, propagate,
value_numbering, constants, tree_shake) are restated in `optimize()`.

Known, documented differences from a real build-circuit run:
  * witness_signals = [1, main outputs, main inputs, remaining signals in
    execution order] minus, at the default opt_level 2, every non-main signal
    that a `<==` defines by a LINEAR expression of other signals: circom's
    constraint simplifier, which build-circuit runs at full strength
    (build-circuit.rs:2189-2205), substitutes exactly such linear constraints
    away, so what is left are the inputs/outputs of main, the signals defined
    by quadratic constraints and the `<--` hints.  The real simplifier also
    works on explicit `===` constraints and picks its own substitution order;
    the selection here is an approximation of its result, not a replica.
    opt_level 1 keeps every distinct declared signal (circom --O1-like).
  * node order is this executor's data-flow order; any topological order is
    valid for the evaluator.
"""
from __future__ import annotations

import os
import random
import re
import sys
from typing import Dict, List, Optional, Tuple

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import pyoracle as po  # noqa: E402  (tooling may use the oracle's op semantics)

P = po.M

# --------------------------------------------------------------------------
# lexer
# --------------------------------------------------------------------------

_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<hex>0x[0-9a-fA-F]+)
  | (?P<num>[0-9]+)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"[^"]*")
  | (?P<op><==|==>|<--|-->|===|\*\*=|<<=|>>=|\*\*|<<|>>|<=|>=|==|!=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|\\=|%=|&=|\|=|\^=|[-+*/\\%&|^~!<>=?:;,.(){}\[\]])
""", re.X | re.S)


def lex(src: str, fname: str):
    toks = []
    pos = 0
    line = 1
    n = len(src)
    while pos < n:
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"{fname}:{line}: bad character {src[pos]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind == "hex":
            toks.append(("num", int(text, 16), line))
        elif kind == "num":
            toks.append(("num", int(text), line))
        elif kind == "id":
            toks.append(("id", text, line))
        elif kind == "str":
            toks.append(("str", text[1:-1], line))
        elif kind == "op":
            toks.append(("op", text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(("eof", None, line))
    return toks


# --------------------------------------------------------------------------
# parser
# --------------------------------------------------------------------------

ASSIGN_OPS = {"=", "<==", "<--", "+=", "-=", "*=", "/=", "\\=", "%=", "**=", "<<=", ">>=", "&=", "|=", "^="}
# circom precedence, loosest first (comparison binds looser than bitwise ops)
BIN_LEVELS = [
    ["||"], ["&&"], ["==", "!=", "<", ">", "<=", ">="], ["|"], ["^"], ["&"],
    ["<<", ">>"], ["+", "-"], ["*", "/", "\\", "%"],
]
BINOP_TO_DUO = {"*": 0, "/": 1, "+": 2, "-": 3, "**": 4, "\\": 5, "%": 6, "==": 7, "!=": 8, "<": 9, ">": 10,
                "<=": 11, ">=": 12, "&&": 13, "||": 14, "<<": 15, ">>": 16, "|": 17, "&": 18, "^": 19}


class Parser:
    def __init__(self, toks, fname):
        self.t = toks
        self.i = 0
        self.fname = fname

    # token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, val, k=0):
        tk = self.t[self.i + k]
        return tk[0] in ("op", "id") and tk[1] == val

    def next(self):
        tk = self.t[self.i]
        self.i += 1
        return tk

    def expect(self, val):
        tk = self.next()
        if tk[1] != val:
            raise SyntaxError(f"{self.fname}:{tk[2]}: expected {val!r}, got {tk[1]!r}")
        return tk

    def ident(self):
        tk = self.next()
        if tk[0] != "id":
            raise SyntaxError(f"{self.fname}:{tk[2]}: expected identifier, got {tk[1]!r}")
        return tk[1]

    # top level
    def parse_file(self):
        includes, templates, functions, main = [], {}, {}, None
        while self.peek()[0] != "eof":
            if self.at("pragma"):
                while not self.at(";"):
                    self.next()
                self.next()
            elif self.at("include"):
                self.next()
                includes.append(self.next()[1])
                self.expect(";")
            elif self.at("template"):
                self.next()
                while self.at("custom") or self.at("parallel"):
                    self.next()
                name = self.ident()
                params = self.param_list()
                body = self.block()
                templates[name] = (params, body)
            elif self.at("function"):
                self.next()
                name = self.ident()
                params = self.param_list()
                body = self.block()
                functions[name] = (params, body)
            elif self.at("component"):
                self.next()
                self.expect("main")
                public = []
                if self.at("{"):
                    self.next()
                    self.expect("public")
                    self.expect("[")
                    while not self.at("]"):
                        public.append(self.ident())
                        if self.at(","):
                            self.next()
                    self.expect("]")
                    self.expect("}")
                self.expect("=")
                e = self.expression()
                self.expect(";")
                main = (public, e)
            else:
                tk = self.peek()
                raise SyntaxError(f"{self.fname}:{tk[2]}: unexpected {tk[1]!r} at top level")
        return includes, templates, functions, main

    def param_list(self):
        self.expect("(")
        ps = []
        while not self.at(")"):
            ps.append(self.ident())
            if self.at(","):
                self.next()
        self.expect(")")
        return ps

    def block(self):
        self.expect("{")
        stmts = []
        while not self.at("}"):
            stmts.append(self.statement())
        self.expect("}")
        return ("block", stmts)

    def stmt_or_block(self):
        if self.at("{"):
            return self.block()
        return self.statement()

    def dims(self):
        d = []
        while self.at("["):
            self.next()
            d.append(self.expression())
            self.expect("]")
        return d

    def statement(self):
        tk = self.peek()
        line = tk[2]
        if self.at("{"):
            return self.block()
        if tk[0] == "id":
            kw = tk[1]
            if kw == "signal":
                self.next()
                kind = "inter"
                if self.at("input") or self.at("output"):
                    kind = self.next()[1]
                if self.at("{"):           # tags
                    while not self.at("}"):
                        self.next()
                    self.next()
                names = []
                while True:
                    nm = self.ident()
                    names.append((nm, self.dims()))
                    if self.at(","):
                        self.next()
                        continue
                    break
                init = None
                if self.at("<==") or self.at("<--"):
                    op = self.next()[1]
                    init = (op, self.expression())
                self.expect(";")
                return ("sigdecl", kind, names, init, line)
            if kw == "var":
                self.next()
                s = self.var_decl_rest(line)
                self.expect(";")
                return s
            if kw == "component":
                self.next()
                items = []
                while True:
                    nm = self.ident()
                    d = self.dims()
                    init = None
                    if self.at("="):
                        self.next()
                        init = self.expression()
                    items.append((nm, d, init))
                    if self.at(","):
                        self.next()
                        continue
                    break
                self.expect(";")
                return ("compdecl", items, line)
            if kw == "if":
                self.next()
                self.expect("(")
                c = self.expression()
                self.expect(")")
                th = self.stmt_or_block()
                el = None
                if self.at("else"):
                    self.next()
                    el = self.stmt_or_block()
                return ("if", c, th, el, line)
            if kw == "for":
                self.next()
                self.expect("(")
                if self.at("var"):
                    self.next()
                    init = self.var_decl_rest(line)
                else:
                    init = self.simple_statement()
                self.expect(";")
                cond = self.expression()
                self.expect(";")
                step = self.simple_statement()
                self.expect(")")
                body = self.stmt_or_block()
                return ("for", init, cond, step, body, line)
            if kw == "while":
                self.next()
                self.expect("(")
                c = self.expression()
                self.expect(")")
                body = self.stmt_or_block()
                return ("while", c, body, line)
            if kw == "return":
                self.next()
                e = self.expression()
                self.expect(";")
                return ("return", e, line)
            if kw == "assert":
                self.next()
                self.expect("(")
                e = self.expression()
                self.expect(")")
                self.expect(";")
                return ("assert", e, line)
            if kw == "log":
                self.next()
                self.expect("(")
                depth = 1
                while depth:
                    t2 = self.next()
                    if t2[1] == "(":
                        depth += 1
                    elif t2[1] == ")":
                        depth -= 1
                self.expect(";")
                return ("nop",)
        s = self.simple_statement()
        self.expect(";")
        return s

    def var_decl_rest(self, line):
        items = []
        while True:
            nm = self.ident()
            d = self.dims()
            init = None
            if self.at("="):
                self.next()
                init = self.expression()
            items.append((nm, d, init))
            if self.at(","):
                self.next()
                continue
            break
        return ("vardecl", items, line)

    def simple_statement(self):
        line = self.peek()[2]
        e = self.expression()
        tk = self.peek()
        if tk[0] == "op":
            v = tk[1]
            if v in ASSIGN_OPS:
                self.next()
                rhs = self.expression()
                return ("subst", e, v, rhs, line)
            if v in ("==>", "-->"):
                self.next()
                tgt = self.expression()
                return ("subst", tgt, "<==" if v == "==>" else "<--", e, line)
            if v == "===":
                self.next()
                rhs = self.expression()
                return ("constraint", e, rhs, line)
            if v == "++":
                self.next()
                return ("subst", e, "+=", ("num", 1), line)
            if v == "--":
                self.next()
                return ("subst", e, "-=", ("num", 1), line)
        return ("expr", e, line)

    # expressions
    def expression(self):
        if self.at("parallel"):
            self.next()
        c = self.binary(0)
        if self.at("?"):
            self.next()
            a = self.expression()
            self.expect(":")
            b = self.expression()
            return ("tern", c, a, b)
        return c

    def binary(self, lvl):
        if lvl == len(BIN_LEVELS):
            return self.power()
        ops = BIN_LEVELS[lvl]
        l = self.binary(lvl + 1)
        while self.peek()[0] == "op" and self.peek()[1] in ops:
            op = self.next()[1]
            r = self.binary(lvl + 1)
            l = ("bin", op, l, r)
        return l

    def power(self):
        l = self.unary()
        while self.at("**"):
            self.next()
            r = self.unary()
            l = ("bin", "**", l, r)
        return l

    def unary(self):
        if self.peek()[0] == "op" and self.peek()[1] in ("-", "!", "~"):
            op = self.next()[1]
            e = self.unary()
            return ("un", op, e)
        return self.postfix()

    def postfix(self):
        tk = self.next()
        if tk[0] == "num":
            return ("num", tk[1] % P)
        if tk[0] == "op" and tk[1] == "(":
            e = self.expression()
            if self.at(","):
                items = [e]
                while self.at(","):
                    self.next()
                    items.append(self.expression())
                self.expect(")")
                return ("tuple", items)
            self.expect(")")
            return e
        if tk[0] == "op" and tk[1] == "[":
            items = []
            while not self.at("]"):
                items.append(self.expression())
                if self.at(","):
                    self.next()
            self.expect("]")
            return ("arr", items)
        if tk[0] != "id":
            raise SyntaxError(f"{self.fname}:{tk[2]}: unexpected token {tk[1]!r} in expression")
        name = tk[1]
        if self.at("("):
            args = self.call_args()
            if self.at("("):               # anonymous component  T(params)(inputs)
                inputs = self.call_args(named=True)
                e = ("anon", name, [a for _, a in args], inputs, tk[2])
            else:
                e = ("call", name, [a for _, a in args], tk[2])
                return e
        else:
            e = ("var", name, [], tk[2])
        # accessors
        acc = []
        while True:
            if self.at("["):
                self.next()
                acc.append(("idx", self.expression()))
                self.expect("]")
            elif self.at("."):
                self.next()
                acc.append(("mem", self.ident()))
            else:
                break
        if e[0] == "var":
            return ("var", name, acc, tk[2])
        if acc:
            return ("access", e, acc, tk[2])
        return e

    def call_args(self, named=False):
        self.expect("(")
        args = []
        while not self.at(")"):
            nm = None
            if named and self.peek()[0] == "id" and self.peek(1)[0] == "op" and self.peek(1)[1] in ("<==", "<--"):
                nm = self.next()[1]
                self.next()
            args.append((nm, self.expression()))
            if self.at(","):
                self.next()
        self.expect(")")
        return args


# --------------------------------------------------------------------------
# program loading
# --------------------------------------------------------------------------

class Program:
    def __init__(self, lib_dirs: List[str]):
        self.templates: Dict[str, tuple] = {}
        self.functions: Dict[str, tuple] = {}
        self.main = None
        self.lib_dirs = lib_dirs
        self.loaded = set()

    def resolve_include(self, inc: str, cur_dir: str) -> str:
        cand = [os.path.join(cur_dir, inc)] + [os.path.join(d, inc) for d in self.lib_dirs]
        marker = "circomlib/circuits/"
        if marker in inc:                      # node_modules/circomlib/circuits/x -> -l <circomlib>/x
            tail = inc.split(marker, 1)[1]
            cand += [os.path.join(d, tail) for d in self.lib_dirs]
        for c in cand:
            if os.path.isfile(c):
                return os.path.realpath(c)
        raise FileNotFoundError(f"include {inc!r} not found from {cur_dir}")

    def load(self, path: str, is_root=True):
        path = os.path.realpath(path)
        if path in self.loaded:
            return
        self.loaded.add(path)
        with open(path) as f:
            src = f.read()
        includes, templates, functions, main = Parser(lex(src, path), path).parse_file()
        for inc in includes:
            self.load(self.resolve_include(inc, os.path.dirname(path)), False)
        self.templates.update(templates)
        self.functions.update(functions)
        if main is not None and is_root:
            self.main = main


# --------------------------------------------------------------------------
# symbolic executor
# --------------------------------------------------------------------------

class Sym:
    """A value that depends on input signals: reference to a graph node.  `deg` is the degree of the EXPRESSION that
    produced it in the signals it reads (circom's classification of `<==` right-hand sides): 1 = linear, 2 = quadratic,
    3 = not quadratic (only legal under `<--`).  A signal read has degree 1 whatever defined the signal."""
    __slots__ = ("id", "deg")

    def __init__(self, i, deg=1):
        self.id = i
        self.deg = deg


def _deg(v):
    return v.deg if isinstance(v, Sym) else 0

    def __repr__(self):
        return f"n{self.id}"


class Builder:
    def __init__(self):
        self.nodes: List[tuple] = []
        self.const_ids: Dict[int, int] = {}
        self.cse: Dict[tuple, int] = {}

    def _push(self, node):
        i = self.cse.get(node)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(node)
            self.cse[node] = i
        return i

    def input(self, idx) -> Sym:
        return Sym(self._push((po.K_INPUT, idx)))

    def node_of(self, v) -> int:
        if isinstance(v, Sym):
            return v.id
        return self._push((po.K_CONST, v % P))

    def duo(self, op, a, b):
        if not isinstance(a, Sym) and not isinstance(b, Sym):
            return po.eval_duo(op, a, b, "circom")
        da, db = _deg(a), _deg(b)
        if op in (2, 3):                       # Add, Sub
            deg = max(da, db)
        elif op == 0:                          # Mul
            deg = min(da + db, 3)
        elif op == 1 and db == 0:              # Div by a constant
            deg = da
        else:
            deg = 3
        return Sym(self._push((po.K_DUO, op, self.node_of(a), self.node_of(b))), deg)

    def uno(self, op, a):
        if not isinstance(a, Sym):
            return po.eval_uno(op, a, "circom")
        return Sym(self._push((po.K_UNO, op, a.id)), a.deg if op == 0 else 3)

    def tern(self, c, a, b):
        if not isinstance(c, Sym):
            return a if c != 0 else b
        return Sym(self._push((po.K_TRES, 0, c.id, self.node_of(a), self.node_of(b))), 3)


def _prod(dims):
    n = 1
    for d in dims:
        n *= d
    return n


class Signal:
    __slots__ = ("name", "kind", "dims", "vals", "owner", "lin")

    def __init__(self, name, kind, dims, owner):
        self.name = name
        self.kind = kind
        self.dims = dims
        self.vals = [None] * _prod(dims)
        self.lin = [False] * len(self.vals)     # defined by `<==` with a linear (or constant) right-hand side
        self.owner = owner


class Comp:
    __slots__ = ("tname", "params", "inputs", "in_order", "pending", "done", "signals", "out_order", "path")

    def __init__(self, tname, params, path):
        self.tname = tname
        self.params = params
        self.inputs: Dict[str, Signal] = {}
        self.in_order: List[str] = []
        self.pending = 0
        self.done = False
        self.signals: Dict[str, Signal] = {}
        self.out_order: List[str] = []
        self.path = path


class ReturnEx(Exception):
    def __init__(self, v):
        self.v = v


class CircomError(Exception):
    pass


def _zeros(dims):
    if not dims:
        return 0
    return [_zeros(dims[1:]) for _ in range(dims[0])]


def _nones(dims):
    if not dims:
        return None
    return [_nones(dims[1:]) for _ in range(dims[0])]


def _copy(v):
    if isinstance(v, list):
        return [_copy(x) for x in v]
    return v


def _freeze(v):
    if isinstance(v, list):
        return tuple(_freeze(x) for x in v)
    if isinstance(v, Sym):
        raise TypeError
    return v


class Env:
    __slots__ = ("vars", "comp", "comps", "is_function")

    def __init__(self, comp, is_function=False):
        self.vars: Dict[str, object] = {}
        self.comp = comp
        self.comps: Dict[str, object] = {}
        self.is_function = is_function


class Executor:
    def __init__(self, prog: Program, opt_level=2):
        self.prog = prog
        self.opt_level = opt_level      # 1: every declared signal is a witness signal; 2: signals defined by a linear `<==` are not
        self.b = Builder()
        self.signal_order: List[Signal] = []
        self.constraints: List[Tuple[int, int, str]] = []   # (node_a, node_b, where)
        self.fn_cache: Dict[tuple, object] = {}
        self.prescan_cache: Dict[tuple, list] = {}
        self.n_components = 0
        sys.setrecursionlimit(100000)

    # ---- entry -----------------------------------------------------------
    def run_main(self):
        public, e = self.prog.main
        if e[0] != "call":
            raise CircomError("main must be a template call")
        _, tname, args, _ = e
        genv = Env(None)
        params = [self.eval(a, genv) for a in args]
        one = self.b.input(0)
        comp = self.instantiate(tname, params, "main")
        # main inputs -> Input nodes in declaration order, slot 0 is the constant 1
        input_map = {}
        off = 1
        for nm in comp.in_order:
            sig = comp.inputs[nm]
            n = len(sig.vals)
            input_map[nm] = (off, n)
            for k in range(n):
                sig.vals[k] = self.b.input(off + k)
            off += n
        comp.pending = 0
        self.run_template(comp)
        # witness order
        wit = [one.id]
        seen = {one.id}
        main_sigs = [comp.signals[n] for n in comp.out_order] + [comp.inputs[n] for n in comp.in_order]
        for sig in main_sigs:
            for v in sig.vals:
                if v is None:
                    raise CircomError(f"main signal {sig.name} unassigned")
                wit.append(self.b.node_of(v))
                seen.add(wit[-1])
        self.n_fixed = len(wit)
        n_all = 1 + sum(len(s.vals) for s in main_sigs)
        unassigned = 0
        n_linear = 0
        for sig in self.signal_order:
            if sig.owner is comp and (sig.kind != "inter"):
                continue
            for k, v in enumerate(sig.vals):
                n_all += 1
                if v is None:
                    unassigned += 1
                    continue
                if self.opt_level >= 2 and sig.lin[k]:
                    n_linear += 1              # circom --O2: the linear constraint defining it is substituted away
                    continue
                nid = self.b.node_of(v)
                if nid not in seen:
                    seen.add(nid)
                    wit.append(nid)
        self.stats = {"signals_declared": n_all, "signals_unassigned": unassigned, "signals_linear_eliminated": n_linear,
                      "opt_level": self.opt_level, "components": self.n_components}
        return self.b.nodes, wit, input_map

    # ---- components ------------------------------------------------------
    def prescan(self, tname, params):
        """Find the input signals (names, dims) of a template instance without running it."""
        try:
            key = (tname, _freeze(params))
        except TypeError:
            key = None
        if key is not None and key in self.prescan_cache:
            return self.prescan_cache[key]
        pnames, body = self.prog.templates[tname]
        env = Env(None)
        for n, v in zip(pnames, params):
            env.vars[n] = v
        found = []
        for st in body[1]:
            k = st[0]
            if k == "vardecl":
                for nm, dims, init in st[1]:
                    try:
                        d = [self.eval_int(x, env) for x in dims]
                        if init is not None:
                            env.vars[nm] = _copy(self.eval(init, env))
                        else:
                            env.vars[nm] = _zeros(d)
                    except Exception:
                        env.vars.pop(nm, None)
            elif k == "subst" and st[2] == "=" and st[1][0] == "var" and not st[1][2]:
                try:
                    env.vars[st[1][1]] = _copy(self.eval(st[3], env))
                except Exception:
                    env.vars.pop(st[1][1], None)
            elif k == "sigdecl" and st[1] == "input":
                for nm, dims in st[2]:
                    found.append((nm, [self.eval_int(x, env) for x in dims]))
        if key is not None:
            self.prescan_cache[key] = found
        return found

    def instantiate(self, tname, params, path) -> Comp:
        if tname not in self.prog.templates:
            raise CircomError(f"unknown template {tname}")
        pnames, _ = self.prog.templates[tname]
        if len(pnames) != len(params):
            raise CircomError(f"template {tname}: expected {len(pnames)} params, got {len(params)}")
        comp = Comp(tname, params, path)
        self.n_components += 1
        for nm, dims in self.prescan(tname, params):
            sig = Signal(nm, "input", dims, comp)
            comp.inputs[nm] = sig
            comp.in_order.append(nm)
            comp.pending += len(sig.vals)
        if comp.pending == 0 and path != "main":
            self.run_template(comp)
        return comp

    def run_template(self, comp: Comp):
        if comp.done:
            raise CircomError(f"component {comp.path} executed twice")
        comp.done = True
        pnames, body = self.prog.templates[comp.tname]
        env = Env(comp)
        for n, v in zip(pnames, comp.params):
            env.vars[n] = v
        try:
            self.exec_block(body, env)
        except ReturnEx:
            raise CircomError("return inside template")

    # ---- statements ------------------------------------------------------
    def exec_block(self, blk, env):
        for st in blk[1]:
            self.exec(st, env)

    def exec(self, st, env: Env):
        k = st[0]
        if k == "subst":
            self.exec_subst(st, env)
        elif k == "sigdecl":
            self.exec_sigdecl(st, env)
        elif k == "vardecl":
            for nm, dims, init in st[1]:
                if init is not None:
                    env.vars[nm] = _copy(self.eval(init, env))
                else:
                    env.vars[nm] = _zeros([self.eval_int(x, env) for x in dims])
        elif k == "compdecl":
            for nm, dims, init in st[1]:
                d = [self.eval_int(x, env) for x in dims]
                env.comps[nm] = _nones(d)
                if init is not None:
                    self.assign_component(env, nm, [], init)
        elif k == "block":
            self.exec_block(st, env)
        elif k == "if":
            c = self.eval(st[1], env)
            if isinstance(c, Sym):
                raise CircomError(f"line {st[4]}: signal-dependent if condition is not supported")
            if c != 0:
                self.exec(st[2], env)
            elif st[3] is not None:
                self.exec(st[3], env)
        elif k == "for":
            self.exec(st[1], env)
            while True:
                c = self.eval(st[2], env)
                if isinstance(c, Sym):
                    raise CircomError(f"line {st[5]}: signal-dependent loop condition")
                if c == 0:
                    break
                self.exec(st[4], env)
                self.exec(st[3], env)
        elif k == "while":
            while True:
                c = self.eval(st[1], env)
                if isinstance(c, Sym):
                    raise CircomError(f"line {st[3]}: signal-dependent loop condition")
                if c == 0:
                    break
                self.exec(st[2], env)
        elif k == "constraint":
            a = self.eval(st[1], env)
            b = self.eval(st[2], env)
            self.add_constraint(a, b, f"{env.comp.path if env.comp else '?'}:{st[3]}")
        elif k == "return":
            raise ReturnEx(self.eval(st[1], env))
        elif k == "assert":
            c = self.eval(st[1], env)
            if not isinstance(c, Sym) and c == 0:
                raise CircomError(f"line {st[2]}: assert failed at compile time")
        elif k == "expr":
            self.eval(st[1], env)
        elif k == "nop":
            pass
        else:
            raise CircomError(f"unknown statement {k}")

    def add_constraint(self, a, b, where):
        if isinstance(a, list) or isinstance(b, list):
            if not (isinstance(a, list) and isinstance(b, list) and len(a) == len(b)):
                raise CircomError(f"{where}: array constraint shape mismatch")
            for x, y in zip(a, b):
                self.add_constraint(x, y, where)
            return
        if not isinstance(a, Sym) and not isinstance(b, Sym):
            if a != b:
                raise CircomError(f"{where}: constant constraint violated")
            return
        self.constraints.append((self.b.node_of(a), self.b.node_of(b), where))

    def exec_sigdecl(self, st, env: Env):
        _, kind, names, init, line = st
        comp = env.comp
        if comp is None:
            raise CircomError("signal declaration outside template")
        for nm, dims in names:
            d = [self.eval_int(x, env) for x in dims]
            if kind == "input":
                sig = comp.inputs.get(nm)
                if sig is None or sig.dims != d:
                    raise CircomError(f"{comp.path}: input {nm} not pre-scanned correctly")
            else:
                sig = Signal(nm, kind, d, comp)
                if kind == "output":
                    comp.out_order.append(nm)
            comp.signals[nm] = sig
            self.signal_order.append(sig)
        if init is not None:
            if len(names) != 1:
                raise CircomError("initialised multi-signal declaration")
            sig = comp.signals[names[0][0]]
            val = self.eval(init[1], env)
            self.assign_signal(sig, 0, sig.dims, val, f"{comp.path}:{line}", init[0])

    def assign_signal(self, sig: Signal, off, dims, val, where, op="<--"):
        if dims:
            if not isinstance(val, list) or len(val) != dims[0]:
                raise CircomError(f"{where}: array shape mismatch assigning {sig.name}")
            stride = _prod(dims[1:])
            for i, v in enumerate(val):
                self.assign_signal(sig, off + i * stride, dims[1:], v, where, op)
            return
        if isinstance(val, (list, tuple)) or val is None:
            raise CircomError(f"{where}: scalar signal {sig.name} assigned a non-scalar")
        if sig.vals[off] is not None:
            raise CircomError(f"{where}: signal {sig.name}[{off}] assigned twice")
        sig.vals[off] = val
        sig.lin[off] = op in ("<==", "==>") and _deg(val) <= 1
        if sig.kind == "input":
            owner = sig.owner
            if owner.pending > 0:
                owner.pending -= 1
                if owner.pending == 0:
                    self.run_template(owner)

    def assign_component(self, env, name, idxs, init_expr):
        if init_expr[0] != "call":
            raise CircomError("component must be initialised with a template call")
        _, tname, args, _ = init_expr
        params = [self.eval(a, env) for a in args]
        path = f"{env.comp.path}.{name}" + "".join(f"[{i}]" for i in idxs)
        comp = self.instantiate(tname, params, path)
        if not idxs:
            env.comps[name] = comp
        else:
            arr = env.comps[name]
            for i in idxs[:-1]:
                arr = arr[i]
            arr[idxs[-1]] = comp

    def exec_subst(self, st, env: Env):
        _, lhs, op, rhs, line = st
        where = f"{env.comp.path if env.comp else 'fn'}:{line}"
        if lhs[0] == "tuple":
            val = self.eval(rhs, env)
            if not isinstance(val, tuple) or len(val) != len(lhs[1]):
                raise CircomError(f"{where}: tuple assignment arity mismatch")
            for l, v in zip(lhs[1], val):
                self.store(l, op, v, env, where)
            return
        if lhs[0] != "var":
            raise CircomError(f"{where}: bad assignment target")
        name = lhs[1]
        # component instantiation  c = T(..) / c[i] = T(..)
        if op == "=" and name in env.comps and all(a[0] == "idx" for a in lhs[2]) and rhs[0] == "call" \
                and rhs[1] in self.prog.templates:
            idxs = [self.eval_int(a[1], env) for a in lhs[2]]
            self.assign_component(env, name, idxs, rhs)
            return
        if op in ("=", "<==", "<--"):
            val = self.eval(rhs, env)
        else:
            cur = self.eval(lhs, env)
            val = self.binop(op[:-1], cur, self.eval(rhs, env))
            op = "="
        self.store(lhs, op, val, env, where)

    def store(self, lhs, op, val, env: Env, where):
        if lhs[0] != "var":
            raise CircomError(f"{where}: bad assignment target")
        name, acc = lhs[1], lhs[2]
        if name == "_":
            return
        if name in env.vars:
            if op != "=":
                raise CircomError(f"{where}: {op} on a var")
            if not acc:
                env.vars[name] = _copy(val)
                return
            arr = env.vars[name]
            idxs = [self.eval_int(a[1], env) for a in acc]
            for i in idxs[:-1]:
                arr = arr[i]
            arr[idxs[-1]] = _copy(val)
            return
        sig, off, dims = self.resolve_signal(name, acc, env, where, allow_tag=True)
        if sig is None:
            return                     # tag assignment: ignored
        if op == "=":
            raise CircomError(f"{where}: '=' on signal {name}")
        self.assign_signal(sig, off, dims, val, where, op)

    def resolve_signal(self, name, acc, env: Env, where, allow_tag=False):
        """name + accessors -> (Signal, flat offset, remaining dims)."""
        comp = env.comp
        i = 0
        if comp is not None and name in comp.signals:
            sig = comp.signals[name]
        elif name in env.comps:
            c = env.comps[name]
            while isinstance(c, list):
                if i >= len(acc) or acc[i][0] != "idx":
                    raise CircomError(f"{where}: component array {name} needs an index")
                c = c[self.eval_int(acc[i][1], env)]
                i += 1
            if c is None:
                raise CircomError(f"{where}: component {name} used before instantiation")
            if i >= len(acc) or acc[i][0] != "mem":
                raise CircomError(f"{where}: component {name} used as a value")
            sname = acc[i][1]
            i += 1
            if sname in c.inputs:
                sig = c.inputs[sname]
            elif sname in c.signals:
                sig = c.signals[sname]
            else:
                if not c.done:
                    raise CircomError(f"{where}: {c.path}.{sname} read before the component ran "
                                      f"(pending inputs: {c.pending})")
                raise CircomError(f"{where}: component {c.path} has no signal {sname}")
        else:
            raise CircomError(f"{where}: unknown name {name}")
        off = 0
        dims = sig.dims
        while i < len(acc):
            a = acc[i]
            if a[0] == "mem":
                if allow_tag and i == len(acc) - 1:
                    return None, 0, []
                raise CircomError(f"{where}: tag access {name}.{a[1]} not supported")
            if not dims:
                raise CircomError(f"{where}: too many indices for {name}")
            ix = self.eval_int(a[1], env)
            if ix < 0 or ix >= dims[0]:
                raise CircomError(f"{where}: index {ix} out of range for {name} (dim {dims[0]})")
            off += ix * _prod(dims[1:])
            dims = dims[1:]
            i += 1
        return sig, off, dims

    # ---- expressions -----------------------------------------------------
    def eval_int(self, e, env) -> int:
        v = self.eval(e, env)
        if isinstance(v, Sym) or isinstance(v, list):
            raise CircomError("expected a compile-time constant")
        return v

    def read_signal(self, sig, off, dims, where):
        if dims:
            stride = _prod(dims[1:])
            return [self.read_signal(sig, off + i * stride, dims[1:], where) for i in range(dims[0])]
        v = sig.vals[off]
        if v is None:
            raise CircomError(f"{where}: signal {sig.owner.path}.{sig.name}[{off}] read before assignment")
        return Sym(v.id, 1) if isinstance(v, Sym) else v

    def binop(self, op, a, b):
        if isinstance(a, list) or isinstance(b, list):
            raise CircomError(f"operator {op} on arrays")
        return self.b.duo(BINOP_TO_DUO[op], a, b)

    def eval(self, e, env: Env):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            name, acc = e[1], e[2]
            if name in env.vars:
                v = env.vars[name]
                for a in acc:
                    if a[0] != "idx":
                        raise CircomError(f"line {e[3]}: member access on var {name}")
                    v = v[self.eval_int(a[1], env)]
                return v
            where = f"{env.comp.path if env.comp else 'fn'}:{e[3]}"
            sig, off, dims = self.resolve_signal(name, acc, env, where)
            return self.read_signal(sig, off, dims, where)
        if k == "bin":
            op = e[1]
            a = self.eval(e[2], env)
            # compile-time short circuit keeps `i < n && arr[i]` patterns safe
            if not isinstance(a, (Sym, list)):
                if op == "&&" and a == 0:
                    return 0
                if op == "||" and a != 0:
                    return 1
            return self.binop(op, a, self.eval(e[3], env))
        if k == "un":
            a = self.eval(e[2], env)
            if e[1] == "-":
                return self.b.uno(0, a)
            if e[1] == "!":
                return self.b.uno(2, a)
            return self.b.uno(3, a)
        if k == "tern":
            c = self.eval(e[1], env)
            if not isinstance(c, Sym):
                return self.eval(e[2], env) if c != 0 else self.eval(e[3], env)
            return self.b.tern(c, self.eval(e[2], env), self.eval(e[3], env))
        if k == "arr":
            return [self.eval(x, env) for x in e[1]]
        if k == "call":
            return self.call(e, env)
        if k == "anon":
            return self.anon(e, env)
        if k == "access":
            v = self.eval(e[1], env)
            for a in e[2]:
                if a[0] != "idx":
                    raise CircomError("member access on expression result")
                v = v[self.eval_int(a[1], env)]
            return v
        if k == "tuple":
            return tuple(self.eval(x, env) for x in e[1])
        raise CircomError(f"unknown expression {k}")

    def call(self, e, env):
        _, name, args, line = e
        if name in self.prog.templates:
            raise CircomError(f"line {line}: template {name} used as a function")
        if name not in self.prog.functions:
            raise CircomError(f"line {line}: unknown function {name}")
        argv = [self.eval(a, env) for a in args]
        try:
            key = (name, _freeze(argv))
        except TypeError:
            key = None
        if key is not None and key in self.fn_cache:
            return self.fn_cache[key]
        pnames, body = self.prog.functions[name]
        fenv = Env(None, True)
        for n, v in zip(pnames, argv):
            fenv.vars[n] = _copy(v)
        try:
            self.exec_block(body, fenv)
        except ReturnEx as r:
            if key is not None:
                self.fn_cache[key] = r.v
            return r.v
        raise CircomError(f"function {name} did not return")

    def anon(self, e, env: Env):
        _, tname, pargs, inputs, line = e
        params = [self.eval(a, env) for a in pargs]
        where = f"{env.comp.path}:{line}"
        comp = self.instantiate(tname, params, f"{env.comp.path}.<{tname}@{line}>")
        named = [nm for nm, _ in inputs if nm is not None]
        if named and len(named) != len(inputs):
            raise CircomError(f"{where}: mixed named/positional anonymous inputs")
        if len(inputs) != len(comp.in_order):
            raise CircomError(f"{where}: {tname} expects {len(comp.in_order)} inputs, got {len(inputs)}")
        vals = [self.eval(x, env) for _, x in inputs]
        for k2, (nm, _) in enumerate(inputs):
            sname = nm if nm is not None else comp.in_order[k2]
            if sname not in comp.inputs:
                raise CircomError(f"{where}: {tname} has no input {sname}")
            sig = comp.inputs[sname]
            self.assign_signal(sig, 0, sig.dims, vals[k2], where)
        if not comp.done:
            raise CircomError(f"{where}: anonymous component {tname} did not run")
        outs = []
        for nm in comp.out_order:
            sig = comp.signals[nm]
            outs.append(self.read_signal(sig, 0, sig.dims, where))
        if len(outs) == 1:
            return outs[0]
        if not outs:
            return None
        return tuple(outs)


# --------------------------------------------------------------------------
# graph::optimize restated (src/graph.rs:358-365)
# --------------------------------------------------------------------------

def _operands(node):
    k = node[0]
    if k == po.K_UNO:
        return (node[2],)
    if k == po.K_DUO:
        return (node[2], node[3])
    if k == po.K_TRES:
        return (node[2], node[3], node[4])
    return ()


def _with_operands(node, ops):
    k = node[0]
    if k == po.K_UNO:
        return (k, node[1], ops[0])
    if k == po.K_DUO:
        return (k, node[1], ops[0], ops[1])
    if k == po.K_TRES:
        return (k, node[1], ops[0], ops[1], ops[2])
    return node


def tree_shake(nodes, outputs):
    """src/graph.rs:430-498: drop nodes unreachable from the outputs; order preserved."""
    used = bytearray(len(nodes))
    for o in outputs:
        used[o] = 1
    for i in range(len(nodes) - 1, -1, -1):
        if used[i]:
            for a in _operands(nodes[i]):
                used[a] = 1
    renum = [0] * len(nodes)
    out_nodes = []
    for i, n in enumerate(nodes):
        if used[i]:
            renum[i] = len(out_nodes)
            out_nodes.append(n)
    out_nodes = [_with_operands(n, [renum[a] for a in _operands(n)]) for n in out_nodes]
    return out_nodes, [renum[o] for o in outputs]


def _random_eval(nodes, rng):
    """src/graph.rs:500-538: Add/Sub/Mul algebraic, everything else a random function of its operands."""
    vals = []
    inputs: Dict[int, int] = {}
    prf: Dict[tuple, int] = {}
    for n in nodes:
        k = n[0]
        if k == po.K_CONST:
            v = n[1]
        elif k == po.K_INPUT:
            v = inputs.get(n[1])
            if v is None:
                v = inputs[n[1]] = rng.randrange(P)
        elif k == po.K_DUO and n[1] in (0, 2, 3):
            a, b = vals[n[2]], vals[n[3]]
            v = a * b % P if n[1] == 0 else (a + b) % P if n[1] == 2 else (a - b) % P
        else:
            key = (k, n[1]) + tuple(vals[a] for a in _operands(n))
            v = prf.get(key)
            if v is None:
                v = prf[key] = rng.randrange(P)
        vals.append(v)
    return vals


def value_numbering(nodes, outputs, rng):
    """src/graph.rs:540-578."""
    vals = _random_eval(nodes, rng)
    first: Dict[int, int] = {}
    renum = []
    for i, v in enumerate(vals):
        renum.append(first.setdefault(v, i))
    nodes = [_with_operands(n, [renum[a] for a in _operands(n)]) for n in nodes]
    return nodes, [renum[o] for o in outputs]


def constants_pass(nodes, rng):
    """src/graph.rs:580-600: nodes with the same value under two random evaluations become constants."""
    va = _random_eval(nodes, rng)
    vb = _random_eval(nodes, rng)
    out = []
    for i, n in enumerate(nodes):
        if n[0] != po.K_CONST and va[i] == vb[i]:
            out.append((po.K_CONST, va[i]))
        else:
            out.append(n)
    return out


def optimize(nodes, outputs, seed=0x5EED):
    rng = random.Random(seed)
    nodes, outputs = tree_shake(nodes, outputs)
    # propagate (src/graph.rs:393-428) already happened in Builder.duo/uno/tern
    nodes, outputs = value_numbering(nodes, outputs, rng)
    nodes = constants_pass(nodes, rng)
    nodes, outputs = tree_shake(nodes, outputs)
    return nodes, outputs


def normalize_layout(nodes, outputs):
    """Put the Input nodes first as one contiguous ascending run (what calc_witness's
    get_inputs_size relies on, src/lib.rs:138-152); everything else keeps its order."""
    order = sorted((i for i, n in enumerate(nodes) if n[0] == po.K_INPUT), key=lambda i: nodes[i][1])
    order += [i for i, n in enumerate(nodes) if n[0] != po.K_INPUT]
    renum = [0] * len(nodes)
    for new, old in enumerate(order):
        renum[old] = new
    out = [_with_operands(nodes[old], [renum[a] for a in _operands(nodes[old])]) for old in order]
    return out, [renum[o] for o in outputs]


def dedupe_outputs(outputs, n_fixed):
    """After value numbering several signals can share a node: keep the first n_fixed
    entries (1, main outputs, main inputs) as they are and drop later duplicates."""
    seen = set()
    res = []
    for i, o in enumerate(outputs):
        if i < n_fixed:
            res.append(o)
            seen.add(o)
        elif o not in seen:
            seen.add(o)
            res.append(o)
    return res


def compile_circuit(path: str, lib_dirs: List[str], check_inputs: Optional[dict] = None, seed=0x5EED, opt_level=2):
    """Returns dict(nodes, witness, input_map, stats).  If check_inputs is given
    ({name: [ints]}), every `===` of the sources is verified on the unoptimised graph."""
    prog = Program(lib_dirs)
    prog.load(path)
    if prog.main is None:
        raise CircomError("no main component")
    ex = Executor(prog, opt_level)
    nodes, wit, input_map = ex.run_main()
    stats = dict(ex.stats)
    stats["nodes_unoptimized"] = len(nodes)
    stats["constraints_eq"] = len(ex.constraints)
    if check_inputs is not None:
        buf = [0] * (1 + sum(ln for _, ln in input_map.values()))
        buf[0] = 1
        for k, vals in check_inputs.items():
            off, ln = input_map[k]
            assert ln == len(vals), f"input {k}: expected {ln} values"
            buf[off:off + ln] = vals
        _, values = po.evaluate(nodes, buf, wit, "circom", return_values=True)
        bad = [(w, values[a], values[b]) for a, b, w in ex.constraints if values[a] != values[b]]
        stats["constraints_violated"] = len(bad)
        stats["violations"] = bad[:10]
    return {"nodes": nodes, "witness": wit, "input_map": input_map, "stats": stats,
            "n_fixed": ex.n_fixed, "executor": ex}


def build_graph(path: str, lib_dirs: List[str], check_inputs=None, seed=0x5EED, opt_level=2):
    res = compile_circuit(path, lib_dirs, check_inputs, seed, opt_level)
    nodes, wit = res["nodes"], res["witness"]
    n_fixed = res["n_fixed"]          # 1 + main outputs + main inputs: never deduplicated
    onodes, owit = optimize(nodes, wit, seed)
    onodes, owit = normalize_layout(onodes, owit)
    owit = dedupe_outputs(owit, n_fixed)
    stats = res["stats"]
    stats["nodes"] = len(onodes)
    stats["witness_len"] = len(owit)
    ops = {}
    for n in onodes:
        if n[0] == po.K_DUO:
            nm = po.DUO_OPS[n[1]]
        elif n[0] == po.K_UNO:
            nm = po.UNO_OPS[n[1]]
        elif n[0] == po.K_TRES:
            nm = "TernCond"
        elif n[0] == po.K_CONST:
            nm = "Constant"
        else:
            nm = "Input"
        ops[nm] = ops.get(nm, 0) + 1
    stats["op_histogram"] = ops
    data = po.serialize_graph(onodes, owit, res["input_map"])
    return data, onodes, owit, res["input_map"], stats


if __name__ == "__main__":
    import argparse
    import json
    import time

    ap = argparse.ArgumentParser(description="circom subset -> wtns.graph.001")
    ap.add_argument("circuit")
    ap.add_argument("out")
    ap.add_argument("-l", action="append", default=[], dest="libs")
    ap.add_argument("--inputs", help="inputs.json used to verify every === of the sources")
    ap.add_argument("--O1", action="store_const", const=1, default=2, dest="opt_level",
                    help="keep every declared signal in the witness (default --O2-like: signals defined by a linear <== are dropped)")
    a = ap.parse_args()
    chk = None
    if a.inputs:
        chk = po.deserialize_inputs(open(a.inputs).read())
    t0 = time.time()
    data, _, _, _, st = build_graph(a.circuit, a.libs, chk, opt_level=a.opt_level)
    with open(a.out, "wb") as f:
        f.write(data)
    st["seconds"] = round(time.time() - t0, 2)
    print(json.dumps(st, default=str))
