"""f95c -- a small Fortran 95 -> C++ translator (TEST INFRASTRUCTURE ONLY).

Why it exists: neither this image nor the GPU boxes hold a Fortran compiler, so the reference cannot be built the
usual way and the CPU oracle (oracle/beom_oracle.c, a restatement written by hand) had nothing of the reference to be
checked against but analytical solutions at per-cent tolerances.  This translator knows nothing about the model: it
parses the language subset the reference is written in -- modules, subroutines with keyword / optional arguments,
declarations with bounds, DO / IF / WHERE, whole-array and array-section expressions, SUM / ANY / ALL / MINVAL /
MAXVAL / SIZE, character variables, unformatted direct-access and list-directed I/O, OpenMP PARALLEL DO -- and emits
the same statements in the same order as C++ (run-time support: f95rt.h).  The reference's sources are read where they
lie (/root/reference) at build time and the output goes to oracle/_ref/ (git-ignored): no reference source is copied
into the repository.  oracle/refbuild.py drives it; tests/test_reference_pin.py compares what the translated
reference computes with the hand-written oracle, bit for bit.

Anything outside the subset stops the translation with the file and line of the statement (no guessing).
"""
from __future__ import annotations

import re
import sys

# ------------------------------------------------------------------------------------------------ source lines


class FError(Exception):
    pass


def _strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def logical_lines(text: str, fname: str):
    """-> [(file, lineno, statement text)]; OpenMP sentinels come out as statements starting with '!$omp'."""
    res = []
    cur, cur_no = None, 0
    omp, omp_no = None, 0
    for no, raw in enumerate(text.split("\n"), 1):
        raw = raw.replace("\t", " ")
        s = raw.strip()
        if s.lower().startswith("!$omp"):
            body = s[5:].strip()
            if body.startswith("&"):
                body = body[1:].strip()
                if omp is None:
                    raise FError("%s:%d: OpenMP continuation without a start" % (fname, no))
                omp += " " + body
            else:
                if omp is not None:
                    res.append((fname, omp_no, "!$omp " + omp))
                omp, omp_no = body, no
            if omp.endswith("&"):
                omp = omp[:-1].rstrip().rstrip(",")
                continue
            res.append((fname, omp_no, "!$omp " + omp))
            omp = None
            continue
        code = _strip_comment(raw).strip()
        if not code:
            continue
        if cur is not None:
            if code.startswith("&"):
                code = code[1:].lstrip()
            cur += " " + code
        else:
            cur, cur_no = code, no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        for part in _split_semicolons(cur):
            if part.strip():
                res.append((fname, cur_no, part.strip()))
        cur = None
    return res


def _split_semicolons(s: str):
    parts, q, buf = [], None, []
    for ch in s:
        if q:
            buf.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            buf.append(ch)
        elif ch == ";":
            parts.append("".join(buf))
            buf = []
        else:
            buf.append(ch)
    parts.append("".join(buf))
    return parts


# ------------------------------------------------------------------------------------------------ tokens

_DOTOPS = ("and", "or", "not", "eqv", "neqv", "true", "false", "eq", "ne", "lt", "le", "gt", "ge")
_TOKEN = re.compile(
    r"""\s*(?:
      (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
    | (?P<dot>\.(?:%s)\.)
    | (?P<num>(?:\d+\.(?!(?:%s)\.)\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)
    | (?P<name>[A-Za-z][A-Za-z0-9_]*)
    | (?P<op>\(/|/\)|\*\*|//|==|/=|<=|>=|=>|::|[-+*/(),:=<>%%])
    )"""
    % ("|".join(_DOTOPS), "|".join(_DOTOPS)),
    re.X | re.I,
)


def tokenize(s: str, where=""):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise FError("%s: cannot tokenize at %r" % (where, s[pos:pos + 30]))
        pos = m.end()
        if m.group("str") is not None:
            t = m.group("str")
            q = t[0]
            toks.append(("str", t[1:-1].replace(q + q, q)))
        elif m.group("dot") is not None:
            toks.append(("op", m.group("dot").lower()))
        elif m.group("num") is not None:
            toks.append(("num", m.group("num").lower()))
        elif m.group("name") is not None:
            toks.append(("name", m.group("name").lower()))
        else:
            toks.append(("op", m.group("op")))
    # "(/" directly followed by something that makes it a division inside parentheses never occurs in the subset
    return toks


# ------------------------------------------------------------------------------------------------ expressions
# AST: ('num', text) ('str', s) ('log', bool) ('name', id) ('ref', id, [args]) ('un', op, e) ('bin', op, l, r)
#      ('cplx', a, b) ('actor', [e]) ('paren', e); args: expr | ('kw', name, expr) | ('sec', lo|None, hi|None) | ('star',)

_REL = {"==": "==", "/=": "/=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
        ".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


class Parser:
    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def at(self, val):
        tok = self.peek()
        return tok[0] in ("op", "name") and tok[1] == val

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise FError("%s: expected %r, found %r" % (self.where, val, self.peek()[1]))

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing, lowest first
    def expr(self):
        e = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            e = ("bin", op, e, self.p_or())
        return e

    def p_or(self):
        e = self.p_and()
        while self.at(".or."):
            self.next()
            e = ("bin", ".or.", e, self.p_and())
        return e

    def p_and(self):
        e = self.p_not()
        while self.at(".and."):
            self.next()
            e = ("bin", ".and.", e, self.p_not())
        return e

    def p_not(self):
        if self.at(".not."):
            self.next()
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        e = self.p_cat()
        if self.peek()[0] == "op" and self.peek()[1] in _REL:
            op = _REL[self.next()[1]]
            e = ("bin", op, e, self.p_cat())
        return e

    def p_cat(self):
        e = self.p_add()
        while self.at("//"):
            self.next()
            e = ("bin", "//", e, self.p_add())
        return e

    def p_add(self):
        if self.peek() == ("op", "-") or self.peek() == ("op", "+"):
            op = self.next()[1]
            e = ("un", op, self.p_mul())
        else:
            e = self.p_mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            e = ("bin", op, e, self.p_mul())
        return e

    def p_mul(self):
        e = self.p_pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            e = ("bin", op, e, self.p_pow())
        return e

    def p_pow(self):
        e = self.p_primary()
        if self.at("**"):
            self.next()
            # right associative; a signed exponent in parentheses is a primary anyway
            if self.peek() in (("op", "-"), ("op", "+")):
                op = self.next()[1]
                r = ("un", op, self.p_pow())
            else:
                r = self.p_pow()
            e = ("bin", "**", e, r)
        return e

    def p_primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.next()
            return ("num", val)
        if kind == "str":
            self.next()
            return ("str", val)
        if kind == "op" and val in (".true.", ".false."):
            self.next()
            return ("log", val == ".true.")
        if kind == "op" and val == "(/":
            self.next()
            items = [self.expr()]
            while self.accept(","):
                items.append(self.expr())
            self.expect("/)")
            return ("actor", items)
        if kind == "op" and val == "(":
            self.next()
            e = self.expr()
            if self.accept(","):
                b = self.expr()
                self.expect(")")
                return ("cplx", e, b)
            self.expect(")")
            return ("paren", e)
        if kind == "name":
            self.next()
            if self.at("("):
                args = self.arglist()
                node = ("ref", val, args)
                if self.at("("):  # substring of an array element: not in the subset
                    raise FError("%s: %s(...)(...) is outside the subset" % (self.where, val))
                return node
            return ("name", val)
        raise FError("%s: unexpected %r in expression" % (self.where, val))

    def arglist(self):
        """'(' already peeked; parses up to the matching ')'."""
        self.expect("(")
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(","):
                continue
            self.expect(")")
            return args

    def arg(self):
        if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
            name = self.next()[1]
            self.next()
            if self.at("*"):
                self.next()
                return ("kw", name, ("star",))
            return ("kw", name, self.expr())
        if self.at("*") and self.peek(1)[1] in (",", ")"):
            self.next()
            return ("star",)
        lo = None
        if not self.at(":"):
            lo = self.expr()
            if not self.at(":"):
                return lo
        self.expect(":")
        hi = None
        if not (self.at(",") or self.at(")")):
            hi = self.expr()
        if self.at(":"):
            raise FError("%s: strided sections are outside the subset" % self.where)
        return ("sec", lo, hi)


# ------------------------------------------------------------------------------------------------ symbols

CTYPE = {"int": "int", "r4": "float", "r8": "double", "log": "bool", "c8": "std::complex<double>", "char": "FStr"}
KINDS = {"r4": "r4", "r8": "r8", "rw": "r8", "i4": "int"}
CXX_RESERVED = {"int", "float", "double", "char", "long", "short", "signed", "unsigned", "do", "if", "else", "for",
                "while", "switch", "case", "default", "break", "continue", "return", "new", "delete", "this", "class",
                "struct", "union", "template", "typename", "auto", "register", "static", "const", "void", "bool",
                "true", "false", "operator", "namespace", "using", "public", "private", "protected", "virtual",
                "friend", "inline", "extern", "goto", "try", "catch", "throw", "enum", "typedef", "sizeof", "and",
                "or", "not", "xor", "main", "unix", "linux", "errno", "stdin", "stdout", "stderr", "NULL", "line"}


def cname(n: str) -> str:
    return n + "_" if n in CXX_RESERVED else n


class Sym:
    def __init__(self, name, ftype, charlen=None, dims=None, attrs=(), init=None, intent=None):
        self.name, self.ftype, self.charlen = name, ftype, charlen
        self.dims = dims  # None = scalar; list of (lo, hi) ASTs, (None, None) = deferred / assumed
        self.attrs, self.init, self.intent = set(attrs), init, intent
        self.dummy = False
        self.is_result = False

    @property
    def rank(self):
        return len(self.dims) if self.dims is not None else 0

    @property
    def ctype(self):
        return CTYPE[self.ftype]

    @property
    def optional(self):
        return "optional" in self.attrs


class Proc:
    def __init__(self, name, kind, args, where):
        self.name, self.kind, self.args, self.where = name, kind, args, where
        self.syms = {}
        self.body = []
        self.result = None


INTRINSIC_ELEMENTAL = {"abs", "max", "min", "sqrt", "cos", "sin", "exp", "mod", "sign", "real", "int", "nint", "aimag",
                       "trim", "len_trim", "adjustl", "tiny", "huge"}
INTRINSIC_REDUCE = {"sum", "minval", "maxval", "any", "all", "count"}


# ------------------------------------------------------------------------------------------------ translation unit


class Unit:
    def __init__(self):
        self.module_syms = {}      # name -> Sym (all modules share one namespace here: private_mod uses all of shared_mod)
        self.module_order = []     # declaration order
        self.procs = {}            # name -> Proc
        self.proc_order = []
        self.program = None        # Proc of the main program
        self.dump_names = []       # module variables that go into the state dump
        self.uid = 0

    # ------------------------------------------------------------------ parsing of files
    def parse_file(self, text, fname, dump=False):
        lines = logical_lines(text, fname)
        i = 0
        cur = None  # current Proc (or None at module level)
        stack = []  # block nesting inside a procedure body: lists being filled
        in_module = False
        while i < len(lines):
            fn, no, s = lines[i]
            i += 1
            where = "%s:%d" % (fn, no)
            if s.startswith("!$omp"):
                if cur is not None:
                    stack[-1].append(("omp", where, s[5:].strip().lower()))
                continue
            toks = tokenize(s, where)
            k0 = toks[0][1] if toks[0][0] == "name" else None
            k1 = toks[1][1] if len(toks) > 1 and toks[1][0] == "name" else None
            # ---- program units
            if cur is None:
                if k0 == "module" and len(toks) == 2:
                    in_module = True
                    continue
                if k0 == "end" and k1 in ("module", "program"):
                    in_module = False
                    continue
                if k0 == "program":
                    cur = Proc("main_program", "program", [], where)
                    self.program = cur
                    stack = [cur.body]
                    continue
                if k0 in ("use", "implicit", "private", "public", "contains"):
                    continue
                if k0 in ("subroutine", "function") or (k0 == "recursive"):
                    p = Parser(toks, where)
                    kind = p.next()[1]
                    name = p.next()[1]
                    args = []
                    if p.at("("):
                        for a in p.arglist():
                            if a[0] != "name":
                                raise FError("%s: dummy argument list" % where)
                            args.append(a[1])
                    cur = Proc(name, kind, args, where)
                    if kind == "function":
                        cur.result = name
                    self.procs[name] = cur
                    self.proc_order.append(name)
                    stack = [cur.body]
                    continue
                if self._is_decl(toks):
                    for sym in self._parse_decl(toks, where):
                        if sym.name in self.module_syms:
                            raise FError("%s: %s declared twice" % (where, sym.name))
                        self.module_syms[sym.name] = sym
                        self.module_order.append(sym.name)
                        if dump and "parameter" not in sym.attrs and sym.ftype in ("int", "r4", "r8", "log"):
                            self.dump_names.append(sym.name)
                    continue
                raise FError("%s: unexpected statement at module level: %s" % (where, s))
            # ---- inside a procedure
            if k0 == "end" and k1 in ("subroutine", "function", "program"):
                if len(stack) != 1:
                    raise FError("%s: unclosed block at end of %s" % (where, cur.name))
                cur = None
                continue
            if k0 in ("use", "implicit"):
                continue
            if self._is_decl(toks) and not self._looks_like_assignment(toks):
                for sym in self._parse_decl(toks, where):
                    cur.syms[sym.name] = sym
                continue
            self._parse_exec(toks, where, stack, s)
        if cur is not None:
            raise FError("%s: file ends inside %s" % (fname, cur.name))

    @staticmethod
    def _is_decl(toks):
        if toks[0][0] != "name":
            return False
        if toks[0][1] in ("integer", "real", "logical", "character", "complex"):
            return any(t == ("op", "::") for t in toks)
        return False

    @staticmethod
    def _looks_like_assignment(toks):
        return len(toks) > 1 and toks[1] == ("op", "=")

    def _parse_decl(self, toks, where):
        p = Parser(toks, where)
        base = p.next()[1]
        ftype, charlen = None, None
        if base == "integer":
            ftype = "int"
            if p.at("("):
                p.arglist()
        elif base == "logical":
            ftype = "log"
        elif base == "real":
            ftype = "r4"
            if p.at("("):
                a = p.arglist()
                k = a[0][2] if a[0][0] == "kw" else a[0]
                if k[0] != "name" or k[1] not in KINDS:
                    raise FError("%s: real kind %r" % (where, k))
                ftype = KINDS[k[1]]
        elif base == "complex":
            ftype = "c8"
            if p.at("("):
                p.arglist()
        elif base == "character":
            ftype = "char"
            charlen = ("num", "1")
            if p.at("("):
                a = p.arglist()
                charlen = a[0][2] if a[0][0] == "kw" else a[0]
        attrs, intent, dimattr = [], None, None
        while p.accept(","):
            a = p.next()[1]
            if a == "intent":
                intent = "".join(x[1] for x in p.arglist() if x[0] == "name")
            elif a == "dimension":
                dimattr = self._dims(p.arglist(), where)
            else:
                attrs.append(a)
        p.expect("::")
        syms = []
        while True:
            name = p.next()[1]
            dims = dimattr
            if p.at("("):
                dims = self._dims(p.arglist(), where)
            init = None
            if p.accept("="):
                init = p.expr()
            syms.append(Sym(name, ftype, charlen, dims, attrs, init, intent))
            if not p.accept(","):
                break
        if not p.done():
            raise FError("%s: trailing tokens in declaration" % where)
        return syms

    @staticmethod
    def _dims(args, where):
        dims = []
        for a in args:
            if a[0] == "sec":
                dims.append((a[1], a[2]))
            elif a[0] in ("kw", "star"):
                raise FError("%s: array spec" % where)
            else:
                dims.append((("num", "1"), a))
        return dims

    # ------------------------------------------------------------------ executable statements -> nested lists
    def _parse_exec(self, toks, where, stack, text):
        p = Parser(toks, where)
        k0 = toks[0][1] if toks[0][0] == "name" else None
        k1 = toks[1][1] if len(toks) > 1 and toks[1][0] == "name" else None
        out = stack[-1]

        def is_assign():
            if toks[0][0] != "name":
                return False
            if len(toks) > 1 and toks[1] == ("op", "="):
                return True
            if len(toks) > 1 and toks[1] == ("op", "("):
                depth = 0
                for j in range(1, len(toks)):
                    if toks[j] in (("op", "("), ("op", "(/")):
                        depth += 1
                    elif toks[j] in (("op", ")"), ("op", "/)")):
                        depth -= 1
                        if depth == 0:
                            return j + 1 < len(toks) and toks[j + 1] == ("op", "=")
            return False

        if is_assign() and k0 not in ("if", "elseif", "where"):
            lhs = p.p_primary()
            p.expect("=")
            rhs = p.expr()
            if not p.done():
                raise FError("%s: trailing tokens after assignment" % where)
            out.append(("assign", where, lhs, rhs))
            return
        # block closers / continuations
        if (k0 == "end" and k1 == "if") or k0 == "endif":
            blk = stack.pop()
            return
        if (k0 == "end" and k1 == "do") or k0 == "enddo":
            stack.pop()
            return
        if k0 == "else" and k1 == "if" or k0 == "elseif":
            p.next()
            if k0 == "else":
                p.next()
            p.expect("(")
            cond = p.expr()
            p.expect(")")
            p.expect("then")
            stack.pop()
            ifnode = stack[-1][-1]
            body = []
            ifnode[2].append((cond, body))
            stack.append(body)
            return
        if k0 == "else" and len(toks) == 1:
            stack.pop()
            ifnode = stack[-1][-1]
            body = []
            ifnode[2].append((None, body))
            stack.append(body)
            return
        if k0 == "if":
            p.next()
            p.expect("(")
            cond = p.expr()
            p.expect(")")
            if p.accept("then"):
                body = []
                out.append(("if", where, [(cond, body)]))
                stack.append(body)
                return
            body = []
            out.append(("if", where, [(cond, body)]))
            stack.append(body)
            self._parse_exec(toks[p.i:], where, stack, text)
            stack.pop()
            return
        if k0 == "where":
            p.next()
            p.expect("(")
            cond = p.expr()
            p.expect(")")
            if p.done():
                raise FError("%s: WHERE blocks are outside the subset" % where)
            lhs = p.p_primary()
            p.expect("=")
            rhs = p.expr()
            out.append(("where", where, cond, lhs, rhs))
            return
        if k0 == "do":
            p.next()
            if p.done():
                body = []
                out.append(("doforever", where, body))
                stack.append(body)
                return
            if p.at("while"):
                p.next()
                p.expect("(")
                cond = p.expr()
                p.expect(")")
                body = []
                out.append(("dowhile", where, cond, body))
                stack.append(body)
                return
            var = p.next()[1]
            p.expect("=")
            a = p.expr()
            p.expect(",")
            b = p.expr()
            c = None
            if p.accept(","):
                c = p.expr()
            body = []
            out.append(("do", where, var, a, b, c, body))
            stack.append(body)
            return
        if k0 == "call":
            p.next()
            name = p.next()[1]
            args = p.arglist() if p.at("(") else []
            out.append(("call", where, name, args))
            return
        if k0 in ("exit", "cycle", "return", "stop", "continue") and len(toks) == 1:
            out.append((k0, where))
            return
        if k0 in ("allocate", "deallocate"):
            p.next()
            out.append((k0, where, p.arglist()))
            return
        if k0 in ("open", "close", "inquire", "read", "write"):
            p.next()
            ctl = p.arglist()
            items = []
            if not p.done():
                items.append(p.expr())
                while p.accept(","):
                    items.append(p.expr())
            out.append(("io", where, k0, ctl, items))
            return
        raise FError("%s: statement outside the subset: %s" % (where, text))

    # ------------------------------------------------------------------ emission
    def fresh(self, stem):
        self.uid += 1
        return "%s%d_" % (stem, self.uid)


class Scope:
    def __init__(self, unit: Unit, proc: Proc | None):
        self.unit, self.proc = unit, proc

    def sym(self, name):
        if self.proc is not None and name in self.proc.syms:
            return self.proc.syms[name]
        return self.unit.module_syms.get(name)


class Emitter:
    def __init__(self, unit: Unit, uninit=None, env_params=(), trace=(), externs=(), hooks_include=None):
        self.u = unit
        self.externs = set(externs)          # procedures defined in C++ by hooks_include (the drop-in build's binding)
        self.hooks_include = hooks_include   # included inside namespace ref after the module variables
        self.uninit = dict(uninit or {})
        self.env_params = set(env_params)  # module scalars / strings a run may override through F95_<NAME> (harness)
        self.trace = set(trace)            # procedures whose entries are time-stamped (harness)
        self.out = []
        self.scope = Scope(unit, None)
        self.ind = 0

    def w(self, s=""):
        self.out.append("  " * self.ind + s)

    # ---------------------------------------------------------- expressions
    def rank(self, e) -> int:
        k = e[0]
        if k in ("num", "str", "log", "cplx"):
            return 0
        if k == "name":
            s = self.scope.sym(e[1])
            return s.rank if s is not None else 0
        if k == "paren":
            return self.rank(e[1])
        if k == "un":
            return self.rank(e[2])
        if k == "bin":
            return max(self.rank(e[2]), self.rank(e[3]))
        if k == "actor":
            return 1
        if k == "ref":
            name, args = e[1], e[2]
            s = self.scope.sym(name)
            if s is not None and s.rank > 0:
                return sum(1 for a in args if a[0] == "sec")
            if s is not None and s.ftype == "char":
                return 0
            if name in INTRINSIC_REDUCE:
                dim = [a for a in args if a[0] == "kw" and a[1] == "dim"]
                if dim:
                    return self.rank(args[0]) - 1
                return 0
            if name in ("size", "present", "allocated"):
                return 0
            return max([self.rank(a[2] if a[0] == "kw" else a) for a in args if a[0] != "sec"] + [0])
        raise FError("rank of %r" % (e,))

    def first_leaf(self, e):
        """First array-valued leaf (whole array or sectioned reference) of an elemental expression."""
        k = e[0]
        if k == "name":
            s = self.scope.sym(e[1])
            return e if (s is not None and s.rank > 0) else None
        if k == "paren":
            return self.first_leaf(e[1])
        if k == "un":
            return self.first_leaf(e[2])
        if k == "bin":
            return self.first_leaf(e[2]) or self.first_leaf(e[3])
        if k == "ref":
            s = self.scope.sym(e[1])
            if s is not None and s.rank > 0:
                return e if any(a[0] == "sec" for a in e[2]) else None
            if e[1] in INTRINSIC_REDUCE or e[1] in ("size",):
                if any(a[0] == "kw" and a[1] == "dim" for a in e[2]) and e[1] in INTRINSIC_REDUCE:
                    return ("reduced", e)
                return None
            for a in e[2]:
                a = a[2] if a[0] == "kw" else a
                if a[0] in ("sec", "star"):
                    continue
                r = self.first_leaf(a)
                if r:
                    return r
        return None

    def leaf_dims(self, leaf):
        """[(lo_cxx, extent_cxx)] of the section dimensions of a leaf, in order."""
        if leaf[0] == "reduced":
            inner = leaf[1]
            dim = [a for a in inner[2] if a[0] == "kw" and a[1] == "dim"][0][2]
            d = int(dim[1])
            src = self.leaf_dims(self.first_leaf(inner[2][0]))
            return src[:d - 1] + src[d:]
        if leaf[0] == "name":
            s = self.scope.sym(leaf[1])
            n = self.vname(leaf[1])
            return [("%s.lb(%d)" % (n, d + 1), "%s.ext(%d)" % (n, d + 1)) for d in range(s.rank)]
        name, args = leaf[1], leaf[2]
        n = self.vname(name)
        dims = []
        for d, a in enumerate(args):
            if a[0] != "sec":
                continue
            lo = self.ex(a[1]) if a[1] is not None else "%s.lb(%d)" % (n, d + 1)
            hi = self.ex(a[2]) if a[2] is not None else "%s.ub(%d)" % (n, d + 1)
            dims.append((lo, "((long)(%s) - (long)(%s) + 1)" % (hi, lo)))
        return dims

    def vname(self, name):
        """C++ spelling of a variable reference."""
        s = self.scope.sym(name)
        if s is not None and s.dummy and s.optional and s.rank == 0:
            return "(*%s__p)" % cname(name)
        if s is not None and s.is_result:
            return cname(name) + "__r"
        return cname(name)

    def num(self, text):
        m = re.match(r"^([0-9.]+(?:[ed][+-]?\d+)?)(?:_(\w+))?$", text)
        if not m:
            raise FError("number %r" % text)
        body, kind = m.group(1), m.group(2)
        is_real = ("." in body) or ("e" in body) or ("d" in body)
        if not is_real:
            return str(int(body))
        if "d" in body:
            return body.replace("d", "e")
        if kind is None or KINDS.get(kind) == "r4":
            if "." not in body and "e" in body:
                body = body.replace("e", ".e")
            return body + "f"
        if KINDS.get(kind) == "r8":
            return body
        raise FError("kind %r" % kind)

    def ex(self, e, lv=None):
        """lv: loop variables bound to the section dimensions of array leaves (None = scalar context)."""
        k = e[0]
        if k == "num":
            return self.num(e[1])
        if k == "str":
            return 'std::string("%s")' % e[1].replace("\\", "\\\\").replace('"', '\\"')
        if k == "log":
            return "true" if e[1] else "false"
        if k == "cplx":
            return "std::complex<double>(%s, %s)" % (self.ex(e[1]), self.ex(e[2]))
        if k == "paren":
            return "(%s)" % self.ex(e[1], lv)
        if k == "name":
            s = self.scope.sym(e[1])
            if s is None:
                raise FError("undeclared name %r" % e[1])
            if s.rank > 0:
                if lv is None:
                    return self.vname(e[1])  # whole array as an actual argument / I/O item
                n = self.vname(e[1])
                idx = ["(%s.lb(%d) + %s)" % (n, d + 1, lv[d]) for d in range(s.rank)]
                return "%s(%s)" % (n, ", ".join(idx))
            return self.vname(e[1])
        if k == "un":
            op = {"-": "-", "+": "+", ".not.": "!"}[e[1]]
            return "(%s(%s))" % (op, self.ex(e[2], lv))
        if k == "bin":
            op, l, r = e[1], self.ex(e[2], lv), self.ex(e[3], lv)
            if op == "**":
                return "f_pow(%s, %s)" % (l, r)
            if op == "//":
                return "f_cat(%s, %s)" % (l, r)
            if op == "==":
                return "f_eq(%s, %s)" % (l, r)
            if op == "/=":
                return "f_ne(%s, %s)" % (l, r)
            cop = {".and.": "&&", ".or.": "||", ".eqv.": "==", ".neqv.": "!="}.get(op, op)
            return "(%s %s %s)" % (l, cop, r)
        if k == "ref":
            return self.ex_ref(e, lv)
        raise FError("expression %r" % (e,))

    def ex_ref(self, e, lv):
        name, args = e[1], e[2]
        s = self.scope.sym(name)
        if s is not None and s.rank > 0:
            n = self.vname(name)
            idx, m = [], 0
            for d, a in enumerate(args):
                if a[0] == "sec":
                    if lv is None:
                        raise FError("array section of %s in a scalar context" % name)
                    lo = self.ex(a[1]) if a[1] is not None else "%s.lb(%d)" % (n, d + 1)
                    idx.append("(%s + %s)" % (lo, lv[m]))
                    m += 1
                else:
                    idx.append(self.ex(a, lv=None) if self.rank(a) == 0 else self.ex(a, lv))
            if len(idx) != s.rank:
                raise FError("%s: %d subscripts for rank %d" % (name, len(idx), s.rank))
            return "%s(%s)" % (n, ", ".join(idx))
        if s is not None and s.ftype == "char":
            if len(args) == 1 and args[0][0] == "sec":
                lo = self.ex(args[0][1]) if args[0][1] is not None else "1"
                hi = self.ex(args[0][2]) if args[0][2] is not None else "(long)std::string(%s).size()" % self.vname(name)
                return "f_substr(%s, %s, %s)" % (self.vname(name), lo, hi)
            raise FError("character reference %s(...)" % name)
        pos = [a for a in args if a[0] != "kw"]
        kw = {a[1]: a[2] for a in args if a[0] == "kw"}
        if name in ("real", "int", "nint"):
            kind = kw.get("kind") or (pos[1] if len(pos) > 1 else None)
            x = self.ex(pos[0], lv)
            if name == "real":
                if kind is None:
                    return "f_real(%s)" % x
                return "((%s)(%s))" % (CTYPE[KINDS[kind[1]]], x)
            if name == "int":
                return "((int)(%s))" % x
            return "f_nint(%s)" % x
        if name in ("abs", "max", "min", "sqrt", "cos", "sin", "exp", "mod", "sign", "aimag", "trim", "len_trim",
                    "adjustl", "tiny", "huge", "floor", "ceiling", "log", "tanh", "atan"):
            return "f_%s(%s)" % (name, ", ".join(self.ex(a, lv) for a in pos))
        if name == "present":
            return "(%s__p != nullptr)" % cname(pos[0][1])
        if name == "allocated":
            return "%s.allocated()" % self.vname(pos[0][1])
        if name == "size":
            leaf = pos[0]
            dims = self.leaf_dims(leaf if leaf[0] in ("name", "ref") else self.first_leaf(leaf))
            if "dim" in kw or len(pos) > 1:
                d = int((kw.get("dim") or pos[1])[1])
                return "((int)(%s))" % dims[d - 1][1]
            return "((int)(%s))" % " * ".join("(%s)" % x[1] for x in dims)
        if name in INTRINSIC_REDUCE:
            return self.ex_reduce(name, pos, kw, lv)
        if name in self.u.procs and self.u.procs[name].kind == "function":
            return "%s(%s)" % (cname(name), ", ".join(self.call_args(name, args)))
        raise FError("unknown function or array %r" % name)

    def ex_reduce(self, name, pos, kw, outer_lv):
        arg = pos[0]
        mask = kw.get("mask") or (pos[2] if len(pos) > 2 else None)
        dim = kw.get("dim") or (pos[1] if len(pos) > 1 and name not in ("any", "all", "count") else None)
        if name in ("any", "all", "count") and len(pos) > 1:
            dim = pos[1]
        leaf = self.first_leaf(arg) or (self.first_leaf(mask) if mask else None)
        if leaf is None:
            raise FError("%s() of a scalar" % name)
        dims = self.leaf_dims(leaf)
        n = len(dims)
        vars_ = [self.u.fresh("r") for _ in range(n)]
        if dim is not None:
            d = int(dim[1])
            if outer_lv is None or len(outer_lv) != n - 1:
                raise FError("%s(dim=) outside a matching array context" % name)
            inner = list(outer_lv[:d - 1]) + [vars_[d - 1]] + list(outer_lv[d - 1:])
            loop_ids = [d - 1]
        else:
            inner = vars_
            loop_ids = list(range(n))
        body = self.ex(arg, inner)
        mk = self.ex(mask, inner) if mask is not None else None
        acc = self.u.fresh("acc")
        # declare the loop variables first so that decltype() of the element expression is well formed
        decl = "".join("long %s = 0; " % inner[i] for i in loop_ids)
        loops_open, loops_close = "", ""
        for i in reversed(loop_ids):  # first dimension innermost: array-element order
            loops_open += "for (%s = 0; %s < (long)(%s); ++%s) { " % (inner[i], inner[i], dims[i][1], inner[i])
            loops_close += "} "
        guard = "if (%s) " % mk if mk else ""
        ty = "std::decay_t<decltype(%s)>" % body
        if name == "sum":
            core = "%s %s = 0; %s%s%s = %s + (%s); %s" % (ty, acc, loops_open, guard, acc, acc, body, loops_close)
        elif name == "minval":
            core = "%s %s = std::numeric_limits<%s>::max(); %s%sif ((%s) < %s) %s = (%s); %s" % (
                ty, acc, ty, loops_open, guard, body, acc, acc, body, loops_close)
        elif name == "maxval":
            core = "%s %s = std::numeric_limits<%s>::lowest(); %s%sif ((%s) > %s) %s = (%s); %s" % (
                ty, acc, ty, loops_open, guard, body, acc, acc, body, loops_close)
        elif name == "any":
            core = "bool %s = false; %sif (%s) %s = true; %s" % (acc, loops_open, body, acc, loops_close)
        elif name == "all":
            core = "bool %s = true; %sif (!(%s)) %s = false; %s" % (acc, loops_open, body, acc, loops_close)
        else:
            core = "int %s = 0; %sif (%s) %s += 1; %s" % (acc, loops_open, body, acc, loops_close)
        return "[&]{ %s%sreturn %s; }()" % (decl, core, acc)

    def call_args(self, name, args):
        proc = self.u.procs.get(name)
        if proc is None:
            raise FError("call of unknown procedure %r" % name)
        actual = {}
        for i, a in enumerate(args):
            if a[0] == "kw":
                if a[1] not in proc.args:
                    raise FError("%s has no dummy argument %s" % (name, a[1]))
                actual[a[1]] = a[2]
            else:
                actual[proc.args[i]] = a
        out = []
        for d in proc.args:
            ds = proc.syms[d]
            a = actual.get(d)
            if a is None:
                if not ds.optional:
                    raise FError("call %s: argument %s missing" % (name, d))
                out.append("nullptr")
                continue
            if ds.optional:
                # the actual may itself be an optional dummy of the caller
                if a[0] == "name":
                    cs = self.scope.sym(a[1])
                    if cs is not None and cs.dummy and cs.optional:
                        out.append("%s__p" % cname(a[1]))
                        continue
                    if ds.rank > 0 or ds.intent != "in":
                        out.append("&%s" % self.ex(a))
                        continue
                tmp = self.ex(a)
                out.append("&(const %s&)(%s)" % (ds.ctype, tmp) if ds.rank == 0 else "&%s" % tmp)
                continue
            out.append(self.ex(a))
        return out

    # ---------------------------------------------------------- declarations
    def bounds(self, dims):
        out = []
        for lo, hi in dims:
            if hi is None:
                return None
            out.append("B(%s, %s)" % (self.ex(lo) if lo is not None else "1", self.ex(hi)))
        return ", ".join(out)

    def declare(self, s: Sym, static: bool):
        n = cname(s.name)
        st = "static " if static else ""
        if s.ftype == "char":
            ln = self.ex(s.charlen)
            if s.rank:
                raise FError("character arrays are outside the subset (%s)" % s.name)
            if s.init is not None:
                init = self.ex(s.init)
                if self.scope.proc is None and s.name in self.env_params:
                    init = 'env_str("F95_%s", %s)' % (s.name.upper(), init)
                self.w("%sFStr %s(%s, %s);" % (st, n, ln, init))
            else:
                self.w("%sFStr %s(%s);" % (st, n, ln))
            return
        if s.rank == 0:
            const = "const " if "parameter" in s.attrs else ""
            if s.init is not None:
                init = self.ex(s.init)
                if self.scope.proc is None and s.name in self.env_params:
                    init = 'env_num("F95_%s", %s)' % (s.name.upper(), init)
                self.w("%s%s%s %s = %s;" % (st, const, s.ctype, n, init))
            else:
                # Fortran leaves an unassigned local undefined; here it is 0 unless the caller names the value it is to hold
                v0 = self.uninit.get(s.name) if self.scope.proc is not None else None
                self.w("%s%s %s = %s;" % (st, s.ctype, n, v0 or ("0" if s.ftype != "c8" else "0.0")))
            return
        b = self.bounds(s.dims)
        if b is None:  # allocatable
            self.w("%sArr<%s, %d> %s;" % (st, s.ctype, s.rank, n))
        else:
            self.w("%sArr<%s, %d> %s(%s);" % (st, s.ctype, s.rank, n, b))
        if s.init is not None:
            if s.init[0] == "actor":
                lo = self.ex(s.dims[0][0]) if s.dims[0][0] is not None else "1"
                sets = "; ".join("%s(%s + %d) = %s" % (n, lo, i, self.ex(v)) for i, v in enumerate(s.init[1]))
                self.w("%sconst int %s__i = [%s]{ %s; return 0; }();" % (st, n, "" if static else "&", sets))
            else:
                self.w("%sconst int %s__i = (%s.fill(%s), 0);" % (st, n, n, self.ex(s.init)))

    # ---------------------------------------------------------- statements
    def loops(self, dims, lv):
        """Open nested loops over section dimensions (first dimension innermost); returns the closing count."""
        for i in reversed(range(len(dims))):
            self.w("for (long %s = 0; %s < (long)(%s); ++%s) {" % (lv[i], lv[i], dims[i][1], lv[i]))
            self.ind += 1
        return len(dims)

    def close(self, n):
        for _ in range(n):
            self.ind -= 1
            self.w("}")

    def stmt(self, st):
        k, where = st[0], st[1]
        try:
            getattr(self, "s_" + k)(st)
        except FError as err:
            raise FError("%s: %s" % (where, err)) from None

    def block(self, body):
        pend = None
        for st in body:
            if st[0] == "omp":
                d = st[2]
                if d.startswith("parallel do"):
                    pend = "#pragma omp parallel for" + d[len("parallel do"):]
                elif d.startswith("end parallel do"):
                    pend = None
                else:
                    raise FError("%s: OpenMP directive outside the subset: %s" % (st[1], d))
                continue
            if pend is not None:
                if st[0] != "do":
                    raise FError("%s: PARALLEL DO not followed by a DO loop" % st[1])
                self.s_do(st, pragma=pend)
                pend = None
                continue
            self.stmt(st)

    def s_assign(self, st):
        _, where, lhs, rhs = st
        r = self.rank(lhs)
        if r == 0:
            ls = self.scope.sym(lhs[1])
            if ls is None:
                raise FError("assignment to undeclared %r" % lhs[1])
            self.w("%s = %s;" % (self.ex(lhs), self.ex(rhs)))
            return
        if self.rank(rhs) not in (0, r):
            raise FError("rank mismatch in array assignment")
        dims = self.leaf_dims(lhs)
        lv = [self.u.fresh("k") for _ in dims]
        self.w("{")
        self.ind += 1
        n = self.loops(dims, lv)
        self.w("%s = %s;" % (self.ex(lhs, lv), self.ex(rhs, lv)))
        self.close(n)
        self.ind -= 1
        self.w("}")

    def s_where(self, st):
        _, where, cond, lhs, rhs = st
        dims = self.leaf_dims(self.first_leaf(cond))
        lv = [self.u.fresh("k") for _ in dims]
        self.w("{")
        self.ind += 1
        n = self.loops(dims, lv)
        self.w("if (%s) %s = %s;" % (self.ex(cond, lv), self.ex(lhs, lv), self.ex(rhs, lv)))
        self.close(n)
        self.ind -= 1
        self.w("}")

    def s_if(self, st):
        for i, (cond, body) in enumerate(st[2]):
            if i == 0:
                self.w("if (%s) {" % self.ex(cond))
            elif cond is not None:
                self.w("} else if (%s) {" % self.ex(cond))
            else:
                self.w("} else {")
            self.ind += 1
            self.block(body)
            self.ind -= 1
        self.w("}")

    def s_do(self, st, pragma=None):
        _, where, var, a, b, c, body = st
        v = self.vname(var)
        n, k, a_, s_ = self.u.fresh("n"), self.u.fresh("t"), self.u.fresh("a"), self.u.fresh("s")
        self.w("{")
        self.ind += 1
        self.w("const long %s = %s, %s = %s;" % (a_, self.ex(a), s_, self.ex(c) if c is not None else "1"))
        self.w("const long %s = ((long)(%s) - %s + %s) / %s > 0 ? ((long)(%s) - %s + %s) / %s : 0;" % (
            n, self.ex(b), a_, s_, s_, self.ex(b), a_, s_, s_))
        if pragma:
            pragma = re.sub(r"private\(([^)]*)\)",
                            lambda m: "private(%s)" % ", ".join(cname(x.strip()) for x in m.group(1).split(",")), pragma)
            pragma = pragma.replace(",", ", ").replace("),  private", ") private").replace(") ,", ")")
            pragma = re.sub(r"\)\s*,\s*private", ") private", pragma)
            self.out.append(pragma)
        self.w("for (long %s = 0; %s < %s; ++%s) {" % (k, k, n, k))
        self.ind += 1
        self.w("%s = (int)(%s + %s * %s);" % (v, a_, k, s_))
        self.block(body)
        self.ind -= 1
        self.w("}")
        self.w("%s = (int)(%s + %s * %s);" % (v, a_, n, s_))
        self.ind -= 1
        self.w("}")

    def s_doforever(self, st):
        self.w("for (;;) {")
        self.ind += 1
        self.block(st[2])
        self.ind -= 1
        self.w("}")

    def s_dowhile(self, st):
        self.w("while (%s) {" % self.ex(st[2]))
        self.ind += 1
        self.block(st[3])
        self.ind -= 1
        self.w("}")

    def s_call(self, st):
        _, where, name, args = st
        if name in self.externs:
            self.w("%s(%s);" % (cname(name), ", ".join(self.ex(a) for a in args)))
            return
        self.w("%s(%s);" % (cname(name), ", ".join(self.call_args(name, args))))

    def s_exit(self, st):
        self.w("break;")

    def s_cycle(self, st):
        self.w("continue;")

    def s_continue(self, st):
        self.w(";")

    def s_return(self, st):
        p = self.scope.proc
        if p is not None and p.kind == "function":
            self.w("return %s__r;" % cname(p.name))
        else:
            self.w("return;")

    def s_stop(self, st):
        self.w("f95_stop();")

    def s_allocate(self, st):
        for a in st[2]:
            if a[0] != "ref":
                raise FError("allocate item")
            s = self.scope.sym(a[1])
            b = []
            for d in a[2]:
                if d[0] == "sec":
                    b.append("B(%s, %s)" % (self.ex(d[1]), self.ex(d[2])))
                else:
                    b.append("B(1, %s)" % self.ex(d))
            if s is None or len(b) != s.rank:
                raise FError("allocate(%s)" % a[1])
            self.w("%s.allocate(%s);" % (self.vname(a[1]), ", ".join(b)))

    def s_deallocate(self, st):
        for a in st[2]:
            self.w("%s.deallocate();" % self.vname(a[1]))

    def s_io(self, st):
        _, where, kind, ctl, items = st
        io = self.u.fresh("io")
        pos = [a for a in ctl if a[0] != "kw"]
        kw = {a[1]: a[2] for a in ctl if a[0] == "kw"}
        if pos:
            kw.setdefault("unit", pos[0])
        if len(pos) > 1:
            kw.setdefault("fmt", pos[1])
        self.w("{")
        self.ind += 1
        self.w("Io %s;" % io)
        internal = None
        if "unit" in kw:
            us = self.scope.sym(kw["unit"][1]) if kw["unit"][0] == "name" else None
            if us is not None and us.ftype == "char":
                internal = kw["unit"]
            else:
                self.w("%s.unit = %s;" % (io, self.ex(kw["unit"])))
        for key in ("file", "form", "access", "status", "action", "position"):
            if key in kw and kind == "open":
                self.w("%s.%s = %s;" % (io, key, self.ex(kw[key])))
        if "status" in kw and kind == "close":
            pass
        if "recl" in kw:
            self.w("%s.recl = %s;" % (io, self.ex(kw["recl"])))
        if "rec" in kw:
            self.w("%s.rec = %s; %s.has_rec = true;" % (io, self.ex(kw["rec"]), io))
        if "iostat" in kw:
            self.w("%s.has_iostat = true;" % io)
        if "fmt" in kw and kw["fmt"][0] == "str":
            self.w("%s.fmt = %s;" % (io, self.ex(kw["fmt"])))
        if kind == "open":
            self.w("f_open(%s);" % io)
        elif kind == "close":
            self.w("f_close(%s);" % io)
        elif kind == "inquire":
            if "iolength" in kw:
                terms = []
                for it in items:
                    leaf = it
                    if leaf[0] == "ref":
                        leaf = ("name", leaf[1])
                    terms.append("(long)%s.bytes()" % self.vname(leaf[1]))
                self.w("%s = (int)(%s);" % (self.ex(kw["iolength"]), " + ".join(terms)))
            elif "exist" in kw:
                self.w("%s = file_exists(%s);" % (self.ex(kw["exist"]), self.ex(kw["file"])))
            elif "opened" in kw:
                self.w("%s = f_opened(%s.unit);" % (self.ex(kw["opened"]), io))
            elif "action" in kw:
                self.w("%s = f_unit_action(%s.unit);" % (self.ex(kw["action"]), io))
            else:
                raise FError("inquire form outside the subset")
        elif kind in ("read", "write"):
            direct = "rec" in kw
            if direct:
                x = self.u.fresh("x")
                self.w("DirectXfer %s(%s, %s);" % (x, io, "true" if kind == "write" else "false"))
                for it in items:
                    if it[0] == "ref" and all(a == ("sec", None, None) for a in it[2]):
                        it = ("name", it[1])
                    if it[0] == "name" and self.scope.sym(it[1]).rank > 0:
                        self.w("%s.item(%s);" % (x, self.vname(it[1])))
                    elif self.rank(it) == 0:
                        self.w("%s.scalar(%s);" % (x, self.ex(it)))
                    else:
                        raise FError("partial array sections in a direct-access I/O list are outside the subset")
                self.w("%s.end();" % x)
            elif kind == "write":
                o = self.u.fresh("o")
                if internal is not None:
                    tmp = self.u.fresh("s")
                    self.w("std::string %s;" % tmp)
                    self.w("ListOut %s(%s, &%s);" % (o, io, tmp))
                else:
                    self.w("ListOut %s(%s);" % (o, io))
                for it in items:
                    r = self.rank(it)
                    if r == 0:
                        self.w("%s.put(%s);" % (o, self.ex(it)))
                    else:
                        dims = self.leaf_dims(self.first_leaf(it))
                        lv = [self.u.fresh("k") for _ in dims]
                        n = self.loops(dims, lv)
                        self.w("%s.put(%s);" % (o, self.ex(it, lv)))
                        self.close(n)
                self.w("%s.end();" % o)
                if internal is not None:
                    self.w("%s = %s;" % (self.ex(internal), tmp))
            else:
                o = self.u.fresh("i")
                self.w("ListIn %s(%s);" % (o, io))
                for it in items:
                    self.w("%s.get(%s);" % (o, self.ex(it)))
                self.w("%s.end();" % o)
        if "iostat" in kw:
            self.w("%s = %s.iostat;" % (self.ex(kw["iostat"]), io))
        self.ind -= 1
        self.w("}")

    # ---------------------------------------------------------- procedures
    def signature(self, p: Proc):
        parts = []
        for a in p.args:
            s = p.syms.get(a)
            if s is None:
                raise FError("%s: dummy %s is not declared" % (p.where, a))
            s.dummy = True
            n = cname(a)
            if s.rank > 0:
                ty = "Arr<%s, %d>" % (s.ctype, s.rank)
                parts.append("%s* %s__p" % (ty, n) if s.optional else "%s& %s__a" % (ty, n))
            elif s.ftype == "char":
                parts.append("const std::string* %s__p" % n if s.optional else "const std::string& %s" % n)
            elif s.optional:
                parts.append("%s%s* %s__p" % ("const " if s.intent == "in" else "", s.ctype, n))
            elif s.intent == "in":
                parts.append("const %s %s" % (s.ctype, n))
            else:
                parts.append("%s& %s" % (s.ctype, n))
        ret = "void"
        if p.kind == "function":
            ret = p.syms[p.name].ctype
        return "%s %s(%s)" % (ret, cname(p.name), ", ".join(parts))

    def procedure(self, p: Proc):
        self.scope = Scope(self.u, p)
        self.w(self.signature(p) + " {")
        self.ind += 1
        for name, s in p.syms.items():
            if s.dummy:
                if s.rank > 0:
                    n = cname(name)
                    ty = "Arr<%s, %d>" % (s.ctype, s.rank)
                    src = "(*%s__p)" % n if s.optional else "%s__a" % n
                    b = self.bounds(s.dims)
                    if b is None:  # assumed shape: lower bounds 1
                        b = ", ".join("B(1, %s.ext(%d))" % (src, d + 1) for d in range(s.rank))
                    if s.optional:
                        self.w("%s %s = %s__p ? %s.rebound(%s) : %s();" % (ty, n, n, src, b, ty))
                    else:
                        self.w("%s %s = %s.rebound(%s);" % (ty, n, src, b))
                continue
            if p.kind == "function" and name == p.name:
                s.is_result = True
                self.w("%s %s__r = 0;" % (s.ctype, cname(name)))
                continue
            self.declare(s, static="save" in s.attrs or "parameter" in s.attrs or s.init is not None)
        if p.name in self.trace:
            self.w('trace_mark("%s");' % p.name)
        self.block(p.body)
        if p.kind == "function":
            self.w("return %s__r;" % cname(p.name))
        self.ind -= 1
        self.w("}")
        self.w()
        self.scope = Scope(self.u, None)

    def translate(self, dump_path_expr: str):
        u = self.u
        self.w('#include "f95rt.h"')
        self.w("using namespace f95;")
        self.w("namespace ref {")
        self.w("[[noreturn]] void f95_stop();")
        skip = set(KINDS)
        for name in u.module_order:
            s = u.module_syms[name]
            if name in skip:
                continue
            self.declare(s, static=True)
        self.w()
        for name in u.proc_order:
            p = u.procs[name]
            self.scope = Scope(u, p)
            self.w(self.signature(p) + ";")
        self.scope = Scope(u, None)
        if self.hooks_include:
            self.w('#include "%s"' % self.hooks_include)
        self.w()
        for name in u.proc_order:
            self.procedure(u.procs[name])
        # the harness's state dump + STOP
        self.w("void f95_stop() {")
        self.ind += 1
        if self.trace:
            self.w('trace_mark("stop");')
            self.w('trace_write(f_cat(f_trim(odir), std::string("ref_trace.txt")));')
        self.w("if (!std::getenv(\"F95_NO_DUMP\")) {")
        self.w("Dump dump__(%s);" % dump_path_expr)
        for name in u.dump_names:
            self.w('dump__.put("%s", %s);' % (name, cname(name)))
        self.w("}")
        self.w("std::fflush(nullptr);")
        self.w("std::exit(errc != 0 ? 1 : 0);")
        self.ind -= 1
        self.w("}")
        self.w()
        p = u.program
        self.scope = Scope(u, p)
        self.w("void main_program() {")
        self.ind += 1
        for name, s in p.syms.items():
            self.declare(s, static=False)
        self.block(p.body)
        self.w("f95_stop();")
        self.ind -= 1
        self.w("}")
        self.w("}  // namespace ref")
        self.w("int main() { ref::main_program(); return 0; }")
        return "\n".join(self.out) + "\n"


def translate(sources, dump_path_expr='f_cat(f_trim(odir), std::string("ref_dump.bin"))', uninit=None, env_params=(),
              trace=(), externs=(), hooks_include=None):
    """sources: [(text, file name, dump its module variables?)] in dependency order -> C++ text.
    uninit: {local name: C++ literal} -- the value a local that the source reads before assigning it holds on entry
    (undefined in Fortran; 0 here unless named).
    env_params: module parameters / strings whose compiled-in value a run may replace through the environment variable
    F95_<NAME> (the timing harness sets the directories and the run length that way; the pin builds use none).
    trace: procedures whose entries are time-stamped into <odir>/ref_trace.txt (timing harness).
    externs / hooks_include: procedure names the source CALLs that are defined in C++ by the named header, which is included
    after the module variables (the drop-in build: the reference driving libbeom_gpu.so through its C ABI)."""
    u = Unit()
    for text, fname, dump in sources:
        u.parse_file(text, fname, dump)
    return Emitter(u, uninit, env_params, trace, externs, hooks_include).translate(dump_path_expr)


if __name__ == "__main__":
    srcs = [(open(f).read(), f, i > 0) for i, f in enumerate(sys.argv[1:])]
    sys.stdout.write(translate(srcs))
