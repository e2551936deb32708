"""A small interpreter for the GLSL 4.30 subset the reference's seven shaders use (test infrastructure only).

Purpose: pin the oracle to the reference's OWN shader text.  Nothing of the reference can run in this image (no GL
stack, no GPU here), so the programmable stages -- Shader/Voxelization.{vs,gs,fs}, Shader/VoxelConeTracing.{vs,fs},
Shader/Shadow.vs -- are executed from their source files, statement by statement, by this interpreter; the fixed-function
stages between them (rasterisation, attribute interpolation, texture filtering, mip generation) are supplied by the
caller as Python callbacks written from the GL 4.3 specification.  tests/golden/make_reference_shader_vectors.py drives
it over /root/reference and commits the resulting vectors; tests/test_reference_glsl.py compares the oracle (CPU) and the
CUDA path (GPU) with them.

Scope: global in / out / uniform / const declarations incl. layout qualifiers and interface blocks (with or without an
instance name, arrayed for geometry-shader inputs), global arrays with array constructors, functions, if / else / for /
while / return / discard, the usual expression grammar with swizzles (read and write), compound assignment, ++ / --, the
ternary operator, constructors, `.length()`, and the built-ins those shaders call.  Arithmetic is carried out in ONE
floating-point type chosen at construction (numpy float32 to mimic the GPU, float64 for a stability cross-check).
Not a general GLSL implementation: anything outside the subset raises GlslError instead of guessing.
"""
from __future__ import annotations

import re

import numpy as np


class GlslError(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class Discard(Exception):
    pass


TYPES = {"void", "float", "int", "bool", "vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4", "mat3", "mat4",
         "sampler2D", "sampler3D", "image3D"}
QUALIFIERS = {"in", "out", "uniform", "const", "flat"}
_TOKEN = re.compile(r"""
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fF]?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\+\+|--|\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||[-+*/%<>=!?:.,;()\[\]{}])
  | (?P<ws>\s+)
""", re.X)


def tokenize(src: str):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#[^\n]*", " ", src, flags=re.M)
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise GlslError(f"cannot tokenise at {src[pos:pos + 20]!r}")
        pos = m.end()
        if m.lastgroup != "ws":
            out.append((m.lastgroup, m.group()))
    out.append(("eof", ""))
    return out


class Mat:
    """Square matrix in maths form a[row][col]; GLSL constructors fill it column by column."""

    def __init__(self, a):
        self.a = a


class Block:
    """Instance of an interface block (or one element of gl_in[] / an arrayed block)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


_SWZ = {c: i for s in ("xyzw", "rgba", "stpq") for i, c in enumerate(s)}


# ----------------------------------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k][1]

    def kind(self, k=0):
        return self.t[self.i + k][0]

    def next(self):
        v = self.t[self.i][1]
        self.i += 1
        return v

    def expect(self, s):
        if self.peek() != s:
            raise GlslError(f"expected {s!r}, found {self.peek()!r} (token {self.i})")
        self.i += 1

    def accept(self, s):
        if self.peek() == s:
            self.i += 1
            return True
        return False

    # ---- top level
    def program(self):
        decls = []
        while self.kind() != "eof":
            decls.append(self.external())
        return decls

    def qualifiers(self):
        q = []
        while True:
            if self.peek() == "layout":
                self.next(); self.expect("(")
                depth = 1
                while depth:
                    v = self.next()
                    depth += (v == "(") - (v == ")")
            elif self.peek() in QUALIFIERS:
                q.append(self.next())
            else:
                return q

    def external(self):
        q = self.qualifiers()
        if self.accept(";"):                                   # `layout (triangles) in;`
            return ("nop",)
        if self.kind() == "id" and self.peek() not in TYPES and self.peek(1) == "{":     # interface block
            name = self.next(); self.expect("{")
            members = []
            while not self.accept("}"):
                self.qualifiers()
                ty = self.next(); mname = self.next(); self.expect(";")
                members.append((ty, mname))
            inst, arrayed = None, False
            if self.kind() == "id":
                inst = self.next()
                if self.accept("["):
                    self.expect("]"); arrayed = True
            self.expect(";")
            return ("block", q, name, members, inst, arrayed)
        ty = self.next()
        if ty not in TYPES:
            raise GlslError(f"unknown type {ty!r}")
        name = self.next()
        if self.accept("("):                                   # function definition
            params = []
            while not self.accept(")"):
                self.qualifiers()
                pty = self.next(); pname = self.next()
                params.append((pty, pname))
                self.accept(",")
            return ("func", ty, name, params, self.block())
        self.i -= 1
        return ("global", q, self.declarators(ty))

    def declarators(self, ty):
        """name [ '[' n ']' ] [ '=' expr ] { ',' ... } ';'"""
        out = []
        while True:
            name = self.next()
            size = None
            if self.accept("["):
                size = None if self.peek() == "]" else self.expr()
                self.expect("]")
                size = size or ("int", 0)
            init = self.assignment() if self.accept("=") else None
            out.append((ty, name, size, init))
            if not self.accept(","):
                break
        self.expect(";")
        return out

    # ---- statements
    def block(self):
        self.expect("{")
        body = []
        while not self.accept("}"):
            body.append(self.statement())
        return ("block", body)

    def statement(self):
        p = self.peek()
        if p == "{":
            return self.block()
        if p == "if":
            self.next(); self.expect("("); c = self.expr(); self.expect(")")
            a = self.statement()
            b = self.statement() if self.accept("else") else None
            return ("if", c, a, b)
        if p == "while":
            self.next(); self.expect("("); c = self.expr(); self.expect(")")
            return ("while", c, self.statement())
        if p == "for":
            self.next(); self.expect("(")
            init = self.statement()                            # declaration or expression statement (eats ';')
            cond = self.expr(); self.expect(";")
            step = self.expr(); self.expect(")")
            return ("for", init, cond, step, self.statement())
        if p == "return":
            self.next()
            e = None if self.peek() == ";" else self.expr()
            self.expect(";")
            return ("return", e)
        if p == "discard":
            self.next(); self.expect(";")
            return ("discard",)
        if p == ";":
            self.next()
            return ("nop",)
        if p in QUALIFIERS or (p in TYPES and self.kind(1) == "id"):
            self.qualifiers()
            return ("decl", self.declarators(self.next()))
        e = self.expr(); self.expect(";")
        return ("expr", e)

    # ---- expressions
    def expr(self):
        return self.assignment()

    def assignment(self):
        lhs = self.ternary()
        if self.peek() in ("=", "+=", "-=", "*=", "/="):
            op = self.next()
            return ("assign", op, lhs, self.assignment())
        return lhs

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.assignment(); self.expect(":")
            return ("ternary", c, a, self.assignment())
        return c

    LEVELS = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]

    def binary(self, level):
        if level == len(self.LEVELS):
            return self.unary()
        lhs = self.binary(level + 1)
        while self.peek() in self.LEVELS[level] and self.kind() == "op":
            op = self.next()
            lhs = ("bin", op, lhs, self.binary(level + 1))
        return lhs

    def unary(self):
        p = self.peek()
        if p in ("-", "+", "!"):
            self.next()
            return ("un", p, self.unary())
        if p in ("++", "--"):
            self.next()
            return ("incdec", p, self.unary(), True)
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            if self.accept("."):
                e = ("member", e, self.next())
            elif self.accept("["):
                i = self.expr(); self.expect("]")
                e = ("index", e, i)
            elif self.peek() == "(" and e[0] in ("name", "member"):
                self.next()
                e = ("call", e, self.args())
            elif self.peek() in ("++", "--"):
                e = ("incdec", self.next(), e, False)
            else:
                return e

    def args(self):
        a = []
        while not self.accept(")"):
            a.append(self.assignment())
            self.accept(",")
        return a

    def primary(self):
        k, v = self.t[self.i]
        if k == "num":
            self.i += 1
            if re.fullmatch(r"\d+", v):
                return ("int", int(v))
            return ("float", float(v.rstrip("fF")))
        if v == "(":
            self.next(); e = self.expr(); self.expect(")")
            return e
        if v in ("true", "false"):
            self.i += 1
            return ("bool", v == "true")
        if k == "id":
            self.i += 1
            if v in TYPES:
                if self.accept("["):                           # array constructor  float[](...)
                    if self.peek() != "]":
                        self.expr()
                    self.expect("]"); self.expect("(")
                    return ("array", v, self.args())
                self.expect("(")
                return ("ctor", v, self.args())
            return ("name", v)
        raise GlslError(f"unexpected token {v!r}")


# ------------------------------------------------------------------------------------------------ interpreter
class Program:
    """One shader stage.  `globals` holds uniforms / inputs / outputs by name; `hooks` the fixed-function built-ins
    (texture, textureLod, imageStore, EmitVertex, EndPrimitive) the caller supplies."""

    def __init__(self, source: str, dtype=np.float32):
        self.F = dtype
        self.ast = Parser(tokenize(source)).program()
        self.funcs, self.globals, self.hooks = {}, {}, {}
        self.decl = {}                                         # name -> (qualifiers, type)
        self.blocks = {}                                       # instance name -> (members, arrayed)
        self.scopes = []
        self._pending = []
        for d in self.ast:
            if d[0] == "func":
                self.funcs[d[2]] = d
            elif d[0] == "global":
                for ty, name, size, init in d[2]:
                    self.decl[name] = (d[1], ty)
                    if init is not None:
                        self._pending.append((None if size else ty, name, init))
                    else:
                        self.globals[name] = [] if size else self.default(ty)
            elif d[0] == "block":
                _, q, bname, members, inst, arrayed = d
                if inst is None:
                    for ty, name in members:
                        self.decl[name] = (q, ty)
                        self.globals[name] = self.default(ty)
                else:
                    self.blocks[inst] = (members, arrayed)
                    self.decl[inst] = (q, bname)
                    self.globals[inst] = [] if arrayed else self.new_block(inst)
        for ty, name, init in self._pending:                   # initialisers may use earlier globals
            self.globals[name] = self.ev(init) if ty is None else self.coerce(ty, self.ev(init))

    # ---- values
    def default(self, ty):
        F = self.F
        if ty == "float":
            return F(0)
        if ty == "int":
            return 0
        if ty == "bool":
            return False
        if ty in ("vec2", "vec3", "vec4"):
            return np.zeros(int(ty[3]), dtype=F)
        if ty in ("ivec2", "ivec3", "ivec4"):
            return np.zeros(int(ty[4]), dtype=np.int64)
        if ty in ("mat3", "mat4"):
            return Mat(np.zeros((int(ty[3]),) * 2, dtype=F))
        return None                                            # samplers / images: opaque, set by the caller

    def new_block(self, inst):
        return Block(**{name: self.default(ty) for ty, name in self.blocks[inst][0]})

    def coerce(self, ty, v):
        """Value semantics + the implicit int -> float conversion of an initialiser / argument."""
        if ty == "float":
            return self.F(v)
        if ty == "int":
            if isinstance(v, (float, np.floating)):
                raise GlslError("float assigned to int")
            return int(v)
        if isinstance(v, np.ndarray):
            return v.astype(self.F if ty.startswith("vec") else np.int64, copy=True)
        if isinstance(v, Mat):
            return Mat(v.a.copy())
        if isinstance(v, list):
            return list(v)
        return v

    def set_uniform(self, name, v):
        q, ty = self.decl[name]
        if ty in ("mat3", "mat4"):
            n = int(ty[3])
            v = Mat(np.asarray(v, dtype=np.float64).reshape(n, n).T.astype(self.F))      # column-major in
        elif ty.startswith(("vec", "ivec")):
            v = np.asarray(v).astype(self.F if ty.startswith("vec") else np.int64)
        elif ty == "float":
            v = self.F(v)
        elif ty == "int":
            v = int(v)
        self.globals[name] = v

    # ---- scopes
    def lookup(self, name):
        for s in reversed(self.scopes):
            if name in s:
                return s
        if name in self.globals:
            return self.globals
        raise GlslError(f"undeclared identifier {name!r}")

    def run(self, entry="main", *args):
        saved, self.scopes = self.scopes, []
        try:
            return self.call_user(entry, list(args))
        finally:
            self.scopes = saved

    def call_user(self, name, args):
        _, rty, _, params, body = self.funcs[name]
        if len(args) != len(params):
            raise GlslError(f"{name}: {len(args)} arguments for {len(params)} parameters")
        frame = {pn: self.coerce(pt, a) for (pt, pn), a in zip(params, args)}
        saved, self.scopes = self.scopes, [frame]              # GLSL has no closures: a fresh scope chain per call
        try:
            self.exec(body)
            return None
        except _Return as r:
            return None if rty == "void" else self.coerce(rty, r.value)
        finally:
            self.scopes = saved

    # ---- statements
    def exec(self, s):
        k = s[0]
        if k == "block":
            self.scopes.append({})
            try:
                for x in s[1]:
                    self.exec(x)
            finally:
                self.scopes.pop()
        elif k == "expr":
            self.ev(s[1])
        elif k == "decl":
            for ty, name, size, init in s[1]:
                if size is not None:
                    raise GlslError("local arrays are outside the subset")
                self.scopes[-1][name] = self.default(ty) if init is None else self.coerce(ty, self.ev(init))
        elif k == "if":
            if self.truth(self.ev(s[1])):
                self.exec(s[2])
            elif s[3] is not None:
                self.exec(s[3])
        elif k == "while":
            while self.truth(self.ev(s[1])):
                self.exec(s[2])
        elif k == "for":
            self.scopes.append({})
            try:
                self.exec(s[1])
                while self.truth(self.ev(s[2])):
                    self.exec(s[4])
                    self.ev(s[3])
            finally:
                self.scopes.pop()
        elif k == "return":
            raise _Return(None if s[1] is None else self.ev(s[1]))
        elif k == "discard":
            raise Discard()
        elif k != "nop":
            raise GlslError(f"statement {k}")

    @staticmethod
    def truth(v):
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        raise GlslError("condition is not a bool")

    # ---- expressions
    def ev(self, e):
        k = e[0]
        if k == "float":
            return self.F(e[1])
        if k in ("int", "bool"):
            return e[1]
        if k == "name":
            return self.lookup(e[1])[e[1]]
        if k == "bin":
            return self.binop(e[1], e[2], e[3])
        if k == "un":
            v = self.ev(e[2])
            if e[1] == "!":
                return not self.truth(v)
            if isinstance(v, Mat):
                return Mat(-v.a)
            return -v if e[1] == "-" else v
        if k == "ternary":
            return self.ev(e[2]) if self.truth(self.ev(e[1])) else self.ev(e[3])
        if k == "member":
            return self.member(self.ev(e[1]), e[2])
        if k == "index":
            base, i = self.ev(e[1]), self.ev(e[2])
            if not isinstance(i, (int, np.integer)):
                raise GlslError("non-integer index")
            if isinstance(base, Mat):
                return base.a[:, i].copy()                     # m[i] is column i
            return base[i]
        if k == "call":
            return self.call(e[1], e[2])
        if k == "ctor":
            return self.construct(e[1], [self.ev(a) for a in e[2]])
        if k == "array":
            return [self.coerce(e[1], self.ev(a)) for a in e[2]]
        if k == "assign":
            return self.assign(e[1], e[2], e[3])
        if k == "incdec":
            old = self.ev(e[2])
            new = old + 1 if e[1] == "++" else old - 1
            self.store(e[2], new)
            return new if e[3] else old
        raise GlslError(f"expression {k}")

    def member(self, obj, field):
        if isinstance(obj, Block):
            return getattr(obj, field)
        if isinstance(obj, np.ndarray):
            idx = [_SWZ[c] for c in field]
            if max(idx) >= len(obj):
                raise GlslError(f"swizzle .{field} on a {len(obj)}-vector")
            return obj[idx[0]] if len(idx) == 1 else obj[idx].copy()
        raise GlslError(f".{field} on {type(obj).__name__}")

    def arith(self, v):
        """operand of an arithmetic operator -> (value, is_float)"""
        if isinstance(v, (bool, np.bool_)):
            raise GlslError("arithmetic on bool")
        if isinstance(v, Mat):
            return v, True
        if isinstance(v, np.ndarray):
            return v, v.dtype != np.int64
        if isinstance(v, (int, np.integer)):
            return int(v), False
        return v, True

    def binop(self, op, ea, eb):
        if op == "&&":
            return self.truth(self.ev(ea)) and self.truth(self.ev(eb))
        if op == "||":
            return self.truth(self.ev(ea)) or self.truth(self.ev(eb))
        a, b = self.ev(ea), self.ev(eb)
        if op in ("==", "!=", "<", ">", "<=", ">="):
            if isinstance(a, (np.ndarray, Mat)) or isinstance(b, (np.ndarray, Mat)):
                raise GlslError("vector comparison is outside the subset")
            a, fa = self.arith(a); b, fb = self.arith(b)
            if fa != fb:
                a, b = self.F(a), self.F(b)
            return bool({"==": a == b, "!=": a != b, "<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op])
        return self.arith_op(op, a, b)

    def arith_op(self, op, a, b):
        F = self.F
        a, fa = self.arith(a); b, fb = self.arith(b)
        if fa != fb:                                           # implicit int -> float (GLSL 4.30, 4.1.10)
            if not fa:
                a = a.astype(F) if isinstance(a, np.ndarray) else F(a)
            else:
                b = b.astype(F) if isinstance(b, np.ndarray) else F(b)
        am, bm = isinstance(a, Mat), isinstance(b, Mat)
        if am or bm:
            if op == "*":
                if am and bm:
                    return Mat((a.a @ b.a).astype(F))
                if am and isinstance(b, np.ndarray):
                    return self.mat_vec(a.a, b)
                if bm and isinstance(a, np.ndarray):
                    return self.mat_vec(b.a.T, a)
                return Mat(a.a * b) if am else Mat(a * b.a)
            raise GlslError(f"matrix {op} is outside the subset")
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if not (fa or fb):
                if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                    raise GlslError("integer vector division is outside the subset")
                q = abs(a) // abs(b)
                return q if (a >= 0) == (b >= 0) else -q
            return a / b
        if op == "%":
            if fa or fb:
                raise GlslError("% on floats")
            return a % b
        raise GlslError(f"operator {op}")

    def mat_vec(self, m, v):
        """row-by-row sum of products in the working type, left to right (no BLAS, no wider accumulator)"""
        out = np.empty(len(v), dtype=self.F)
        for r in range(len(v)):
            acc = m[r, 0] * v[0]
            for c in range(1, len(v)):
                acc = acc + m[r, c] * v[c]
            out[r] = acc
        return out

    # ---- assignment
    def assign(self, op, lhs, rhs):
        v = self.ev(rhs)
        if op != "=":
            v = self.arith_op(op[0], self.ev(lhs), v)
        self.store(lhs, v)
        return v

    def type_like(self, old, v):
        if isinstance(old, np.ndarray):
            if not isinstance(v, np.ndarray) or v.shape != old.shape:
                raise GlslError("vector size mismatch in assignment")
            return v.astype(old.dtype, copy=True)
        if isinstance(old, (float, np.floating)):
            return self.F(v)
        if isinstance(old, (int, np.integer)) and not isinstance(old, (bool, np.bool_)):
            if isinstance(v, (float, np.floating)):
                raise GlslError("float assigned to int")
            return int(v)
        if isinstance(old, Mat):
            return Mat(v.a.copy())
        return v

    def store(self, lhs, v):
        k = lhs[0]
        if k == "name":
            scope = self.lookup(lhs[1])
            scope[lhs[1]] = self.type_like(scope[lhs[1]], v)
        elif k == "member":
            obj = self.ev(lhs[1])
            if isinstance(obj, Block):
                setattr(obj, lhs[2], self.type_like(getattr(obj, lhs[2]), v))
            elif isinstance(obj, np.ndarray):
                idx = [_SWZ[c] for c in lhs[2]]
                if len(set(idx)) != len(idx):
                    raise GlslError("repeated component in a swizzle store")
                if lhs[1][0] not in ("name", "member", "index"):
                    raise GlslError("swizzle store into a temporary")
                obj[idx] = v                                   # arrays are held by reference in their scope
            else:
                raise GlslError("member store")
        elif k == "index":
            base = self.ev(lhs[1])
            base[self.ev(lhs[2])] = v
        else:
            raise GlslError("not an l-value")

    # ---- calls
    def construct(self, ty, args):
        F = self.F
        if ty == "float":
            return F(args[0])
        if ty == "int":
            return int(args[0])                                # truncation toward zero
        if ty.startswith(("vec", "ivec")):
            n = int(ty[-1])
            integer = ty.startswith("i")
            flat = []
            for a in args:
                flat.extend(a.tolist() if isinstance(a, np.ndarray) else [a])
            if len(args) == 1 and not isinstance(args[0], np.ndarray):
                flat = flat * n
            if len(flat) < n or (len(flat) > n and len(args) != 1):
                raise GlslError(f"{ty}: wrong number of components")
            if integer:
                return np.array([int(x) for x in flat[:n]], dtype=np.int64)        # float -> int truncates
            return np.array(flat[:n], dtype=F)                 # .tolist() widened exactly; the cast back is exact too
        if ty in ("mat3", "mat4"):
            n = int(ty[3])
            if len(args) == n and all(isinstance(a, np.ndarray) and len(a) == n for a in args):
                return Mat(np.stack(args, axis=1).astype(F))   # arguments are the columns
            if len(args) == 1 and isinstance(args[0], Mat) and args[0].a.shape[0] >= n:
                return Mat(args[0].a[:n, :n].copy())
            raise GlslError(f"{ty} constructor form is outside the subset")
        raise GlslError(f"constructor {ty}")

    def call(self, callee, arg_nodes):
        if callee[0] == "member":
            if callee[2] == "length" and not arg_nodes:
                return len(self.ev(callee[1]))
            raise GlslError(f"method .{callee[2]}()")
        name = callee[1]
        args = [self.ev(a) for a in arg_nodes]
        if name in self.funcs:
            return self.call_user(name, args)
        if name in self.hooks:
            return self.hooks[name](*args)
        fn = getattr(self, "bi_" + name, None)
        if fn is None:
            raise GlslError(f"unknown function {name!r}")
        return fn(*args)

    # ---- built-ins (GLSL 4.30 section 8), in the working type
    def _f(self, v):
        if isinstance(v, np.ndarray):
            return v.astype(self.F) if v.dtype != self.F else v
        return self.F(v)

    def bi_dot(self, a, b):
        a, b = self._f(a), self._f(b)
        acc = a[0] * b[0]
        for i in range(1, len(a)):
            acc = acc + a[i] * b[i]
        return acc

    def bi_length(self, v):
        return np.sqrt(self.bi_dot(v, v))

    def bi_normalize(self, v):
        v = self._f(v)
        return v / self.bi_length(v)

    def bi_cross(self, a, b):
        a, b = self._f(a), self._f(b)
        return np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], dtype=self.F)

    def bi_abs(self, v):
        return abs(v) if not isinstance(v, np.ndarray) else np.abs(v)

    def _minmax(self, a, b, fn):
        a, fa = self.arith(a); b, fb = self.arith(b)
        if fa != fb:
            raise GlslError("min/max with mixed int and float arguments")
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return fn(a, b)
        return a if fn(a, b) == a else b

    def bi_max(self, a, b):
        return self._minmax(a, b, np.maximum)

    def bi_min(self, a, b):
        return self._minmax(a, b, np.minimum)

    def bi_log2(self, v):
        return np.log2(self._f(v))

    def bi_pow(self, a, b):
        return np.power(self._f(a), self._f(b))

    def bi_reflect(self, i, n):
        i, n = self._f(i), self._f(n)
        return i - self.F(2) * self.bi_dot(n, i) * n

    def bi_transpose(self, m):
        return Mat(m.a.T.copy())

    def bi_inverse(self, m):
        return Mat(np.linalg.inv(m.a.astype(np.float64)).astype(self.F))
