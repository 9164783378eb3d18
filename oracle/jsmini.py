"""jsmini — a small ES5-subset interpreter, used ONLY to execute the reference's own JavaScript
(/root/reference/src/ccv.js, camshift.js, whitebalance.js, cascade.js) in this container, where no
JS engine exists, so that the C oracle can be pinned against outputs of the REAL reference source.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  It is used by tools/make_goldens.py to
produce tests/golden/*.json; it cannot travel to the GPU box (it needs /root/reference).

Supported: var/function/closures/this/new, if/for/while/break/continue/return, the ES5 expression
grammar used by those files (assignment incl. compound, ?:, || &&, comparisons, arithmetic, bitwise,
shifts, ++/--, member/index/call, array & object literals), Numbers as IEEE doubles, ToInt32 shifts,
Arrays, Uint8ClampedArray, Math, and a canvas shim (`CanvasShim`) whose drawImage is the resampler
DEFINED in oracle/ht_oracle.h (the one thing the reference leaves to the browser).
"""
import math
import re
import numpy as np

# ------------------------------------------------------------------------------------------------
# values


class JSUndefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


undefined = JSUndefined()
null = None


class JSObject:
    __slots__ = ("props", "proto")

    def __init__(self, proto=None):
        self.props = {}
        self.proto = proto

    def get(self, key):
        o = self
        while o is not None:
            p = o.props
            if key in p:
                return p[key]
            o = o.proto
        return undefined

    def set(self, key, val):
        self.props[key] = val


class JSArray(JSObject):
    __slots__ = ("items",)

    def __init__(self, items=None):
        JSObject.__init__(self)
        self.items = items if items is not None else []

    def get(self, key):
        if key.__class__ is float or key.__class__ is int:
            i = int(key)
            if i == key and 0 <= i < len(self.items):
                return self.items[i]
            return undefined
        if key == "length":
            return float(len(self.items))
        if key == "push":
            return NativeFunction(lambda this, args: self._push(args))
        if key == "pop":
            return NativeFunction(lambda this, args: self.items.pop() if self.items else undefined)
        if key == "unshift":
            return NativeFunction(lambda this, args: self._unshift(args))
        if key == "splice":        # headDiagonal.splice(0,1) - src/main.js:275
            return NativeFunction(lambda this, args: self._splice(args))
        return JSObject.get(self, key)

    def _push(self, args):
        self.items.extend(args)
        return float(len(self.items))

    def _unshift(self, args):
        self.items[0:0] = list(args)
        return float(len(self.items))

    def _splice(self, args):
        n = len(self.items)
        start = int(to_number(args[0])) if args else 0
        start = max(n + start, 0) if start < 0 else min(start, n)
        cnt = n - start if len(args) < 2 else max(0, min(int(to_number(args[1])), n - start))
        removed = self.items[start:start + cnt]
        self.items[start:start + cnt] = list(args[2:])
        return JSArray(removed)

    def set(self, key, val):
        if key.__class__ is float or key.__class__ is int:
            i = int(key)
            if i == key and i >= 0:
                n = len(self.items)
                if i < n:
                    self.items[i] = val
                else:
                    self.items.extend([undefined] * (i - n))
                    self.items.append(val)
                return
        if key == "length":
            del self.items[int(val):]
            return
        JSObject.set(self, key, val)


class JSUint8ClampedArray(JSObject):
    """Typed array over a numpy uint8 buffer; stores apply ES ToUint8Clamp (round half to even)."""
    __slots__ = ("buf",)

    def __init__(self, buf):
        JSObject.__init__(self)
        self.buf = buf

    def get(self, key):
        if key.__class__ is float or key.__class__ is int:
            i = int(key)
            if i == key and 0 <= i < self.buf.shape[0]:
                return float(self.buf[i])
            return undefined
        if key == "length":
            return float(self.buf.shape[0])
        return JSObject.get(self, key)

    def set(self, key, val):
        if key.__class__ is float or key.__class__ is int:
            i = int(key)
            if i == key and 0 <= i < self.buf.shape[0]:
                v = to_number(val)
                if v != v or v <= 0:
                    b = 0
                elif v >= 255:
                    b = 255
                else:
                    f = math.floor(v)
                    d = v - f
                    if d < 0.5:
                        b = int(f)
                    elif d > 0.5:
                        b = int(f) + 1
                    else:
                        b = int(f) if int(f) % 2 == 0 else int(f) + 1
                self.buf[i] = b
            return
        JSObject.set(self, key, val)


class NativeFunction(JSObject):
    __slots__ = ("fn",)

    def __init__(self, fn):
        JSObject.__init__(self)
        self.fn = fn

    def get(self, key):
        if key == "apply":        # Math.max.apply(null, array) — src/facetrackr.js:84-86
            return NativeFunction(lambda this, a: self.call(a[0] if a else undefined,
                                                            list(a[1].items) if len(a) > 1 else []))
        return JSObject.get(self, key)

    def call(self, this, args):
        return self.fn(this, args)

    def construct(self, args):
        return self.fn(None, args)


class JSDate(JSObject):
    """`new Date()` — only what the reference uses: subtraction and getTime()."""
    __slots__ = ("ms",)

    def __init__(self, ms):
        JSObject.__init__(self)
        self.ms = float(ms)
        self.props["getTime"] = NativeFunction(lambda this, a: self.ms)


GLOBAL_THIS = JSObject()


class ReturnEx(Exception):
    def __init__(self, v):
        self.v = v


class BreakEx(Exception):
    pass


class ContinueEx(Exception):
    pass


class Env:
    __slots__ = ("vars", "parent", "this")

    def __init__(self, parent, this):
        self.vars = {}
        self.parent = parent
        self.this = this

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


class JSFunction(JSObject):
    __slots__ = ("params", "body", "env", "hoisted", "name")

    def __init__(self, params, body, env, hoisted, name=None):
        JSObject.__init__(self)
        self.params, self.body, self.env, self.hoisted, self.name = params, body, env, hoisted, name
        self.props["prototype"] = JSObject()

    def call(self, this, args):
        if this is undefined or this is None:
            this = GLOBAL_THIS          # sloppy-mode functions see the global object as `this`
        env = Env(self.env, this)
        v = env.vars
        for n in self.hoisted:
            v[n] = undefined
        na = len(args)
        for i, p in enumerate(self.params):
            v[p] = args[i] if i < na else undefined
        if self.name:
            v.setdefault(self.name, self)
        try:
            self.body(env)
        except ReturnEx as r:
            return r.v
        return undefined

    def construct(self, args):
        obj = JSObject(self.props.get("prototype"))
        r = self.call(obj, args)
        return r if isinstance(r, JSObject) else obj

    def get(self, key):
        if key == "bind":          # function(){...}.bind(this) - src/main.js:77,305
            return NativeFunction(lambda this, a: NativeFunction(
                lambda this2, a2, f=self, bt=(a[0] if a else undefined), ba=list(a[1:]): f.call(bt, ba + list(a2))))
        if key == "call":
            return NativeFunction(lambda this, a: self.call(a[0] if a else undefined, list(a[1:])))
        if key == "apply":
            return NativeFunction(lambda this, a: self.call(a[0] if a else undefined,
                                                            list(a[1].items) if len(a) > 1 and isinstance(a[1], JSArray) else []))
        return JSObject.get(self, key)


# ------------------------------------------------------------------------------------------------
# conversions / operators

def to_number(v):
    c = v.__class__
    if c is float:
        return v
    if c is int:
        return float(v)
    if c is bool:
        return 1.0 if v else 0.0
    if v is undefined:
        return math.nan
    if v is None:
        return 0.0
    if c is JSDate:
        return v.ms
    if c is str:
        try:
            return float(v) if v.strip() else 0.0
        except ValueError:
            return math.nan
    return math.nan


def to_int32(v):
    v = to_number(v)
    if v != v or v in (math.inf, -math.inf):
        return 0
    i = int(v) & 0xFFFFFFFF
    return i - 0x100000000 if i >= 0x80000000 else i


def to_bool(v):
    c = v.__class__
    if c is bool:
        return v
    if c is float:
        return not (v == 0.0 or v != v)
    if v is undefined or v is None:
        return False
    if c is str:
        return len(v) > 0
    return True


def js_add(a, b):
    if a.__class__ is float and b.__class__ is float:
        return a + b
    if isinstance(a, str) or isinstance(b, str):
        return js_str(a) + js_str(b)
    return to_number(a) + to_number(b)


def js_str(v):
    if isinstance(v, str):
        return v
    if v.__class__ is float:
        if v == int(v) and abs(v) < 1e21:
            return str(int(v))
        return repr(v)
    if v is undefined:
        return "undefined"
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    return "[object Object]"


def js_div(a, b):
    a, b = to_number(a), to_number(b)
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf
    return a / b


def js_mul(a, b):
    a, b = to_number(a), to_number(b)
    try:
        return a * b
    except OverflowError:
        return math.inf if (a > 0) == (b > 0) else -math.inf


def js_lt(a, b):     # a < b with undefined -> NaN -> false
    a, b = to_number(a), to_number(b)
    return a < b


def js_eq(a, b):     # loose equality for the cases the reference uses (numbers, strings, undefined/null, objects)
    if a is undefined or a is None:
        return b is undefined or b is None
    if b is undefined or b is None:
        return False
    if isinstance(a, JSObject) or isinstance(b, JSObject):
        return a is b
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    return to_number(a) == to_number(b)


def js_seq(a, b):
    if a.__class__ is float or a.__class__ is int:
        return (b.__class__ is float or b.__class__ is int) and float(a) == float(b)
    if isinstance(a, JSObject):
        return a is b
    if a is undefined:
        return b is undefined
    if a is None:
        return b is None
    return a.__class__ is b.__class__ and a == b


def get_member(obj, key):
    if isinstance(obj, JSObject):
        return obj.get(key)
    if isinstance(obj, str):
        if key == "length":
            return float(len(obj))
        return undefined
    if obj is undefined or obj is None:
        raise RuntimeError(f"TypeError: cannot read property {key!r} of {obj!r}")
    return undefined


def call_function(f, this, args):
    if isinstance(f, (JSFunction, NativeFunction)):
        return f.call(this, args)
    raise RuntimeError(f"TypeError: {f!r} is not a function")


# ------------------------------------------------------------------------------------------------
# tokenizer

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')
  | (?P<op>>>>=|===|!==|>>>|<<=|>>=|\+\+|--|&&|\|\||==|!=|<=|>=|\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|[{}()\[\];,<>+\-*/%&|^!~?:=.])
""", re.X | re.S)

KEYWORDS = {"var", "function", "if", "else", "for", "while", "do", "break", "continue", "return", "new", "this",
            "true", "false", "null", "undefined", "typeof", "in", "try", "catch", "finally"}


def tokenize(src):
    out, pos = [], 0
    n = len(src)
    while pos < n:
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"jsmini: cannot tokenize at {pos}: {src[pos:pos + 30]!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        t = m.group(k)
        if k == "id" and t in KEYWORDS:
            k = "kw"
        out.append((k, t))
    out.append(("eof", ""))
    return out


# ------------------------------------------------------------------------------------------------
# parser -> closures.  Every compiled node is a Python callable taking an Env.

class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0
        self.scopes = []   # stack of sets: names declared with var/function in the enclosing function

    # -- token helpers
    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, v):
        k, t = self.t[self.i]
        return t == v and k in ("op", "kw")

    def eat(self, v):
        if self.at(v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.eat(v):
            raise SyntaxError(f"jsmini: expected {v!r}, got {self.peek()} near token {self.i}")

    # -- program / statements
    def parse_program(self):
        self.scopes.append(set())
        stmts = []
        while self.peek()[0] != "eof":
            stmts.append(self.statement())
        hoisted = self.scopes.pop()
        body = self.block_of(stmts)
        return body, hoisted

    @staticmethod
    def block_of(stmts):
        if len(stmts) == 1:
            return stmts[0]
        stmts = tuple(stmts)

        def run(env):
            for s in stmts:
                s(env)
        return run

    def statement(self):
        k, t = self.peek()
        if k == "op" and t == "{":
            self.next()
            stmts = []
            while not self.at("}"):
                stmts.append(self.statement())
            self.expect("}")
            return self.block_of(stmts) if stmts else (lambda env: None)
        if k == "op" and t == ";":
            self.next()
            return lambda env: None
        if k == "kw":
            if t == "var":
                s = self.var_decl()
                self.eat(";")
                return s
            if t == "function":
                self.next()
                name = self.next()[1]
                fn = self.function_rest(name)
                self.scopes[-1].add(name)

                def decl(env, name=name, fn=fn):
                    env.vars[name] = fn(env)
                return decl
            if t == "if":
                self.next()
                self.expect("(")
                c = self.expression()
                self.expect(")")
                a = self.statement()
                b = self.statement() if self.eat("else") else None
                if b is None:
                    def if1(env):
                        if to_bool(c(env)):
                            a(env)
                    return if1

                def if2(env):
                    if to_bool(c(env)):
                        a(env)
                    else:
                        b(env)
                return if2
            if t == "for":
                return self.for_stmt()
            if t == "while":
                self.next()
                self.expect("(")
                c = self.expression()
                self.expect(")")
                body = self.statement()

                def wh(env):
                    while to_bool(c(env)):
                        try:
                            body(env)
                        except ContinueEx:
                            continue
                        except BreakEx:
                            break
                return wh
            if t == "try":                      # src/main.js:310-325 (starter)
                self.next()
                body = self.statement()
                handler = param = finalizer = None
                if self.at("catch"):
                    self.next()
                    self.expect("(")
                    param = self.next()[1]
                    self.expect(")")
                    self.scopes[-1].add(param)
                    handler = self.statement()
                if self.at("finally"):
                    self.next()
                    finalizer = self.statement()

                def tr(env, body=body, handler=handler, param=param, finalizer=finalizer):
                    try:
                        body(env)
                    except (ReturnEx, BreakEx, ContinueEx):
                        raise
                    except Exception as ex:     # JS-level errors surface as Python exceptions in this interpreter
                        if handler is None:
                            raise
                        env.vars[param] = str(ex)
                        handler(env)
                    finally:
                        if finalizer is not None:
                            finalizer(env)
                return tr
            if t == "break":
                self.next()
                self.eat(";")

                def br(env):
                    raise BreakEx()
                return br
            if t == "continue":
                self.next()
                self.eat(";")

                def co(env):
                    raise ContinueEx()
                return co
            if t == "return":
                self.next()
                if self.at(";") or self.at("}"):
                    self.eat(";")

                    def r0(env):
                        raise ReturnEx(undefined)
                    return r0
                e = self.expression()
                self.eat(";")

                def r1(env):
                    raise ReturnEx(e(env))
                return r1
        e = self.expression()
        self.eat(";")
        return e

    def var_decl(self):
        self.expect("var")
        decls = []
        while True:
            name = self.next()[1]
            self.scopes[-1].add(name)
            init = self.assignment() if self.eat("=") else None
            decls.append((name, init))
            if not self.eat(","):
                break
        decls = tuple(decls)

        def run(env):
            for name, init in decls:
                if init is not None:
                    v = init(env)
                    e = env.lookup(name) or env
                    e.vars[name] = v
        return run

    def for_stmt(self):
        self.expect("for")
        self.expect("(")
        init = None
        if not self.at(";"):
            init = self.var_decl() if self.at("var") else self.expression()
        self.expect(";")
        cond = None if self.at(";") else self.expression()
        self.expect(";")
        step = None if self.at(")") else self.expression()
        self.expect(")")
        body = self.statement()

        def run(env):
            if init is not None:
                init(env)
            while cond is None or to_bool(cond(env)):
                try:
                    body(env)
                except ContinueEx:
                    pass
                except BreakEx:
                    break
                if step is not None:
                    step(env)
        return run

    def function_rest(self, name=None):
        self.expect("(")
        params = []
        while not self.at(")"):
            params.append(self.next()[1])
            self.eat(",")
        self.expect(")")
        self.expect("{")
        self.scopes.append(set())
        stmts = []
        while not self.at("}"):
            stmts.append(self.statement())
        self.expect("}")
        hoisted = tuple(self.scopes.pop() - set(params))
        body = self.block_of(stmts) if stmts else (lambda env: None)
        params = tuple(params)
        return lambda env: JSFunction(params, body, env, hoisted, name)

    # -- expressions
    def expression(self):
        e = self.assignment()
        if self.at(","):
            parts = [e]
            while self.eat(","):
                parts.append(self.assignment())
            parts = tuple(parts)

            def seq(env):
                v = undefined
                for p in parts:
                    v = p(env)
                return v
            return seq
        return e

    ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>="}

    def assignment(self):
        start = self.i
        left = self.conditional()
        k, t = self.peek()
        if k == "op" and t in self.ASSIGN_OPS:
            self.next()
            ref = getattr(left, "ref", None)
            if ref is None:
                raise SyntaxError(f"jsmini: invalid assignment target near token {start}")
            right = self.assignment()
            binop = None if t == "=" else BINOPS[t[:-1]]
            kind = ref[0]
            if kind == "var":
                name = ref[1]

                def assign_var(env):
                    if binop is None:
                        v = right(env)
                    else:
                        e0 = env.lookup(name)
                        cur = e0.vars[name] if e0 else undefined
                        v = binop(cur, right(env))
                    e = env.lookup(name)
                    if e is None:          # implicit global
                        e = env
                        while e.parent is not None:
                            e = e.parent
                    e.vars[name] = v
                    return v
                return assign_var
            objf, keyf = ref[1], ref[2]

            def assign_member(env):
                obj = objf(env)            # reference (object, key) is evaluated before the right-hand side
                key = keyf(env)
                if binop is None:
                    v = right(env)
                else:
                    v = binop(get_member(obj, key), right(env))
                obj.set(key, v)
                return v
            return assign_member
        return left

    def conditional(self):
        c = self.binary(0)
        if self.eat("?"):
            a = self.assignment()
            self.expect(":")
            b = self.assignment()
            return lambda env: a(env) if to_bool(c(env)) else b(env)
        return c

    PREC = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!=", "===", "!=="), ("<", ">", "<=", ">="),
            ("<<", ">>", ">>>"), ("+", "-"), ("*", "/", "%")]

    def binary(self, level):
        if level == len(self.PREC):
            return self.unary()
        left = self.binary(level + 1)
        ops = self.PREC[level]
        while True:
            k, t = self.peek()
            if k == "op" and t in ops:
                self.next()
                right = self.binary(level + 1)
                left = self.make_binary(t, left, right)
            else:
                return left

    @staticmethod
    def make_binary(op, a, b):
        if op == "||":
            def f(env):
                v = a(env)
                return v if to_bool(v) else b(env)
            return f
        if op == "&&":
            def f(env):
                v = a(env)
                return b(env) if to_bool(v) else v
            return f
        fn = BINOPS[op]
        return lambda env: fn(a(env), b(env))

    def unary(self):
        k, t = self.peek()
        if k == "op":
            if t == "!":
                self.next()
                e = self.unary()
                return lambda env: not to_bool(e(env))
            if t == "-":
                self.next()
                e = self.unary()
                return lambda env: -to_number(e(env))
            if t == "+":
                self.next()
                e = self.unary()
                return lambda env: to_number(e(env))
            if t == "~":
                self.next()
                e = self.unary()
                return lambda env: float(~to_int32(e(env)))
            if t in ("++", "--"):
                self.next()
                e = self.unary()
                return self.incdec(e, 1.0 if t == "++" else -1.0, prefix=True)
        if k == "kw" and t == "typeof":
            self.next()
            e = self.unary()

            def ty(env):
                try:
                    v = e(env)
                except RuntimeError:
                    return "undefined"
                if v is undefined:
                    return "undefined"
                if isinstance(v, (JSFunction, NativeFunction)):
                    return "function"
                if v.__class__ is float:
                    return "number"
                if isinstance(v, str):
                    return "string"
                if isinstance(v, bool):
                    return "boolean"
                return "object"
            return ty
        return self.postfix()

    def incdec(self, e, delta, prefix):
        ref = getattr(e, "ref", None)
        if ref is None:
            raise SyntaxError("jsmini: invalid ++/-- target")
        if ref[0] == "var":
            name = ref[1]

            def f(env):
                sc = env.lookup(name)
                old = to_number(sc.vars[name])
                sc.vars[name] = old + delta
                return old + delta if prefix else old
            return f
        objf, keyf = ref[1], ref[2]

        def g(env):
            obj, key = objf(env), keyf(env)
            old = to_number(get_member(obj, key))
            obj.set(key, old + delta)
            return old + delta if prefix else old
        return g

    def postfix(self):
        e = self.call_member()
        k, t = self.peek()
        if k == "op" and t in ("++", "--"):
            self.next()
            return self.incdec(e, 1.0 if t == "++" else -1.0, prefix=False)
        return e

    def arguments(self):
        args = []
        self.expect("(")
        while not self.at(")"):
            args.append(self.assignment())
            self.eat(",")
        self.expect(")")
        return tuple(args)

    def call_member(self):
        if self.at("new"):
            self.next()
            # `new X.Y(args)` / `(new Date).getTime()`
            target = self.member_only()
            args = self.arguments() if self.at("(") else ()

            def mk(env):
                f = target(env)
                a = [x(env) for x in args]
                if isinstance(f, (JSFunction, NativeFunction)):
                    return f.construct(a)
                raise RuntimeError("TypeError: not a constructor")
            e = mk
        else:
            e = self.primary()
        while True:
            if self.at("."):
                self.next()
                name = self.next()[1]
                e = self.member(e, lambda env, name=name: name)
            elif self.at("["):
                self.next()
                key = self.expression()
                self.expect("]")
                e = self.member(e, key)
            elif self.at("("):
                args = self.arguments()
                ref = getattr(e, "ref", None)
                if ref is not None and ref[0] == "member":
                    objf, keyf = ref[1], ref[2]

                    def mcall(env, objf=objf, keyf=keyf, args=args):
                        obj = objf(env)
                        f = get_member(obj, keyf(env))
                        return call_function(f, obj, [a(env) for a in args])
                    e = mcall
                else:
                    def fcall(env, fe=e, args=args):
                        return call_function(fe(env), undefined, [a(env) for a in args])
                    e = fcall
            else:
                return e

    def member_only(self):
        e = self.primary()
        while True:
            if self.at("."):
                self.next()
                name = self.next()[1]
                e = self.member(e, lambda env, name=name: name)
            elif self.at("["):
                self.next()
                key = self.expression()
                self.expect("]")
                e = self.member(e, key)
            else:
                return e

    @staticmethod
    def member(objf, keyf):
        def get(env):
            return get_member(objf(env), keyf(env))
        get.ref = ("member", objf, keyf)
        return get

    def primary(self):
        k, t = self.next()
        if k == "num":
            v = float(int(t, 16)) if t[:2] in ("0x", "0X") else float(t)
            return lambda env: v
        if k == "str":
            s = bytes(t[1:-1], "utf-8").decode("unicode_escape")
            return lambda env: s
        if k == "id":
            name = t

            def var(env):
                e = env
                while e is not None:
                    vs = e.vars
                    if name in vs:
                        return vs[name]
                    e = e.parent
                raise RuntimeError(f"ReferenceError: {name} is not defined")
            var.ref = ("var", name)
            return var
        if k == "kw":
            if t == "this":
                return lambda env: env.this
            if t == "true":
                return lambda env: True
            if t == "false":
                return lambda env: False
            if t == "null":
                return lambda env: None
            if t == "undefined":
                return lambda env: undefined
            if t == "function":
                name = None
                if self.peek()[0] == "id":
                    name = self.next()[1]
                return self.function_rest(name)
        if k == "op":
            if t == "(":
                e = self.expression()
                self.expect(")")
                return e
            if t == "[":
                items = []
                while not self.at("]"):
                    items.append(self.assignment())
                    self.eat(",")
                self.expect("]")
                items = tuple(items)
                return lambda env: JSArray([x(env) for x in items])
            if t == "{":
                props = []
                while not self.at("}"):
                    kk, kt = self.next()
                    key = bytes(kt[1:-1], "utf-8").decode("unicode_escape") if kk == "str" else (
                        js_str(float(kt)) if kk == "num" else kt)
                    self.expect(":")
                    props.append((key, self.assignment()))
                    self.eat(",")
                self.expect("}")
                props = tuple(props)

                def obj(env):
                    o = JSObject()
                    for key, v in props:
                        o.props[key] = v(env)
                    return o
                return obj
        raise SyntaxError(f"jsmini: unexpected token {(k, t)} at {self.i}")


def _shift(fn):
    return lambda a, b: float(fn(to_int32(a), to_int32(b) & 31))


BINOPS = {
    "+": js_add,
    "-": lambda a, b: to_number(a) - to_number(b),
    "*": js_mul,
    "/": js_div,
    "%": lambda a, b: math.fmod(to_number(a), to_number(b)) if to_number(b) != 0 else math.nan,
    "<": js_lt,
    ">": lambda a, b: js_lt(b, a),
    "<=": lambda a, b: (lambda x, y: x <= y)(to_number(a), to_number(b)),
    ">=": lambda a, b: (lambda x, y: x >= y)(to_number(a), to_number(b)),
    "==": js_eq,
    "!=": lambda a, b: not js_eq(a, b),
    "===": js_seq,
    "!==": lambda a, b: not js_seq(a, b),
    "&": lambda a, b: float(to_int32(a) & to_int32(b)),
    "|": lambda a, b: float(to_int32(a) | to_int32(b)),
    "^": lambda a, b: float(to_int32(a) ^ to_int32(b)),
    "<<": lambda a, b: float(to_int32(to_int32(a) << (to_int32(b) & 31))),
    ">>": _shift(lambda x, s: x >> s),
    ">>>": lambda a, b: float((to_int32(a) & 0xFFFFFFFF) >> (to_int32(b) & 31)),
}


# ------------------------------------------------------------------------------------------------
# runtime: Math, Array, canvas shim

def _math():
    m = JSObject()

    def fn(f):
        return NativeFunction(lambda this, args: f(*[to_number(a) for a in args]))

    def js_floor(x):
        return x if (x != x or x in (math.inf, -math.inf)) else float(math.floor(x))

    def js_sqrt(x):
        return math.nan if (x != x or x < 0) else math.sqrt(x)

    def js_log(x):
        if x != x or x < 0:
            return math.nan
        return -math.inf if x == 0 else math.log(x)

    def js_min(*a):
        r = math.inf
        for x in a:
            if x != x:
                return math.nan
            r = min(r, x)
        return r

    def js_max(*a):
        r = -math.inf
        for x in a:
            if x != x:
                return math.nan
            r = max(r, x)
        return r

    m.props.update({
        "floor": fn(js_floor), "sqrt": fn(js_sqrt), "log": fn(js_log), "pow": fn(lambda a, b: math.pow(a, b)),
        "min": fn(js_min), "max": fn(js_max), "abs": fn(abs), "atan2": fn(math.atan2),
        "round": fn(lambda x: float(math.floor(x + 0.5))), "PI": math.pi,
        "atan": fn(math.atan), "sin": fn(math.sin), "cos": fn(math.cos), "tan": fn(math.tan),
    })
    return m


def _array_ctor():
    def ctor(this, args):
        if len(args) == 1 and args[0].__class__ is float:
            return JSArray([undefined] * int(args[0]))
        return JSArray(list(args))
    return NativeFunction(ctor)


def shim_draw(src, sx, sy, sw, sh, dst, dw, dh):
    """Canvas-shim drawImage on (H,W,4) uint8 arrays — the resampler DEFINED in oracle/ht_oracle.h.
    Written independently of the C oracle (numpy int64), channel-wise; source assumed opaque."""
    if dw <= 0 or dh <= 0 or sw <= 0 or sh <= 0:
        return
    X = np.arange(dw, dtype=np.int64)
    Y = np.arange(dh, dtype=np.int64)
    un = (2 * X + 1) * sw - dw
    vn = (2 * Y + 1) * sh - dh
    x0 = np.floor_divide(un, 2 * dw)
    y0 = np.floor_divide(vn, 2 * dh)
    fx = (un - x0 * 2 * dw)[None, :, None]
    fy = (vn - y0 * 2 * dh)[:, None, None]
    xa = np.clip(x0, 0, sw - 1) + sx
    xb = np.clip(x0 + 1, 0, sw - 1) + sx
    ya = np.clip(y0, 0, sh - 1) + sy
    yb = np.clip(y0 + 1, 0, sh - 1) + sy
    s = src.astype(np.int64)
    Dx, Dy = 2 * dw, 2 * dh
    num = ((Dx - fx) * (Dy - fy) * s[ya][:, xa] + fx * (Dy - fy) * s[ya][:, xb] +
           (Dx - fx) * fy * s[yb][:, xa] + fx * fy * s[yb][:, xb])
    dst[:dh, :dw, :] = ((num + 2 * dw * dh) // (4 * dw * dh)).astype(np.uint8)


class CanvasShim(JSObject):
    """document.createElement('canvas') stand-in: width/height properties, 2D context with getImageData /
    putImageData / drawImage / createImageData.  New canvases are transparent black."""

    def __init__(self, pixels=None):
        JSObject.__init__(self)
        self.pix = pixels if pixels is not None else np.zeros((0, 0, 4), np.uint8)
        self._ctx = None

    def get(self, key):
        if key == "width":
            return float(self.pix.shape[1])
        if key == "height":
            return float(self.pix.shape[0])
        if key == "getContext":
            return NativeFunction(lambda this, args: self.context())
        if key == "tagName":
            return "CANVAS"
        return JSObject.get(self, key)

    def set(self, key, val):
        if key == "width":
            self.pix = np.zeros((self.pix.shape[0], int(to_number(val)), 4), np.uint8)   # resizing clears
            return
        if key == "height":
            self.pix = np.zeros((int(to_number(val)), self.pix.shape[1], 4), np.uint8)
            return
        JSObject.set(self, key, val)

    def context(self):
        if self._ctx is None:
            c = JSObject()
            c.props["getImageData"] = NativeFunction(lambda this, a: self.get_image_data(*[int(to_number(v)) for v in a]))
            c.props["putImageData"] = NativeFunction(lambda this, a: self.put_image_data(a[0], int(to_number(a[1])), int(to_number(a[2]))))
            c.props["drawImage"] = NativeFunction(lambda this, a: self.draw_image(a))
            c.props["createImageData"] = NativeFunction(lambda this, a: self.make_image_data(
                np.zeros((int(to_number(a[1])), int(to_number(a[0])), 4), np.uint8)))
            self._ctx = c
        return self._ctx

    @staticmethod
    def make_image_data(arr):
        o = JSObject()
        o.props["width"] = float(arr.shape[1])
        o.props["height"] = float(arr.shape[0])
        o.props["data"] = JSUint8ClampedArray(arr.reshape(-1))
        return o

    def get_image_data(self, x, y, w, h):
        if w <= 0 or h <= 0:
            raise RuntimeError("IndexSizeError: getImageData with a 0-sized rectangle")
        H, W = self.pix.shape[:2]
        out = np.zeros((h, w, 4), np.uint8)                     # outside the canvas: transparent black
        x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + w, W), min(y + h, H)
        if x1 > x0 and y1 > y0:
            out[y0 - y:y1 - y, x0 - x:x1 - x] = self.pix[y0:y1, x0:x1]
        return self.make_image_data(out)

    def put_image_data(self, img, x, y):
        w, h = int(img.get("width")), int(img.get("height"))
        data = img.get("data").buf.reshape(h, w, 4)
        H, W = self.pix.shape[:2]
        x1, y1 = min(x + w, W), min(y + h, H)
        self.pix[y:y1, x:x1] = data[:y1 - y, :x1 - x]

    def draw_image(self, a):
        src = a[0]
        n = [to_number(v) for v in a[1:]]
        if len(n) == 8:
            sx, sy, sw, sh, dx, dy, dw, dh = [int(v) for v in n]
        elif len(n) == 4:
            dx, dy, dw, dh = [int(v) for v in n]
            sx, sy, sw, sh = 0, 0, src.pix.shape[1], src.pix.shape[0]
        else:
            dx, dy = int(n[0]), int(n[1])
            sx, sy, sw, sh = 0, 0, src.pix.shape[1], src.pix.shape[0]
            dw, dh = sw, sh
        assert dx == 0 and dy == 0, "the reference only draws at the origin"
        shim_draw(src.pix, sx, sy, sw, sh, self.pix, dw, dh)
        return undefined


class Interpreter:
    def __init__(self):
        self.genv = Env(None, undefined)
        g = self.genv.vars
        g["Math"] = _math()
        g["Array"] = _array_ctor()
        g["undefined"] = undefined
        g["NaN"] = math.nan
        g["Infinity"] = math.inf
        doc = JSObject()
        doc.props["createElement"] = NativeFunction(lambda this, a: CanvasShim())
        self.events = []                 # document.dispatchEvent log (facetrackingEvent / headtrackingEvent ...)
        self.now_ms = 1.0e12             # `new Date()` clock (tests may advance it)

        def create_event(this, a):
            e = JSObject()
            e.props["initEvent"] = NativeFunction(lambda th, aa: e.props.__setitem__("type", aa[0]) or undefined)
            return e
        doc.props["createEvent"] = NativeFunction(create_event)
        def dispatch(this, a):       # log a snapshot: src/main.js re-dispatches one mutated Event object
            e = JSObject()
            e.props.update(a[0].props)
            self.events.append(e)
            return True
        doc.props["dispatchEvent"] = NativeFunction(dispatch)
        self.timers = []                 # window.setTimeout log: [id, callback, ms, cleared]

        def set_timeout(this, a):
            self.timers.append([float(len(self.timers) + 1), a[0], to_number(a[1]) if len(a) > 1 else 0.0, False])
            return self.timers[-1][0]

        def clear_timeout(this, a):
            for t in self.timers:
                if a and t[0] == a[0]:
                    t[3] = True
            return undefined
        win = JSObject()
        win.props["setTimeout"] = NativeFunction(set_timeout)
        win.props["clearTimeout"] = NativeFunction(clear_timeout)
        g["window"] = win
        fproto = JSObject()
        fproto.props["bind"] = NativeFunction(lambda this, a: undefined)   # truthy: the bind polyfill of main.js is not needed
        fn_global = JSObject()
        fn_global.props["prototype"] = fproto
        g["Function"] = fn_global
        g["Date"] = NativeFunction(lambda this, a: JSDate(self.now_ms))
        g["document"] = doc
        g["headtrackr"] = JSObject()

    def run(self, src):
        body, hoisted = Parser(src).parse_program()
        for n in hoisted:
            self.genv.vars.setdefault(n, undefined)
        body(self.genv)

    def get(self, path):
        v = self.genv.vars[path[0]]
        for p in path[1:]:
            v = get_member(v, p)
        return v

    @staticmethod
    def call(f, this=undefined, *args):
        return call_function(f, this, list(args))


def to_py(v):
    """JS value -> plain Python (lists / dicts / floats)."""
    if isinstance(v, JSArray):
        return [to_py(x) for x in v.items]
    if isinstance(v, JSUint8ClampedArray):
        return v.buf.copy()
    if isinstance(v, (JSFunction, NativeFunction)):
        return "<function>"
    if isinstance(v, JSObject):
        return {k: to_py(x) for k, x in v.props.items() if not isinstance(x, (JSFunction, NativeFunction))}
    if v is undefined:
        return None
    return v
