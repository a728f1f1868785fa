"""Loads the reference's jsonnet experiment configs unchanged (``config/pretrain/*.jsonnet``) and exposes them
through the accessor surface the reference code uses on pyhocon's ``ConfigTree``
(``get_int / get_float / get_string / get_bool / get_list / get_config / put``; framework/config.py:44-75).

``_jsonnet`` and ``pyhocon`` are not installable here, so this module contains a small evaluator for the jsonnet
subset those files use: ``local x = import "...";``, object literals with ``:``, ``::`` (hidden) and ``+:`` fields,
object inheritance (``base { ... }`` / ``a + b``) with late-bound ``self`` / ``$`` / ``super``, field access, indexing,
``if/then/else``, ``std.length``, arithmetic / comparison, arrays, strings, numbers, booleans, null and comments.
Composition follows framework/config.py:14-62: ``base + arg0 + arg1 ...`` with ``add`` bound to the sibling
``addition.libsonnet`` inside each ``-x`` snippet.
"""
import json
import re
from pathlib import Path
from typing import Any, Dict, List, Optional

# ------------------------------------------------------------------------------------------------ tokenizer
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|\#[^\n]*|/\*.*?\*/)
  | (?P<num>\d+\.?\d*(?:[eE][+-]?\d+)?)
  | (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>:::|::|==|!=|<=|>=|&&|\|\||[{}\[\]().,;:+\-*/%<>=!$])
""", re.X | re.S)


def _tokenize(text: str):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"jsonnet: unexpected character {text[pos]!r} at offset {pos}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


# ------------------------------------------------------------------------------------------------ parser -> AST tuples
class _Parser:
    def __init__(self, text):
        self.toks = _tokenize(text)
        self.i = 0

    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"jsonnet: expected {val!r}, got {self.peek()[1]!r}")

    def parse(self):
        e = self.expr()
        if self.peek()[0] != "eof":
            raise SyntaxError(f"jsonnet: trailing input at {self.peek()[1]!r}")
        return e

    def expr(self):
        if self.peek() == ("id", "local"):
            self.next()
            name = self.next()[1]
            self.expect("=")
            val = self.expr()
            self.expect(";")
            return ("local", name, val, self.expr())
        if self.peek() == ("id", "if"):
            self.next()
            c = self.expr()
            self.expect("then")
            a = self.expr()
            b = ("lit", None)
            if self.accept("else"):
                b = self.expr()
            return ("if", c, a, b)
        return self.binary(0)

    _LEVELS = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]

    def binary(self, level):
        if level == len(self._LEVELS):
            return self.unary()
        left = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self._LEVELS[level]:
            op = self.next()[1]
            left = ("bin", op, left, self.binary(level + 1))
        return left

    def unary(self):
        if self.peek() == ("op", "-"):
            self.next()
            return ("neg", self.unary())
        if self.peek() == ("op", "!"):
            self.next()
            return ("not", self.unary())
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            if self.accept("."):
                e = ("field", e, self.next()[1])
            elif self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.accept("("):
                args = []
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
                e = ("call", e, args)
            elif self.peek() == ("op", "{"):
                e = ("bin", "+", e, self.primary())  # `base { ... }` is sugar for `base + { ... }`
            else:
                return e

    def primary(self):
        kind, val = self.next()
        if kind == "num":
            return ("lit", float(val) if re.search(r"[.eE]", val) else int(val))
        if kind == "str":
            body = val[1:-1]
            return ("lit", json.loads('"' + body.replace('"', '\\"') + '"') if val[0] == "'" else json.loads(val))
        if kind == "id":
            if val in ("true", "false"):
                return ("lit", val == "true")
            if val == "null":
                return ("lit", None)
            if val == "import":
                return ("import", self.primary()[1])
            if val == "self":
                return ("self",)
            if val == "super":
                return ("super",)
            return ("var", val)
        if val == "$":
            return ("dollar",)
        if val == "(":
            e = self.expr()
            self.expect(")")
            return e
        if val == "[":
            items = []
            while not self.accept("]"):
                items.append(self.expr())
                self.accept(",")
            return ("array", items)
        if val == "{":
            fields = []
            while not self.accept("}"):
                k, name = self.next()
                if k == "str":
                    name = name[1:-1]
                plus = self.accept("+")
                sep = self.next()[1]
                if sep not in (":", "::", ":::"):
                    raise SyntaxError(f"jsonnet: bad field separator {sep!r} after {name!r}")
                fields.append((name, plus, sep, self.expr()))
                self.accept(",")
            return ("object", fields)
        raise SyntaxError(f"jsonnet: unexpected token {val!r}")


# ------------------------------------------------------------------------------------------------ evaluator
class _Obj:
    """A jsonnet object: a stack of layers (base first), each layer maps name -> (plus, sep, expr, env)."""

    def __init__(self, layers):
        self.layers = layers

    def names(self):
        seen = []
        for layer in self.layers:
            for n in layer:
                if n not in seen:
                    seen.append(n)
        return seen

    def hidden(self, name):
        hid = False
        for layer in self.layers:
            if name in layer:
                sep = layer[name][1]
                hid = True if sep == "::" else (False if sep == ":::" else hid)
        return hid

    def get(self, name, top=None, self_obj=None):
        self_obj = self_obj or self
        top = len(self.layers) - 1 if top is None else top
        for i in range(top, -1, -1):
            if name in self.layers[i]:
                plus, _, expr, env = self.layers[i][name]
                scope = dict(env)
                scope["self"] = self_obj
                scope["super"] = (self, i - 1, self_obj)
                if "$" not in env:
                    scope["$"] = self_obj  # this literal is the outermost object of its file
                val = _eval(expr, scope)
                if plus and self.has(name, i - 1):
                    val = _add(self.get(name, i - 1, self_obj), val)
                return val
        raise KeyError(name)

    def has(self, name, top=None):
        top = len(self.layers) - 1 if top is None else top
        return any(name in self.layers[i] for i in range(top, -1, -1))


def _add(a, b):
    if isinstance(a, _Obj) and isinstance(b, _Obj):
        return _Obj(a.layers + b.layers)
    if isinstance(a, str) or isinstance(b, str):
        return str(a) + str(b)
    return a + b


def _eval(e, env) -> Any:
    t = e[0]
    if t == "lit":
        return e[1]
    if t == "var":
        if e[1] == "std":
            return "std"
        if e[1] not in env:
            raise NameError(f"jsonnet: unknown variable {e[1]!r}")
        v = env[e[1]]
        return v() if callable(v) else v
    if t == "local":
        scope = dict(env)
        cache = {}

        def thunk(expr=e[2], scope=scope):
            if "v" not in cache:
                cache["v"] = _eval(expr, scope)
            return cache["v"]
        scope[e[1]] = thunk
        return _eval(e[3], scope)
    if t == "import":
        return env["__import__"](e[1], env["__dir__"])
    if t == "object":
        inner = {k: v for k, v in env.items() if k not in ("self", "super")}
        if "self" in env:      # nested literal: `$` stays bound to the enclosing file's root
            inner["$"] = env["$"]
        layer = {name: (plus, sep, expr, inner) for name, plus, sep, expr in e[1]}
        return _Obj([layer])
    if t == "array":
        return [_eval(x, env) for x in e[1]]
    if t == "self":
        return env["self"]
    if t == "dollar":
        return env["$"]
    if t == "field":
        if e[1] == ("super",):
            obj, top, self_obj = env["super"]
            return obj.get(e[2], top, self_obj)
        base = _eval(e[1], env)
        if base == "std":
            return ("std", e[2])
        return base.get(e[2])
    if t == "index":
        base, idx = _eval(e[1], env), _eval(e[2], env)
        return base.get(idx) if isinstance(base, _Obj) else base[idx]
    if t == "call":
        fn = _eval(e[1], env)
        args = [_eval(a, env) for a in e[2]]
        if fn == ("std", "length"):
            return len(args[0].names()) if isinstance(args[0], _Obj) else len(args[0])
        raise NotImplementedError(f"jsonnet: call of {fn!r} is outside the supported subset")
    if t == "if":
        return _eval(e[2], env) if _eval(e[1], env) else _eval(e[3], env)
    if t == "neg":
        return -_eval(e[1], env)
    if t == "not":
        return not _eval(e[1], env)
    if t == "bin":
        op = e[1]
        if op == "&&":
            return _eval(e[2], env) and _eval(e[3], env)
        if op == "||":
            return _eval(e[2], env) or _eval(e[3], env)
        a, b = _eval(e[2], env), _eval(e[3], env)
        if op == "+":
            return _add(a, b)
        if op in ("==", "!="):
            eq = _manifest(a) == _manifest(b)
            return eq if op == "==" else not eq
        return {"-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b, "%": lambda: a % b, "<": lambda: a < b,
                ">": lambda: a > b, "<=": lambda: a <= b, ">=": lambda: a >= b}[op]()
    raise NotImplementedError(f"jsonnet: node {t}")


def _manifest(v):
    if isinstance(v, _Obj):
        return {n: _manifest(v.get(n)) for n in v.names() if not v.hidden(n)}
    if isinstance(v, list):
        return [_manifest(x) for x in v]
    return v


def evaluate_file(path, ext_config: Optional[List[str]] = None) -> Dict[str, Any]:
    """Evaluates ``base + arg0 + arg1 + ...`` (framework/config.py:14-62) and returns plain python data."""
    path = Path(path)
    ext_config = list(ext_config or [])

    def importer(rel, cur_dir):
        full = (Path(cur_dir) / rel).resolve() if not str(rel).startswith("/") else Path(rel)
        return _eval(_Parser(full.read_text()).parse(), {"__import__": importer, "__dir__": str(full.parent)})

    root_env = {"__import__": importer, "__dir__": str(path.parent.resolve())}
    value = importer(path.name, root_env["__dir__"])
    for snippet in ext_config:
        env = dict(root_env)
        addition = path.with_name("addition.libsonnet")
        env["add"] = (lambda a=addition: importer(a.name, str(a.parent.resolve())))
        value = _add(value, _eval(_Parser(snippet).parse(), env))
    return _manifest(value)


# ------------------------------------------------------------------------------------------------ ConfigTree-like view
class Config:
    """Dict-backed view with pyhocon's accessor names; dotted paths address nested tables."""
    _MISSING = object()

    def __init__(self, data: Optional[Dict[str, Any]] = None):
        self._d = data if data is not None else {}

    def _find(self, key, default=_MISSING):
        cur = self._d
        for part in str(key).split("."):
            if isinstance(cur, Config):
                cur = cur._d
            if not isinstance(cur, dict) or part not in cur:
                if default is self._MISSING:
                    raise KeyError(f"No configuration setting found for key {key}")
                return default
            cur = cur[part]
        return cur

    def get(self, key, default=_MISSING):
        v = self._find(key, default)
        return Config(v) if isinstance(v, dict) else v

    def get_int(self, key, default=_MISSING):
        v = self._find(key, default)
        return v if v is None else int(v)

    def get_float(self, key, default=_MISSING):
        v = self._find(key, default)
        return v if v is None else float(v)

    def get_string(self, key, default=_MISSING):
        v = self._find(key, default)
        if isinstance(v, bool):
            return str(v).lower()
        return v if v is None else str(v)

    def get_bool(self, key, default=_MISSING):
        v = self._find(key, default)
        if isinstance(v, str):
            return {"true": True, "yes": True, "on": True, "false": False, "no": False, "off": False}[v.lower()]
        return v if v is None else bool(v)

    def get_list(self, key, default=_MISSING):
        v = self._find(key, default)
        return v if v is None else list(v)

    def get_config(self, key, default=_MISSING):
        v = self._find(key, default)
        if isinstance(v, Config):
            return v
        if v is None or isinstance(v, dict):
            return Config(v) if v is not None else None
        raise TypeError(f"{key} is not a table")

    def put(self, key, value):
        parts = str(key).split(".")
        cur = self._d
        for part in parts[:-1]:
            cur = cur.setdefault(part, {})
        cur[parts[-1]] = value

    def __getitem__(self, key):
        return self.get(key)

    def __contains__(self, key):
        return self._find(key, None) is not None or self._find(key, 0) != 0

    def keys(self):
        return self._d.keys()

    def items(self):
        return ((k, self.get(k)) for k in self._d)

    def as_plain_ordered_dict(self):
        return json.loads(json.dumps(self._d))

    def __repr__(self):
        return f"Config({json.dumps(self._d, indent=1)})"


def get_config(config_path, ext_config: Optional[List[str]] = None) -> Config:
    """Counterpart of framework/config.py:44-75 (``args.config`` + ``args.ext_config``)."""
    return Config(evaluate_file(config_path, ext_config))


def trim_moco_k(k: int, batch_size: int, world_size: int) -> int:
    """utils/moco.py:8-10: largest multiple of the global batch not above k."""
    total = batch_size * world_size
    return k // total * total
