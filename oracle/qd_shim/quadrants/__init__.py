"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the `quadrants` package (the Taichi-like JIT DSL of the reference, pinned at
quadrants==0.5.2 in /root/reference/pyproject.toml and absent from this image) so that the reference's OWN kernel sources under
/root/reference/genesis import and EXECUTE on the CPU as plain Python.  Nothing here restates any reference algorithm: it only
gives `@qd.func` / `@qd.kernel` bodies the value semantics the DSL defines:

  * `qd.func` / `qd.kernel` re-compile the decorated function from its source after an AST pass that
      - copies matrix values on assignment to a local name (DSL locals own their data; Python would alias),
      - rewrites `qd.atomic_*(target, v)` into read-modify-write of the target expression, returning the old value,
      - maps builtin `min` / `max` / `abs` to their element-wise DSL versions,
    and pass matrix arguments BY VALUE unless the parameter is annotated `qd.template()` (by reference, as in the DSL);
  * `qd.field / Vector.field / Matrix.field / ndarray` are numpy-backed; indexing a vector / matrix field yields a VIEW so that
    `f[i][k] = x` and `f[i] += v` store through, as in the DSL;
  * `qd.static`, `qd.loop_config`, `qd.ndrange`, `qd.grouped`, math functions, dtypes-as-casts.

All floating point is float64 (`qd.f32` does not round) -- fixtures produced through this shim pin the fp64 instantiation of
oracle/rigid_oracle.hpp.  Used only by tools/make_ref_golden.py (run in the build container, where /root/reference exists).
"""

import ast
import functools
import inspect
import itertools
import math as _pymath
import sys
import textwrap
import types as _pytypes

import numpy as np

np.seterr(all="ignore")

__version__ = "0.5.2-shim"


# ------------------------------------------------------------------------------------------------ dtypes
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.np_dtype = name, np_dtype

    def __call__(self, x=0):
        if isinstance(x, Mat):
            return Mat._wrap(x.a.astype(self.np_dtype))
        if self.np_dtype is np.int64 and isinstance(x, (float, np.floating)) and not np.isfinite(x):
            return np.int64(0)  # float -> int of NaN / inf: 0 as on CUDA (cvt.rzi), where the reference's GPU backend runs
        return self.np_dtype(x)

    def __repr__(self):
        return f"qd.{self.name}"


f32 = float32 = DType("f32", np.float64)  # no rounding: fixtures are fp64 throughout
f64 = float64 = DType("f64", np.float64)
f16 = DType("f16", np.float64)
i8 = int8 = DType("i8", np.int64)
i16 = int16 = DType("i16", np.int64)
i32 = int32 = DType("i32", np.int64)
i64 = int64 = DType("i64", np.int64)
u8 = uint8 = DType("u8", np.int64)
u16 = uint16 = DType("u16", np.int64)
u32 = uint32 = DType("u32", np.int64)
u64 = uint64 = DType("u64", np.int64)
u1 = DType("u1", np.bool_)

cpu, gpu, cuda, metal, amdgpu, vulkan = "cpu", "gpu", "cuda", "metal", "amdgpu", "vulkan"
ERROR, WARN, INFO, DEBUG, TRACE = range(5)


def _np_dtype(dt):
    if dt is None:
        return None
    if isinstance(dt, DType):
        return dt.np_dtype
    return np.dtype(dt).type


def init(*a, **k):
    pass


def reset():
    pass


def sync():
    pass


# ------------------------------------------------------------------------------------------------ Vector / Matrix values
def _raw(x):
    return x.a if type(x) is Mat else x


class Mat:
    """Vector (1-D) or matrix (2-D) value backed by a numpy array (possibly a view into a field's storage)."""

    __slots__ = ("a",)
    __array_priority__ = 1000

    def __init__(self, data=None, dt=None, **_):
        if type(data) is Mat:
            a = data.a.copy()
        else:
            a = np.array([[_raw(e) for e in row] if isinstance(row, (list, tuple)) else _raw(row) for row in data])
        npdt = _np_dtype(dt)
        if npdt is not None and a.dtype != npdt:
            a = a.astype(npdt)
        elif a.dtype == np.float32:
            a = a.astype(np.float64)
        elif a.dtype.kind in "iu" and a.dtype != np.int64:
            a = a.astype(np.int64)
        self.a = a

    @staticmethod
    def _wrap(a):
        m = Mat.__new__(Mat)
        m.a = a
        return m

    def copy(self):
        return Mat._wrap(self.a.copy())

    # ---- constructors
    @staticmethod
    def zero(dt, n, m=None):
        return Mat._wrap(np.zeros((n,) if m is None else (n, m), _np_dtype(dt)))

    @staticmethod
    def one(dt, n, m=None):
        return Mat._wrap(np.ones((n,) if m is None else (n, m), _np_dtype(dt)))

    @staticmethod
    def identity(dt, n):
        return Mat._wrap(np.eye(n, dtype=_np_dtype(dt)))

    @staticmethod
    def unit(n, i, dt=None):
        a = np.zeros(n, _np_dtype(dt) or np.float64)
        a[i] = 1
        return Mat._wrap(a)

    @staticmethod
    def rows(rows):
        return Mat._wrap(np.stack([np.asarray(_raw(r)) for r in rows], 0))

    @staticmethod
    def cols(cols):
        return Mat._wrap(np.stack([np.asarray(_raw(c)) for c in cols], 1))

    @staticmethod
    def field(*args, **kw):
        # Vector.field(n, dtype, shape) / Matrix.field(n, m, dtype, shape): told apart by the position of the dtype
        args = list(args)
        if len(args) >= 2 and isinstance(args[1], DType) or ("dtype" in kw and len(args) == 1):
            n = args.pop(0)
            vec = (n,)
        else:
            n = kw.pop("n", None) if not args else args.pop(0)
            m = kw.pop("m", None) if not args else args.pop(0)
            vec = (n, m)
        dtype = kw.pop("dtype", None) if not args else args.pop(0)
        shape = kw.pop("shape", None) if not args else args.pop(0)
        return Field(dtype, shape, vec)

    ndarray = field

    # ---- shape
    @property
    def n(self):
        return self.a.shape[0]

    @property
    def m(self):
        return self.a.shape[1] if self.a.ndim > 1 else 1

    def get_shape(self):
        return self.a.shape

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        if self.a.ndim == 1:
            return iter(self.a)
        return (Mat._wrap(r) for r in self.a)

    # ---- element access
    def __getitem__(self, i):
        if type(i) is tuple:
            i = tuple(int(_raw(k)) if not isinstance(k, slice) else k for k in i)
        elif not isinstance(i, slice):
            i = int(_raw(i))
        r = self.a[i]
        return Mat._wrap(r) if isinstance(r, np.ndarray) else r

    def __setitem__(self, i, v):
        if type(i) is tuple:
            i = tuple(int(_raw(k)) if not isinstance(k, slice) else k for k in i)
        elif not isinstance(i, slice):
            i = int(_raw(i))
        self.a[i] = _raw(v)

    def __getattr__(self, name):
        if name and all(c in "xyzw" for c in name):
            idx = ["xyzw".index(c) for c in name]
            return self.a[idx[0]] if len(idx) == 1 else Mat._wrap(self.a[idx].copy())
        if name and all(c in "rgba" for c in name):
            idx = ["rgba".index(c) for c in name]
            return self.a[idx[0]] if len(idx) == 1 else Mat._wrap(self.a[idx].copy())
        raise AttributeError(name)

    # ---- arithmetic
    def __add__(self, o): return Mat._wrap(self.a + _raw(o))
    def __radd__(self, o): return Mat._wrap(_raw(o) + self.a)
    def __sub__(self, o): return Mat._wrap(self.a - _raw(o))
    def __rsub__(self, o): return Mat._wrap(_raw(o) - self.a)
    def __mul__(self, o): return Mat._wrap(self.a * _raw(o))
    def __rmul__(self, o): return Mat._wrap(_raw(o) * self.a)
    def __truediv__(self, o): return Mat._wrap(self.a / _raw(o))
    def __rtruediv__(self, o): return Mat._wrap(_raw(o) / self.a)
    def __floordiv__(self, o): return Mat._wrap(self.a // _raw(o))
    def __mod__(self, o): return Mat._wrap(self.a % _raw(o))
    def __pow__(self, o): return Mat._wrap(self.a ** _raw(o))
    def __neg__(self): return Mat._wrap(-self.a)
    def __pos__(self): return self
    def __abs__(self): return Mat._wrap(np.abs(self.a))
    def __matmul__(self, o): return Mat._wrap(self.a @ _raw(o))
    def __rmatmul__(self, o): return Mat._wrap(_raw(o) @ self.a)
    def __and__(self, o): return Mat._wrap(self.a & _raw(o))
    def __or__(self, o): return Mat._wrap(self.a | _raw(o))
    def __invert__(self): return Mat._wrap(~self.a)

    def _inplace(self, r):
        if self.a.dtype.kind in "iub" and np.asarray(r).dtype.kind == "f":
            r = np.trunc(r)
        self.a[...] = r
        return self

    def __iadd__(self, o): return self._inplace(self.a + _raw(o))
    def __isub__(self, o): return self._inplace(self.a - _raw(o))
    def __imul__(self, o): return self._inplace(self.a * _raw(o))
    def __itruediv__(self, o): return self._inplace(self.a / _raw(o))

    def __lt__(self, o): return Mat._wrap(self.a < _raw(o))
    def __le__(self, o): return Mat._wrap(self.a <= _raw(o))
    def __gt__(self, o): return Mat._wrap(self.a > _raw(o))
    def __ge__(self, o): return Mat._wrap(self.a >= _raw(o))
    def __eq__(self, o): return Mat._wrap(self.a == _raw(o))
    def __ne__(self, o): return Mat._wrap(self.a != _raw(o))
    __hash__ = None

    def __bool__(self):
        if self.a.size == 1:
            return bool(self.a.reshape(-1)[0])
        raise TypeError("a vector has no truth value in the DSL (use .all() / .any())")

    def all(self): return bool(self.a.all())
    def any(self): return bool(self.a.any())

    # ---- linear algebra
    def dot(self, o): return np.dot(self.a, _raw(o))
    def cross(self, o):
        a, b = self.a, _raw(o)
        if a.shape[0] == 2:
            return a[0] * b[1] - a[1] * b[0]
        return Mat._wrap(np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]))
    def norm(self, eps=0.0): return np.sqrt(np.sum(self.a * self.a) + eps)
    def norm_sqr(self): return np.sum(self.a * self.a)
    def norm_inv(self, eps=0.0): return 1.0 / np.sqrt(np.sum(self.a * self.a) + eps)
    def normalized(self, eps=0.0): return Mat._wrap(self.a / np.sqrt(np.sum(self.a * self.a) + eps))
    def outer_product(self, o): return Mat._wrap(np.outer(self.a, _raw(o)))
    def transpose(self): return Mat._wrap(self.a.T.copy())
    def inverse(self): return Mat._wrap(np.linalg.inv(self.a))
    def determinant(self): return np.linalg.det(self.a)
    def trace(self): return np.trace(self.a)
    def sum(self): return self.a.sum()
    def max(self): return self.a.max()
    def min(self): return self.a.min()
    def fill(self, v): self.a[...] = _raw(v)
    def cast(self, dt): return Mat._wrap(self.a.astype(_np_dtype(dt)))
    def to_numpy(self): return self.a.copy()
    def to_list(self): return self.a.tolist()

    def __repr__(self):
        return f"qd.Mat({self.a!r})"


Vector = Matrix = Mat


# ------------------------------------------------------------------------------------------------ fields
class Field:
    """Scalar / vector / matrix field (or ndarray) over a dense index shape, numpy-backed."""

    def __init__(self, dtype, shape, vec=(), **_):
        if isinstance(dtype, _VecType):  # qd.field(dtype=<vector / matrix type>, shape=...)
            dtype, vec = dtype.dtype, dtype.shape
        if shape is None:
            shape = ()
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.vec = tuple(int(v) for v in vec)
        self.dtype = dtype
        self.np = np.zeros(self.shape + self.vec, _np_dtype(dtype) or np.float64)

    @staticmethod
    def _idx(i):
        t = type(i)
        if t is tuple or t is list:
            return tuple(int(_raw(k)) for k in i)
        if i is None:
            return ()
        if t is Mat:
            return tuple(int(k) for k in i.a)
        return int(i)

    def __getitem__(self, i):
        r = self.np[Field._idx(i)]
        return Mat._wrap(r) if self.vec else r

    def __setitem__(self, i, v):
        v = _raw(v)
        if self.np.dtype.kind in "iu" and np.asarray(v).dtype.kind == "f":
            v = np.trunc(v)
        self.np[Field._idx(i)] = v

    def fill(self, v):
        self.np[...] = _raw(v)

    def from_numpy(self, arr):
        self.np[...] = np.asarray(arr).reshape(self.np.shape)

    def to_numpy(self, *a, **k):
        return self.np.copy()

    def from_torch(self, t):
        self.from_numpy(t.detach().cpu().numpy())

    def to_torch(self, *a, **k):
        import torch

        return torch.from_numpy(self.np.copy())

    def copy_from(self, other):
        self.np[...] = other.np

    @property
    def grad(self):
        return self

    def __repr__(self):
        return f"qd.Field(shape={self.shape}, vec={self.vec}, dtype={self.dtype})"


def field(dtype=None, shape=None, **kw):
    return Field(dtype, shape, (), **kw)


def ndarray(dtype=None, shape=None, **kw):
    return Field(dtype, shape, (), **kw)


ScalarField = MatrixField = ScalarNdarray = VectorNdarray = MatrixNdarray = Ndarray = Field


# ------------------------------------------------------------------------------------------------ annotations / types
class _Template:
    def __call__(self, *a, **k):
        return self

    def __repr__(self):
        return "qd.template()"


template = _Template()


class _VecType:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = shape, dtype

    def __call__(self, *vals):
        if len(vals) == 1 and isinstance(vals[0], (list, tuple, Mat, np.ndarray)):
            vals = vals[0]
        a = np.array([_raw(v) for v in vals], _np_dtype(self.dtype)).reshape(self.shape)
        return Mat._wrap(a)


class _Types:
    @staticmethod
    def ndarray(*a, **k):
        return template

    @staticmethod
    def vector(n, dtype=None):
        return _VecType((n,), dtype)

    @staticmethod
    def matrix(n, m, dtype=None):
        return _VecType((n, m), dtype)

    NDArray = template


types = _Types()


# ------------------------------------------------------------------------------------------------ control / ranges
def static(x, *more):
    return x if not more else (x,) + more


def loop_config(**k):
    pass


def _rng(a):
    if isinstance(a, (tuple, list)):
        return range(int(a[0]), int(a[1]))
    return range(int(a))


def ndrange(*args):
    if len(args) == 1:
        return ((i,) for i in _rng(args[0])) if False else _rng(args[0])
    return itertools.product(*[_rng(a) for a in args])


def grouped(it):
    if isinstance(it, Field):
        it = np.ndindex(*it.shape)
    for i in it:
        yield Mat._wrap(np.array(i if isinstance(i, tuple) else (i,), np.int64))


def data_oriented(cls):
    return cls


def cast(x, dt):
    return dt(x)


def select(c, a, b):
    if isinstance(c, Mat):
        return Mat._wrap(np.where(c.a, _raw(a), _raw(b)))
    return a if c else b


def random(dt=None):
    return np.random.random()


# ------------------------------------------------------------------------------------------------ math
def _lift(fn):
    def g(*xs):
        r = fn(*[_raw(x) for x in xs])
        return Mat._wrap(r) if isinstance(r, np.ndarray) and r.ndim > 0 else r

    return g


sqrt, sin, cos, tan, asin, acos, exp, log, floor, ceil, tanh = (_lift(f) for f in (np.sqrt, np.sin, np.cos, np.tan, np.arcsin, np.arccos, np.exp, np.log, np.floor, np.ceil, np.tanh))
atan2 = _lift(np.arctan2)
abs = _lift(np.abs)  # noqa: A001
pow = _lift(np.power)  # noqa: A001
rsqrt = _lift(lambda x: 1.0 / np.sqrt(x))
round = _lift(np.round)  # noqa: A001


def _reduce(fn):
    def g(*xs):
        r = _raw(xs[0])
        for x in xs[1:]:
            r = fn(r, _raw(x))
        return Mat._wrap(r) if isinstance(r, np.ndarray) and r.ndim > 0 else r

    return g


max = _reduce(np.fmax)  # noqa: A001  (NaN-ignoring like fmaxf / llvm.maxnum, which the DSL lowers to)
min = _reduce(np.fmin)  # noqa: A001
raw_mod = _lift(np.fmod)
raw_div = _lift(lambda a, b: np.trunc(a / b))


class _Math:
    inf = float("inf")
    nan = float("nan")
    pi = _pymath.pi
    e = _pymath.e
    sqrt, sin, cos, tan, asin, acos, exp, log, floor, ceil, atan2, pow = (staticmethod(f) for f in (sqrt, sin, cos, tan, asin, acos, exp, log, floor, ceil, atan2, pow))
    isnan = staticmethod(_lift(np.isnan))
    isinf = staticmethod(_lift(np.isinf))
    clamp = staticmethod(_lift(lambda x, lo, hi: np.fmin(np.fmax(x, lo), hi)))
    sign = staticmethod(_lift(np.sign))
    radians = staticmethod(_lift(np.radians))
    degrees = staticmethod(_lift(np.degrees))
    max = staticmethod(max)
    min = staticmethod(min)
    dot = staticmethod(lambda a, b: a.dot(b))
    cross = staticmethod(lambda a, b: a.cross(b))
    length = staticmethod(lambda a: a.norm())
    normalize = staticmethod(lambda a: a.normalized())
    vec2 = _VecType((2,), f64)
    vec3 = _VecType((3,), f64)
    vec4 = _VecType((4,), f64)
    ivec2 = _VecType((2,), i32)
    ivec3 = _VecType((3,), i32)
    ivec4 = _VecType((4,), i32)
    mat2 = _VecType((2, 2), f64)
    mat3 = _VecType((3, 3), f64)
    mat4 = _VecType((4, 4), f64)


math = _Math()


# ------------------------------------------------------------------------------------------------ atomics (after the AST pass)
def _atomic(op, container, index, value):
    """read-modify-write of container[index]; returns the old value (the DSL's atomic_* return the old value)."""
    old = container[index]
    if type(old) is Mat:
        old = old.copy()
    v = _raw(value)
    cur = _raw(old)
    if op == "add": new = cur + v
    elif op == "sub": new = cur - v
    elif op == "max": new = np.maximum(cur, v)
    elif op == "min": new = np.minimum(cur, v)
    elif op == "or": new = cur | v
    elif op == "and": new = cur & v
    elif op == "xor": new = cur ^ v
    else: raise NotImplementedError(op)
    container[index] = new
    return old


def _atomic_value(op, cur, value):
    v, c = _raw(value), _raw(cur)
    if op == "add": r = c + v
    elif op == "sub": r = c - v
    elif op == "max": r = np.maximum(c, v)
    elif op == "min": r = np.minimum(c, v)
    elif op == "or": r = c | v
    elif op == "and": r = c & v
    elif op == "xor": r = c ^ v
    else: raise NotImplementedError(op)
    return Mat._wrap(r) if isinstance(r, np.ndarray) and r.ndim > 0 else r


def _val(x):
    """value semantics of `name = expr`: matrices are copied (DSL locals own their storage)."""
    return Mat._wrap(x.a.copy()) if type(x) is Mat else x


def _plain_atomic(op):
    def g(target, value):  # un-rewritten call (Python scope): cannot store through, only meaningful for its return value
        raise RuntimeError(f"qd.atomic_{op} reached without the AST pass (called from Python scope?)")

    return g


atomic_add, atomic_sub, atomic_max, atomic_min, atomic_or, atomic_and, atomic_xor = (_plain_atomic(o) for o in ("add", "sub", "max", "min", "or", "and", "xor"))


class _Rewrite(ast.NodeTransformer):
    ATOMICS = {"atomic_add": "add", "atomic_sub": "sub", "atomic_max": "max", "atomic_min": "min", "atomic_or": "or", "atomic_and": "and", "atomic_xor": "xor"}
    BUILTINS = {"max": "__qd_max", "min": "__qd_min", "abs": "__qd_abs"}

    def __init__(self):
        self.depth = 0

    def visit_FunctionDef(self, node):
        self.depth += 1
        if self.depth == 1:
            node.decorator_list = []
            # annotations are evaluated at definition time and may reference names that only exist for typing: drop them
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs:
                a.annotation = None
        self.generic_visit(node)
        self.depth -= 1
        return node

    def visit_Assign(self, node):
        self.generic_visit(node)
        if all(isinstance(t, ast.Name) for t in node.targets) and not isinstance(node.value, (ast.Constant, ast.Tuple)):
            node.value = ast.Call(func=ast.Name(id="__qd_val", ctx=ast.Load()), args=[node.value], keywords=[])
        elif isinstance(node.value, ast.Tuple):
            # `a[i], a[j] = a[j], a[i]`: the DSL loads every right-hand value before storing; field reads are views here
            node.value.elts = [e if isinstance(e, ast.Constant) else ast.Call(func=ast.Name(id="__qd_val", ctx=ast.Load()), args=[e], keywords=[])
                               for e in node.value.elts]
        return node

    def visit_AnnAssign(self, node):
        self.generic_visit(node)
        node.annotation = ast.Constant(value=None)
        if node.value is not None and isinstance(node.target, ast.Name) and not isinstance(node.value, ast.Constant):
            node.value = ast.Call(func=ast.Name(id="__qd_val", ctx=ast.Load()), args=[node.value], keywords=[])
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        f = node.func
        if isinstance(f, ast.Attribute) and f.attr in self.ATOMICS and isinstance(f.value, ast.Name) and f.value.id in ("qd", "ti"):
            op = self.ATOMICS[f.attr]
            tgt, val = node.args[0], node.args[1]
            if isinstance(tgt, ast.Subscript):
                return ast.Call(func=ast.Name(id="__qd_atomic", ctx=ast.Load()),
                                args=[ast.Constant(value=op), tgt.value, tgt.slice, val], keywords=[])
            if isinstance(tgt, ast.Name):
                # (old, (name := op(name, v)))[0]
                new = ast.Call(func=ast.Name(id="__qd_atomic_value", ctx=ast.Load()),
                               args=[ast.Constant(value=op), ast.Name(id=tgt.id, ctx=ast.Load()), val], keywords=[])
                tup = ast.Tuple(elts=[ast.Name(id=tgt.id, ctx=ast.Load()), ast.NamedExpr(target=ast.Name(id=tgt.id, ctx=ast.Store()), value=new)], ctx=ast.Load())
                return ast.Subscript(value=tup, slice=ast.Constant(value=0), ctx=ast.Load())
            raise NotImplementedError(f"atomic target {ast.dump(tgt)}")
        if isinstance(f, ast.Name) and f.id in self.BUILTINS:
            node.func = ast.Name(id=self.BUILTINS[f.id], ctx=ast.Load())
        return node


_HELPERS = {"__qd_val": _val, "__qd_atomic": _atomic, "__qd_atomic_value": _atomic_value, "__qd_max": max, "__qd_min": min, "__qd_abs": abs}


def _compile(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    # parameters annotated qd.template() (or anything that is not a value type) are passed by reference
    by_ref = []
    for a in fdef.args.args:
        ann = ast.unparse(a.annotation) if a.annotation is not None else ""
        by_ref.append("template" in ann)
    tree = _Rewrite().visit(tree)
    ast.fix_missing_locations(tree)
    ast.increment_lineno(tree, fn.__code__.co_firstlineno - 1)  # tracebacks point at the reference file's real lines
    g = fn.__globals__
    for k, v in _HELPERS.items():
        g.setdefault(k, v)
    code = compile(tree, inspect.getsourcefile(fn) or "<qd_shim>", "exec")
    ns = {}
    exec(code, g, ns)
    new = ns[fdef.name]
    if fn.__defaults__:
        new.__defaults__ = fn.__defaults__
    if fn.__kwdefaults__:
        new.__kwdefaults__ = fn.__kwdefaults__
    return new, by_ref


def _decorate(fn):
    state = {}

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if "f" not in state:
            state["f"], state["by_ref"] = _compile(fn)
        by_ref = state["by_ref"]
        if args:
            args = tuple(a if (type(a) is not Mat or (i < len(by_ref) and by_ref[i])) else Mat._wrap(a.a.copy()) for i, a in enumerate(args))
        if kwargs:
            kwargs = {k: (Mat._wrap(v.a.copy()) if type(v) is Mat else v) for k, v in kwargs.items()}
        return state["f"](*args, **kwargs)

    wrapper.__qd_shim__ = True
    wrapper._primal = fn
    return wrapper


def func(fn=None, **kw):
    if fn is None:
        return lambda f: _decorate(f)
    return _decorate(fn)


def kernel(fn=None, **kw):
    if fn is None:
        return lambda f: _decorate(f)
    return _decorate(fn)


pyfunc = func
real_func = func


class _PerfDispatch:
    """qd.perf_dispatch: a prototype with several registered implementations; the real package times them and keeps the fastest
    compatible one.  The shim calls the FIRST registered compatible implementation (registration order = the order of the
    reference source: the monolithic solver body comes first)."""

    def __init__(self, proto):
        self.proto, self.impls = proto, []
        functools.update_wrapper(self, proto)

    def register(self, fn=None, is_compatible=None, **_):
        def deco(f):
            self.impls.append((is_compatible, f))
            return f

        return deco(fn) if fn is not None else deco

    def __call__(self, *args, **kwargs):
        for ok, f in self.impls:
            if ok is None or ok(*args, **kwargs):
                return f(*args, **kwargs)
        raise RuntimeError(f"no compatible implementation registered for {self.proto.__name__}")


def perf_dispatch(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return _PerfDispatch(a[0])
    return lambda proto: _PerfDispatch(proto)


# ------------------------------------------------------------------------------------------------ misc namespaces used at import time
class _NS(_pytypes.SimpleNamespace):
    def __getattr__(self, name):
        raise AttributeError(f"quadrants shim: '{name}' is not provided")


def _not_on_cpu(*a, **k):
    raise NotImplementedError("GPU-only construct (qd.simt) reached: the shim only runs the reference's CPU code path")


simt = _NS(block=_NS(sync=_not_on_cpu, SharedArray=_not_on_cpu), warp=_NS(sync=_not_on_cpu, shfl_down_f32=_not_on_cpu))
lang = _NS(impl=_NS(get_max_shared_memory_bytes=lambda: 48 * 1024, current_cfg=lambda: _NS(arch="cpu")), util=_NS(), misc=_NS())
_logging = _NS(debug=print, trace=print, info=print, warn=print, error=print, critical=print)


def graph_do_while(*a, **k):
    raise NotImplementedError("qd.graph_do_while (CUDA graph loop) is GPU-only")
