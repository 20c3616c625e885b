"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

A NumPy restatement of the reference's gradient-tape hot path (dfdx/Yota.jl v0.8.5):
trace -> chainrules_transform! -> back! -> finalize_grad!, executed op-by-op, unfused, in
the reference's op order and with the reference's materialised temporaries.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline leg / --impl reference)
may import this module.

Parity status: the reference itself cannot run here (no Julia in the image) and its
arithmetic lives in un-vendored packages (ChainRules >=1.43, NNlib 0.8/0.9, Base,
OpenBLAS -- /root/reference/Project.toml:18-24).  This oracle is PINNED against every
known-answer the reference's tests hold for the path (tests/golden/known_answers.json,
from test/test_helpers.jl:5-25, test/test_grad.jl:71-95,175-182,253-261) and against
central finite differences at the reference's own tolerance (src/gradcheck.jl:3-9,
rtol=atol=1e-5, Float64) on the reference's gradcheck cases (test/test_grad.jl:28-68,
129-173).  For Float32 at benchmark scale there are no reference fixtures:
"parity unpinned" beyond those.

Arrays are Julia arrays: column-major (NumPy order='F'), 1-based indices, broadcast
aligned from the FIRST dimension.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# tangent markers (ChainRulesCore ZeroTangent / NoTangent / Thunk)
# --------------------------------------------------------------------------------------


class _Marker:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name + "()"


ZeroTangent = _Marker("ZeroTangent")
NoTangent = _Marker("NoTangent")


def iszerotangent(x):  # src/utils.jl:21-23
    return x is ZeroTangent or x is NoTangent


class Thunk:
    """ChainRulesCore.@thunk: a lazy tangent forced by `unthunk`."""

    def __init__(self, fn):
        self.fn = fn
        self._v = None
        self._done = False

    def force(self):
        if not self._done:
            self._v = self.fn()
            self._done = True
            self.fn = None
        return self._v


def unthunk(x):
    return x.force() if isinstance(x, Thunk) else x


class Colon:
    def __repr__(self):
        return ":"


colon = Colon()


class JlRange:
    """a:b (inclusive, 1-based)."""

    def __init__(self, a, b):
        self.a, self.b = int(a), int(b)

    def __len__(self):
        return max(0, self.b - self.a + 1)


# --------------------------------------------------------------------------------------
# Julia array helpers
# --------------------------------------------------------------------------------------


def jl_array(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    return np.asfortranarray(a)


def _pad(a, nd):
    a = np.asarray(a)
    return a.reshape(a.shape + (1,) * (nd - a.ndim), order="F") if a.ndim < nd else a


def jl_broadcast(fn, *xs):
    """Base.broadcast: dims aligned from the first dim, missing trailing dims = 1."""
    nd = max((np.ndim(x) for x in xs), default=0)
    ys = [_pad(x, nd) if isinstance(x, np.ndarray) else x for x in xs]
    out = fn(*ys)
    return np.asfortranarray(out) if isinstance(out, np.ndarray) and out.ndim > 1 else out


def unbroadcast(x, d):
    """ChainRules `unbroadcast(x, dx)`; same spec as src/helpers.jl:63-74."""
    if iszerotangent(d):
        return d
    if not isinstance(x, np.ndarray) or x.ndim == 0:
        if isinstance(d, np.ndarray) and d.ndim > 0:
            return d.dtype.type(d.sum()) if d.size else d.dtype.type(0)
        return d
    d = np.asarray(d)
    if x.shape == d.shape:
        return d
    if x.size == d.size:
        return d.reshape(x.shape, order="F")
    dims = tuple(i for i in range(d.ndim) if i >= x.ndim or x.shape[i] == 1)
    s = d.sum(axis=dims, keepdims=True)
    return np.asfortranarray(s.reshape(x.shape, order="F"))


def _jl_index_lists(shape, I):
    """Translate Julia index tuple -> (per-dim 0-based index arrays, kept-dims mask, work shape).

    A single index on an N-d array is linear indexing (column-major)."""
    if len(I) == 1 and len(shape) != 1:
        shape = (int(np.prod(shape)),)
    if len(I) != len(shape):
        # trailing singleton indices are legal in Julia; only the common forms are needed
        raise IndexError(f"index {I} does not match dims {shape}")
    lists, keep = [], []
    for n, ix in zip(shape, I):
        if type(ix).__name__ == "Colon":  # duck-typed: any array package's `:` / `a:b`
            lists.append(np.arange(n, dtype=np.int64)); keep.append(True)
        elif hasattr(ix, "a") and hasattr(ix, "b"):
            lists.append(np.arange(ix.a - 1, ix.b, dtype=np.int64)); keep.append(True)
        elif isinstance(ix, (int, np.integer)):
            lists.append(np.array([int(ix) - 1], dtype=np.int64)); keep.append(False)
        else:
            v = np.asarray(ix)
            assert v.ndim == 1 and np.issubdtype(v.dtype, np.integer), "index vectors must be 1-d ints"
            lists.append(v.astype(np.int64) - 1); keep.append(True)
        if lists[-1].size and (lists[-1].min() < 0 or lists[-1].max() >= n):
            raise IndexError("BoundsError")
    return lists, keep, shape


def _linear_dest(shape, lists):
    """Column-major linear destination index for every element of the gathered block,
    laid out column-major over the (unsqueezed) block."""
    strides = np.cumprod((1,) + tuple(shape[:-1])).astype(np.int64)
    L = np.zeros((1,) * len(lists), dtype=np.int64)
    for d, (l, s) in enumerate(zip(lists, strides)):
        sh = [1] * len(lists)
        sh[d] = l.size
        L = L + (l * s).reshape(sh)
    return L


def jl_getindex(x, I):
    lists, keep, shape = _jl_index_lists(x.shape, I)
    L = _linear_dest(shape, lists)
    flat = x.reshape(-1, order="F")
    y = flat[L.reshape(-1, order="F")].reshape(L.shape, order="F")
    out_shape = tuple(n for n, k in zip(L.shape, keep) if k)
    if not out_shape:
        return y.reshape(()).copy()[()] if x.dtype.kind != "f" else x.dtype.type(y.reshape(-1)[0])
    return np.asfortranarray(y.reshape(out_shape, order="F"))


def scatter_add_seq(dx, dy, L):
    """dx[L[k]] += dy[k] for k in column-major order of dy -- the sequential left fold
    of NNlib.scatter!(+, dx, dy, idx) (src/helpers.jl:7-11) and ChainRules' ∇getindex!
    (SURVEY App. A.8).  np.add.at is unbuffered and processes its index array in order."""
    flat = np.ascontiguousarray(dx.reshape(-1, order="F"))
    src = np.asarray(dy, dtype=dx.dtype).reshape(-1, order="F")
    np.add.at(flat, L.reshape(-1, order="F"), src)
    return flat.reshape(dx.shape, order="F")


def jl_ungetindex(x_shape, dtype, dy, I):
    """ungetindex(x, dy, I...) = ungetindex!(zero(x), x, dy, I...)  (src/helpers.jl:30-33)."""
    lists, keep, shape = _jl_index_lists(x_shape, I)
    L = _linear_dest(shape, lists)
    dx = np.zeros(int(np.prod(x_shape)), dtype=dtype)
    dyv = np.asarray(unthunk(dy), dtype=dtype)
    if dyv.ndim == 0:  # scalar dy boxed via similar(dx,(1,)) + fill!  (src/helpers.jl:16-20)
        dyv = np.full(L.size, dyv, dtype=dtype)
    np.add.at(dx, L.reshape(-1, order="F"), dyv.reshape(-1, order="F"))
    return np.asfortranarray(dx.reshape(x_shape, order="F"))


# --------------------------------------------------------------------------------------
# rrules (ChainRules / NNlib / Yota) restated.  Each returns (value, pullback); the
# pullback returns one tangent per (f, args...) position exactly like ChainRules.
# --------------------------------------------------------------------------------------

# scalar-rule table (SURVEY §8(a) "Scalar-rule table"): derivative given x and output Ω
_SCALAR_RULES = {
    "tanh": (np.tanh, lambda x, o: 1 - o * o),
    "exp": (np.exp, lambda x, o: o),
    "log": (np.log, lambda x, o: 1 / x),
    "sin": (np.sin, lambda x, o: np.cos(x)),
    "cos": (np.cos, lambda x, o: -np.sin(x)),
    "abs2": (lambda x: x * x, lambda x, o: 2 * x),
    "sqrt": (np.sqrt, lambda x, o: 1 / (2 * o)),
    "neg": (lambda x: -x, lambda x, o: -np.ones_like(x)),
    "relu": (lambda x: np.maximum(x, 0), lambda x, o: (x > 0).astype(x.dtype)),
    "sigmoid": (lambda x: 1 / (1 + np.exp(-x)), lambda x, o: o * (1 - o)),
}


def rrule_matmul(A, B):
    """ChainRules rrule(*, A::VecOrMat, B::VecOrMat): dA = @thunk(Ȳ*B'), dB = @thunk(A'*Ȳ)."""
    y = np.asfortranarray(A @ B)

    def pb(dy):
        dy = unthunk(dy)
        if B.ndim == 1 and A.ndim == 2:
            dA = Thunk(lambda: np.asfortranarray(np.outer(dy, B)))
            dB = Thunk(lambda: A.T @ dy)
        else:
            dA = Thunk(lambda: np.asfortranarray(dy @ B.T))
            dB = Thunk(lambda: np.asfortranarray(A.T @ dy))
        return (NoTangent, dA, dB)

    return y, pb


def rrule_bc_add(a, b):
    y = jl_broadcast(lambda p, q: p + q, a, b)

    def pb(d):
        d = unthunk(d)
        return (NoTangent, NoTangent, unbroadcast(a, d), unbroadcast(b, d))

    return y, pb


def rrule_bc_sub(a, b):
    y = jl_broadcast(lambda p, q: p - q, a, b)

    def pb(d):
        d = unthunk(d)
        db = unbroadcast(b, d)
        return (NoTangent, NoTangent, unbroadcast(a, d), db if iszerotangent(db) else -db)

    return y, pb


def rrule_bc_mul(a, b):
    y = jl_broadcast(lambda p, q: p * q, a, b)

    def pb(d):
        d = unthunk(d)
        da = Thunk(lambda: unbroadcast(a, jl_broadcast(lambda p, q: p * q, d, b)))
        db = Thunk(lambda: unbroadcast(b, jl_broadcast(lambda p, q: p * q, d, a)))
        return (NoTangent, NoTangent, da, db)

    return y, pb


def rrule_bc_div(a, b):
    y = jl_broadcast(lambda p, q: p / q, a, b)

    def pb(d):
        d = unthunk(d)
        da = Thunk(lambda: unbroadcast(a, jl_broadcast(lambda p, q: p / q, d, b)))
        db = Thunk(lambda: unbroadcast(b, jl_broadcast(lambda p, q, r: -(p * (q / r)), d, y, b)))
        return (NoTangent, NoTangent, da, db)

    return y, pb


def rrule_bc_square(x):
    """broadcasted(literal_pow, ^, x, Val(2)): pullback 2 .* Δ .* x."""
    y = x * x

    def pb(d):
        d = unthunk(d)
        two = x.dtype.type(2) if isinstance(x, np.ndarray) else 2
        return (NoTangent, NoTangent, NoTangent, jl_broadcast(lambda p, q: two * p * q, d, x), NoTangent)

    return y, pb


def rrule_bc_pow(x, p):
    """x .^ p with a float exponent constant: generic scalar rule p*x^(p-1)."""
    y = jl_broadcast(lambda q: q ** x.dtype.type(p), x)

    def pb(d):
        d = unthunk(d)
        t = x.dtype.type
        return (NoTangent, NoTangent, jl_broadcast(lambda dd, q: dd * (t(p) * q ** t(p - 1)), d, x), NoTangent)

    return y, pb


def rrule_bc_unary(name, x):
    f, df = _SCALAR_RULES[name]
    y = f(x)
    if isinstance(x, np.ndarray):
        y = y.astype(x.dtype, copy=False)

    def pb(d):
        d = unthunk(d)
        g = jl_broadcast(lambda dd, xx, oo: dd * df(xx, oo), d, x, y)
        if isinstance(x, np.ndarray):
            g = np.asarray(g, dtype=x.dtype)
        return (NoTangent, NoTangent, unbroadcast(x, g))

    return y, pb


def rrule_materialize(x):  # src/rulesets.jl:27-29
    return x, lambda dy: (NoTangent, dy)


def rrule_sum(x, dims=None):
    """ChainRules rrule(sum, x; dims): broadcast-back of ȳ."""
    if dims is None:
        y = x.dtype.type(x.sum())

        def pb(d):
            d = unthunk(d)
            return (NoTangent, Thunk(lambda: np.full(x.shape, d, dtype=x.dtype, order="F")))

        return y, pb
    ax = tuple(d - 1 for d in (dims if isinstance(dims, (tuple, list)) else (dims,)))
    y = np.asfortranarray(x.sum(axis=ax, keepdims=True))

    def pb(d):
        d = unthunk(d)
        return (NoTangent, Thunk(lambda: np.asfortranarray(np.broadcast_to(_pad(d, x.ndim), x.shape).astype(x.dtype))))

    return y, pb


def rrule_mean(x, dims=None):
    """Statistics.mean rrule: sum pullback scaled by 1/n."""
    if dims is None:
        n = x.size
    else:
        dd = dims if isinstance(dims, (tuple, list)) else (dims,)
        n = int(np.prod([x.shape[d - 1] for d in dd]))
    ys, pbs = rrule_sum(x, dims)
    y = ys / x.dtype.type(n)

    def pb(d):
        d = unthunk(d)
        return pbs(d / x.dtype.type(n))

    return y, pb


def rrule_getindex(x, *I):
    y = jl_getindex(x, I)

    def pb(d):
        dx = Thunk(lambda: jl_ungetindex(x.shape, x.dtype, d, I))
        return (NoTangent, dx) + (NoTangent,) * len(I)

    return y, pb


def rrule_logsoftmax(x, dims=1):
    """NNlib.logsoftmax + ∇logsoftmax_data: dy .- sum(dy; dims) .* exp.(y)."""
    ax = dims - 1
    m = x.max(axis=ax, keepdims=True)
    s = x - m
    y = np.asfortranarray(s - np.log(np.exp(s).sum(axis=ax, keepdims=True)))

    def pb(d):
        d = unthunk(d)
        return (NoTangent, np.asfortranarray(d - d.sum(axis=ax, keepdims=True) * np.exp(y)))

    return y, pb


def rrule_transpose(x):
    y = np.asfortranarray(x.T) if x.ndim == 2 else np.asfortranarray(x.reshape(1, -1))

    def pb(d):
        d = unthunk(d)
        return (NoTangent, np.asfortranarray(d.T) if x.ndim == 2 else np.asfortranarray(d.reshape(-1)))

    return y, pb


def rrule_reshape(x, dims):
    """ChainRules rrule(reshape, A, dims...): column-major reshape; the pullback reshapes back."""
    y = np.asfortranarray(x.reshape(tuple(int(d) for d in dims), order="F"))

    def pb(d):
        return (NoTangent, np.asfortranarray(unthunk(d).reshape(x.shape, order="F"))) + (NoTangent,) * len(dims)

    return y, pb


def rrule_cat(dim, *xs):
    nd = max(dim, max(x.ndim for x in xs))
    ps = [_pad(x, nd) for x in xs]
    y = np.asfortranarray(np.concatenate(ps, axis=dim - 1))

    def pb(d):
        d = _pad(unthunk(d), nd)
        outs, off = [], 0
        for x, p in zip(xs, ps):
            sl = [slice(None)] * nd
            sl[dim - 1] = slice(off, off + p.shape[dim - 1])
            off += p.shape[dim - 1]
            outs.append(np.asfortranarray(d[tuple(sl)].reshape(x.shape, order="F")))
        return (NoTangent,) + tuple(outs)

    return y, pb


def rrule_scalar(op, a, b=None):
    """Scalar rules used at the loss tail (-s, s/n, s*c, a+b, a-b)."""
    if op == "neg":
        return -a, lambda d: (NoTangent, -unthunk(d))
    if op == "add":
        return a + b, lambda d: (NoTangent, unthunk(d), unthunk(d))
    if op == "sub":
        return a - b, lambda d: (NoTangent, unthunk(d), -unthunk(d))
    if op == "mul":
        return a * b, lambda d: (NoTangent, unthunk(d) * b, unthunk(d) * a)
    if op == "div":
        y = a / b
        # @scalar_rule x / y (one(x) / y, -(Ω / y))
        return y, lambda d: (NoTangent, unthunk(d) * (1 / b), unthunk(d) * -(y / b))
    raise KeyError(op)


# --------------------------------------------------------------------------------------
# the tape (Umlaut.Tape restated at the granularity src/grad.jl needs)
# --------------------------------------------------------------------------------------


class _Op:
    __slots__ = ("kind", "val", "fn", "args", "kw", "pb", "is_rule")

    def __init__(self, kind, val, fn=None, args=(), kw=None):
        self.kind, self.val, self.fn, self.args, self.kw = kind, val, fn, args, kw or {}
        self.pb = None


class OV:
    """A traced oracle value (Umlaut.Variable bound to a tape op)."""

    __array_priority__ = 1000

    def __init__(self, tape, id_, lazy=False):
        self.tape, self.id, self.lazy = tape, id_, lazy

    @property
    def val(self):
        return self.tape.ops[self.id].val

    @property
    def shape(self):
        return np.shape(self.val)

    # ---- dispatch protocol shared with the product DSL (no import either way) ----
    def __jl__(self, name, *args, **kw):
        return self.tape.call(name, *args, **kw)

    def __matmul__(self, o):
        return self.tape.call("matmul", self, o)

    def __add__(self, o):
        return self.tape.call("add", self, o)

    def __radd__(self, o):
        return self.tape.call("add", o, self)

    def __sub__(self, o):
        return self.tape.call("sub", self, o)

    def __rsub__(self, o):
        return self.tape.call("sub", o, self)

    def __mul__(self, o):
        return self.tape.call("mul", self, o)

    def __rmul__(self, o):
        return self.tape.call("mul", o, self)

    def __truediv__(self, o):
        return self.tape.call("div", self, o)

    def __neg__(self):
        return self.tape.call("neg", self)

    def __pow__(self, p):
        return self.tape.call("pow", self, p)

    def __getitem__(self, I):
        return self.tape.call("getindex", self, *(I if isinstance(I, tuple) else (I,)))


_BCAST = {"add", "sub", "mul", "div", "pow"} | set(_SCALAR_RULES)


class OracleTape:
    """Forward trace: every primitive call is recorded AND executed (Umlaut `mkcall`,
    SURVEY App. C), through its rrule (chainrules_transform!, src/grad.jl:134-159)."""

    def __init__(self, dtype):
        self.ops = []
        self.dtype = dtype
        self.inputs = []

    def push(self, op):
        self.ops.append(op)
        return OV(self, len(self.ops) - 1)

    def input(self, val):
        v = self.push(_Op("input", val))
        self.inputs.append(v.id)
        return v

    def _arg(self, a):
        return a

    def _materialize(self, v):
        """tanh.(...) consumed by a non-broadcast call: Broadcast.materialize with Yota's
        identity rrule (src/rulesets.jl:27-29)."""
        if isinstance(v, OV) and v.lazy:
            y, pb = rrule_materialize(v.val)
            op = _Op("call", y, "materialize", (v,))
            op.pb = pb
            return self.push(op)
        return v

    def call(self, name, *args, **kw):
        vals = [a.val if isinstance(a, OV) else a for a in args]
        isarr = any(isinstance(v, np.ndarray) and v.ndim > 0 for v in vals)
        lazy = False
        if name in _BCAST and isarr:
            # dotted call: rrule(broadcasted, f, args...) -- lazy for + - * / ^2
            t = self.dtype
            vals = [t(v) if isinstance(v, (int, float)) and not isinstance(v, bool) and name != "pow" else v for v in vals]
            if name == "add":
                y, pb = rrule_bc_add(*vals)
            elif name == "sub":
                y, pb = rrule_bc_sub(*vals)
            elif name == "mul":
                y, pb = rrule_bc_mul(*vals)
            elif name == "div":
                y, pb = rrule_bc_div(*vals)
            elif name == "pow":
                p = vals[1]
                if isinstance(p, (int, np.integer)) and p == 2:
                    y, pb = rrule_bc_square(vals[0])
                else:
                    y, pb = rrule_bc_pow(vals[0], float(p))
            else:
                y, pb = rrule_bc_unary(name, vals[0])
            lazy = name in ("add", "sub", "mul", "div") or (name == "pow" and vals[1] == 2 and isinstance(vals[1], (int, np.integer)))
            op = _Op("call", y, "bc_" + name, tuple(args), kw)
        else:
            args = tuple(self._materialize(a) for a in args)
            vals = [a.val if isinstance(a, OV) else a for a in args]
            if name == "matmul":
                y, pb = rrule_matmul(*vals)
            elif name == "sum":
                y, pb = rrule_sum(vals[0], kw.get("dims"))
            elif name == "mean":
                y, pb = rrule_mean(vals[0], kw.get("dims"))
            elif name == "getindex":
                y, pb = rrule_getindex(vals[0], *vals[1:])
            elif name == "logsoftmax":
                y, pb = rrule_logsoftmax(vals[0], kw.get("dims", 1))
            elif name in ("transpose", "adjoint"):
                y, pb = rrule_transpose(vals[0])
            elif name == "reshape":
                y, pb = rrule_reshape(vals[0], vals[1:])
            elif name in ("vcat", "hcat", "cat"):
                dim = {"vcat": 1, "hcat": 2}.get(name) or kw["dims"]
                y, pb = rrule_cat(dim, *vals)
            elif name in ("neg", "add", "sub", "mul", "div"):
                y, pb = rrule_scalar(name, *vals)
            elif name == "pow":
                a, p = vals
                y = a ** p
                pb = (lambda a, p: lambda d: (NoTangent, unthunk(d) * p * a ** (p - 1), NoTangent))(a, p)
            elif name in _SCALAR_RULES:
                f, df = _SCALAR_RULES[name]
                a = vals[0]
                y = f(a)
                pb = (lambda a, y, df: lambda d: (NoTangent, unthunk(d) * df(a, y)))(a, y, df)
            elif name == "materialize":
                y, pb = rrule_materialize(vals[0])
            else:
                raise NotImplementedError(f"No derivative rule found for op {name}")  # src/grad.jl:179
            op = _Op("call", y, name, args, kw)
        op.pb = pb
        v = self.push(op)
        v.lazy = lazy
        return v


def _todo_list(tape, res_id):
    """src/grad.jl:98-121: DFS from the result over Call args; ids sorted descending."""
    seen = set()
    stack = [res_id]
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        seen.add(i)
        for a in tape.ops[i].args:
            if isinstance(a, OV) and a.id not in seen and tape.ops[a.id].kind == "call":
                stack.append(a.id)
    return sorted(seen, reverse=True)


def grad(f, *args, seed=1, dtype=np.float32):
    """Yota.grad(f, args...; seed) restated (src/grad.jl:349-362 over :242-289).

    Returns (val, (ZeroTangent, g_arg1, ...)) -- gradients for EVERY argument
    (finalize_grad!, src/grad.jl:230-239)."""
    dtype = np.dtype(dtype).type
    tape = OracleTape(dtype)
    ins = []
    for a in args:
        if isinstance(a, np.ndarray):
            if a.dtype.kind == "f":
                a = jl_array(a, dtype)
            else:
                a = jl_array(a)
        elif isinstance(a, float):
            a = dtype(a)
        ins.append(tape.input(a))
    res = f(*ins)
    if not isinstance(res, OV):
        # constant result: tuple of ZeroTangent for each argument (src/grad.jl:268-276)
        return res, (ZeroTangent,) + tuple(ZeroTangent for _ in ins)
    if tape.ops[res.id].kind == "input":
        # function returning its input: the seed is the derivative (src/grad.jl:277-286)
        return res.val, (ZeroTangent,) + tuple(seed if v.id == res.id else ZeroTangent for v in ins)
    res = tape._materialize(res)
    zval = res.val
    # ---- back! (src/grad.jl:198-227) ----
    if isinstance(seed, (int, float)) and seed == 1 and np.ndim(zval) > 0:
        raise ValueError("Gradient of a vector-valued function requires a seed")
    if isinstance(seed, str) and seed == "auto":
        seed = np.ones(np.shape(zval), dtype=dtype, order="F") if np.ndim(zval) > 0 else dtype(1)
    elif isinstance(seed, np.ndarray):
        seed = jl_array(seed, dtype)
    derivs = {res.id: seed}
    for y in _todo_list(tape, res.id):
        # step_back! (src/grad.jl:165-189)
        if y not in derivs:
            continue
        dy = derivs[y]
        if iszerotangent(dy):
            continue
        op = tape.ops[y]
        dxs = op.pb(dy)
        for i, x in enumerate(op.args):
            if isinstance(x, OV):
                dx = _tangent_for(op, dxs, i)
                # set_or_add_deriv! (src/grad.jl:83-95): broadcast(+, dx, old_dx)
                if x.id not in derivs:
                    derivs[x.id] = dx
                else:
                    old = derivs[x.id]
                    if iszerotangent(dx):
                        derivs[x.id] = old
                    elif iszerotangent(old):
                        derivs[x.id] = dx
                    else:
                        derivs[x.id] = jl_broadcast(lambda p, q: p + q, unthunk(dx), unthunk(old))
    grads = tuple(unthunk(derivs[i]) if i in derivs else ZeroTangent for i in tape.inputs)
    return zval, (ZeroTangent,) + grads


def _tangent_for(op, dxs, i):
    """Map argument position -> tangent position (src/grad.jl:176,182-186): ChainRules
    returns one tangent per (f, args...); bc rules carry an extra slot for `broadcasted`'s
    `f`, literal_pow one more for `^`."""
    n_extra = len(dxs) - len(op.args)
    if op.fn == "bc_pow" and len(dxs) == 5:  # (bcasted, literal_pow, ^, x, Val)
        return dxs[3] if i == 0 else NoTangent
    return dxs[n_extra + i]


def ngradient(f, *args, eps=None):
    """5-point central finite differences in Float64 (FiniteDifferences.central_fdm(5,1),
    src/gradcheck.jl:3-6) of the *primal* f evaluated through this module's rules."""
    outs = []
    base = [np.array(a, dtype=np.float64, order="F") if isinstance(a, np.ndarray) and a.dtype.kind == "f" else a for a in args]

    def fval(xs):
        v, _ = _primal(f, xs)
        return float(v)

    for n, a in enumerate(base):
        if not (isinstance(a, np.ndarray) and a.dtype.kind == "f"):
            outs.append(None)
            continue
        g = np.zeros_like(a)
        flat_g = g.reshape(-1, order="F")
        h = 1e-3 if eps is None else eps
        for k in range(a.size):
            def at(delta):
                xs = list(base)
                b = a.copy(order="F")
                bf = b.reshape(-1, order="F")
                bf[k] += delta
                xs[n] = bf.reshape(a.shape, order="F")
                return fval(xs)
            flat_g[k] = (at(-2 * h) - 8 * at(-h) + 8 * at(h) - at(2 * h)) / (12 * h)
        outs.append(flat_g.reshape(a.shape, order="F"))
    return outs


def _primal(f, xs):
    tape = OracleTape(np.float64)
    ins = [tape.input(x) for x in xs]
    r = f(*ins)
    return (r.val if isinstance(r, OV) else r), tape


def gradcheck(f, *args, atol=1e-5, rtol=1e-5):
    """src/gradcheck.jl:9-23 (Float64)."""
    _, g = grad(f, *args, dtype=np.float64)
    ng = ngradient(f, *args)
    ok = True
    for a, b in zip(g[1:], ng):
        if b is None or a is NoTangent:
            continue
        a = np.zeros_like(b) if a is ZeroTangent else a
        ok &= bool(np.allclose(a, b, rtol=rtol, atol=atol))
    return ok


def update(x, gx, fn=None):
    """update!(x::Array, gx, fn) : x .= fn(x, gx), default x .- gx (src/update.jl:42-49)."""
    x[...] = (x - gx) if fn is None else fn(x, gx)
    return x
