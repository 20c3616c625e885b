"""Linear tape in the form the reference's tape builder produces.

Restates only what the backend consumes of Umlaut's `Tape`:
`Input`, `Constant`, `Call(fn, args)` ops addressed by `Variable` ids, plus the tracer that
records primitive calls.  Values on the tape are *abstract* (`AVal`: dims + dtype): unlike
`Umlaut.mkcall` nothing is executed while tracing -- the cold call runs the compiled
program instead (same result as `tape.retval`, /root/reference/src/grad.jl:357-360).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


class _Marker:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name + "()"


ZeroTangent = _Marker("ZeroTangent")
NoTangent = _Marker("NoTangent")


def iszerotangent(x):  # src/utils.jl:21-23
    return x is ZeroTangent or x is NoTangent


class Colon:
    def __repr__(self):
        return ":"


colon = Colon()


class JlRange:
    """a:b, inclusive and 1-based."""

    def __init__(self, a, b):
        self.a, self.b = int(a), int(b)

    def __len__(self):
        return max(0, self.b - self.a + 1)

    def __repr__(self):
        return f"{self.a}:{self.b}"


@dataclass(frozen=True)
class AVal:
    """Abstract array value: Julia dims (column-major) and element type."""

    dims: tuple
    dtype: str = "f32"  # "f32" | "i64"
    lazy: bool = False  # result of a dotted call that has not gone through `materialize` yet

    @property
    def ndim(self):
        return len(self.dims)


@dataclass
class Op:
    id: int
    val: object = None
    line: str = ""


@dataclass
class Input(Op):
    pass


@dataclass
class Constant(Op):
    pass


@dataclass
class Call(Op):
    fn: object = None
    args: tuple = ()
    kw: dict = field(default_factory=dict)


class V:
    """Umlaut.Variable: a reference to a tape op by id."""

    __slots__ = ("id",)

    def __init__(self, id_):
        self.id = id_

    def __repr__(self):
        return f"%{self.id}"

    def __eq__(self, o):
        return isinstance(o, V) and o.id == self.id

    def __hash__(self):
        return hash(("V", self.id))


class GradCtx:
    """src/grad.jl:25-32"""

    def __init__(self):
        self.pullbacks = {}  # primal var id -> pullback var id
        self.derivs = {}  # primal var id -> derivative var id


class Tape:
    def __init__(self):
        self.ops = []
        self.result = None
        self.meta = {}
        self.c = GradCtx()
        self.captured = []  # host index vectors captured as constants: (Constant id, np.ndarray)

    def __len__(self):
        return len(self.ops)

    def __getitem__(self, v):
        return self.ops[v.id if isinstance(v, V) else v]

    def push(self, op):
        op.id = len(self.ops)
        self.ops.append(op)
        return V(op.id)

    def inputs(self):
        return [V(op.id) for op in self.ops if isinstance(op, Input)]

    def __repr__(self):
        lines = []
        for op in self.ops:
            if isinstance(op, Input):
                lines.append(f"  inp %{op.id}::{_fmt(op.val)}")
            elif isinstance(op, Constant):
                lines.append(f"  const %{op.id} = {_fmt(op.val)}")
            else:
                fn = op.fn if not callable(op.fn) else getattr(op.fn, "__name__", op.fn)
                args = ", ".join(map(_fmt, op.args))
                lines.append(f"  %{op.id} = {fn}({args})" + (f"  # {op.line}" if op.line else ""))
        return "Tape{GradCtx}\n" + "\n".join(lines)


def _fmt(x):
    if isinstance(x, AVal):
        return ("Broadcasted" if x.lazy else "Array") + f"{list(x.dims)}" if x.dims else "Float32"
    if isinstance(x, np.ndarray):
        return f"<{x.dtype}{list(x.shape)}>"
    return repr(x)


# ---------------------------------------------------------------------------------------
# tracer
# ---------------------------------------------------------------------------------------

_BCAST = {"add", "sub", "mul", "div", "pow", "neg", "tanh", "exp", "log", "sin", "cos", "abs2", "sqrt",
          "relu", "sigmoid"}


class TVar:
    """A traced value: what the user's function sees in place of a device array."""

    __array_priority__ = 1000

    def __init__(self, tracer, v):
        self._t, self.v = tracer, v

    @property
    def aval(self):
        return self._t.tape[self.v].val

    @property
    def shape(self):
        return self.aval.dims

    size = shape

    def __jl__(self, name, *args, **kw):
        return self._t.call(name, *args, **kw)

    def __matmul__(self, o):
        return self._t.call("matmul", self, o)

    def __add__(self, o):
        return self._t.call("add", self, o)

    def __radd__(self, o):
        return self._t.call("add", o, self)

    def __sub__(self, o):
        return self._t.call("sub", self, o)

    def __rsub__(self, o):
        return self._t.call("sub", o, self)

    def __mul__(self, o):
        return self._t.call("mul", self, o)

    def __rmul__(self, o):
        return self._t.call("mul", o, self)

    def __truediv__(self, o):
        return self._t.call("div", self, o)

    def __neg__(self):
        return self._t.call("neg", self)

    def __pow__(self, p):
        return self._t.call("pow", self, p)

    def __getitem__(self, I):
        return self._t.call("getindex", self, *(I if isinstance(I, tuple) else (I,)))


class Tracer:
    """Records primitive calls (the `isprimitive(::GradCtx, ...)` set restricted to the ops
    with a backend lowering, src/grad.jl:61-69).  Dotted calls become
    `broadcasted(f, args...)`; a lazy Broadcasted consumed by a non-broadcast primitive is
    forced through `materialize` first, as Julia's dot-fusion lowering does."""

    def __init__(self):
        from . import rules

        self.tape = Tape()
        self.rules = rules

    def input(self, aval):
        return TVar(self, self.tape.push(Input(0, aval)))

    def _unwrap(self, a):
        if isinstance(a, TVar):
            return a.v
        if isinstance(a, (list, tuple)) and a and all(isinstance(i, (int, np.integer)) for i in a):
            a = np.asarray(a, dtype=np.int64)
        if isinstance(a, np.ndarray):
            # host index vector: captured as a tape Constant, uploaded once by the program
            v = self.tape.push(Constant(0, AVal(tuple(a.shape), "i64" if a.dtype.kind in "iu" else "f32")))
            self.tape.captured.append((v.id, np.ascontiguousarray(a)))
            return v
        return a

    def _aval(self, a):
        return self.tape[a].val if isinstance(a, V) else a

    def _materialize(self, a):
        if isinstance(a, V) and isinstance(self.tape[a].val, AVal) and self.tape[a].val.lazy:
            av = self.tape[a].val
            return self.tape.push(Call(0, AVal(av.dims, av.dtype), fn="materialize", args=(a,)))
        return a

    def call(self, name, *args, **kw):
        args = tuple(self._unwrap(a) for a in args)
        avals = [self._aval(a) for a in args]
        isarr = any(isinstance(a, AVal) and a.ndim > 0 for a in avals)
        if name in _BCAST and isarr:
            if name == "pow":
                p = avals[1]
                if isinstance(p, (int, np.integer)) and not isinstance(p, bool) and p == 2:
                    # x .^ 2  ->  broadcasted(literal_pow, ^, x, Val(2))   
                    fn, fargs, lazy = "broadcasted", ("literal_pow", "^", args[0], "Val(2)"), True
                else:
                    fn, fargs, lazy = "broadcasted", ("^", args[0], float(p)), True
            else:
                fn, fargs, lazy = "broadcasted", (name,) + args, True
            val = self.rules.infer(fn, [self._aval(a) for a in fargs], kw)
            val = AVal(val.dims, val.dtype, lazy)
            return TVar(self, self.tape.push(Call(0, val, fn=fn, args=fargs, kw=kw)))
        args = tuple(self._materialize(a) for a in args)
        avals = [self._aval(a) for a in args]
        val = self.rules.infer(name, avals, kw)
        return TVar(self, self.tape.push(Call(0, val, fn=name, args=args, kw=kw)))
