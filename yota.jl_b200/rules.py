"""Rule table: for every primitive the backend can lower, (a) result shape inference for
tracing, (b) the abstract tangent tuple of its pullback, (c) the op-program lowering of
`rrule(cfg, f, args...)` and of `pb(dy)`.

The math restates the third-party rules the reference dispatches to at
/root/reference/src/grad.jl:147 (ChainRules `*`, `broadcasted`, `sum`, `mean`, `getindex`;
NNlib `logsoftmax`) and Yota's own `materialize` rule (src/rulesets.jl:27-29); SURVEY.md
§8(a) rows 7-12 and Appendix A list them.  Tangent tuples have one entry per
`(f, args...)` position exactly like ChainRules, because the tape addresses them with
`getfield(dxs, i)` (src/grad.jl:176-186).
"""
from __future__ import annotations

import numpy as np

from . import ir
from .tape import AVal, Colon, JlRange, NoTangent, V, ZeroTangent, iszerotangent

UNARY = {
    "tanh": ir.EW_TANH, "exp": ir.EW_EXP, "log": ir.EW_LOG, "sin": ir.EW_SIN, "cos": ir.EW_COS,
    "abs2": ir.EW_SQUARE, "sqrt": ir.EW_SQRT, "relu": ir.EW_RELU, "sigmoid": ir.EW_SIGMOID, "neg": ir.EW_NEG,
}
BINARY = {"add": ir.EW_ADD, "sub": ir.EW_SUB, "mul": ir.EW_MUL, "div": ir.EW_DIV}
SCALAR_FNS = set(UNARY) | set(BINARY) | {"pow"}


class NoRuleError(Exception):
    pass


def _dims(a):
    return a.dims if isinstance(a, AVal) else ()


def _isnum(a):
    return isinstance(a, (int, float, np.number)) and not isinstance(a, (bool, ir.TId))


# ---------------------------------------------------------------------------------------
# (a) shape inference
# ---------------------------------------------------------------------------------------


def _index_out_dims(xdims, I):
    if len(I) == 1 and len(xdims) != 1:
        xdims = (int(np.prod(xdims)),)
    if len(I) != len(xdims):
        raise IndexError(f"BoundsError: {len(I)} indices for an array of dims {xdims}")
    out = []
    for n, ix in zip(xdims, I):
        if isinstance(ix, Colon):
            out.append(n)
        elif isinstance(ix, JlRange):
            if ix.a < 1 or ix.b > n:
                raise IndexError("BoundsError")
            out.append(len(ix))
        elif _isnum(ix):
            if not 1 <= int(ix) <= n:
                raise IndexError("BoundsError")
        elif isinstance(ix, AVal):
            if ix.dtype != "i64" or ix.ndim != 1:
                raise TypeError("index vectors must be 1-d Int64")
            out.append(ix.dims[0])
        else:
            raise TypeError(f"unsupported index {ix!r}")
    return tuple(out)


def infer(fn, avals, kw):
    if fn == "broadcasted":
        op = avals[0]
        if op == "literal_pow":
            return AVal(_dims(avals[2]))
        if op == "^":
            return AVal(_dims(avals[1]))
        if op in UNARY:
            return AVal(_dims(avals[1]))
        if op in BINARY:
            return AVal(ir.bcast_dims(_dims(avals[1]), _dims(avals[2])))
        raise NoRuleError(f"No derivative rule found for op broadcasted({op}, ...)")
    if fn == "materialize":
        return AVal(_dims(avals[0]))
    if fn == "matmul":
        a, b = _dims(avals[0]), _dims(avals[1])
        if len(a) != 2 or len(b) not in (1, 2) or a[1] != b[0]:
            raise ValueError(f"DimensionMismatch: matrix A has dimensions {a}, B has dimensions {b}")
        return AVal((a[0],) + tuple(b[1:]))
    if fn in ("sum", "mean"):
        d = _dims(avals[0])
        dims = kw.get("dims")
        if dims is None:
            return AVal(())
        dd = (dims,) if _isnum(dims) else tuple(dims)
        return AVal(tuple(1 if (i + 1) in dd else n for i, n in enumerate(d)))
    if fn == "getindex":
        return AVal(_index_out_dims(_dims(avals[0]), avals[1:]))
    if fn == "logsoftmax":
        if kw.get("dims", 1) != 1:
            raise NoRuleError("logsoftmax: only dims=1 has a backend lowering")
        return AVal(_dims(avals[0]))
    if fn in ("transpose", "adjoint"):
        d = _dims(avals[0])
        return AVal((1, d[0]) if len(d) == 1 else (d[1], d[0]))
    if fn == "reshape":
        d = tuple(int(n) for n in avals[1:])
        if int(np.prod(d)) != int(np.prod(_dims(avals[0]))):
            raise ValueError(f"DimensionMismatch: new dimensions {d} must be consistent with array size {_dims(avals[0])}")
        return AVal(d, avals[0].dtype)
    if fn in ("vcat", "hcat", "cat"):
        dim = {"vcat": 1, "hcat": 2}.get(fn) or int(kw["dims"])
        ds = [_dims(a) for a in avals]
        nd = max(dim, max(len(d) for d in ds))
        ds = [d + (1,) * (nd - len(d)) for d in ds]
        for d in ds:
            if any(d[i] != ds[0][i] for i in range(nd) if i != dim - 1):
                raise ValueError("DimensionMismatch in cat")
        out = list(ds[0])
        out[dim - 1] = sum(d[dim - 1] for d in ds)
        return AVal(tuple(out))
    if fn in SCALAR_FNS:  # scalar rules at the loss tail
        if any(_dims(a) != () for a in avals if isinstance(a, AVal)):
            raise NoRuleError(f"{fn}: expected scalars")
        return AVal(())
    raise NoRuleError(f"No derivative rule found for op {fn}")  # src/grad.jl:179-180


def fargs_of(call_args):
    """`y_fargs = tape[rr].args[2:end]` (src/grad.jl:176): the (f, args...) of an rrule call
    whose args are (cfg, f, args...)."""
    return call_args[1:]


def pb_infer(fn, avals, kw, dy):
    """Abstract tangents returned by the pullback: one per (f, args...) position."""

    def like(a):
        if isinstance(a, AVal):
            return NoTangent if a.dtype == "i64" else AVal(a.dims)
        return NoTangent if not _isnum(a) else AVal(())

    if fn == "broadcasted":
        op = avals[0]
        if op == "literal_pow":
            return (NoTangent, NoTangent, NoTangent, like(avals[2]), NoTangent)
        if op == "^":
            return (NoTangent, NoTangent, like(avals[1]), NoTangent)
        return (NoTangent, NoTangent) + tuple(like(a) for a in avals[1:])
    if fn == "getindex":
        return (NoTangent, like(avals[0])) + (NoTangent,) * (len(avals) - 1)
    if fn == "pow":
        return (NoTangent, like(avals[0]), NoTangent)
    return (NoTangent,) + tuple(like(a) for a in avals)


# ---------------------------------------------------------------------------------------
# (c) lowering
# ---------------------------------------------------------------------------------------


def _t(b, v):
    """IR tensor id for a lowered value (numbers become 0-dim constants)."""
    return b.const(float(v)) if _isnum(v) else v


def unbroadcast(b, xdims, d):
    """ChainRules `unbroadcast(x, Δ)` (spec: src/helpers.jl:63-74)."""
    if iszerotangent(d):
        return d
    dd = b.dims(d)
    xdims = tuple(xdims)
    if xdims == dd:
        return d
    nx = int(np.prod(xdims)) if xdims else 1
    if nx == (int(np.prod(dd)) if dd else 1):
        return b.op(ir.OP_RESHAPE, (d,), xdims)
    mask = 0
    for i in range(len(dd)):
        if i >= len(xdims) or xdims[i] == 1:
            mask |= 1 << i
    return b.op(ir.OP_SUM, (d,), xdims, iattr=(mask,), fattr=(1.0,))


def _index_spec(b, I_vals):
    iattr, vec_ins = [len(I_vals)], []
    for ix in I_vals:
        if isinstance(ix, Colon):
            iattr += [ir.IDX_COLON, 0, 0]
        elif isinstance(ix, JlRange):
            iattr += [ir.IDX_RANGE, ix.a, ix.b]
        elif _isnum(ix):
            iattr += [ir.IDX_INT, int(ix), int(ix)]
        else:  # IR tensor id of an Int64 vector
            iattr += [ir.IDX_VEC, len(vec_ins), 0]
            vec_ins.append(ix)
    return iattr, vec_ins


def lower_fwd(b, fn, vals, kw, out):
    """Emit the forward of `rrule(cfg, fn, vals...)`; returns (value id, saved-for-pullback)."""
    od = out.dims
    if fn == "broadcasted":
        op = vals[0]
        if op == "literal_pow":
            x = _t(b, vals[2])
            return b.op(ir.EW_SQUARE, (x,), od), ("sq", x)
        if op == "^":
            x, p = _t(b, vals[1]), float(vals[2])
            return b.op(ir.EW_POWC, (x,), od, fattr=(p,)), ("powc", x, p)
        if op in UNARY:
            x = _t(b, vals[1])
            y = b.op(UNARY[op], (x,), od)
            return y, ("un", op, x, y)
        xa, xb = _t(b, vals[1]), _t(b, vals[2])
        y = b.op(BINARY[op], (xa, xb), od)
        return y, ("bin", op, xa, xb, y)
    if fn == "materialize":
        return vals[0], ("id",)
    if fn == "matmul":
        A, B = vals
        return b.op(ir.OP_MATMUL, (A, B), od, iattr=(0, 0)), ("mm", A, B)
    if fn in ("sum", "mean"):
        x = vals[0]
        xd = b.dims(x)
        dims = kw.get("dims")
        dd = tuple(range(1, len(xd) + 1)) if dims is None else ((dims,) if _isnum(dims) else tuple(dims))
        mask = sum(1 << (d - 1) for d in dd)
        n = int(np.prod([xd[d - 1] for d in dd])) if dd else 1
        y = b.op(ir.OP_SUM, (x,), od, iattr=(mask,), fattr=(1.0,))
        if fn == "mean":  # mean = sum / n  (Statistics; rounding kept as a separate divide)
            y = b.op(ir.EW_DIV, (y, b.const(n)), od)
        return y, ("sum", xd, n if fn == "mean" else None)
    if fn == "getindex":
        x = vals[0]
        iattr, vec = _index_spec(b, vals[1:])
        y = b.op(ir.OP_GETINDEX, (x, *vec), od, iattr=iattr)
        return y, ("idx", b.dims(x), iattr, vec)
    if fn == "logsoftmax":
        y = b.op(ir.OP_LOGSOFTMAX, (vals[0],), od)
        return y, ("lsm", y)
    if fn in ("transpose", "adjoint"):
        return b.op(ir.OP_TRANSPOSE, (vals[0],), od), ("tr", b.dims(vals[0]))
    if fn == "reshape":
        return b.op(ir.OP_RESHAPE, (vals[0],), od), ("rs", b.dims(vals[0]), len(vals) - 1)
    if fn in ("vcat", "hcat", "cat"):
        dim = {"vcat": 1, "hcat": 2}.get(fn) or int(kw["dims"])
        return b.op(ir.OP_CAT, tuple(vals), od, iattr=(dim,)), ("cat", dim, [b.dims(v) for v in vals])
    if fn in UNARY:  # scalar rule
        x = _t(b, vals[0])
        y = b.op(UNARY[fn], (x,), ())
        return y, ("un", fn, x, y)
    if fn in BINARY:
        xa, xb = _t(b, vals[0]), _t(b, vals[1])
        y = b.op(BINARY[fn], (xa, xb), ())
        return y, ("sbin", fn, xa, xb, y, _isnum(vals[0]), _isnum(vals[1]))
    if fn == "pow":
        x, p = _t(b, vals[0]), float(vals[1])
        return b.op(ir.EW_POWC, (x,), (), fattr=(p,)), ("powc", x, p)
    raise NoRuleError(f"No derivative rule found for op {fn}")


def _mul(b, x, y):
    return b.op(ir.EW_MUL, (x, y), ir.bcast_dims(b.dims(x), b.dims(y)))


def _dunary(b, op, d, x, y):
    """Δ .* ∂f per the ChainRules @scalar_rule table (SURVEY §8(a) "Scalar-rule table")."""
    xd = b.dims(x)
    if op == "tanh":  # 1 - Ω^2
        g = b.op(ir.EW_RSUBC, (b.op(ir.EW_SQUARE, (y,), xd),), xd, fattr=(1.0,))
    elif op == "exp":  # Ω
        g = y
    elif op == "log":  # inv(x)
        g = b.op(ir.EW_RDIVC, (x,), xd, fattr=(1.0,))
    elif op == "sin":  # cos(x)
        g = b.op(ir.EW_COS, (x,), xd)
    elif op == "cos":  # -sin(x)
        g = b.op(ir.EW_NEG, (b.op(ir.EW_SIN, (x,), xd),), xd)
    elif op == "abs2":  # 2x
        g = b.op(ir.EW_MULC, (x,), xd, fattr=(2.0,))
    elif op == "sqrt":  # inv(2Ω)
        g = b.op(ir.EW_RDIVC, (b.op(ir.EW_MULC, (y,), xd, fattr=(2.0,)),), xd, fattr=(1.0,))
    elif op == "neg":
        return b.op(ir.EW_NEG, (d,), b.dims(d))
    elif op == "relu":  # x > 0
        g = b.op(ir.EW_GT0, (x,), xd)
    elif op == "sigmoid":  # Ω(1-Ω)
        g = _mul(b, y, b.op(ir.EW_RSUBC, (y,), xd, fattr=(1.0,)))
    else:
        raise NoRuleError(op)
    return _mul(b, d, g)


def lower_pb(b, fn, saved, d):
    """Emit `pb(dy)`; returns the tangent tuple (IR ids / NoTangent / ZeroTangent)."""
    N = NoTangent
    k = saved[0]
    if k == "id":
        return (N, d)
    if k == "mm":
        _, A, B = saved
        dA = b.op(ir.OP_MATMUL, (d, B), b.dims(A), iattr=(0, 1))  # Ȳ * B'
        dB = b.op(ir.OP_MATMUL, (A, d), b.dims(B), iattr=(1, 0))  # A' * Ȳ
        return (N, dA, dB)
    if k == "bin":
        _, op, xa, xb, y = saved
        da, db_ = b.dims(xa), b.dims(xb)
        if op == "add":
            return (N, N, unbroadcast(b, da, d), unbroadcast(b, db_, d))
        if op == "sub":
            ub = unbroadcast(b, db_, d)
            return (N, N, unbroadcast(b, da, d), b.op(ir.EW_NEG, (ub,), b.dims(ub)))
        if op == "mul":
            return (N, N, unbroadcast(b, da, _mul(b, d, xb)), unbroadcast(b, db_, _mul(b, d, xa)))
        if op == "div":
            ga = b.op(ir.EW_DIV, (d, xb), b.dims(d))
            q = b.op(ir.EW_DIV, (y, xb), b.dims(y))
            gb = b.op(ir.EW_NEG, (_mul(b, d, q),), b.dims(d))
            return (N, N, unbroadcast(b, da, ga), unbroadcast(b, db_, gb))
    if k == "sq":  # 2 .* Δ .* x
        x = saved[1]
        g = _mul(b, b.op(ir.EW_MULC, (d,), b.dims(d), fattr=(2.0,)), x)
        return (N, N, N, g, N)
    if k == "powc":  # Δ * (p * x^(p-1))
        _, x, p = saved
        xd = b.dims(x)
        g = b.op(ir.EW_MULC, (b.op(ir.EW_POWC, (x,), xd, fattr=(p - 1.0,)),), xd, fattr=(p,))
        g = _mul(b, d, g)
        return (N, N, g, N) if fn == "broadcasted" else (N, g, N)
    if k == "un":
        _, op, x, y = saved
        g = unbroadcast(b, b.dims(x), _dunary(b, op, d, x, y))
        return (N, N, g) if fn == "broadcasted" else (N, g)
    if k == "sum":  # broadcast-back of ȳ (÷ n for mean)
        _, xd, n = saved
        if n is not None:
            d = b.op(ir.EW_DIV, (d, b.const(n)), b.dims(d))
        return (N, b.op(ir.EW_COPY, (d,), xd))
    if k == "idx":
        _, xd, iattr, vec = saved
        dx = b.op(ir.OP_UNGETINDEX, (d, *vec), xd, iattr=iattr)
        return (N, dx) + (N,) * iattr[0]
    if k == "lsm":  # dy .- sum(dy; dims=1) .* exp.(y)
        y = saved[1]
        yd = b.dims(y)
        s = b.op(ir.OP_SUM, (d,), (1,) + tuple(yd[1:]), iattr=(1,), fattr=(1.0,))
        e = b.op(ir.EW_EXP, (y,), yd)
        return (N, b.op(ir.EW_SUB, (d, _mul(b, s, e)), yd))
    if k == "tr":
        return (N, b.op(ir.OP_TRANSPOSE, (d,), saved[1]))
    if k == "rs":
        return (N, b.op(ir.OP_RESHAPE, (d,), saved[1])) + (N,) * saved[2]
    if k == "cat":
        _, dim, dlist = saved
        outs, off = [], 1
        for dd in dlist:
            n = dd[dim - 1] if dim <= len(dd) else 1
            outs.append(b.op(ir.OP_SLICE, (d,), dd, iattr=(dim, off, off + n - 1)))
            off += n
        return (N,) + tuple(outs)
    if k == "sbin":
        _, op, xa, xb, y, ca, cb = saved
        if op == "add":
            return (N, d, d)
        if op == "sub":
            return (N, d, b.op(ir.EW_NEG, (d,), ()))
        if op == "mul":
            return (N, _mul(b, d, xb), _mul(b, d, xa))
        if op == "div":  # (one(x)/y, -(Ω/y))
            ga = _mul(b, d, b.op(ir.EW_RDIVC, (xb,), (), fattr=(1.0,)))
            gb = _mul(b, d, b.op(ir.EW_NEG, (b.op(ir.EW_DIV, (y, xb), ()),), ()))
            return (N, ga, gb)
    raise NoRuleError(f"no pullback lowering for {fn}/{k}")
