"""`gradtape` / `grad` for device arrays: the reference's tape builder, then lowering of the
finished `Tape{GradCtx}` to an op program for libyota_cuda.so.

The tape-building half restates /root/reference/src/grad.jl so that the program fed to
the backend has exactly the reference's op forms and sweep order:
  chainrules_transform!  src/grad.jl:134-159   rrule(cfg, f, args...) + two _getfield
  todo_list              src/grad.jl:98-121    DFS from the result, ids descending
  step_back!             src/grad.jl:165-189   dxs = pb(dy); dx_i = getfield(dxs, i)
  set_or_add_deriv!      src/grad.jl:83-95     broadcast(+, dx, old_dx)
  back!                  src/grad.jl:198-227   seed rules
  finalize_grad!         src/grad.jl:230-239   tuple(derivs...) -> map(unthunk, .) -> tuple(val, .)
In a Julia deployment those functions stay Yota's own; only `lower` + `grad_compile`
(the replacement for src/grad.jl:293-307) are the plugin.
"""
from __future__ import annotations

import numpy as np

from . import ir, rules
from .tape import AVal, Call, Constant, Input, Tape, Tracer, TVar, V, NoTangent, ZeroTangent, iszerotangent

YOTA_RULE_CONFIG = "YotaRuleConfig()"


# ---------------------------------------------------------------------------------------
# tape construction (reference algorithm)
# ---------------------------------------------------------------------------------------


def trace(f, avals):
    tr = Tracer()
    ins = [tr.input(a) if isinstance(a, AVal) else a for a in avals]
    res = f(*ins)
    tape = tr.tape
    if isinstance(res, TVar):
        res = TVar(tr, tr._materialize(res.v))
        tape.result = res.v
    else:
        tape.result = tape.push(Constant(0, res))
    return tape


def chainrules_transform(tape):
    """Replace every primitive call `fn(args...)` by
        t = rrule(cfg, fn, args...); val = _getfield(t, 1); pb = _getfield(t, 2)
    and rebind later uses to `val` (src/grad.jl:134-159, replace!(...; rebind_to=2))."""
    new = Tape()
    new.meta, new.captured = tape.meta, []
    remap = {}

    def rb(a):
        return V(remap[a.id]) if isinstance(a, V) else a

    cap = dict(tape.captured)
    for op in tape.ops:
        if isinstance(op, Call):
            args = tuple(rb(a) for a in op.args)
            rr = new.push(Call(0, ("rrule", op.val), fn="rrule", args=(YOTA_RULE_CONFIG, op.fn) + args, kw=op.kw,
                               line=op.line))
            val = new.push(Call(0, op.val, fn="_getfield", args=(rr, 1)))
            pb = new.push(Call(0, ("pullback", rr.id), fn="_getfield", args=(rr, 2)))
            new.c.pullbacks[val.id] = pb.id
            remap[op.id] = val.id
        else:
            v = new.push(type(op)(0, op.val, op.line))
            remap[op.id] = v.id
            if op.id in cap:
                new.captured.append((v.id, cap[op.id]))
    new.result = rb(tape.result)
    return new


def todo_list(tape):
    result = set()

    def visit(y_id):
        result.add(y_id)
        for x in tape[y_id].args:
            if isinstance(x, V) and x.id not in result and isinstance(tape[x], Call):
                visit(x.id)

    assert isinstance(tape[tape.result], Call), "The tape's result is expected to be a Call"
    import sys

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10 * len(tape) + 1000))
    visit(tape.result.id)
    return [V(i) for i in sorted(result, reverse=True)]


def set_or_add_deriv(tape, x, dx):
    d = tape.c.derivs
    if x.id not in d:
        d[x.id] = dx.id
    else:
        old = V(d[x.id])
        a, b_ = tape[dx].val, tape[old].val
        val = b_ if iszerotangent(a) else a
        new_dx = tape.push(Call(0, val, fn="broadcast", args=("+", dx, old), line=f"updated deriv for {x}"))
        d[x.id] = new_dx.id


def step_back(tape, y):
    if y.id not in tape.c.derivs:
        return
    dy = V(tape.c.derivs[y.id])
    if iszerotangent(tape[dy].val):
        return
    if y.id not in tape.c.pullbacks:
        raise rules.NoRuleError(f"No derivative rule found for op {tape[y]}")
    pb = V(tape.c.pullbacks[y.id])
    rr = tape[y].args[0]
    rr_op = tape[rr]
    y_fargs = rules.fargs_of(rr_op.args)
    fn, fargs = y_fargs[0], y_fargs[1:]
    avals = [tape[a].val if isinstance(a, V) else a for a in fargs]
    tangents = rules.pb_infer(fn, avals, rr_op.kw, tape[dy].val)
    dxs = tape.push(Call(0, tangents, fn=pb, args=(dy,), line=f"pullback for {y}"))
    for i, x in enumerate(y_fargs, start=1):
        if isinstance(x, V):
            dx = tape.push(Call(0, tangents[i - 1], fn="getfield", args=(dxs, i), line=f"d{y}/d{x}"))
            set_or_add_deriv(tape, x, dx)


def back(tape, seed=1):
    z = tape.result
    zval = tape[z].val
    if isinstance(seed, (int, float)) and seed == 1 and zval.ndim > 0:
        raise ValueError("Gradient of a vector-valued function requires a seed")
    if isinstance(seed, str) and seed == "auto":
        seed_val = ("ones", zval) if zval.ndim > 0 else 1
    elif isinstance(seed, (int, float)):
        seed_val = seed
    else:  # array seed: becomes a run-time input of the compiled function (the `seed` kwarg)
        seed_val = ("array", AVal(tuple(seed.shape)))
        if tuple(seed.shape) != zval.dims:
            raise ValueError("seed must have the dims of the result")
    dy = tape.push(Constant(0, seed_val, line="seed"))
    tape.meta["seed"] = dy
    tape.c.derivs[z.id] = dy.id
    for y in todo_list(tape):
        step_back(tape, y)


def finalize_grad(tape):
    derivs = [V(tape.c.derivs[v.id]) if v.id in tape.c.derivs else ZeroTangent for v in tape.inputs()]
    t = tape.push(Call(0, None, fn="tuple", args=tuple(derivs)))
    u = tape.push(Call(0, None, fn="map", args=("unthunk", t)))
    tape.result = tape.push(Call(0, None, fn="tuple", args=(tape.result, u)))


def gradtape_(tape, seed=1):
    tape = chainrules_transform(tape)
    back(tape, seed=seed)
    finalize_grad(tape)
    return tape


def gradtape(f, *avals, seed=1):
    """gradtape(f, args...; seed=1) over abstract arguments (src/grad.jl:260-289)."""
    tape = trace(f, avals)
    res = tape[tape.result]
    if isinstance(res, Call):
        return gradtape_(tape, seed=seed)
    if isinstance(res, Constant):  # function returning a constant (src/grad.jl:268-276)
        tape.meta["seed"] = tape.push(Constant(0, seed, line="seed"))
        t = tape.push(Constant(0, tuple(ZeroTangent for _ in tape.inputs())))
        tape.result = tape.push(Call(0, None, fn="tuple", args=(tape.result, t)))
        return tape
    # function returning its input: the seed is its derivative (src/grad.jl:277-286)
    s = tape.push(Constant(0, seed, line="seed"))
    tape.meta["seed"] = s
    vals = tuple(s if v.id == tape.result.id else ZeroTangent for v in tape.inputs())
    t = tape.push(Call(0, None, fn="tuple", args=vals))
    tape.result = tape.push(Call(0, None, fn="tuple", args=(tape.result, t)))
    return tape


# ---------------------------------------------------------------------------------------
# lowering: Tape{GradCtx} -> op program (the plugin half)
# ---------------------------------------------------------------------------------------


class Lowered:
    """Result of `lower`: the op program plus how to bind it."""

    def __init__(self):
        self.prog = ir.IRProgram()
        self.arg_slots = []  # per tape Input: in slot or None (non-array argument)
        self.captured = []  # (in slot, host ndarray) constants uploaded once
        self.seed_slot = None  # in slot of an array seed
        self.val = None  # ("out", slot, dims) | ("host", python value) | ("arg", input position)
        self.grads = []  # per tape Input: ("out", slot, dims) | ZeroTangent | NoTangent | ("seed",) | ("host", v)


def lower(tape):
    """Walk the finished gradient tape and emit the op program (what `b200_compile(tape)` does
    in the Julia glue).  Every tape statement form of SURVEY §3.2 is handled; anything else
    raises NoRuleError -- there is no fallback."""
    L = Lowered()
    b = L.prog
    env = {}
    cap = dict(tape.captured)
    n_in = 0
    seed_id = tape.meta["seed"].id
    for op in tape.ops:
        i = op.id
        if isinstance(op, Input):
            if isinstance(op.val, AVal):
                env[i] = b.tensor(op.val.dims, ir.I64 if op.val.dtype == "i64" else ir.F32, ir.T_INPUT, n_in)
                L.arg_slots.append(n_in)
                n_in += 1
            else:
                env[i] = op.val
                L.arg_slots.append(None)
        elif isinstance(op, Constant):
            if i == seed_id:
                s = op.val
                if isinstance(s, tuple) and s[0] == "ones":
                    env[i] = b.op(ir.EW_COPY, (b.tensor((), ir.F32, ir.T_SEED),), s[1].dims)
                elif isinstance(s, tuple) and s[0] == "array":
                    env[i] = b.tensor(s[1].dims, ir.F32, ir.T_INPUT, n_in)
                    L.seed_slot = n_in
                    n_in += 1
                else:
                    env[i] = b.tensor((), ir.F32, ir.T_SEED)
            elif i in cap:
                env[i] = b.tensor(op.val.dims, ir.I64 if op.val.dtype == "i64" else ir.F32, ir.T_INPUT, n_in)
                L.captured.append((n_in, cap[i]))
                n_in += 1
            else:
                env[i] = op.val
        else:
            fn = op.fn
            if fn == "rrule":
                f, fargs = op.args[1], op.args[2:]
                vals = [env[a.id] if isinstance(a, V) else a for a in fargs]
                env[i] = ("rr", f) + rules.lower_fwd(b, f, vals, op.kw, op.val[1])
            elif fn == "_getfield":
                rr = env[op.args[0].id]
                env[i] = rr[2] if op.args[1] == 1 else ("pb", rr[1], rr[3])
            elif isinstance(fn, V):  # pb(dy)
                _, f, saved = env[fn.id]
                d = env[op.args[0].id]
                env[i] = rules.lower_pb(b, f, saved, rules._t(b, d) if not iszerotangent(d) else d)
            elif fn == "getfield":
                env[i] = env[op.args[0].id][op.args[1] - 1]
            elif fn == "broadcast":  # tape gradient accumulation (src/grad.jl:91)
                x, y = env[op.args[1].id], env[op.args[2].id]
                if iszerotangent(x):
                    env[i] = y
                elif iszerotangent(y):
                    env[i] = x
                else:
                    env[i] = b.op(ir.EW_ADD, (x, y), ir.bcast_dims(b.dims(x), b.dims(y)))
            elif fn == "tuple":
                env[i] = tuple(env[a.id] if isinstance(a, V) else a for a in op.args)
            elif fn == "map":  # map(unthunk, t): thunks are already plain dataflow here
                env[i] = env[op.args[1].id]
            else:
                raise rules.NoRuleError(f"No derivative rule found for op {op}")
    L.prog.n_inputs = n_in
    val, grads = env[tape.result.id]
    n_out = 0

    def out_of(x):
        nonlocal n_out
        if iszerotangent(x):
            return x
        assert isinstance(x, ir.TId), f"unexpected tape value {x!r}"
        if b.tensors[x].kind == ir.T_SEED:  # derivative is the seed itself
            x = b.op(ir.EW_COPY, (x,), ())
        tid = b.mark_output(x, n_out)
        n_out += 1
        return ("out", n_out - 1, b.tensors[tid].dims)

    L.val = out_of(val)
    for g in grads:
        L.grads.append(out_of(g))
    return L
