"""Op-by-op execution on the device: the operator seam of the boundary (Seam 2).

The reference runs every statement of a tape eagerly when it is traced and built -- `mkcall`
executes each primitive (/root/reference/src/grad.jl:145-147 the `rrule`s, :173 the pullbacks),
which is how a cold `grad` call already returns `tape.retval` (:357-360) -- and `Umlaut.play!`
re-executes a finished tape statement by statement (test/test_grad.jl:205, 279-282).  With
`CuArray` arguments each of those statements is one or a few kernel launches.

This module is that path for `B200Array`s: `rrule(f, args...)` runs the forward of ONE rule as its
own tiny device program and returns `(value, pullback)`; the pullback closure runs ONE pullback as
its own program; `play(tape, args...)` interprets a finished gradient tape with them.  Nothing is
computed on the host: every statement goes through `yota_program_build` / `yota_program_run` (a
single-op program uses the same kernels as the fused plan, just one launch per op).  The fused,
graph-replayed program of `CompiledGrad` remains the warm path; tests check
compiled == play == oracle, the analogue of test/test_grad.jl:271-289.

The Julia glue mirrors it with `rrule(::YotaRuleConfig, f, ::B200Array...)` methods
(julia/YotaB200.jl, section "operator seam").
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ffi, ir, rules
from .device import B200Array, context
from .tape import AVal, Call, Constant, Input, V, NoTangent, ZeroTangent, iszerotangent

_PROGRAMS = {}


class DeviceProgram:
    """An op program with array inputs/outputs bound positionally (one launch list, no graph:
    a cold statement runs once)."""

    def __init__(self, prog: ir.IRProgram, ctx, flags):
        self.ctx, self.prog = ctx, prog
        ops, tensors = ffi.pack_program(prog)
        h = C.c_void_p()
        ffi.check(ffi.lib().yota_program_build(ctx.h, ops, len(prog.ops), tensors, len(prog.tensors),
                                               flags | ffi.FLAG_NO_GRAPH, C.byref(h)))
        self.h = h
        self.out_dims = {t.slot: t.dims for t in prog.tensors if t.kind == ir.T_OUTPUT}
        self._in = (C.c_void_p * max(1, prog.n_inputs))()
        self._out = (C.c_void_p * max(1, prog.n_outputs))()

    def __call__(self, ins, seed=1.0):
        for i, a in enumerate(ins):
            self._in[i] = a.ptr
        outs = [B200Array(self.out_dims[s], "f32", self.ctx) for s in range(self.prog.n_outputs)]
        for s, o in enumerate(outs):
            self._out[s] = o.ptr
        ffi.check(ffi.lib().yota_program_run(self.h, self._in, self.prog.n_inputs, self._out, self.prog.n_outputs, float(seed)))
        return outs

    def __del__(self):
        try:
            if self.h and ffi._lib is not None:
                ffi._lib.yota_program_destroy(self.h)
        except Exception:
            pass
        self.h = None


def _sig(a):
    if isinstance(a, B200Array):
        return ("arr", a.dims, a.dtype)
    if isinstance(a, AVal):
        return ("aval", a.dims, a.dtype)
    if isinstance(a, (list, tuple)):
        return tuple(_sig(x) for x in a)
    return repr(a)


def _aval(a):
    return AVal(a.dims, a.dtype) if isinstance(a, B200Array) else a


def rrule(fn, *args, kw=None, flags=0, ctx=None, out=None):
    """rrule(YotaRuleConfig(), fn, args...) on device arrays: executes the rule's forward NOW (one
    single-rule device program) and returns (value, pullback).  `pullback(dy)` executes the rule's
    pullback as its own program and returns the tangent tuple, one entry per (fn, args...)
    position (NoTangent / ZeroTangent / B200Array), as ChainRules does."""
    kw = kw or {}
    ctx = ctx or context()
    avals = [_aval(a) for a in args]
    out = out or rules.infer(fn, avals, kw)
    key = ("fwd", fn, _sig(args), tuple(sorted(kw.items())), flags)
    ent = _PROGRAMS.get(key)
    arrays = [a for a in args if isinstance(a, B200Array)]
    if ent is None:
        b = ir.IRProgram()
        vals, n = [], 0
        for a in args:
            if isinstance(a, B200Array):
                vals.append(b.tensor(a.dims, ir.I64 if a.dtype == "i64" else ir.F32, ir.T_INPUT, n))
                n += 1
            else:
                vals.append(a)
        b.n_inputs = n
        y, saved = rules.lower_fwd(b, fn, vals, kw, out)
        y_out = b.mark_output(y, 0)
        ent = (DeviceProgram(b, ctx, flags), b, y, saved, y_out)
        _PROGRAMS[key] = ent
    prog, b, y, saved, _ = ent
    val = prog(arrays)[0]

    def pullback(dy):
        if iszerotangent(dy):
            return tuple(ZeroTangent for _ in range(len(args) + 1))
        return _run_pullback(fn, b, y, saved, arrays, val, dy, key, flags, ctx)

    return val, pullback


def _run_pullback(fn, b, y, saved, arrays, val, dy, fkey, flags, ctx):
    key = ("pb", fkey, _sig(dy))
    ent = _PROGRAMS.get(key)
    if ent is None:
        b2 = ir.IRProgram()
        binds = []  # what each input slot of the pullback program is bound to: ("dy",) | ("arg", i) | ("val",)
        memo = {}

        def inp(dims, dtype, what):
            t = b2.tensor(dims, dtype, ir.T_INPUT, len(binds))
            binds.append(what)
            return t

        def remap(x):
            if isinstance(x, ir.TId):
                if x in memo:
                    return memo[x]
                t = b.tensors[x]
                if t.kind == ir.T_CONST:
                    r = b2.const(t.cval)
                elif t.kind == ir.T_INPUT:
                    r = inp(t.dims, t.dtype, ("arg", t.slot))
                elif int(x) == int(y) or t.kind == ir.T_OUTPUT:
                    r = inp(t.dims, t.dtype, ("val",))
                else:
                    raise rules.NoRuleError(f"{fn}: pullback needs an intermediate of the forward rule that is not kept")
                memo[x] = r
                return r
            if isinstance(x, tuple):
                return tuple(remap(e) for e in x)
            if isinstance(x, list):
                return [remap(e) for e in x]
            return x

        saved2 = remap(saved)
        d2 = inp(dy.dims, ir.F32, ("dy",)) if isinstance(dy, B200Array) else b2.const(float(dy))
        b2.n_inputs = len(binds)
        tangents = rules.lower_pb(b2, fn, saved2, d2)
        outs = []
        for t in tangents:
            if isinstance(t, ir.TId):
                if b2.tensors[t].kind == ir.T_CONST:  # the tangent is a constant (e.g. the seed itself)
                    t = b2.op(ir.EW_COPY, (t,), ())
                b2.mark_output(t, len(outs))
                outs.append(True)
        ent = (DeviceProgram(b2, ctx, flags) if outs else None, binds, tangents)
        _PROGRAMS[key] = ent
    prog, binds, tangents = ent
    ins = [dy if w[0] == "dy" else val if w[0] == "val" else arrays[w[1]] for w in binds]
    res = iter(prog(ins)) if prog is not None else iter(())
    return tuple(next(res) if isinstance(t, ir.TId) else t for t in tangents)


def _add(x, y, flags, ctx):
    """broadcast(+, x, y): the tape's gradient accumulation (src/grad.jl:91), one device op."""
    key = ("add", _sig(x), _sig(y), flags)
    prog = _PROGRAMS.get(key)
    if prog is None:
        b = ir.IRProgram()
        a = b.tensor(x.dims, ir.F32, ir.T_INPUT, 0)
        c = b.tensor(y.dims, ir.F32, ir.T_INPUT, 1)
        b.n_inputs = 2
        b.mark_output(b.op(ir.EW_ADD, (a, c), ir.bcast_dims(x.dims, y.dims)), 0)
        prog = _PROGRAMS[key] = DeviceProgram(b, ctx, flags)
    return prog([x, y])[0]


def play(tape, *args, seed=None, flags=0, ctx=None):
    """Umlaut.play!(tape, f, args...): interpret a finished gradient tape statement by statement on
    the device.  Returns (val, (ZeroTangent, grads...)) like the compiled function."""
    ctx = ctx or context()
    env = {}
    it = iter(args)
    cap = dict(tape.captured)
    seed_id = tape.meta["seed"].id
    for op in tape.ops:
        i = op.id
        if isinstance(op, Input):
            a = next(it)
            env[i] = a if isinstance(a, B200Array) or not isinstance(a, np.ndarray) else B200Array.from_host(a, ctx)
        elif isinstance(op, Constant):
            if i == seed_id:
                s = op.val
                if isinstance(s, tuple) and s[0] == "ones":
                    env[i] = B200Array.from_host(np.ones(s[1].dims, np.float32), ctx)
                elif isinstance(s, tuple) and s[0] == "array":
                    env[i] = seed if isinstance(seed, B200Array) else B200Array.from_host(np.asarray(seed, np.float32), ctx)
                else:
                    env[i] = float(seed) if isinstance(seed, (int, float)) else float(s)
            elif i in cap:
                env[i] = B200Array.from_host(cap[i], ctx)
            else:
                env[i] = op.val
        else:
            fn = op.fn
            if fn == "rrule":
                f, fargs = op.args[1], op.args[2:]
                vals = [env[a.id] if isinstance(a, V) else a for a in fargs]
                env[i] = rrule(f, *vals, kw=op.kw, flags=flags, ctx=ctx, out=op.val[1])
            elif fn == "_getfield":
                env[i] = env[op.args[0].id][op.args[1] - 1]
            elif isinstance(fn, V):  # pb(dy)
                env[i] = env[fn.id](env[op.args[0].id])
            elif fn == "getfield":
                env[i] = env[op.args[0].id][op.args[1] - 1]
            elif fn == "broadcast":
                x, y = env[op.args[1].id], env[op.args[2].id]
                env[i] = y if iszerotangent(x) else x if iszerotangent(y) else _add(x, y, flags, ctx)
            elif fn == "tuple":
                env[i] = tuple(env[a.id] if isinstance(a, V) else a for a in op.args)
            elif fn == "map":
                env[i] = env[op.args[1].id]
            else:
                raise rules.NoRuleError(f"No derivative rule found for op {op}")
    val, grads = env[tape.result.id]
    grads = tuple(B200Array.from_host(np.asarray(g, np.float32), ctx) if isinstance(g, float) else g for g in grads)
    if isinstance(val, B200Array) and val.dims == ():
        val = float(val.to_host())
    return val, (ZeroTangent,) + grads
