"""User-facing API with the reference's surface: `grad`, `gradtape`, `compile`, `reset`,
`update`.

  grad(f, args...; seed=1)   /root/reference/src/grad.jl:349-362  (GRAD_CACHE, warm path)
  grad_compile(tape)         src/grad.jl:293-307                  -> CompiledGrad
  reset!()                   src/grad.jl:312
  update!(x, gx, fn)         src/update.jl:42-49

Device arrays (`B200Array`) in -> device arrays out.  Host NumPy arrays are accepted too:
they are uploaded, the program runs on the GPU, and the results are downloaded -- this is
the end-to-end path `bench.py` times as `e2e`.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ffi, ir
from .device import B200Array, context
from .grad import gradtape as _gradtape_abstract, lower
from .tape import AVal, Call, Constant, Input, NoTangent, ZeroTangent

GRAD_CACHE = {}

_MODE_FLAGS = {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16, "tf32x3": ffi.FLAG_GEMM_TF32X3}


def default_flags():
    f = _MODE_FLAGS[os.environ.get("YOTA_GEMM_MODE", "fp32")]
    if os.environ.get("YOTA_NO_FUSE"):
        f |= ffi.FLAG_NO_FUSE
    if os.environ.get("YOTA_NO_GRAPH"):
        f |= ffi.FLAG_NO_GRAPH
    if os.environ.get("YOTA_NO_STATIC"):
        f |= ffi.FLAG_NO_STATIC
    return f


def reset():
    """reset!(): drop all compiled gradient functions."""
    GRAD_CACHE.clear()


def _aval_of(a):
    if isinstance(a, B200Array):
        return AVal(a.dims, a.dtype)
    if isinstance(a, np.ndarray):
        return AVal(tuple(a.shape), "i64" if a.dtype.kind in "iu" else "f32")
    if isinstance(a, (float, np.floating)):
        return AVal(())
    raise TypeError(f"grad: unsupported argument type {type(a).__name__}")


def gradtape(f, *args, seed=1):
    """gradtape(f, args...; seed=1): the gradient tape for these argument shapes."""
    return _gradtape_abstract(f, *[a if isinstance(a, AVal) else _aval_of(a) for a in args], seed=seed)


class CompiledGrad:
    """The compiled tape: what `grad_compile` returns and `GRAD_CACHE` stores.  Calling it
    performs one forward + pullback sweep on the device."""

    def __init__(self, tape, ctx=None, flags=None):
        self.ctx = ctx or context()
        self.tape = tape
        self.flags = default_flags() if flags is None else flags
        self.low = lower(tape)
        ops, tensors = ffi.pack_program(self.low.prog)
        h = C.c_void_p()
        ffi.check(ffi.lib().yota_program_build(self.ctx.h, ops, len(self.low.prog.ops), tensors,
                                               len(self.low.prog.tensors), self.flags, C.byref(h)))
        self.h = h
        self.n_in, self.n_out = self.low.prog.n_inputs, self.low.prog.n_outputs
        self._captured = [(slot, B200Array.from_host(a, self.ctx)) for slot, a in self.low.captured]
        self._in = (C.c_void_p * max(1, self.n_in))()
        self._out = (C.c_void_p * max(1, self.n_out))()
        for slot, arr in self._captured:
            self._in[slot] = arr.ptr
        self._out_dims = {}
        for spec in [self.low.val] + self.low.grads:
            if isinstance(spec, tuple) and spec[0] == "out":
                self._out_dims[spec[1]] = spec[2]
        s = tape[tape.meta["seed"]].val
        self.default_seed = float(s) if isinstance(s, (int, float)) else 1.0

    # -- raw run: device arrays in, device arrays out, asynchronous -------------------
    def run(self, args, seed=None, out=None):
        """args: B200Arrays (one per tape input); seed: number or B200Array.  `out` may
        supply preallocated output arrays (list indexed by out slot) to avoid allocation."""
        low = self.low
        for a, slot in zip(args, low.arg_slots):
            self._in[slot] = a.ptr
        seed_f = self.default_seed
        if low.seed_slot is not None:
            self._in[low.seed_slot] = seed.ptr
        elif seed is not None and not isinstance(seed, str):
            seed_f = float(seed)
        outs = out if out is not None else [B200Array(self._out_dims[s], "f32", self.ctx) for s in range(self.n_out)]
        for s, o in enumerate(outs):
            self._out[s] = o.ptr
        ffi.check(ffi.lib().yota_program_run(self.h, self._in, self.n_in, self._out, self.n_out, seed_f))

        def pick(spec):
            return outs[spec[1]] if isinstance(spec, tuple) else spec

        return pick(low.val), tuple(pick(g) for g in low.grads)

    def __call__(self, f, *args, seed=None):
        host = any(isinstance(a, (np.ndarray, float, np.floating)) for a in args)
        dargs = [a if isinstance(a, B200Array) else B200Array.from_host(np.asarray(a, dtype=None), self.ctx) for a in args]
        if self.low.seed_slot is not None and not isinstance(seed, B200Array):
            seed = B200Array.from_host(np.asarray(seed, dtype=np.float32), self.ctx)
        val, grads = self.run(dargs, seed)
        if val.dims == () or host:
            val = val.to_host()
            if np.ndim(val) == 0:
                val = float(val)
        if host:
            grads = tuple(g.to_host() if isinstance(g, B200Array) else g for g in grads)
        return val, (ZeroTangent,) + grads

    def describe(self):
        n = ffi.lib().yota_program_describe(self.h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        ffi.lib().yota_program_describe(self.h, buf, n + 1)
        return buf.value.decode()

    def stats(self):
        v = [C.c_int64() for _ in range(5)]
        ffi.check(ffi.lib().yota_program_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("steps", "launches", "workspace_bytes", "fused_regions", "static_kernels"), (x.value for x in v)))

    def profile(self, args, seed=None, out=None):
        """Per-step device times (ms) of one serialised run: [(label, ms), ...]."""
        self.run(args, seed, out)  # binds pointers
        cap = 4096
        ms = (C.c_float * cap)()
        names = C.create_string_buffer(1 << 20)
        n = ffi.lib().yota_program_profile(self.h, self._in, self.n_in, self._out, self.n_out,
                                           self.default_seed if seed is None or self.low.seed_slot is not None else float(seed),
                                           ms, cap, names, len(names))
        if n < 0:
            ffi.check(n)
        labels = names.value.decode().split("\n")
        return [(labels[i] if i < len(labels) else "?", ms[i]) for i in range(n)]

    def timeline(self, args, seed=None, out=None, reps=3):
        """Both-stream timeline of one eager run: [(label, dev_start_ms, dev_end_ms, host_enqueue_ms), ...]."""
        self.run(args, seed, out)  # binds pointers
        cap = 8192
        rec = (C.c_float * (3 * cap))()
        names = C.create_string_buffer(1 << 20)
        n = ffi.lib().yota_program_timeline(self.h, self._in, self.n_in, self._out, self.n_out,
                                            self.default_seed if seed is None or self.low.seed_slot is not None else float(seed),
                                            reps, rec, cap, names, len(names))
        if n < 0:
            ffi.check(n)
        labels = names.value.decode().split("\n")
        return [(labels[i] if i < len(labels) else "?", rec[3 * i], rec[3 * i + 1], rec[3 * i + 2]) for i in range(n)]

    def set_update(self, arg_indices, lr=1.0, optimizer="sgd", beta=(0.9, 0.999), eps=1e-8):
        """Fold update!(arg, grad) into every run for these argument positions (0-based):
        arg .= arg .- lr .* grad after the sweep (and after the all-reduce, if marked).
        optimizer="adam": Optimisers.jl's Adam with program-owned moments and a device-side step
        counter (one optimiser step per run; call again to reset the state)."""
        ins = [self.low.arg_slots[i] for i in arg_indices]
        outs = []
        for i in arg_indices:
            g = self.low.grads[i]
            if not (isinstance(g, tuple) and g[0] == "out"):
                raise ValueError(f"argument {i} has no array gradient")
            outs.append(g[1])
        a = (C.c_int32 * max(1, len(ins)))(*ins)
        b = (C.c_int32 * max(1, len(outs)))(*outs)
        if optimizer == "adam":
            ffi.check(ffi.lib().yota_program_set_update_adam(self.h, a, b, len(ins), float(lr), float(beta[0]), float(beta[1]), float(eps)))
        elif optimizer == "sgd":
            ffi.check(ffi.lib().yota_program_set_update(self.h, a, b, len(ins), float(lr)))
        else:
            raise ValueError(f"unknown optimizer {optimizer!r}")

    def set_allreduce(self, out_slots):
        arr = (C.c_int32 * max(1, len(out_slots)))(*out_slots)
        ffi.check(ffi.lib().yota_program_set_allreduce(self.h, arr, len(out_slots)))

    def __del__(self):
        try:
            if self.h and ffi._lib is not None:
                ffi._lib.yota_program_destroy(self.h)
        except Exception:
            pass
        self.h = None


def grad_compile(tape, ctx=None, flags=None):
    return CompiledGrad(tape, ctx, flags)


compile = grad_compile  # noqa: A001  (Umlaut.compile / Yota's `compile` export)


def _seed_key(seed):
    if isinstance(seed, str):
        return seed
    if isinstance(seed, (int, float)):
        return "num"
    return ("arr", tuple(seed.shape))


def grad(f, *args, seed=1, flags=None, cold_opwise=False):
    """grad(f, args...; seed=1) -> (val, (ZeroTangent(), dargs...)).

    cold_opwise=True mirrors the reference's cold call: on a cache miss the tape is executed
    statement by statement on the device (eager.play, the analogue of the eager `mkcall`s at
    src/grad.jl:145-147,173 and of `tape.retval`, :357-360) while the fused program is compiled
    and cached for the warm calls.

    Gradients are returned for every argument (finalize_grad!, src/grad.jl:230-239);
    arguments the result does not depend on get ZeroTangent(), index arguments NoTangent()."""
    avals = [_aval_of(a) for a in args]
    flags = default_flags() if flags is None else flags
    # key: function + argument element types AND dims (the reference keys on types only,
    # src/grad.jl:352; a shape-specialised program must key on sizes -- SURVEY App. A.12)
    key = (f, tuple((a.dims, a.dtype) for a in avals), _seed_key(seed), flags)
    gf = GRAD_CACHE.get(key)
    if gf is None:
        tape = _gradtape_abstract(f, *avals, seed=seed)
        res = tape[tape[tape.result].args[0]]
        if not isinstance(res, Call):
            gf = _TrivialGrad(tape, res)
        else:
            gf = CompiledGrad(tape, flags=flags)
        GRAD_CACHE[key] = gf
        if cold_opwise and isinstance(gf, CompiledGrad):
            from .eager import play

            dargs = [a if isinstance(a, B200Array) or not isinstance(a, np.ndarray) else B200Array.from_host(a) for a in args]
            return play(tape, *dargs, seed=seed, flags=flags & ~ffi.FLAG_NO_GRAPH)
    return gf(f, *args, seed=seed)


class _TrivialGrad:
    """f returns a constant or one of its inputs (src/grad.jl:268-286): no device work."""

    def __init__(self, tape, res):
        self.tape, self.res = tape, res
        self.pos = [v.id for v in tape.inputs()]

    def __call__(self, f, *args, seed=1):
        if isinstance(self.res, Constant):
            return self.res.val, (ZeroTangent,) + tuple(ZeroTangent for _ in args)
        k = self.pos.index(self.res.id)
        return args[k], (ZeroTangent,) + tuple(seed if i == k else ZeroTangent for i in range(len(args)))


def update_adam(x, gx, m, v, t, lr=1e-3, beta=(0.9, 0.999), eps=1e-8):
    """One Adam step on device arrays (Optimisers.jl semantics; t >= 1 is the step number)."""
    assert all(isinstance(a, B200Array) for a in (x, gx, m, v)) and x.dims == gx.dims == m.dims == v.dims
    ffi.check(ffi.lib().yota_update_adam(x.ctx.h, x.ptr, gx.ptr, m.ptr, v.ptr, x.numel, float(lr), float(beta[0]),
                                         float(beta[1]), float(eps), int(t)))
    return x


def update(x, gx, lr=1.0):
    """update!(x, gx): x .= x .- gx on the device (src/update.jl:42-49; `lr` generalises the
    reference's `fn` to x .- lr .* gx)."""
    assert isinstance(x, B200Array) and isinstance(gx, B200Array) and x.dims == gx.dims
    ffi.check(ffi.lib().yota_update_sgd(x.ctx.h, x.ptr, gx.ptr, x.numel, float(lr)))
    return x
