"""Randomised host-side check (CPU only): random loss functions over the supported op vocabulary
are traced into reference-form tapes, lowered to op programs and (a) interpreted by the test-only
NumPy interpreter against the oracle in Float64, (b) planned by libyota_cuda under every planner
configuration -- the build must succeed and describe itself, fused or not.  Deterministic seeds."""
import ctypes as C

import numpy as np
import pytest

import yota_b200 as Y
from test_host import _lowered_vs_oracle
from yota_b200 import ffi, jl
from yota_b200.grad import lower

UNARY = [jl.tanh, jl.sin, jl.cos, jl.sigmoid, jl.abs2, lambda v: v ** 2, lambda v: jl.exp(v * 0.25)]


def make_loss(seed, R=4, Cc=3):
    """A random loss over (W: R x R, x: R x Cc, b: R, idx: Int64 column indices)."""
    plan = np.random.default_rng(seed)
    steps = [(int(plan.integers(0, 8)), int(plan.integers(0, 1 << 30))) for _ in range(int(plan.integers(2, 7)))]
    tail = int(plan.integers(0, 4))

    def loss(W, x, b, idx):
        vals = [x]  # all R x (something) matrices derived from x
        for kind, salt in steps:
            v = vals[salt % len(vals)]
            u = vals[(salt >> 8) % len(vals)]
            if kind == 0:
                vals.append(UNARY[salt % len(UNARY)](v))
            elif kind == 1:
                vals.append(W @ v)
            elif kind == 2:
                vals.append(v + b)
            elif kind == 3:
                vals.append(v * b)
            elif kind == 4 and v.shape == u.shape:
                vals.append(v * u if salt & 1 else v - u)
            elif kind == 5:
                vals.append(v[jl.colon, idx])
            elif kind == 6 and v.shape[1] >= 2:
                vals.append(v[jl.colon, jl.range_(1, 2)])
            else:
                vals.append(jl.tanh(v) + v)
        out = vals[-1]
        if tail == 0:
            return jl.sum(out)
        if tail == 1:
            return jl.mean(out ** 2)
        if tail == 2:
            return jl.sum(jl.sum(out, dims=1))
        return jl.sum(out * out) + jl.sum(vals[len(vals) // 2])

    return loss


def _args(seed, R=4, Cc=3):
    r = np.random.default_rng(1000 + seed)
    return (r.random((R, R)) - 0.5, r.random((R, Cc)), r.random(R), r.integers(1, 3, size=5).astype(np.int64))  # every derived matrix keeps >= 2 columns


@pytest.mark.parametrize("seed", range(60))
def test_random_programs_lower_like_the_oracle(seed):
    _lowered_vs_oracle(make_loss(seed), _args(seed), rtol=1e-9)


@pytest.mark.parametrize("seed", range(0, 60, 3))
def test_random_programs_plan_under_every_configuration(seed):
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    L = lower(Y.gradtape(make_loss(seed), *_args(seed)))
    ops, tens = ffi.pack_program(L.prog)
    for flags in (0, ffi.FLAG_NO_FUSE, ffi.FLAG_NO_STATIC, ffi.FLAG_NO_GRAPH, ffi.FLAG_GEMM_TF32, ffi.FLAG_GEMM_BF16):
        h = C.c_void_p()
        ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
        n = lib.yota_program_describe(h, None, 0)
        assert n > 0
        st = [C.c_int64() for _ in range(5)]
        ffi.check(lib.yota_program_stats(h, *[C.byref(s) for s in st]))
        assert st[1].value <= st[0].value * 8 and st[2].value >= 0
        lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)


# ---- a wider vocabulary (CPU only: lowering + planner) ---------------------------------------
def make_loss_wide(seed):
    plan = np.random.default_rng(7000 + seed)
    steps = [(int(plan.integers(0, 12)), int(plan.integers(0, 1 << 30))) for _ in range(int(plan.integers(3, 9)))]
    tail = int(plan.integers(0, 5))

    def loss(W, x, b, idx):
        vals = [x]
        for kind, salt in steps:
            v = vals[salt % len(vals)]
            u = vals[(salt >> 8) % len(vals)]
            if kind == 0:
                vals.append(UNARY[salt % len(UNARY)](v))
            elif kind == 1 and v.shape[0] == W.shape[1]:
                vals.append(W @ v)
            elif kind == 2 and v.shape[0] == b.shape[0]:
                vals.append(v + b if salt & 1 else v * b)
            elif kind == 3 and v.shape == u.shape:
                vals.append(v * u if salt & 1 else v + u)
            elif kind == 4:
                vals.append(v[jl.colon, idx])
            elif kind == 5 and v.shape[1] >= 2:
                vals.append(v[jl.colon, jl.range_(1, 2)])
            elif kind == 6 and v.shape[0] == u.shape[0]:
                vals.append(jl.hcat(v, u))
            elif kind == 7 and v.shape[1] == u.shape[1] and v.shape[0] + u.shape[0] <= 12:
                vals.append(jl.vcat(v, u))
            elif kind == 8:
                vals.append(jl.logsoftmax(v))
            elif kind == 9 and v.shape[0] == v.shape[1]:
                vals.append(jl.transpose(v) + v)
            elif kind == 10:
                vals.append(v / (u * u + 1.5) if v.shape == u.shape else v / 2.0)
            else:
                vals.append(jl.sqrt(v * v + 1.0))
        out = vals[-1]
        if tail == 0:
            return jl.sum(out)
        if tail == 1:
            return jl.mean(out ** 2.0)
        if tail == 2:
            return jl.sum(jl.mean(out, dims=1))
        if tail == 3:
            return jl.sum(out[1, jl.colon]) + jl.sum(out[jl.colon, 1] ** 2)
        return jl.sum(out * out) + jl.sum(vals[len(vals) // 2])

    return loss


@pytest.mark.parametrize("seed", range(80))
def test_random_programs_wide_vocabulary(seed):
    L = _lowered_vs_oracle(make_loss_wide(seed), _args(seed), rtol=1e-9)
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    ops, tens = ffi.pack_program(L.prog)
    for flags in (0, ffi.FLAG_NO_FUSE, ffi.FLAG_GEMM_TF32):
        h = C.c_void_p()
        ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
        lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)


@pytest.mark.parametrize("seed", range(120))
def test_random_programs_plan_at_tile_aligned_shapes(seed):
    """Plan-only (abstract arrays, no data): the same random programs at shapes where the
    tensor-core GEMMs, their fused epilogues, the K-major / BF16 shadows and the vectorised
    regions are all eligible -- the planner must produce a program under every flag."""
    R, Cc = 256, 512
    av = [Y.AVal((R, R)), Y.AVal((R, Cc)), Y.AVal((R,)), Y.AVal((640,), "i64")]
    f = make_loss(seed) if seed % 2 else make_loss_wide(seed)
    try:
        L = lower(Y.gradtape(f, *av))
    except Exception as e:  # shape-dependent branches of the generator (vcat limits) may not apply
        pytest.skip(f"generator: {e}")
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    ops, tens = ffi.pack_program(L.prog)
    descs = {}
    for flags in (0, ffi.FLAG_NO_FUSE, ffi.FLAG_NO_STATIC, ffi.FLAG_GEMM_TF32, ffi.FLAG_GEMM_BF16,
                  ffi.FLAG_GEMM_TF32 | ffi.FLAG_NO_FUSE):
        h = C.c_void_p()
        ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
        n = lib.yota_program_describe(h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib.yota_program_describe(h, buf, n + 1)
        descs[flags] = buf.value.decode()
        lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)
    assert descs[0].count("step ") <= descs[ffi.FLAG_NO_FUSE].count("step "), "fusion never adds steps"
