"""GPU parity: the CUDA path, called through the public API / C ABI, against the oracle.

FP32 gate: rtol 1e-5 relative to max|g| or twice the oracle's own Float32
rounding error, both measured against the Float64 oracle.  Index / scatter pullbacks and
fused-vs-unfused runs: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import cases
import gpu_util as U
import yota_b200 as Y
from oracle import c_oracle
from oracle import yota_oracle as O
from yota_b200 import ffi, jl

pytestmark = pytest.mark.gpu
F32 = np.float32
# random programs whose gate is looser than the reference's 1e-5, with the reason (filled from a
# device run that printed every seed's worst error: see the test)
FUZZ_RTOL = {}


def f32(args):
    return [a.astype(F32) if isinstance(a, np.ndarray) and a.dtype.kind == "f" else a for a in args]


# ---- the reference's own test cases, on the device ---------------------------------------
@pytest.mark.parametrize("name,f,args", cases.gradcheck_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_reference_cases(name, f, args):
    U.check_parity(f, f32(args))


@pytest.mark.parametrize("flags", [ffi.FLAG_NO_FUSE, ffi.FLAG_NO_GRAPH, ffi.FLAG_NO_STATIC])
def test_reference_cases_other_plans(flags):
    for name, f, args in cases.gradcheck_cases():
        U.check_parity(f, f32(args), flags=flags)


def test_getindex_pullbacks_exact():  # test/test_grad.jl:87-95
    x = cases.rng(0).random((3, 4, 5)).astype(F32)
    dx = Y.cu(x)
    for f, sel in ((lambda x: x[1], (0, 0, 0)), (lambda x: x[1, 2, 1], (0, 1, 0))):
        val, g = Y.grad(f, dx)
        e = np.zeros_like(x)
        e[sel] = 1
        assert np.array_equal(g[1].to_host(), e) and val == float(x[sel])
    _, g = Y.grad(lambda x: jl.sum(x[jl.colon, 1, jl.colon]), dx)
    e = np.zeros_like(x)
    e[:, 0, :] = 1
    assert np.array_equal(g[1].to_host(), e)


def test_ungetindex_known_answers_on_device():  # test/test_helpers.jl:11-25,74-86
    ctx = Y.context()
    dy = Y.cu(np.ones((4, 3), F32))
    idx = Y.cu(np.array([1, 3, 1], np.int64))
    dx = Y.B200Array((4, 5))
    ffi.check(ffi.lib().yota_scatter_add_cols(ctx.h, dx.ptr, 4, 5, dy.ptr, idx.ptr, 3))
    assert np.array_equal(dx.to_host(), np.array([[2, 0, 1, 0, 0]] * 4, F32))
    _, g = Y.grad(lambda x: jl.sum(x[[1, 3, 1]]), Y.cu(np.zeros(4, F32)))
    assert np.array_equal(g[1].to_host(), [2, 0, 1, 0])


def test_seed_forms():  # test/test_grad.jl:253-268
    x = Y.cu(np.array([1, 2, 3], F32))
    for seed in (Y.cu(np.ones(3, F32)), "auto"):
        val, g = Y.grad(lambda x: 2 * x, x, seed=seed)
        assert np.array_equal(val.to_host(), [2, 4, 6]) and np.array_equal(g[1].to_host(), [2, 2, 2])
    a = f32(cases.gradcheck_cases()[1][2])
    v1, g1 = Y.grad(cases.loss_tanh, *map(Y.cu, a))
    v2, g2 = Y.grad(cases.loss_tanh, *map(Y.cu, a), seed=2.0)
    assert v1 == v2 and np.allclose(2 * g1[1].to_host(), g2[1].to_host(), rtol=1e-6)


def test_cached_equals_fresh_and_host_arrays():  # test/test_grad.jl:271-289
    a = f32(cases.gradcheck_cases()[2][2])
    Y.reset()
    r1 = Y.grad(cases.loss_abs2, *map(Y.cu, a))
    r2 = Y.grad(cases.loss_abs2, *map(Y.cu, a))  # cached compiled tape (graph replay)
    r3 = Y.grad(cases.loss_abs2, *map(Y.cu, a))
    rh = Y.grad(cases.loss_abs2, *a)  # host arrays in -> host arrays out
    assert r1[0] == r2[0] == r3[0] == rh[0]
    for x, y, z, h in zip(r1[1][1:], r2[1][1:], r3[1][1:], rh[1][1:]):
        assert np.array_equal(U.bits(x.to_host()), U.bits(y.to_host()))
        assert np.array_equal(U.bits(x.to_host()), U.bits(z.to_host()))
        assert np.array_equal(U.bits(x.to_host()), U.bits(h))


def test_update():  # src/update.jl:42-49
    x, g = Y.cu(np.array([1, 2, 3, 4], F32)), Y.cu(np.array([0.5, 0.5, 1, 1], F32))
    Y.update(x, g)
    assert np.array_equal(x.to_host(), [0.5, 1.5, 2, 3])


def test_transpose_cat():  # test/test_grad.jl:98-104,175-182
    r = cases.rng(5)
    U.check_parity(lambda a, b: jl.sum(jl.vcat(a, b) ** 2), [r.random((2, 3)).astype(F32), r.random((3, 3)).astype(F32)])
    U.check_parity(lambda a, b: jl.sum(jl.hcat(a, b) ** 2), [r.random((2, 3)).astype(F32), r.random((2, 4)).astype(F32)])
    U.check_parity(lambda a: jl.sum(jl.transpose(a) * jl.transpose(a)), [r.random((5, 7)).astype(F32)])


def test_unary_rules():
    r = cases.rng(6)
    x = (r.random((8, 6)) + 0.5).astype(F32)
    for fn in (jl.exp, jl.log, jl.sin, jl.cos, jl.sqrt, jl.relu, jl.sigmoid, jl.abs2, jl.tanh):
        U.check_parity(lambda a: jl.sum(fn(a - 0.7) if fn in (jl.relu,) else fn(a)), [x])
    U.check_parity(lambda a, b: jl.sum(a / b), [x, (r.random((8, 1)) + 1).astype(F32)])


# ---- GEMM --------------------------------------------------------------------------------
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_ffma(tA, tB):
    r = cases.rng(8)
    for m, n, k in [(128, 128, 64), (200, 77, 33), (10, 64, 128), (1, 5, 3), (257, 130, 300), (640, 384, 100), (300, 500, 77)]:
        A = r.standard_normal((k, m) if tA else (m, k)).astype(F32)
        B = r.standard_normal((n, k) if tB else (k, n)).astype(F32)
        ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
        got = U.gemm(0, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max()), (m, n, k)
        C0 = np.asfortranarray(r.standard_normal((m, n)).astype(F32))
        got = U.gemm(0, tA, tB, np.asfortranarray(A), np.asfortranarray(B), C0)
        assert np.allclose(got, ref + C0, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


# ---- the benchmark configs, scaled so the oracle finishes in seconds --------------------
def test_c1_mlp_full_size():
    U.check_parity(cases.c1_mlp, cases.c1_args())


def test_c2_chain_scaled_and_ragged():
    for n, m in [(512, 384), (1024, 1024), (250, 33), (4, 3), (1030, 7)]:
        U.check_parity(cases.c2_chain, cases.c2_args(n, m))


def test_c2_fused_equals_unfused_bitwise():
    a = [Y.cu(x) for x in cases.c2_args(1024, 512)]
    ref = None
    for flags in (0, ffi.FLAG_NO_FUSE, ffi.FLAG_NO_STATIC, ffi.FLAG_NO_GRAPH):
        val, g = Y.grad(cases.c2_chain, *a, flags=flags)
        got = [U.bits(x.to_host()) for x in g[1:3]]  # dx, dW: elementwise, order-independent
        if ref is None:
            ref = got
        for x, y in zip(ref, got):
            assert np.array_equal(x, y), flags


def test_c3_mlp_scaled():
    U.check_parity(cases.make_c3(3, 96), cases.c3_args(3, 64, 96))
    U.check_parity(cases.make_c3(8, 256), cases.c3_args(8, 256, 256))


def test_c4_rnn_scaled():
    U.check_parity(cases.make_c4(5), cases.c4_args(32, 16, 5))
    U.check_parity(cases.make_c4(16), cases.c4_args(128, 32, 16))


def test_c5_embedding_bit_exact():
    for rows, nt, ni, zipf in [(8, 16, 100, False), (256, 1024, 8192, False), (256, 512, 20000, True), (3, 7, 50, False),
                               (64, 5000, 100, False)]:
        E, idx, V = cases.c5_args(rows, nt, ni, zipf=zipf)
        val, g = Y.grad(cases.c5_embed, Y.cu(E), Y.cu(idx), Y.cu(V))
        dE = c_oracle.scatter_add_cols(V, idx, nt)  # dy == V (seed 1): the sequential fold
        assert np.array_equal(U.bits(g[1].to_host()), U.bits(dE)), (rows, nt, ni, zipf)
        assert g[2] is Y.NoTangent
        assert np.array_equal(U.bits(g[3].to_host()), U.bits(c_oracle.gather_cols(E, idx)))  # dV = E[:, idx] exactly
        ref = c_oracle.gather_dot(E, V, idx)
        assert abs(val - ref) <= 1e-5 * abs(ref) + 1e-3


@pytest.mark.parametrize("flags", [0, ffi.FLAG_NO_GRAPH])
def test_index_sort_on_the_side_stream_follows_the_bound_indices(flags):
    """The sort of a scatter-add runs beside the forward pass (side stream / parallel branch of the
    graph).  Replaying ONE compiled program on new index vectors and tangents -- same buffers, new
    contents, and other buffers -- must re-sort every run and stay bit-exact."""
    rows, nt, ni = 64, 700, 9000
    E, idx, V = cases.c5_args(rows, nt, ni)
    dE_, dI_, dV_ = Y.cu(E), Y.cu(idx), Y.cu(V)
    gf = Y.CompiledGrad(Y.gradtape(cases.c5_embed, dE_, dI_, dV_), Y.context(), flags)
    assert "index sort on the side stream" in gf.describe()
    r = cases.rng(33)
    for it in range(6):
        if it % 2:  # other buffers: the graph is re-bound
            dI_, dV_ = Y.cu(idx), Y.cu(V)
        else:  # same buffers, new contents
            dI_.copy_from_host(idx)
            dV_.copy_from_host(V)
        val, g = gf(cases.c5_embed, dE_, dI_, dV_)
        assert np.array_equal(U.bits(g[1].to_host()), U.bits(c_oracle.scatter_add_cols(V, idx, nt))), it
        assert np.array_equal(U.bits(g[3].to_host()), U.bits(c_oracle.gather_cols(E, idx))), it
        idx = r.integers(1, nt + 1, size=ni).astype(np.int64) if it != 3 else np.full(ni, 5, np.int64)
        V = np.asfortranarray(r.standard_normal((rows, ni), dtype=F32))


def test_scatter_negative_zero_and_empty_destinations():
    ctx = Y.context()
    dy = np.asfortranarray(np.array([[-0.0, 1.0, -0.0], [-0.0, -1.0, 2.0]], F32))
    idx = np.array([2, 4, 2], np.int64)
    dx = Y.B200Array((2, 5))
    ffi.check(ffi.lib().yota_scatter_add_cols(ctx.h, dx.ptr, 2, 5, Y.cu(dy).ptr, Y.cu(idx).ptr, 3))
    got = dx.to_host()
    assert np.array_equal(U.bits(got), U.bits(c_oracle.scatter_add_cols(dy, idx, 5)))
    assert not np.signbit(got).any() or got[1, 3] == -1.0  # 0.0f + (-0.0f) = +0.0f; untouched columns are +0.0f


def test_scatter_standalone_large_bit_exact():
    r = cases.rng(9)
    rows, nd, ns = 256, 1 << 14, 1 << 17
    dy = r.standard_normal((rows, ns), dtype=F32)
    idx = r.integers(1, nd + 1, size=ns).astype(np.int64)
    ctx = Y.context()
    dx = Y.B200Array((rows, nd))
    ffi.check(ffi.lib().yota_scatter_add_cols(ctx.h, dx.ptr, rows, nd, Y.cu(dy).ptr, Y.cu(idx).ptr, ns))
    assert np.array_equal(U.bits(dx.to_host()), U.bits(c_oracle.scatter_add_cols(dy, idx, nd)))


def test_general_getindex_forms():  # test/test_helpers.jl:29-71 index forms, now value-checked
    r = cases.rng(10)
    x = r.random((3, 4)).astype(F32)
    forms = [(jl.range_(1, 2), jl.range_(2, 3)), (1, jl.range_(2, 3)), (1, jl.colon), (1, [1, 3]), ([3, 1, 3], jl.colon)]
    for I in forms:
        U.check_parity(lambda a: jl.sum(a[I] * a[I]), [x])
    U.check_parity(lambda a: a[1, 2] * a[5], [x])
    x3 = r.random((4, 3, 5)).astype(F32)
    U.check_parity(lambda a: jl.sum(a[jl.colon, 2, jl.colon] ** 2) + jl.sum(a[jl.colon, jl.colon, 5]), [x3])


def test_deterministic_run_to_run():
    a = [Y.cu(x) for x in cases.c3_args(3, 128, 200)]
    f = cases.make_c3(3, 200)
    r1 = Y.grad(f, *a)
    r2 = Y.grad(f, *a)
    assert r1[0] == r2[0]
    for x, y in zip(r1[1][1:], r2[1][1:]):
        assert np.array_equal(U.bits(x.to_host()), U.bits(y.to_host()))


def test_c5_fused_plans_bit_exact_with_seed():
    """getindex fused into the region as a gather operand, the fold reading V scaled by the seed:
    every plan (fused / static chains off / unfused) gives the bits of the oracle, seed != 1 too."""
    E, idx, V = cases.c5_args(64, 300, 2000)
    for seed in (1, 0.5, -3.0):
        oval, og = O.grad(cases.c5_embed, E, idx, V, seed=seed, dtype=np.float32)
        for flags in (0, ffi.FLAG_NO_STATIC, ffi.FLAG_NO_FUSE):
            val, g = Y.grad(cases.c5_embed, Y.cu(E), Y.cu(idx), Y.cu(V), seed=seed, flags=flags)
            assert np.array_equal(U.bits(g[1].to_host()), U.bits(og[1])), (seed, flags)
            assert np.array_equal(U.bits(g[3].to_host()), U.bits(og[3])), (seed, flags)
            assert abs(val - float(oval)) <= 1e-5 * abs(float(oval)) + 1e-3


def test_c4_rnn_tensor_core_modes():
    """Tile-aligned RNN in TF32 and BF16 modes: the x[:, :, t] views (element offsets into x and
    into its BF16 shadow), in-place C += A*B with batched loads, narrow-tile GEMMs."""
    args = cases.c4_args(128, 128, 6)
    f = cases.make_c4(6)
    v64, g64 = O.grad(f, *U.to_f64(list(args)), dtype=np.float64)
    for flags, tol in ((ffi.FLAG_GEMM_TF32, 5e-3), (ffi.FLAG_GEMM_BF16, 4e-2)):
        val, g = Y.grad(f, *map(Y.cu, args), flags=flags)
        assert abs(val - v64) <= tol * abs(v64), (flags, val, v64)
        for a, b in zip(g[1:], g64[1:]):
            a = a.to_host()
            assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(b)), (flags, np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_examples_converge_on_device():
    """The reference's end-to-end fixtures through the device path (test/test_examples.jl:12-36,
    46-70): 100 warm grad() calls from GRAD_CACHE + update! on device arrays."""
    Yv, X, theta = (a.astype(F32) for a in cases.linreg_data())
    dY, dX, dth = Y.cu(Yv), Y.cu(X), Y.cu(theta)
    first, _ = Y.grad(cases.linreg_obj, dY, dX, dth)
    last = first
    for _ in range(100):
        last, g = Y.grad(cases.linreg_obj, dY, dX, dth)
        Y.update(dth, g[3], 0.01)
    assert last < first / 10
    X, Yv, W_true, b_true, W, b = (a.astype(F32) for a in cases.linreg2_data())
    dW, db, dX, dYv = Y.cu(W), Y.cu(b), Y.cu(X), Y.cu(Yv)
    for _ in range(100):
        _, g = Y.grad(cases.linear2_loss, dW, db, dX, dYv)
        Y.update(dW, g[1], 0.1)
        Y.update(db, g[2], 0.1)
    assert np.allclose(dW.to_host(), W_true, atol=0.1) and np.allclose(db.to_host(), b_true, atol=0.1)


def test_update_folded_into_the_run():
    """yota_program_set_update: update!(x, gx, (x, g) -> x .- lr .* g) (src/update.jl:42-49) applied
    inside the program run -- one eval is one training step; bitwise the host formula."""
    r = cases.rng(12)
    W, b, x = r.random((4, 3)).astype(F32), r.random(4).astype(F32), r.random((3, 5)).astype(F32)
    dW, db, dx = Y.cu(W), Y.cu(b), Y.cu(x)
    gf = Y.CompiledGrad(Y.gradtape(cases.loss_tanh, dW, db, dx))
    gf.set_update([0, 1], lr=0.25)
    lr = F32(0.25)
    for _ in range(3):  # eager run, captured run, graph replay
        val, g = gf.run([dW, db, dx])
        gW, gb = g[0].to_host(), g[1].to_host()
        _, og = O.grad(cases.loss_tanh, W, b, x, dtype=np.float32)
        assert np.allclose(gW, og[1], rtol=1e-5, atol=1e-6) and np.allclose(gb, og[2], rtol=1e-5, atol=1e-6)
        W, b = W - lr * gW, b - lr * gb
        assert np.array_equal(U.bits(dW.to_host()), U.bits(W)) and np.array_equal(U.bits(db.to_host()), U.bits(b))
    gf.set_update([], lr=1.0)
    gf.run([dW, db, dx])
    assert np.array_equal(U.bits(dW.to_host()), U.bits(W))  # marks cleared: parameters untouched


@pytest.mark.parametrize("seed", range(0, 60, 2))
def test_random_programs_on_device(seed):
    """The random loss functions of tests/test_fuzz_host.py through the device path against the
    oracle (Float32 and Float64, the SURVEY §8(d) gate): exercises the region interpreter on
    chains that have no ahead-of-time kernel, gather operands, slice views, small GEMMs,
    in-place accumulation and the unbroadcast sinks in combinations no hand-written case has."""
    import test_fuzz_host as Fz

    args = [a.astype(F32) if a.dtype.kind == "f" else a for a in Fz._args(seed)]
    U.check_parity(Fz.make_loss(seed), args, rtol=FUZZ_RTOL.get(seed, 1e-5))


def test_runtime_index_out_of_range_is_a_bounds_error():
    """A run-time index vector with an entry outside 1..n: the reference throws BoundsError from
    getindex / NNlib.scatter! (src/helpers.jl:7-11).  The device path reports it from the kernel
    that reads the index (stand-alone gather, the fused gather operand, the scatter-add's key
    pass) and the next synchronising call fails with YOTA_E_SHAPE -- forward and pullback alike."""
    r = cases.rng(21)
    E, V = r.random((8, 16)).astype(F32), r.random((8, 12)).astype(F32)
    good = r.integers(1, 17, size=12).astype(np.int64)
    for bad_value in (0, 17, -3, 1 << 40):
        idx = good.copy()
        idx[5] = bad_value
        for f in (cases.c5_embed,  # gather fused into the region + scaled scatter-add
                  lambda E, idx, V: jl.sum(jl.tanh(E[jl.colon, idx]) @ jl.transpose(V))):  # stand-alone gather + scatter
            with pytest.raises(Y.YotaError, match="BoundsError") as ei:
                _, g = Y.grad(f, Y.cu(E), Y.cu(idx), Y.cu(V))
                g[1].to_host()
            assert ei.value.code == -3
    # the flag is cleared by the failing call: the next, valid call is clean
    val, g = Y.grad(cases.c5_embed, Y.cu(E), Y.cu(good), Y.cu(V))
    _, og = O.grad(cases.c5_embed, E, good, V)
    assert np.array_equal(U.bits(g[1].to_host()), U.bits(og[1]))
    # stand-alone operators of the C ABI
    lib, ctx = ffi.lib(), Y.context()
    dx, dy, di = Y.B200Array((8, 16)), Y.cu(V), Y.cu(np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 99], np.int64))
    ffi.check(lib.yota_scatter_add_cols(ctx.h, dx.ptr, 8, 16, dy.ptr, di.ptr, 12))
    with pytest.raises(Y.YotaError, match="BoundsError"):
        ctx.sync()
    ctx.sync()


def test_accumulation_never_lands_in_a_view_of_a_shared_array():
    """a matmul thunk added to x[:, :, t] of a saved activation must not be accumulated
    in place (the view's parent is read again by the tanh pullback)."""
    import test_host

    r = cases.rng(31)
    args = [r.random((8, 4, 3)).astype(F32), r.random((8, 8)).astype(F32), r.random((8, 4)).astype(F32)]
    U.check_parity(test_host.alias_hazard_loss, args)
    big = [r.random((256, 128, 3)).astype(F32) - F32(0.5), (r.random((256, 256)).astype(F32) - F32(0.5)) / 16,
           r.random((256, 128)).astype(F32)]
    for flags in (0, ffi.FLAG_GEMM_TF32):
        U.check_parity(test_host.alias_hazard_loss, big, flags=flags, rtol=1e-5 if flags == 0 else 5e-3)


# ---- the operator seam: statement-by-statement execution on the device (eager.play) --------
def _cmp_play(f, args, seed=1, flags=0, tol=1e-6, bitwise=()):
    d = [Y.cu(a) if isinstance(a, np.ndarray) else a for a in args]
    dseed = Y.cu(np.asarray(seed, np.float32)) if isinstance(seed, np.ndarray) else seed
    tape = Y.gradtape(f, *d, seed=dseed if not isinstance(dseed, Y.B200Array) else seed)
    v1, g1 = Y.play(tape, *d, seed=dseed, flags=flags)
    v2, g2 = Y.grad(f, *d, seed=dseed, flags=flags)
    h = lambda x: x.to_host() if isinstance(x, Y.B200Array) else x  # noqa: E731
    assert np.allclose(h(v1), h(v2), rtol=tol, atol=0)
    for i, (a, b) in enumerate(zip(g1[1:], g2[1:])):
        if isinstance(b, Y.B200Array):
            a, b = a.to_host(), b.to_host()
            if i in bitwise:
                assert np.array_equal(U.bits(a), U.bits(b)), i
            assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), (i, np.abs(a - b).max())
        else:
            assert repr(a) == repr(b), (i, a, b)
    return v1, g1


@pytest.mark.parametrize("name,f,args", cases.gradcheck_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_play_equals_compiled_on_the_reference_cases(name, f, args):
    """compiled == play! == grad (test/test_grad.jl:271-289): the tape interpreted one rule at a
    time on the device (every rrule forward and every pullback its own tiny program) against the
    fused, graph-replayed program."""
    _cmp_play(f, f32(args), tol=2e-5)


def test_play_on_the_configs_and_seed_forms():
    _cmp_play(cases.c1_mlp, list(cases.c1_args(48, 32, 10, 16)), tol=2e-5)
    _cmp_play(cases.make_c3(3, 64), list(cases.c3_args(3, 32, 64)), tol=2e-5)
    _cmp_play(cases.make_c4(4), list(cases.c4_args(32, 8, 4)), tol=2e-5)
    E, idx, V = cases.c5_args(8, 64, 300)
    _cmp_play(cases.c5_embed, [E, idx, V], bitwise=(0,), tol=2e-5)  # scatter-add: the same bits either way
    x = np.array([1.0, 2, 3], F32)
    _cmp_play(lambda x: 2 * x, [x], seed=np.ones(3, F32))
    _cmp_play(lambda x: 2 * x, [x], seed="auto")
    v, g = Y.grad(cases.loss_tanh, *map(Y.cu, f32(cases.gradcheck_cases()[1][2])), cold_opwise=True)
    _, og = O.grad(cases.loss_tanh, *cases.gradcheck_cases()[1][2])
    assert np.allclose(g[1].to_host(), og[1], rtol=1e-5, atol=1e-6)


def test_eager_rrule_is_the_operator_api():
    """rrule(cfg, *, A, B) on device arrays: value now, pullback later (src/grad.jl:147,173)."""
    r = cases.rng(41)
    A, B = r.random((256, 128)).astype(F32), r.random((128, 256)).astype(F32)
    val, pb = Y.rrule("matmul", Y.cu(A), Y.cu(B))
    assert np.allclose(val.to_host(), A @ B, rtol=1e-5)
    dy = r.random((256, 256)).astype(F32)
    _, dA, dB = pb(Y.cu(dy))
    assert np.allclose(dA.to_host(), dy @ B.T, rtol=1e-5) and np.allclose(dB.to_host(), A.T @ dy, rtol=1e-5)
    y, pb = Y.rrule("broadcasted", "tanh", Y.cu(A))
    t = pb(Y.cu(np.ones_like(A)))
    assert t[0] is Y.NoTangent and t[1] is Y.NoTangent
    assert np.allclose(t[2].to_host(), 1 - np.tanh(A) ** 2, rtol=1e-5, atol=1e-6)


# ---- the reference's example models and optimiser on the device ------------------------------
def test_convnet_and_relu_mlp_examples():
    """examples/flux/convnet.jl:53-66 (Conv + relu, MaxPool, flatten, Dense, logitcrossentropy) and
    examples/flux/mlp.jl:52-56,81 through the composite layers: conv = linear-index gather + GEMM +
    permuting gather, so its pullback is the bit-exact scatter-add (col2im) + two GEMMs."""
    U.check_parity(cases.convnet_loss, f32(cases.convnet_args()))
    big = f32(cases.convnet_args(W=18, H=18, C=4, N=16, CO=16, classes=10))
    U.check_parity(cases.convnet_loss, big)
    a = cases.c1_args(48, 32, 10, 64)
    y = np.eye(10, dtype=F32)[:, cases.rng(5).integers(0, 10, 64)]
    U.check_parity(cases.mlp_relu_loss, list(a[:5]) + [np.asfortranarray(y)])
    U.check_parity(cases.sum_of_f, [cases.rng(6).standard_normal((33, 20)).astype(F32)])


def _adam_host(x, g, m, v, t, lr, b1, b2, eps):
    f = F32
    m = f(b1) * m + (f(1) - f(b1)) * g
    v = f(b2) * v + (f(1) - f(b2)) * (g * g)
    c1, c2 = f(1) - f(b1) ** f(t), f(1) - f(b2) ** f(t)
    return x - (m / c1) / (np.sqrt(v / c2) + f(eps)) * f(lr), m, v


def test_adam_standalone_and_folded_into_the_run():
    """Optimisers.Adam (examples/flux/mlp.jl:82): the stand-alone step against the Float32 host
    formula, then folded into the run -- program-owned moments, device-side step counter -- over an
    eager run, the captured run and two graph replays (the counter must advance inside the graph)."""
    r = cases.rng(13)
    n = 1000
    x, g = r.standard_normal(n).astype(F32), r.standard_normal(n).astype(F32)
    m, v = np.zeros(n, F32), np.zeros(n, F32)
    dx, dg, dm, dv = Y.cu(x), Y.cu(g), Y.cu(m), Y.cu(v)
    for t in (1, 2, 3):
        Y.update_adam(dx, dg, dm, dv, t, lr=0.01)
        x, m, v = _adam_host(x, g, m, v, t, 0.01, 0.9, 0.999, 1e-8)
        assert np.allclose(dx.to_host(), x, rtol=2e-6, atol=1e-7) and np.allclose(dm.to_host(), m, rtol=1e-6)
    W, b, xx = r.random((4, 3)).astype(F32), r.random(4).astype(F32), r.random((3, 5)).astype(F32)
    dW, db, dxx = Y.cu(W), Y.cu(b), Y.cu(xx)
    gf = Y.CompiledGrad(Y.gradtape(cases.loss_tanh, dW, db, dxx))
    gf.set_update([0, 1], lr=0.05, optimizer="adam")
    mW, vW, mb, vb = np.zeros_like(W), np.zeros_like(W), np.zeros_like(b), np.zeros_like(b)
    for t in (1, 2, 3, 4):
        _, g_ = gf.run([dW, db, dxx])
        gW, gb = g_[0].to_host(), g_[1].to_host()
        W, mW, vW = _adam_host(W, gW, mW, vW, t, 0.05, 0.9, 0.999, 1e-8)
        b, mb, vb = _adam_host(b, gb, mb, vb, t, 0.05, 0.9, 0.999, 1e-8)
        assert np.allclose(dW.to_host(), W, rtol=5e-6, atol=1e-7), t
        assert np.allclose(db.to_host(), b, rtol=5e-6, atol=1e-7), t
    gf.set_update([], lr=1.0)
    before = dW.to_host()
    gf.run([dW, db, dxx])
    assert np.array_equal(before, dW.to_host())


def test_destination_sharded_scatter_add_is_bit_exact():
    """the scatter-add pullback sharded by destination over G ranks (emulated here
    rank by rank on one GPU; tests/dp_gpu_check.py runs it on real ranks): the shards concatenate
    to the one-GPU result bit for bit, duplicates, a hot destination, ragged splits and all."""
    from yota_b200 import dp

    r = cases.rng(17)
    rows, n_dst, n_src = 64, 1000, 20000
    V = r.standard_normal((rows, n_src)).astype(F32)
    idx = r.integers(1, n_dst + 1, size=n_src).astype(np.int64)
    idx[:500] = 7
    ref = c_oracle.scatter_add_cols(np.asfortranarray(V), idx, n_dst)
    dV, di = Y.cu(V), Y.cu(idx)
    for G in (1, 2, 3, 8):
        parts = []
        for rank in range(G):
            shard, (lo, hi) = dp.scatter_add_cols_sharded(dV, di, n_dst, rank, G)
            assert shard.dims == (rows, hi - lo)
            parts.append(shard.to_host())
        assert np.array_equal(U.bits(np.concatenate(parts, axis=1)), U.bits(ref)), G
    bad = idx.copy()
    bad[3] = n_dst + 1
    with pytest.raises(Y.YotaError, match="BoundsError"):
        dp.scatter_add_cols_sharded(dV, Y.cu(bad), n_dst, 0, 2)[0].to_host()
