"""CPU tests of the host side: tape form (vs the worked tape in SURVEY §8(a)), the tape ->
op-program lowering (checked against the oracle through a test-only NumPy interpreter of the
op program), the planner (fusion decisions) and the C ABI surface.  No device work."""
import ctypes as C
import re

import numpy as np
import pytest

import cases
import ir_interp
import yota_b200 as Y
from oracle import yota_oracle as O
from yota_b200 import ffi, ir, jl
from yota_b200.grad import lower
from yota_b200.tape import Call

A = Y.AVal


def _lowered_vs_oracle(f, args, seed=1, rtol=1e-11):
    tape = Y.gradtape(f, *args, seed=seed)
    L = lower(tape)
    ins = [None] * L.prog.n_inputs
    for a, s in zip(args, L.arg_slots):
        ins[s] = a
    for s, a in L.captured:
        ins[s] = a
    if L.seed_slot is not None:
        ins[L.seed_slot] = np.asarray(seed, dtype=np.float64)
    outs = ir_interp.run(L.prog, ins, seed=float(seed) if isinstance(seed, (int, float)) else 1.0, dtype=np.float64)
    val, g = O.grad(f, *args, seed=seed, dtype=np.float64)
    assert np.allclose(outs[L.val[1]], val, rtol=rtol)
    for spec, gg in zip(L.grads, g[1:]):
        if isinstance(spec, tuple):
            assert outs[spec[1]].shape == np.shape(gg)
            assert np.allclose(outs[spec[1]], gg, rtol=rtol, atol=1e-13)
        else:
            assert repr(spec) == repr(gg)
    return L


def test_tape_has_reference_form():
    """sum(tanh.(W*x .+ b)) -- statement for statement the worked tape of SURVEY §8(a)."""
    tape = Y.gradtape(cases.loss_tanh, A((4, 3)), A((4,)), A((3,)))
    fns = [(op.fn if isinstance(op.fn, str) else "pb") for op in tape.ops if isinstance(op, Call)]
    assert fns == ["rrule", "_getfield", "_getfield"] * 5 + ["pb", "getfield"] * 3 + ["pb", "getfield", "getfield"] * 2 \
        + ["tuple", "map", "tuple"]
    rr = [op.args[1:3] for op in tape.ops if isinstance(op, Call) and op.fn == "rrule"]
    assert [a[0] if a[0] != "broadcasted" else a[1] for a in rr] == ["matmul", "add", "tanh", "materialize", "sum"]
    # getfield indices count f as position 1 (src/grad.jl:176,182-186)
    gf = [op.args[1] for op in tape.ops if isinstance(op, Call) and op.fn == "getfield"]
    assert gf == [2, 2, 3, 3, 4, 2, 3]


def test_accumulation_ops_on_tape():
    """multipath: y used twice -> one broadcast(+, dx, old_dx) per extra use (src/grad.jl:91)."""
    f = lambda x: jl.sum(jl.tanh(x) * x)  # noqa: E731
    tape = Y.gradtape(f, A((3, 4)))
    adds = [op for op in tape.ops if isinstance(op, Call) and op.fn == "broadcast"]
    assert len(adds) == 1 and adds[0].args[0] == "+"


@pytest.mark.parametrize("name,f,args", cases.gradcheck_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_lowering_matches_oracle(name, f, args):
    _lowered_vs_oracle(f, args)


def test_lowering_configs_small():
    f64 = lambda xs: [x.astype(np.float64) if x.dtype.kind == "f" else x for x in xs]  # noqa: E731
    _lowered_vs_oracle(cases.c1_mlp, f64(cases.c1_args(6, 5, 3, 4)))
    _lowered_vs_oracle(cases.c2_chain, f64(cases.c2_args(4, 5)))
    _lowered_vs_oracle(cases.make_c3(3, 4), f64(cases.c3_args(3, 5, 4)))
    L = _lowered_vs_oracle(cases.make_c4(3), f64(cases.c4_args(4, 3, 3)))
    assert sum(o.opcode == ir.EW_ADD for o in L.prog.ops) >= 3 * 2  # accumulation adds survive lowering
    L = _lowered_vs_oracle(cases.c5_embed, f64(cases.c5_args(3, 5, 8)))
    assert L.grads[1] is Y.NoTangent  # index argument


def test_lowering_seed_forms():
    x = np.array([1.0, 2, 3])
    _lowered_vs_oracle(lambda x: 2 * x, (x,), seed=np.ones(3))
    _lowered_vs_oracle(lambda x: 2 * x, (x,), seed="auto")
    _lowered_vs_oracle(cases.loss_tanh, (np.ones((4, 3)), np.ones(4), np.ones(3)), seed=2.5)
    with pytest.raises(ValueError):
        Y.gradtape(lambda x: 2 * x, A((3,)))


def test_trivial_grads_need_no_device():
    _, g = Y.grad(lambda x, y: x, 42.0, 54.0)
    assert g == (Y.ZeroTangent, 1, Y.ZeroTangent)
    _, g = Y.grad(lambda x: 1, 42.0)
    assert g == (Y.ZeroTangent, Y.ZeroTangent)


def test_no_rule_error():
    with pytest.raises(Y.NoRuleError):
        Y.gradtape(lambda x: jl.logsoftmax(x, dims=2), A((3, 4)))


# ---- the C ABI -------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(built):
    hdr = open(ffi.LIB_PATH.replace("yota.jl_b200/lib/libyota_cuda.so", "include/yota_cuda.h")).read()
    declared = set(re.findall(r"\b(yota_[a-z0-9_]+)\s*\(", hdr)) - {"yota_ctx", "yota_program"}
    assert declared == set(ffi.SYMBOLS), declared ^ set(ffi.SYMBOLS)
    for name in declared:
        assert getattr(built, name) is not None
    assert built.yota_abi_version() == 1


def _plan(f, *avals, flags=0, seed=1):
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    L = lower(Y.gradtape(f, *avals, seed=seed))
    ops, tens = ffi.pack_program(L.prog)
    h = C.c_void_p()
    ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
    n = lib.yota_program_describe(h, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.yota_program_describe(h, buf, n + 1)
    st = [C.c_int64() for _ in range(5)]
    ffi.check(lib.yota_program_stats(h, *[C.byref(s) for s in st]))
    lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)
    return buf.value.decode(), [s.value for s in st]


def test_planner_fuses_c2_into_one_pass(built):
    """C2 forward + pullback = ONE region: 2 full reads (x, W), 2 full writes (dx, dW), the
    loss and db as fused sinks -- the 4-pass contract of SURVEY §8(d)."""
    desc, (steps, launches, ws, regions, _) = _plan(cases.c2_chain, A((8192, 8192)), A((8192, 8192)), A((8192,)))
    assert regions == 1 and steps == 3, desc
    assert "+sum+sum" in desc and "scalar_finish" in desc and "rowsum_finish" in desc
    assert ws < 16 << 20  # no 256 MiB temporaries
    unf, (steps_u, *_rest) = _plan(cases.c2_chain, A((8192, 8192)), A((8192, 8192)), A((8192,)), flags=ffi.FLAG_NO_FUSE)
    assert steps_u > 10


def test_planner_never_materialises_pullback_temporaries(built):
    desc, st = _plan(cases.make_c3(2, 64), *[A(a.shape) for a in cases.c3_args(2, 32, 64)])
    assert "region[exp,mul,sub]+sum" in desc, desc  # exp.(logp) is recomputed inside the dz pass, not stored


def test_planner_c3_epilogues_and_k_major_shadows(built):
    """C3 in TF32 mode: bias+tanh and tanh'*delta+db run as GEMM epilogues (their region steps are
    absorbed), and those epilogues write the transposed copies the dW GEMMs read K-major; in BF16
    mode they write the BF16 operand shadows."""
    av = [A((4096, 4096))] * 8 + [A((4096,))] * 8 + [A((4096, 8192))] * 2
    desc, (steps, launches, ws, regions, _) = _plan(cases.make_c3(8, 8192), *av, flags=ffi.FLAG_GEMM_TF32)
    assert desc.count("matmul + epilogue{region[add,tanh]}") == 7 and desc.count("+sum} +finish ->") == 7, desc
    assert desc.count("(absorbed into step") == 15 and desc.count("(rowsum finished inside step") == 7
    assert desc.count("[K-major AB]") == 6 and desc.count("[K-major B]") == 1 and desc.count("[K-major A]") == 1, desc
    assert launches == steps - 22  # absorbed regions and in-kernel finishers launch nothing
    desc16, _ = _plan(cases.make_c3(8, 8192), *av, flags=ffi.FLAG_GEMM_BF16)
    assert desc16.count("+bf16") == 14 and "K-major" not in desc16, desc16
    plain, (steps_p, *_r) = _plan(cases.make_c3(8, 8192), *av, flags=0)  # strict FP32: no tensor-core fusions
    assert "epilogue" not in plain and steps_p == steps


def test_planner_c5_gathers_in_the_region_and_scales_in_the_fold(built):
    """C5: E[:, idx] is never materialised (the region gathers), nor is seed .* V (the fold scales
    on load): one fused pass + the scatter-add, and an arena of partials only."""
    av = [A((256, 1 << 20)), A((1 << 22,), "i64"), A((256, 1 << 22))]
    desc, (steps, launches, ws, regions, _) = _plan(cases.c5_embed, *av)
    assert "(gathered inside step" in desc and " gather " in desc and regions == 1, desc
    assert "region[mul,copy,mul]+sum" in desc  # dG = seed .* V is gone from the region
    assert ws < 128 << 20, ws  # was 2 x 4 GiB of temporaries
    # the index vector is a program input: its sort leaves the step and runs beside the forward pass
    assert desc.count("index sort on the side stream") == 1, desc
    unf, (_s, _l, ws_u, *_r) = _plan(cases.c5_embed, *av, flags=ffi.FLAG_NO_FUSE)
    assert ": getindex ->" in unf and ws_u > 8 << 30


def test_k_split_only_where_tiles_are_few(built):
    """The K split of the 2-CTA kernel is for contractions with few 256 x 256 output tiles and a long
    K; the C3 shapes -- whole batch or an eighth of it (8 GPUs) -- have 64-256 tiles and stay whole."""
    for batch in (8192, 1024):
        av = [A((4096, 4096))] * 8 + [A((4096,))] * 8 + [A((4096, batch))] * 2
        desc, _ = _plan(cases.make_c3(8, batch), *av, flags=ffi.FLAG_GEMM_TF32)
        assert "K split" not in desc and "persistent launch" not in desc, batch
    f = lambda W, X: jl.sum(jl.tanh(W @ X))  # dW = dZ * X': 512 x 256 x 8192 -> 2 tiles
    desc, _ = _plan(f, A((512, 256)), A((256, 8192)), flags=ffi.FLAG_GEMM_TF32)
    assert desc.count("[K split in 4]") == 1, desc
    assert "K split" not in _plan(f, A((512, 256)), A((256, 8192)), flags=0)[0]  # strict FP32: FFMA kernel


def test_planner_forks_small_leaf_steps(built, monkeypatch):
    """C1 is twelve tiny kernels in a row only because a stream is a row: the loss, db2, dW2, db1, dW1
    are leaves of the dataflow and are forked onto side streams (parallel branches of the graph);
    what they read stays allocated until the join.  Big GEMMs (C3) stay in line."""
    av = [A((128, 784)), A((128,)), A((10, 128)), A((10,)), A((784, 64)), A((10, 64))]
    desc, (steps, launches, ws, *_r) = _plan(cases.c1_mlp, *av)
    forked = [l.split(":")[0].strip() for l in desc.splitlines() if "(leaf: on a side stream)" in l]
    assert forked == ["step 4", "step 5", "step 6", "step 9", "step 10"], desc
    monkeypatch.setenv("YOTA_NO_FORK", "1")
    desc0, (_s, _l, ws0, *_r0) = _plan(cases.c1_mlp, *av)
    assert "side stream" not in desc0 and ws0 <= ws  # longer lifetimes cost a little arena
    monkeypatch.delenv("YOTA_NO_FORK")
    av3 = [A((4096, 4096))] * 8 + [A((4096,))] * 8 + [A((4096, 8192))] * 2
    desc3, _ = _plan(cases.make_c3(8, 8192), *av3, flags=ffi.FLAG_GEMM_TF32)
    assert all("matmul" not in l for l in desc3.splitlines() if "side stream" in l), desc3


def test_planner_c4_slices_are_views_and_accumulation_is_in_place(built, monkeypatch):
    av = [A((1024, 1024)), A((1024, 1024)), A((1024,)), A((1024, 128, 16))]
    desc, (steps, *_r) = _plan(cases.make_c4(16), *av)
    assert ": getindex ->" not in desc, "x[:, :, t] must be a view, not a copy"
    assert "unrolled loops: 5 " in desc and steps <= 4 * 16 + 8, desc  # batched: 2 GEMMs + 2 regions per step remain
    # full size: the two batched dW contractions (1024 x 1024 x 16384: 16 tiles for 74 CTA pairs) are cut along K
    av_full = [A((1024, 1024)), A((1024, 1024)), A((1024,)), A((1024, 128, 128))]
    desc_full, _ = _plan(cases.make_c4(128), *av_full, flags=ffi.FLAG_GEMM_TF32)
    assert desc_full.count("[K split in 4]") == 2, [l for l in desc_full.splitlines() if "K split" in l]
    # ... and the forward and backward recurrences (126 + 127 GEMM + epilogue steps) are one persistent launch each
    marks = [l for l in desc_full.splitlines() if "one persistent launch]" in l]
    assert len(marks) == 2 and "recurrence of 126 steps" in marks[0] and "recurrence of 127 steps" in marks[1], marks
    monkeypatch.setenv("YOTA_NO_CHAIN", "1")
    assert "persistent launch" not in _plan(cases.make_c4(128), *av_full, flags=ffi.FLAG_GEMM_TF32)[0]
    monkeypatch.delenv("YOTA_NO_CHAIN")
    # strict FP32 / BF16: no tcgen05 TF32 recurrence kernel
    assert "persistent launch" not in _plan(cases.make_c4(128), *av_full, flags=0)[0]
    # BF16 / 3xTF32 keep whole-array operand shadows: the loop stays unrolled there and the tape's
    # accumulation is done in place (GEMM C += ..., slice +=) -- also the plan with the pass disabled
    for flags, env in ((ffi.FLAG_GEMM_BF16, None), (0, "1")):
        if env:
            monkeypatch.setenv("YOTA_NO_LOOP_BATCHING", env)
        desc, _ = _plan(cases.make_c4(16), *av, flags=flags)
        assert "unrolled loops" not in desc and ": getindex ->" not in desc
        assert desc.count("(+= in place)") >= 3 * 15 - 2 and "slice += in place" in desc, desc


def alias_hazard_loss(X3, W, h):
    """A matmul thunk added to a VIEW of a saved activation: accumulating in place would overwrite
    H[:, :, 2], which the tanh pullback (1 - H^2) and sum(H) read later."""
    H = jl.tanh(X3)
    s = jl.sum(H)
    z = (W @ h) + H[jl.colon, jl.colon, 2]
    return jl.sum(z * z) + s


def test_planner_never_accumulates_into_a_view_of_a_shared_array(built):
    av = [A((8, 4, 3)), A((8, 8)), A((8, 4))]
    desc, _ = _plan(alias_hazard_loss, *av)
    assert "(+= in place)" not in desc, desc
    r = cases.rng(3)
    _lowered_vs_oracle(alias_hazard_loss, (r.random((8, 4, 3)), r.random((8, 8)), r.random((8, 4))))


def test_build_rejects_unsupported_and_bad_programs(built):
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    p = ir.IRProgram()
    x = p.tensor((4, 4), kind=ir.T_INPUT, slot=0)
    y = p.op(999, (x,), (4, 4))
    p.mark_output(y, 0)
    ops, tens = ffi.pack_program(p)
    h = C.c_void_p()
    rc = lib.yota_program_build(ctx, ops, len(p.ops), tens, len(p.tensors), 0, C.byref(h))
    assert rc == -4 and b"no CPU fallback" in lib.yota_last_error()
    p = ir.IRProgram()
    a = p.tensor((4, 3), kind=ir.T_INPUT, slot=0)
    b = p.tensor((5,), kind=ir.T_INPUT, slot=1)
    p.mark_output(p.op(ir.EW_ADD, (a, b), (4, 3)), 0)
    ops, tens = ffi.pack_program(p)
    rc = lib.yota_program_build(ctx, ops, len(p.ops), tens, len(p.tensors), 0, C.byref(h))
    assert rc == -3 and b"DimensionMismatch" in lib.yota_last_error()
    lib.yota_shutdown(ctx)


def test_compute_entry_fails_loudly_without_gpu(built):
    import shutil

    if shutil.which("nvidia-smi"):
        pytest.skip("a GPU is present")
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    p = C.c_void_p()
    assert lib.yota_alloc(ctx, 1024, C.byref(p)) == -1
    assert b"no CPU fallback" in lib.yota_last_error()
    with pytest.raises(Y.YotaError):
        Y.grad(cases.loss_tanh, np.ones((4, 3), np.float32), np.ones(4, np.float32), np.ones(3, np.float32))


# ---- eager.play on the CPU: the statement-by-statement executor with its device programs
# replaced by the test-only NumPy interpreter of the op program (host logic only) -------------
class _FakeArr:
    def __init__(self, a):
        self.a = np.asarray(a) if np.ndim(a) == 0 else np.asfortranarray(a)
        self.dims = tuple(self.a.shape)
        self.dtype = "i64" if self.a.dtype.kind in "iu" else "f32"
        self.ptr = None

    def to_host(self):
        return self.a if self.dims else self.a[()]


class _FakeProgram:
    def __init__(self, prog, ctx, flags):
        self.prog = prog

    def __call__(self, ins, seed=1.0):
        outs = ir_interp.run(self.prog, [x.a for x in ins], seed=seed, dtype=np.float64)
        return [_FakeArr(np.asarray(outs[s])) for s in range(self.prog.n_outputs)]


@pytest.fixture
def fake_device(monkeypatch):
    from yota_b200 import eager

    monkeypatch.setattr(eager, "DeviceProgram", _FakeProgram)
    monkeypatch.setattr(eager, "B200Array", _FakeArr)
    monkeypatch.setattr(eager, "context", lambda: None)
    _FakeArr.from_host = staticmethod(lambda a, ctx=None: _FakeArr(a))
    eager._PROGRAMS.clear()
    yield eager
    eager._PROGRAMS.clear()


@pytest.mark.parametrize("name,f,args", cases.gradcheck_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_play_statement_by_statement_matches_the_oracle(fake_device, name, f, args):
    tape = Y.gradtape(f, *[A(a.shape) for a in args])
    val, g = fake_device.play(tape, *[_FakeArr(a) for a in args])
    oval, og = O.grad(f, *args, dtype=np.float64)
    v = val.to_host() if isinstance(val, _FakeArr) else val
    assert np.allclose(v, oval, rtol=1e-11)
    for a, b in zip(g[1:], og[1:]):
        if isinstance(a, _FakeArr):
            assert np.allclose(a.to_host(), b, rtol=1e-11, atol=1e-13)
        else:
            assert repr(a) == repr(b)


def test_play_configs_and_seeds(fake_device):
    f64 = lambda xs: [x.astype(np.float64) if x.dtype.kind == "f" else x for x in xs]  # noqa: E731
    for f, args in ((cases.c1_mlp, f64(cases.c1_args(6, 5, 3, 4))), (cases.make_c3(3, 4), f64(cases.c3_args(3, 5, 4))),
                    (cases.make_c4(3), f64(cases.c4_args(4, 3, 3))), (cases.c5_embed, f64(cases.c5_args(3, 5, 8)))):
        tape = Y.gradtape(f, *[A(a.shape, "i64" if a.dtype.kind == "i" else "f32") for a in args])
        val, g = fake_device.play(tape, *[_FakeArr(a) for a in args])
        oval, og = O.grad(f, *args, dtype=np.float64)
        assert np.allclose(val.to_host() if isinstance(val, _FakeArr) else val, oval, rtol=1e-11)
        for a, b in zip(g[1:], og[1:]):
            if isinstance(a, _FakeArr):
                assert np.allclose(a.to_host(), b, rtol=1e-11, atol=1e-13)
            else:
                assert repr(a) == repr(b)
    x = np.array([1.0, 2, 3])
    tape = Y.gradtape(lambda x: 2 * x, A((3,)), seed="auto")
    _, g = fake_device.play(tape, _FakeArr(x), seed="auto")
    assert np.allclose(g[1].to_host(), 2 * np.ones(3))


# ---- composite layers (conv / maxpool / logitcrossentropy), reshape, sum(f, x) ----------------
def test_conv_composite_is_nnlib_conv():
    """The gather + GEMM + gather composite against an independent reference: scipy's true
    convolution per (input channel, output channel) pair -- NNlib.conv flips the kernel."""
    from scipy.signal import convolve

    r = cases.rng(3)
    x, w = r.standard_normal((7, 6, 3, 2)), r.standard_normal((3, 2, 3, 4))
    val, g = O.grad(lambda x, w: jl.sum(jl.conv(x, w) ** 2), x, w, dtype=np.float64)
    y = np.zeros((5, 5, 4, 2))
    for n in range(2):
        for co in range(4):
            y[:, :, co, n] = sum(convolve(x[:, :, ci, n], w[:, :, ci, co], mode="valid") for ci in range(3))
    assert np.isclose(val, (y ** 2).sum(), rtol=1e-12)
    # dw against the definition: dw[kw,kh,ci,co] = sum 2 y[ow,oh,co,n] x[ow+KW-1-kw, oh+KH-1-kh, ci, n]
    dw = np.zeros_like(w)
    for kw in range(3):
        for kh in range(2):
            xs = x[2 - kw:2 - kw + 5, 1 - kh:1 - kh + 5]
            dw[kw, kh] = np.einsum("abin,abon->io", xs, 2 * y)
    assert np.allclose(g[2], dw, rtol=1e-11)


def test_maxpool_logitcrossentropy_reshape_sum_of_f_on_the_oracle_and_lowered():
    r = cases.rng(4)
    xm = r.standard_normal((4, 6, 2, 3))
    val, _ = O.grad(lambda x: jl.sum(jl.maxpool2(x) ** 2), xm, dtype=np.float64)
    assert np.isclose(val, (xm.reshape(2, 2, 2, 3, 2, 3, order="F").max(axis=(0, 2)) ** 2).sum(), rtol=1e-12)
    _lowered_vs_oracle(lambda x: jl.sum(jl.maxpool2(x) ** 2), (xm,))
    _lowered_vs_oracle(cases.convnet_loss, cases.convnet_args())
    a = cases.c1_args(6, 5, 4, 7)
    y = np.eye(4)[:, cases.rng(5).integers(0, 4, 7)]
    _lowered_vs_oracle(cases.mlp_relu_loss, [x.astype(np.float64) for x in a[:5]] + [y])
    _lowered_vs_oracle(cases.sum_of_f, (r.standard_normal((3, 4)),))
    _lowered_vs_oracle(lambda x: jl.sum(jl.reshape(x, 6, 2) @ jl.reshape(x, 2, 6)), (r.standard_normal((3, 4)),))
    with pytest.raises(ValueError):
        Y.gradtape(lambda x: jl.sum(jl.reshape(x, 5, 2)), A((3, 4)))


# ---- unrolled-loop batching (C4): the op list AFTER the planner's rewrite, executed by the test
# interpreter, against the oracle ---------------------------------------------------------------
def _planned_ops(f, avals, flags=0):
    """(IRProgram of the rewritten op list, Lowered, describe text)."""
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    L = lower(Y.gradtape(f, *avals))
    ops, tens = ffi.pack_program(L.prog)
    h = C.c_void_p()
    ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
    n = lib.yota_program_dump_ops(h, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.yota_program_dump_ops(h, buf, n + 1)
    nd = lib.yota_program_describe(h, None, 0)
    dbuf = C.create_string_buffer(nd + 1)
    lib.yota_program_describe(h, dbuf, nd + 1)
    lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)
    prog = ir.IRProgram()
    for line in buf.value.decode().splitlines():
        if line.startswith("T "):
            _, tid, dtype, ndim, d0, d1, d2, d3, kind, slot, cval = line.split()
            prog.tensors.append(ir.IRTensor(int(tid), int(dtype), tuple(int(x) for x in (d0, d1, d2, d3))[:int(ndim)], int(kind),
                                            int(slot), float(cval)))
        else:
            head, ia, fa, mem = line[2:].split("|")
            hs = [int(x) for x in head.split()]
            op = ir.IROp(hs[0], tuple(hs[3:3 + hs[2]]), hs[1], tuple(int(x) for x in ia.split()), tuple(float(x) for x in fa.split()))
            op.members = [int(x) for x in mem.split()]
            prog.ops.append(op)
    prog.n_inputs, prog.n_outputs = L.prog.n_inputs, L.prog.n_outputs
    return prog, L, dbuf.value.decode()


@pytest.mark.parametrize("T,H,B", [(4, 8, 4), (7, 16, 8)])
def test_unrolled_rnn_is_batched_and_still_the_same_function(built, T, H, B):
    """C4: the T products Wx*x[:,:,t] become one GEMM, the accumulation chains of dWx, dWh, db and dx
    one contraction each over the concatenated dz_t (placed side by side, no copies); the rewritten
    op list computes what the tape computes (Float64, against the oracle)."""
    f = cases.make_c4(T)
    args = [a.astype(np.float64) for a in cases.c4_args(H, B, T)]
    prog, L, desc = _planned_ops(f, [A(a.shape) for a in args])
    assert "unrolled loops: 5 " in desc, desc
    n_mm = sum(o.opcode == ir.OP_MATMUL for o in prog.ops)
    outs = ir_interp.run(prog, args, dtype=np.float64)
    val, g = O.grad(f, *args, dtype=np.float64)
    assert np.allclose(outs[L.val[1]], val, rtol=1e-11)
    for spec, gg in zip(L.grads, g[1:]):
        assert np.allclose(outs[spec[1]], gg, rtol=1e-9, atol=1e-12), spec
    # the serial part that remains: Wh*h_{t-1} and Wh'*dz_t per step, plus the batched ones
    steps = [l for l in desc.splitlines() if l.strip().startswith("step")]
    assert sum(": matmul" in l for l in steps) == 2 * (T - 1) + 4, desc
    assert "in place" not in desc


def test_loop_batching_leaves_other_programs_alone(built):
    desc, _ = _plan(cases.make_c3(3, 64), *[A(a.shape) for a in cases.c3_args(3, 32, 64)])
    assert "unrolled loops" not in desc
    # an index-VECTOR scatter chain is never batched (its pullback is bit-exact, src/helpers.jl:7-11)
    desc, _ = _plan(lambda E, i1, i2, i3, V: jl.sum((E[colon_, i1] + E[colon_, i2] + E[colon_, i3]) * V),
                    A((8, 16)), A((12,), "i64"), A((12,), "i64"), A((12,), "i64"), A((8, 12)))
    assert "unrolled loops" not in desc and desc.count("ungetindex") == 3, desc


colon_ = jl.colon


def test_allreduce_marks_are_validated_and_destination_shards_cover_the_table(built):
    """Host logic of the multi-GPU paths that needs no device: a marked slot must be written by
    some step; dest_shard partitions 0..n exactly (ragged tails to the first ranks)."""
    from yota_b200 import dp

    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    L = lower(Y.gradtape(cases.loss_tanh, A((4, 3)), A((4,)), A((3,))))
    ops, tens = ffi.pack_program(L.prog)
    h = C.c_void_p()
    ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), 0, C.byref(h)))
    good = (C.c_int32 * 2)(L.grads[0][1], L.grads[1][1])
    assert lib.yota_program_set_allreduce(h, good, 2) == 0
    bad = (C.c_int32 * 1)(99)
    assert lib.yota_program_set_allreduce(h, bad, 1) == -6 and b"bad out slot" in lib.yota_last_error()
    lib.yota_program_destroy(h)
    lib.yota_shutdown(ctx)
    for n_dst, G in ((10, 3), (1 << 20, 8), (7, 8)):
        edges = [dp.dest_shard(n_dst, r, G) for r in range(G)]
        assert edges[0][0] == 0 and edges[-1][1] == n_dst
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
