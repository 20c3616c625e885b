"""tcgen05 / TMEM tensor-core GEMM (TF32 operands read from Float32 storage, FP32 accumulate)
through the C ABI (yota_gemm mode 1) and through grad() with YOTA_FLAG_GEMM_TF32.

Stated tolerance of the tensor-core mode: TF32 keeps 10 mantissa bits, so each product
carries <= 2^-10 relative error; gate = 2e-3 * (|A| |B|) element-wise bound, and gradients
of the MLP configs within 5e-3 of max|g| against the Float64 oracle.  Small-integer inputs
are exactly representable in TF32, so those results must be EXACT -- that pins every
descriptor / swizzle / layout choice of the kernel."""
import numpy as np
import pytest

import cases
import gpu_util as U
import yota_b200 as Y
from oracle import yota_oracle as O
from yota_b200 import ffi, jl

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_tf32_exact_on_small_integers(tA, tB):
    r = cases.rng(20)
    # (128 x 32 / 128 x 64 tiles for the skinny outputs, 128 x 128 and wider for the others)
    for m, n, k in [(128, 128, 32), (128, 256, 64), (384, 640, 160), (1024, 1024, 512), (1024, 128, 1024)]:
        A = r.integers(-3, 4, size=(k, m) if tA else (m, k)).astype(F32)
        B = r.integers(-3, 4, size=(n, k) if tB else (k, n)).astype(F32)
        ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
        got = U.gemm(1, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
        assert np.array_equal(got, ref), (m, n, k)
        C0 = np.asfortranarray(r.integers(-5, 6, size=(m, n)).astype(F32))
        got = U.gemm(1, tA, tB, np.asfortranarray(A), np.asfortranarray(B), C0)  # accumulate: C += A*B
        assert np.array_equal(got, ref + C0)


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0)])
def test_tf32_tolerance_on_random_data(tA, tB):
    r = cases.rng(21)
    m, n, k = 512, 768, 1024
    A = r.standard_normal((k, m) if tA else (m, k)).astype(F32)
    B = r.standard_normal((n, k) if tB else (k, n)).astype(F32)
    a, b = (A.T if tA else A).astype(np.float64), (B.T if tB else B).astype(np.float64)
    ref = a @ b
    got = U.gemm(1, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
    bound = 2e-3 * (np.abs(a) @ np.abs(b))
    assert np.all(np.abs(got - ref) <= bound), float(np.max(np.abs(got - ref) / bound))


def test_ineligible_shapes_are_refused_not_silently_rerouted():
    A = np.ones((100, 64), F32, order="F")
    with pytest.raises(Y.YotaError):
        U.gemm(1, 0, 0, A, np.ones((64, 130), F32, order="F"))


def test_mlp_gradients_in_tf32_mode():
    """C3 at a tile-aligned scale: forward, dW (NT) and dh (TN) all on tensor cores; small
    GEMMs of other configs fall through to the FFMA kernel inside the same program."""
    f, args = cases.make_c3(3, 256), cases.c3_args(3, 256, 256)
    val, g = Y.grad(f, *map(Y.cu, args), flags=ffi.FLAG_GEMM_TF32)
    v64, g64 = O.grad(f, *U.to_f64(args), dtype=np.float64)
    assert abs(val - v64) <= 5e-3 * abs(v64)
    for a, b in zip(g[1:], g64[1:]):
        a = a.to_host()
        assert np.max(np.abs(a - b)) <= 5e-3 * np.max(np.abs(b)), np.max(np.abs(a - b)) / np.max(np.abs(b))


# ---- BF16 operand mode (BF16 shadows of the Float32 arrays, FP32 accumulate in TMEM) -------
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_bf16_exact_on_small_integers(tA, tB):
    r = cases.rng(22)
    for m, n, k in [(128, 128, 64), (128, 256, 128), (384, 640, 192), (1024, 1024, 512)]:
        A = r.integers(-3, 4, size=(k, m) if tA else (m, k)).astype(F32)
        B = r.integers(-3, 4, size=(n, k) if tB else (k, n)).astype(F32)
        ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
        got = U.gemm(2, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
        assert np.array_equal(got, ref), (m, n, k)


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0)])
def test_bf16_tolerance_on_random_data(tA, tB):
    r = cases.rng(23)
    m, n, k = 512, 768, 1024
    A = r.standard_normal((k, m) if tA else (m, k)).astype(F32)
    B = r.standard_normal((n, k) if tB else (k, n)).astype(F32)
    a, b = (A.T if tA else A).astype(np.float64), (B.T if tB else B).astype(np.float64)
    ref = a @ b
    got = U.gemm(2, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
    bound = 1.2e-2 * (np.abs(a) @ np.abs(b))  # two operands rounded to 8 significant bits: 2 * 2^-9 per product
    assert np.all(np.abs(got - ref) <= bound), float(np.max(np.abs(got - ref) / bound))


def test_mlp_gradients_in_bf16_mode():
    f, args = cases.make_c3(3, 256), cases.c3_args(3, 256, 256)
    val, g = Y.grad(f, *map(Y.cu, args), flags=ffi.FLAG_GEMM_BF16)
    v64, g64 = O.grad(f, *U.to_f64(args), dtype=np.float64)
    assert abs(val - v64) <= 2e-2 * abs(v64)
    for a, b in zip(g[1:], g64[1:]):
        a = a.to_host()
        assert np.max(np.abs(a - b)) <= 3e-2 * np.max(np.abs(b)), np.max(np.abs(a - b)) / np.max(np.abs(b))


# ---- 2-CTA (cta_group::2) kernel and fused epilogues at test-sized shapes -------------------
@pytest.fixture
def force_2cta(monkeypatch):
    monkeypatch.setenv("YOTA_GEMM_FORCE_2CTA", "1")
    Y.reset()
    yield
    Y.reset()


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_2cta_exact_on_small_integers(force_2cta, mode, tA, tB):
    r = cases.rng(24)
    # (n / 256 odd: one pair per cluster; even: two pairs per cluster sharing A by TMA multicast)
    for m, n, k in [(256, 256, 64), (512, 768, 192), (1024, 512, 320), (512, 2048, 128), (256, 1024, 2048)]:
        A = r.integers(-3, 4, size=(k, m) if tA else (m, k)).astype(F32)
        B = r.integers(-3, 4, size=(n, k) if tB else (k, n)).astype(F32)
        ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
        got = U.gemm(mode, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
        assert np.array_equal(got, ref), (m, n, k)
        C0 = np.asfortranarray(r.integers(-5, 6, size=(m, n)).astype(F32))
        assert np.array_equal(U.gemm(mode, tA, tB, np.asfortranarray(A), np.asfortranarray(B), C0), ref + C0)


@pytest.mark.parametrize("flags,tol", [(ffi.FLAG_GEMM_TF32, 5e-3), (ffi.FLAG_GEMM_BF16, 3e-2)])
def test_fused_epilogues_match_unfused(force_2cta, monkeypatch, flags, tol):
    """bias+tanh and tanh'*delta+db run inside the GEMM epilogue: same elementwise arithmetic on
    the same accumulators as the separate region pass, so activations-derived gradients agree
    to FP32 summation order (db) / exactly (everything elementwise)."""
    f, args = cases.make_c3(4, 512), cases.c3_args(4, 256, 512)
    dargs = [Y.cu(a) for a in args]
    gf = Y.CompiledGrad(Y.gradtape(f, *dargs), flags=flags)
    assert "epilogue{region[add,tanh]}" in gf.describe() and "epilogue{region[square,rsubc,mul]+sum}" in gf.describe()
    v1, g1 = gf.run(dargs)
    g1 = [x.to_host() for x in g1]
    v1 = float(v1.to_host())
    monkeypatch.setenv("YOTA_NO_EPILOGUE", "1")
    gu = Y.CompiledGrad(Y.gradtape(f, *dargs), flags=flags)
    assert "epilogue" not in gu.describe()
    v2, g2 = gu.run(dargs)
    assert abs(v1 - float(v2.to_host())) <= 1e-6 * abs(v1)
    for a, b in zip(g1, g2):
        b = b.to_host()
        assert np.max(np.abs(a - b)) <= 2e-5 * np.max(np.abs(b)), np.max(np.abs(a - b)) / np.max(np.abs(b))
    v64, g64 = O.grad(f, *U.to_f64(args), dtype=np.float64)
    for a, b in zip(g1, g64[1:]):
        assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(b))


# ---- compensated 3xTF32 mode (mode 3): tensor cores at the reference's own tolerance --------
def test_tensor_core_truncates_float32_to_tf32():
    """The split of the compensated mode assumes the tensor core reads a Float32 as TF32 by
    DROPPING the 13 low mantissa bits (hi = x & 0xFFFFE000, lo = x - hi is then exact).  Probe:
    1 + 2^-11 + 2^-12 is above the rounding midpoint, so truncation gives 1, rounding 1 + 2^-10."""
    m = n = 128
    k = 32
    A = np.zeros((m, k), F32, order="F")
    A[:, 0] = F32(1.0) + F32(2.0 ** -11) + F32(2.0 ** -12)
    B = np.zeros((k, n), F32, order="F")
    B[0, :] = 1
    got = U.gemm(1, 0, 0, A, B)
    assert np.all(got == F32(1.0)), got[0, 0]
    got3 = U.gemm(3, 0, 0, A, B)  # hi + lo restores the value exactly
    assert np.all(got3 == A[0, 0]), (got3[0, 0], A[0, 0])


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_tf32x3_meets_the_fp32_tolerance(tA, tB):
    """Random data at every tile width (1-CTA 128 x 32 ... 256, and the 2-CTA kernel at the last
    shapes): the error against Float64, relative to max|result|, is that of an FP32 dot product
    (<= 1e-5, the reference's rtol; plain TF32: ~1e-3) -- also for all-positive data at K = 8192,
    where the tensor core's truncating accumulation would otherwise bias the sum by ~1e-4 (the
    kernel hands the accumulator to the CUDA cores every 256 K-elements for that reason)."""
    r = cases.rng(25)
    shapes = [(128, 128, 32), (1024, 128, 1024), (384, 640, 160), (512, 768, 1024), (2048, 2560, 512), (1024, 1280, 8192)]
    for m, n, k in shapes:
        for positive in (False, True):
            gen = (lambda sh: r.random(sh)) if positive else (lambda sh: r.standard_normal(sh))
            A = gen((k, m) if tA else (m, k)).astype(F32)
            B = gen((n, k) if tB else (k, n)).astype(F32)
            a, b = (A.T if tA else A).astype(np.float64), (B.T if tB else B).astype(np.float64)
            ref = a @ b
            scale = float(np.abs(ref).max())
            got = U.gemm(3, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
            e3 = float(np.abs(got - ref).max()) / scale
            ffma = U.gemm(0, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
            e0 = float(np.abs(ffma - ref).max()) / scale
            assert e3 <= max(1e-5, 3 * e0), (m, n, k, positive, e3, e0)
            C0 = np.asfortranarray(gen((m, n)).astype(F32))
            got = U.gemm(3, tA, tB, np.asfortranarray(A), np.asfortranarray(B), C0)
            assert float(np.abs(got - (ref + C0)).max()) / scale <= max(1e-5, 3 * e0)


def test_mlp_gradients_in_tf32x3_mode(force_2cta):
    """C3 at a tile-aligned scale through the compensated mode with fused epilogues: the FP32 gate
    of the reference (rtol 1e-5 against Float64 with the Float32 oracle's error as floor)."""
    f, args = cases.make_c3(4, 512), cases.c3_args(4, 256, 512)
    gf = Y.CompiledGrad(Y.gradtape(f, *[Y.cu(a) for a in args]), flags=ffi.FLAG_GEMM_TF32X3)
    assert "epilogue{region[add,tanh]}" in gf.describe()
    U.check_parity(f, list(args), flags=ffi.FLAG_GEMM_TF32X3)


def test_recurrence_epilogues_of_the_1cta_kernel(monkeypatch):
    """C4 at a tile-aligned mid size (H = 256, B = 128, T = 6), TF32: the per-step regions
    h = tanh.(zx .+ Wh*h .+ b) and dz = (Wh'*dz) .* (1 .- h.^2) run as epilogues of the skinny
    128 x 32-tile GEMMs, reading / writing the views of the batched containers.  Same numbers as
    with the regions as separate passes (identical arithmetic on identical accumulators), and
    within the TF32 tolerance of the Float64 oracle."""
    T, H, B = 6, 256, 128
    f, args = cases.make_c4(T), cases.c4_args(H, B, T)
    d = [Y.cu(a) for a in args]
    gf = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    desc = gf.describe()
    # (the last forward step's region also carries the loss and the start of the pullback: not an epilogue)
    assert desc.count("epilogue{region[add,add,tanh]}") == T - 2 and desc.count("epilogue{region[square,rsubc,mul]}") == T - 1, desc
    v1, g1 = gf.run(d)
    g1 = [x.to_host() for x in g1]
    monkeypatch.setenv("YOTA_NO_EPILOGUE1", "1")
    gu = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    assert "epilogue" not in gu.describe()
    v2, g2 = gu.run(d)
    assert abs(float(v1.to_host()) - float(v2.to_host())) <= 1e-6 * abs(float(v2.to_host()))
    for a, b in zip(g1, g2):
        b = b.to_host()
        assert np.max(np.abs(a - b)) <= 1e-6 * np.max(np.abs(b)), np.max(np.abs(a - b)) / np.max(np.abs(b))
    v64, g64 = O.grad(f, *U.to_f64(list(args)), dtype=np.float64)
    for a, b in zip(g1, g64[1:]):
        assert np.max(np.abs(a - b)) <= 5e-3 * np.max(np.abs(b))


@pytest.mark.parametrize("transposed", [False, True])
def test_k_split_of_the_2cta_kernel(transposed, monkeypatch):
    """Few 256 x 256 output tiles, long K (the batched dW of an unrolled recurrence): the planner
    cuts K into slices that the 2-CTA kernel schedules like tiles; the partial tiles are summed in
    slice order by a second launch.  Checked against the un-split plan (same operands, FP32 sums of
    TF32 partial products either way), against the Float64 oracle at the TF32 gate, and for
    run-to-run bit equality (fixed summation order)."""
    r = cases.rng(77)
    M, N, K = 512, 256, 8192
    Cw = r.standard_normal((M, N)).astype(np.float32)
    if transposed:  # dW = dZ * X' of the matmul rrule: both operands MN-major, K = the batch
        W = r.standard_normal((M, N)).astype(np.float32) / 16
        X = r.standard_normal((N, K)).astype(np.float32)
        f = lambda W, X: jl.sum(jl.tanh(W @ X))
        args = [W, X]
    else:
        A = r.standard_normal((M, K)).astype(np.float32) / 8
        B = r.standard_normal((K, N)).astype(np.float32) / 8
        f = lambda A, B, Cw: jl.sum((A @ B) * Cw)
        args = [A, B, Cw]
    d = [Y.cu(np.asfortranarray(a)) for a in args]
    gf = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    assert "[K split in 4]" in gf.describe(), gf.describe()
    v1, g1 = gf.run(d)
    g1 = [x.to_host() for x in g1]
    v1b, g1b = gf.run(d)
    for a, b in zip(g1, g1b):
        assert np.array_equal(U.bits(a), U.bits(b.to_host()))
    monkeypatch.setenv("YOTA_GEMM_NO_KSPLIT", "1")
    gu = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    assert "K split" not in gu.describe()
    v2, g2 = gu.run(d)
    assert abs(float(v1.to_host()) - float(v2.to_host())) <= 1e-5 * abs(float(v2.to_host())) + 1e-3
    v64, g64 = O.grad(f, *U.to_f64(list(args)), dtype=np.float64)
    assert abs(float(v1.to_host()) - v64) <= 5e-3 * np.sqrt(M * N) * 4
    for a, b, b64 in zip(g1, g2, g64[1:]):
        b = b.to_host()
        # (not 1e-5: the tensor core's accumulator truncates, a bias that grows with the number of
        # accumulations per slice -- 1.7e-4 of max|g| at K = 8192 x 3 products)
        assert np.max(np.abs(a - b)) <= 3e-4 * np.max(np.abs(b))
        assert np.max(np.abs(a - b64)) <= 5e-3 * np.max(np.abs(b64))


@pytest.mark.parametrize("T,H,B", [(6, 256, 128), (9, 512, 128)])
def test_recurrence_as_one_persistent_launch(T, H, B, monkeypatch):
    """The per-step GEMM + epilogue launches of an unrolled recurrence run as ONE persistent kernel
    (weights resident in shared memory, steps ordered by per-tile flags).  Same arithmetic on the
    same accumulators as the step-by-step launches: bit-identical results, and fewer launches."""
    f, args = cases.make_c4(T), cases.c4_args(H, B, T)
    d = [Y.cu(a) for a in args]
    gf = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    desc = gf.describe()
    assert desc.count("one persistent launch]") == 2, desc
    v1, g1 = gf.run(d)
    g1 = [x.to_host() for x in g1]
    n1 = gf.stats()["launches"]
    v1b, g1b = gf.run(d)  # replay: the flags are reset inside the run
    for a, b in zip(g1, g1b):
        assert np.array_equal(U.bits(a), U.bits(b.to_host()))
    monkeypatch.setenv("YOTA_NO_CHAIN", "1")
    gu = Y.CompiledGrad(Y.gradtape(f, *d), flags=ffi.FLAG_GEMM_TF32)
    assert "persistent launch" not in gu.describe()
    v2, g2 = gu.run(d)
    assert gu.stats()["launches"] >= n1 + (T - 3) + (T - 2) - 2, (n1, gu.stats())
    assert U.bits(np.float32(v1.to_host())) == U.bits(np.float32(v2.to_host()))
    for a, b in zip(g1, g2):
        assert np.array_equal(U.bits(a), U.bits(b.to_host()))
    v64, g64 = O.grad(f, *U.to_f64(list(args)), dtype=np.float64)
    for a, b in zip(g1, g64[1:]):
        assert np.max(np.abs(a - b)) <= 5e-3 * np.max(np.abs(b))
