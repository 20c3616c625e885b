"""Parity at BASELINE.json's FULL sizes, through the C ABI.

Where the oracle finishes in seconds it is compared directly (C2 closed form in Float64, C4,
C5's plain-C sequential fold); for the 6.6 TFLOP C3 eval the checks are size-independent
properties of the reference path: batch linearity (the identity data parallelism rests on),
the finite-difference directional derivative at the reference's own gradcheck form
(src/gradcheck.jl:3-23), softmax mass conservation, and run-to-run determinism.
"""
import numpy as np
import pytest

import cases
import gpu_util as U
import yota_b200 as Y
from oracle import c_oracle
from oracle import yota_oracle as O
from yota_b200 import ffi

pytestmark = pytest.mark.gpu
F32 = np.float32


def test_c2_full_size_closed_form_and_seed_linearity():
    """sum(tanh.(x .* W .+ b) .^ 2) at 8192 x 8192: every output against the Float64 closed
    form, and grad(seed=2) == 2 * grad(seed=1) bitwise (scaling by 2 is exact)."""
    n = 8192
    x, W, b = cases.c2_args(n)
    dx_, dW_, db_ = Y.cu(x), Y.cu(W), Y.cu(b)
    val, g = Y.grad(cases.c2_chain, dx_, dW_, db_)
    val2, g2 = Y.grad(cases.c2_chain, dx_, dW_, db_, seed=2.0)
    gx, gW, gb = (a.to_host() for a in g[1:4])
    for a, a2 in zip((gx, gW, gb), g2[1:4]):
        assert np.array_equal(U.bits(2 * a), U.bits(a2.to_host()))
    loss, db_ref = 0.0, np.zeros(n)
    worst = 0.0
    for c0 in range(0, n, 1024):  # Float64 closed form, 1024 columns at a time
        xs, Ws = x[:, c0:c0 + 1024].astype(np.float64), W[:, c0:c0 + 1024].astype(np.float64)
        t = np.tanh(xs * Ws + b.astype(np.float64)[:, None])
        gz = 2 * t * (1 - t * t)
        loss += float((t * t).sum())
        db_ref += gz.sum(axis=1)
        for got, ref in ((gx[:, c0:c0 + 1024], gz * Ws), (gW[:, c0:c0 + 1024], gz * xs)):
            worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    assert worst <= 1e-5, worst  # elementwise outputs: the rtol the reference's gradcheck uses
    assert abs(float(val) - loss) <= 1e-5 * loss
    assert np.abs(gb - db_ref).max() <= 1e-5 * np.abs(db_ref).max() + 1e-4  # 8192-term FP32 row sums


def test_c5_full_size_bit_exact():
    """E 256 x 1 048 576, 4 194 304 Int64 indices: dE bitwise equal to the plain-C sequential
    fold (src/helpers.jl:7-11 -> NNlib.scatter!), dV == E[:, idx] exactly on sampled columns."""
    rows, nt, ni = 256, 1 << 20, 1 << 22
    r = cases.rng(55)
    E = np.asfortranarray(r.random((nt, rows), dtype=F32).T - F32(0.5))
    V = np.asfortranarray(r.random((ni, rows), dtype=F32).T - F32(0.5))
    idx = r.integers(1, nt + 1, size=ni).astype(np.int64)
    idx[:1000] = 7  # a hot destination: a 1000-term fold in source order
    dE_, di_, dV_ = Y.cu(E), Y.cu(idx), Y.cu(V)
    val, g = Y.grad(cases.c5_embed, dE_, di_, dV_)
    got = g[1].to_host()
    ref = c_oracle.scatter_add_cols(V, idx, nt)
    assert np.array_equal(U.bits(got), U.bits(ref))
    del got, ref
    gv = g[3].to_host()
    cols = r.integers(0, ni, size=4096)
    assert np.array_equal(U.bits(gv[:, cols]), U.bits(E[:, idx[cols] - 1]))
    ref_val = c_oracle.gather_dot(E, V, idx)
    assert abs(float(val) - ref_val) <= 1e-5 * abs(ref_val) + 0.5  # 2^30 products of magnitude <= 0.25


@pytest.fixture(scope="module")
def c4_oracle():
    args = cases.c4_args(1024, 128, 128)
    f = cases.make_c4(128)
    return args, f, O.grad(f, *args, dtype=np.float32), O.grad(f, *U.to_f64(list(args)), dtype=np.float64)


@pytest.mark.parametrize("mode", ["fp32_batched", "fp32_unrolled", "tf32_batched"])
def test_c4_full_size_vs_oracle(mode, c4_oracle, monkeypatch):
    """Unrolled Elman RNN, H = 1024, T = 128, B = 128: the 127-fold accumulation of Wx, Wh, b and
    the slice-wise dx against the oracle -- with the loop batched by the planner (one K = T*B
    contraction per weight gradient, the default), left unrolled (in-place accumulation in tape
    order), and batched with TF32 contractions (stated tolerance 2e-2 of max|g|: 128 recurrent steps
    compound the 2^-11 operand rounding)."""
    args, f, o32, o64 = c4_oracle
    if mode == "fp32_unrolled":
        monkeypatch.setenv("YOTA_NO_LOOP_BATCHING", "1")
    Y.reset()
    if mode.startswith("fp32"):
        U.check_parity(f, list(args), flags=0, o32=o32, o64=o64)
    else:
        val, g = Y.grad(f, *[Y.cu(a) for a in args], flags=ffi.FLAG_GEMM_TF32)
        assert abs(val - float(o64[0])) <= 2e-2 * abs(float(o64[0]))
        worst = 0.0
        for a, b in zip(g[1:], o64[1][1:]):
            worst = max(worst, float(np.abs(a.to_host() - b).max() / np.abs(b).max()))
        print(f"C4 full size, tf32 batched: worst err / max|g| = {worst:.3e}")
        assert worst <= 2e-2, worst
    Y.reset()


@pytest.fixture(scope="module")
def c3_oracle():
    """The restated reference path on the FULL C3 config (8 x 4096-wide, batch 8192), once per
    session, in Float32 (the reference's own arithmetic) and Float64 (ground truth): ~1-3 min of
    host BLAS.  Only (val, grads) are kept."""
    L, W, B = 8, 4096, 8192
    args = cases.c3_args(L, W, B)
    f = cases.make_c3(L, B)
    v32, g32 = O.grad(f, *args, dtype=np.float32)
    g32 = [np.asarray(g, dtype=np.float32) for g in g32[1:]]
    v64, g64 = O.grad(f, *U.to_f64(list(args)), dtype=np.float64)
    g64 = [np.asarray(g, dtype=np.float64) for g in g64[1:]]
    return args, f, (float(v32), g32), (float(v64), g64)


# mode -> (flags, how it is gated).  "gate": the reference's rtol 1e-5 (against the
# Float64 truth with the Float32 oracle's own error as the floor); numbers: tolerance relative to
# max|g| per tensor, the stated looser tolerance of the tensor-core modes.
C3_MODES = {
    "fp32_ffma": (0, "gate"),
    "tf32x3": (ffi.FLAG_GEMM_TF32X3, "gate"),
    "tf32": (ffi.FLAG_GEMM_TF32, 5e-3),
    "bf16": (ffi.FLAG_GEMM_BF16, 3e-2),
}


@pytest.mark.parametrize("mode", list(C3_MODES))
def test_c3_full_size_vs_oracle(mode, c3_oracle):
    """THE benchmarked plan (bench.py default: 2-CTA tcgen05 GEMMs, fused bias+tanh / tanh'+db
    epilogues, K-major shadows, CUDA graph) at the benchmarked size against the oracle: the loss
    and all 18 gradients (8 W, 8 b, x, Y).  Three calls: eager, captured, replayed -- the replay
    is what bench.py times."""
    args, f, (v32, g32), (v64, g64) = c3_oracle
    flags, how = C3_MODES[mode]
    d = [Y.cu(a) for a in args]
    gf = Y.CompiledGrad(Y.gradtape(f, *d), flags=flags)
    if mode in ("tf32", "bf16", "tf32x3"):
        desc = gf.describe()
        assert desc.count("+ epilogue{") == 15 and "ffma" not in desc, desc
    outs = None
    for _ in range(3):
        val, g = gf.run(d)
        outs = g
    val = float(val.to_host())
    report = {}
    if how == "gate":
        report["val"] = U.gate(val, v32, v64, f"{mode}: val")
    else:
        assert abs(val - v64) <= how * abs(v64), (mode, val, v64)
    for i, (a, b32, b64) in enumerate(zip(outs, g32, g64)):
        a = a.to_host()
        assert a.shape == b64.shape, (i, a.shape, b64.shape)
        if how == "gate":
            report[i] = U.gate(a, b32, b64, f"{mode}: grad[{i}]")
        else:
            scale = float(np.abs(b64).max())
            err = float(np.abs(a - b64).max())
            report[i] = err / scale
            assert err <= how * scale, (mode, i, err / scale)
    print(f"C3 full size, {mode}: worst err / max|g| = {max(report.values()):.3e}")


def _c3(flags, B=8192, L=8, W=4096):
    args = cases.c3_args(L, W, B)
    return args, cases.make_c3(L, B), [Y.cu(a) for a in args]


def test_c3_full_size_batch_linearity_and_determinism():
    """8 x (4096 -> 4096) tanh MLP, softmax-xent, batch 8192, tensor-core contractions: the
    full-batch gradients equal the sum of the two half-batch gradients (the loss is a sum over
    columns divided by a constant), twice the same bits run to run, and the last bias gradient
    sums to zero (softmax mass = one-hot mass)."""
    L, W, B = 8, 4096, 8192
    args, f, d = _c3(ffi.FLAG_GEMM_TF32)
    val, g = Y.grad(f, *d, flags=ffi.FLAG_GEMM_TF32)
    full = [a.to_host() for a in g[1:2 * L + 1]]
    val_b, g_b = Y.grad(f, *d, flags=ffi.FLAG_GEMM_TF32)
    assert val == val_b
    for a, b in zip(full, g_b[1:2 * L + 1]):
        assert np.array_equal(U.bits(a), U.bits(b.to_host()))
    del g, g_b
    halves, vsum = None, 0.0
    for lo in (0, B // 2):
        xs = [Y.cu(np.asfortranarray(a[:, lo:lo + B // 2])) for a in args[-2:]]
        v, gh = Y.grad(f, *(d[:2 * L] + xs), flags=ffi.FLAG_GEMM_TF32)
        vsum += v
        hs = [a.to_host() for a in gh[1:2 * L + 1]]
        halves = hs if halves is None else [p + q for p, q in zip(halves, hs)]
        del gh, xs
    assert abs(vsum - val) <= 1e-4 * abs(val)
    for i, (a, b) in enumerate(zip(full, halves)):
        assert np.abs(a - b).max() <= 5e-3 * np.abs(a).max(), i  # TF32 products, FP32 accumulation
    db_last = full[2 * L - 1]
    assert abs(db_last.sum()) <= 1e-4 * np.abs(db_last).sum()


def test_c3_full_size_directional_derivative_fp32():
    """Strict-FP32 contractions at full size: <grad, d> equals the central finite difference of
    the loss along d (the reference's gradcheck, src/gradcheck.jl:3-23, in directional form)."""
    L = 8
    args, f, d = _c3(0)
    val, g = Y.grad(f, *d, flags=0)
    # direction: the gradient itself, rescaled per tensor to the size of the parameter, so the
    # slope is a sum of squares (no cancellation); step sized for a loss change of ~0.02
    gh = [gi.to_host() for gi in g[1:2 * L + 1]]
    del g
    dirs = [gi * F32((np.abs(a).max() if a.ndim == 2 else 0.01) / max(np.abs(gi).max(), 1e-30)) for gi, a in zip(gh, args[:2 * L])]
    slope = sum(float(np.vdot(gi.astype(np.float64), di.astype(np.float64))) for gi, di in zip(gh, dirs))
    assert slope > 0
    eps = 0.01 / slope
    vals = []
    for s in (+1, -1):
        p = [Y.cu((a + F32(s * eps) * di).astype(F32)) for a, di in zip(args[:2 * L], dirs)]
        v, gg = Y.grad(f, *(p + d[2 * L:]), flags=0)
        vals.append(v)
        del gg, p
    fd = (vals[0] - vals[1]) / (2 * eps)
    assert abs(fd - slope) <= 2e-2 * abs(slope), (fd, slope, eps, val)
