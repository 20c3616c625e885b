"""Helpers for the GPU parity tests (call the product only through its public API / C ABI)."""
import ctypes as C

import numpy as np

import yota_b200 as Y
from oracle import yota_oracle as O
from yota_b200 import ffi


def to_f64(args):
    return [a.astype(np.float64) if isinstance(a, np.ndarray) and a.dtype.kind == "f" else a for a in args]


def gate(gpu, o32, o64, what="", rtol=1e-5):
    """SURVEY §8(d) parity gate for the FP32 path: |gpu - f64| <= max(rtol*scale, 2*|oracle32 - f64|),
    scale = max|f64| (the reference's own FP32 rounding on long sums is already ~1e-5)."""
    gpu, o32, o64 = (np.asarray(x, dtype=np.float64) for x in (gpu, o32, o64))
    assert gpu.shape == o64.shape, (what, gpu.shape, o64.shape)
    scale = max(float(np.max(np.abs(o64))) if o64.size else 0.0, 1e-30)
    e_gpu = float(np.max(np.abs(gpu - o64))) if o64.size else 0.0
    e_o32 = float(np.max(np.abs(o32 - o64))) if o64.size else 0.0
    e_pair = float(np.max(np.abs(gpu - o32))) if o64.size else 0.0
    assert e_gpu <= max(rtol * scale, 2 * e_o32) + 1e-30, \
        f"{what}: err(gpu,f64)={e_gpu:.3e} err(oracle32,f64)={e_o32:.3e} err(gpu,oracle32)={e_pair:.3e} scale={scale:.3e}"
    return e_gpu / scale


def check_parity(f, args, seed=1, flags=None, rtol=1e-5, o32=None, o64=None):
    """Device grad (host arrays moved with cu) vs the oracle in Float32 and Float64."""
    dargs = [Y.cu(a) if isinstance(a, np.ndarray) else a for a in args]
    dseed = Y.cu(np.asarray(seed, np.float32)) if isinstance(seed, np.ndarray) else seed
    val, g = Y.grad(f, *dargs, seed=dseed, flags=flags)
    v32, g32 = o32 if o32 is not None else O.grad(f, *args, seed=seed, dtype=np.float32)
    v64, g64 = o64 if o64 is not None else O.grad(f, *to_f64(args), seed=seed, dtype=np.float64)
    val_h = val.to_host() if isinstance(val, Y.B200Array) else val
    worst = gate(val_h, v32, v64, "val", rtol)
    assert g[0] is Y.ZeroTangent
    for i, (a, b32, b64) in enumerate(zip(g[1:], g32[1:], g64[1:])):
        if isinstance(a, Y.B200Array):
            assert a.dims == np.shape(b64), (i, a.dims, np.shape(b64))
            worst = max(worst, gate(a.to_host(), b32, b64, f"grad[{i}]", rtol))
        else:
            assert repr(a) == repr(b64), (i, a, b64)
    return val, g, worst


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def raw_ctx():
    return Y.context()


def dev_ptr(arr):
    return C.c_void_p(arr.ptr)


def gemm(mode, tA, tB, A, B, Cin=None):
    """yota_gemm on host arrays (column-major): returns op(A) op(B) (+ Cin)."""
    ctx = Y.context()
    m, k = (A.shape[1], A.shape[0]) if tA else A.shape
    n = B.shape[0] if tB else B.shape[1]
    dA, dB = Y.cu(A), Y.cu(B)
    dC = Y.cu(Cin) if Cin is not None else Y.B200Array((m, n))
    ffi.check(ffi.lib().yota_gemm(ctx.h, mode, int(tA), int(tB), m, n, k, dA.ptr, A.shape[0], dB.ptr, B.shape[0],
                                  dC.ptr, m, int(Cin is not None)))
    return dC.to_host()
