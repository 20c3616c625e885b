"""GEMM micro-benchmark (GPU): the three C3 contractions in every mode / kernel variant, timed
with CUDA events on the context stream.  A/B switches are environment variables read at launch:
YOTA_GEMM_1CTA (force the single-CTA tcgen05 kernel), YOTA_TMA_SPREAD (one lane issues all TMA
boxes).  Usage: python tests/gemm_bench.py [reps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import yota_b200 as Y  # noqa: E402
from yota_b200 import ffi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = Y.context()
lib = ffi.lib()


def ev():
    e = C.c_void_p()
    ffi.check(lib.yota_event_create(ctx.h, C.byref(e)))
    return e


e0, e1 = ev(), ev()
rng = np.random.default_rng(0)
shapes = [("NN fwd  W*H", 0, 0, 4096, 8192, 4096), ("NT dW  dZ*H'", 0, 1, 4096, 4096, 8192), ("TN dH  W'*dZ", 1, 0, 4096, 8192, 4096),
          ("NN n=1024", 0, 0, 4096, 1024, 4096), ("NT k=1024", 0, 1, 4096, 4096, 1024)]
variants = [("2cta", {}), ("2cta-spread", {"YOTA_TMA_SPREAD": "1"}), ("1cta", {"YOTA_GEMM_1CTA": "1"}),
            ("1cta-spread", {"YOTA_GEMM_1CTA": "1", "YOTA_TMA_SPREAD": "1"})]
for name, tA, tB, m, n, k in shapes:
    A = Y.cu(rng.standard_normal((k, m) if tA else (m, k), dtype=np.float32))
    B = Y.cu(rng.standard_normal((n, k) if tB else (k, n), dtype=np.float32))
    Cd = Y.B200Array((m, n))
    ref = None
    for mode, mname in ((1, "tf32"), (2, "bf16")):
        for vname, env in variants:
            for kk in ("YOTA_GEMM_1CTA", "YOTA_TMA_SPREAD"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            args = (ctx.h, mode, tA, tB, m, n, k, A.ptr, A.dims[0], B.ptr, B.dims[0], Cd.ptr, m, 0)
            for _ in range(3):
                ffi.check(lib.yota_gemm(*args))
            ctx.sync()
            ffi.check(lib.yota_event_record(ctx.h, e0))
            for _ in range(reps):
                ffi.check(lib.yota_gemm(*args))
            ffi.check(lib.yota_event_record(ctx.h, e1))
            ctx.sync()
            ms = C.c_float()
            ffi.check(lib.yota_event_elapsed_ms(ctx.h, e0, e1, C.byref(ms)))
            t = ms.value / reps
            got = Cd.to_host()
            if ref is None:
                ref = got
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            print(f"{name:14s} {m}x{n}x{k} {mname} {vname:11s} {t*1e3:8.1f} us  {2*m*n*k/t/1e9:8.1f} TFLOP/s  maxdiff-vs-first {err:.2e}"
                  + ("  (bf16 includes the operand casts)" if mode == 2 else ""), flush=True)
