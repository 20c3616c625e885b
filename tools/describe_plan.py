#!/usr/bin/env python
"""Print the planned step list of a BASELINE config (host only, no GPU):
   python tools/describe_plan.py c3 [tf32|bf16|fp32]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import ffi  # noqa: E402
from yota_b200.api import AVal, lower  # noqa: E402


def config(name):
    if name == "c3":
        L, W, B = 8, 4096, int(os.environ.get("BATCH", 8192))
        return cases.make_c3(L, B), [AVal((W, W), "f32")] * L + [AVal((W,), "f32")] * L + [AVal((W, B), "f32")] * 2
    if name == "c4":
        H, B, T = 1024, 128, 128
        return cases.make_c4(T), [AVal((H, H), "f32")] * 2 + [AVal((H,), "f32"), AVal((H, B, T), "f32")]
    if name == "c5":
        return cases.c5_embed, [AVal((256, 1 << 20), "f32"), AVal((1 << 22,), "i64"), AVal((256, 1 << 22), "f32")]
    if name == "c2":
        return cases.c2_chain, [AVal((8192, 8192), "f32")] * 2 + [AVal((8192,), "f32")]
    if name == "c1":
        return cases.c1_mlp, [AVal((128, 784), "f32"), AVal((128,), "f32"), AVal((10, 128), "f32"), AVal((10,), "f32"),
                              AVal((784, 64), "f32"), AVal((10, 64), "f32")]
    raise SystemExit("unknown config")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
    flags = {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16}[mode]
    f, av = config(name)
    lib = ffi.lib()
    ctx = C.c_void_p()
    ffi.check(lib.yota_init(0, C.byref(ctx)))
    L = lower(Y.gradtape(f, *av))
    ops, tens = ffi.pack_program(L.prog)
    h = C.c_void_p()
    ffi.check(lib.yota_program_build(ctx, ops, len(L.prog.ops), tens, len(L.prog.tensors), flags, C.byref(h)))
    n = lib.yota_program_describe(h, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.yota_program_describe(h, buf, n + 1)
    print(buf.value.decode())


if __name__ == "__main__":
    main()
