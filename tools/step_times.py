#!/usr/bin/env python
"""Per-step device times of one config (serialised pass, CUDA events): python tools/step_times.py c3 tf32"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import ffi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
ctx = Y.context(0)
wl = bench.workload(name, 0, 1)
dargs = [Y.cu(x) for x in wl["args"]]
gf = Y.CompiledGrad(Y.gradtape(wl["f"], *dargs), ctx, {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16}[mode])
outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
for _ in range(3):
    gf.run(dargs, out=outs)
best = None
for _ in range(5):
    prof = gf.profile(dargs, out=outs)
    best = prof if best is None else [(l, min(a, b)) for (l, a), (_, b) in zip(best, prof)]
tot = 0.0
for label, ms in best:
    if ms > 0.003:
        print(f"{ms * 1e3:9.1f} us  {label[:110]}")
    tot += ms
print(f"total {tot:.3f} ms  ({os.environ.get('YOTA_NO_TSHADOW') and 'no transposed shadows' or 'default'})")
