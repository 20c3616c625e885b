#!/usr/bin/env python
"""evals/s of one workload in the graph-replayed loop (the product path): python tools/loop_ab.py [steps [workload [gemm mode]]]
Used for A/B runs through environment switches (YOTA_NO_PRESORT, YOTA_PRESORT_REGION_PER_SM)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import ffi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
name = sys.argv[2] if len(sys.argv) > 2 else "c5"
mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
ctx = Y.context(0)
wl = bench.workload(name, 0, 1)
dargs = [Y.cu(x) for x in wl["args"]]
gf = Y.CompiledGrad(Y.gradtape(wl["f"], *dargs), ctx, {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16, "tf32x3": ffi.FLAG_GEMM_TF32X3}[mode])
outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
for _ in range(5):
    gf.run(dargs, out=outs)
best = 1e9
for _ in range(3):
    tm = bench.Timer(ctx, ffi.lib(), ffi)
    tm.start()
    for _ in range(steps):
        gf.run(dargs, out=outs)
    best = min(best, tm.stop() / steps)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("YOTA_"))
print(f"{name} {mode}: {best:.4f} ms/eval  {1e3 / best:.1f} evals/s  [{tag or 'default'}]")
