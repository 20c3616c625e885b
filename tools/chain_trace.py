#!/usr/bin/env python
"""Where a step of the persistent recurrence kernel spends its time: %globaltimer stamps of the
cluster that owns tile 0 (YOTA_CHAIN_TRACE=1, no graph so that the last launch is the backward
recurrence of the last eval).  python tools/chain_trace.py"""
import ctypes as C
import os
import sys

os.environ["YOTA_CHAIN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import ffi  # noqa: E402

ctx = Y.context(0)
wl = bench.workload("c4", 0, 1)
dargs = [Y.cu(x) for x in wl["args"]]
gf = Y.CompiledGrad(Y.gradtape(wl["f"], *dargs), ctx, ffi.FLAG_GEMM_TF32 | ffi.FLAG_NO_GRAPH)
outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
for _ in range(3):
    gf.run(dargs, out=outs)
ctx.sync()
buf = np.zeros((4096, 8), np.uint64)
lib = ffi.lib()
lib.yota_debug_chain_trace.restype = C.c_int
n = lib.yota_debug_chain_trace(buf.ctypes.data_as(C.c_void_p), 4096)
t = buf[:n].astype(np.int64)
names = ["flags seen", "B requested", "products done (leader)", "products done (rank 1)", "partials sent (rank 1)",
         "partials in (leader)", "stored", "published"]
print(f"{n} steps; per-step period {np.median(np.diff(t[5:-5, 7])):.0f} ns (median)")
base = t[:, 0:1]
rel = (t - base)[5:-5]
for k, nm in enumerate(names):
    print(f"  {nm:26s} +{np.median(rel[:, k]):7.0f} ns after this step's flags were seen")
gap = t[6:-5, 0] - t[5:-6, 7]
print(f"  publish(i) -> flags seen(i+1): {np.median(gap):.0f} ns")
