"""Accuracy of the GEMM modes against Float64 on one GPU: signed mean and max error relative to
max|result| for random-sign and all-positive data (the latter exposes the tensor core's truncating
accumulation), at several K.  YOTA_X3_KCHUNK sets the accumulation chunk of mode 3."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as U  # noqa: E402

r = np.random.default_rng(0)
m, n = 512, 768
print("kchunk", os.environ.get("YOTA_X3_KCHUNK", "default"))
for k in (512, 2048, 8192):
    for positive in (False, True):
        A = (r.random((m, k)) if positive else r.standard_normal((m, k))).astype(np.float32)
        B = (r.random((k, n)) if positive else r.standard_normal((k, n))).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        sc = np.abs(ref).max()
        row = []
        for mode in (0, 1, 3):
            got = U.gemm(mode, 0, 0, np.asfortranarray(A), np.asfortranarray(B))
            d = (got - ref) / sc
            row.append(f"mode{mode}: mean {d.mean():+.2e} max {np.abs(d).max():.2e}")
        print(f"K={k:5d} {'pos ' if positive else 'rand'} | " + " | ".join(row))
