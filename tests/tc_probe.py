"""Bring-up probe for the tcgen05 GEMM: exact small-integer check per (transA, transB, shape).
Run under `timeout`; prints one line per config."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gpu_util as U  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = [(128, 128, 32), (128, 256, 64), (256, 128, 96), (256, 512, 256), (1024, 2048, 512), (4096, 4096, 4096)]
rng = np.random.default_rng(0)
for (m, n, k) in shapes:
    for tA in (0, 1):
        for tB in (0, 1):
            A = rng.integers(-3, 4, size=(k, m) if tA else (m, k)).astype(np.float32)
            B = rng.integers(-3, 4, size=(n, k) if tB else (k, n)).astype(np.float32)
            ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
            got = U.gemm(mode, tA, tB, np.asfortranarray(A), np.asfortranarray(B))
            err = np.abs(got - ref).max()
            bad = int((got != ref).sum())
            print(f"mode={mode} m={m} n={n} k={k} tA={tA} tB={tB} maxerr={err:.3g} mismatches={bad}/{got.size}", flush=True)
            if bad and m <= 256:
                i, j = np.argwhere(got != ref)[0]
                print("   first mismatch at", (i, j), "got", got[i, j], "ref", ref[i, j], "; row-ok", int((got[i] == ref[i]).sum()),
                      "col-ok", int((got[:, j] == ref[:, j]).sum()), flush=True)
