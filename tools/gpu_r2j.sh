#!/bin/bash
cat > /tmp/prof.py <<'PY'
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cases, yota_b200 as Y
from yota_b200 import ffi
args = cases.c3_args(8, 4096, 8192)
d = [Y.cu(a) for a in args]
gf = Y.CompiledGrad(Y.gradtape(cases.make_c3(8, 8192), *d), flags=ffi.FLAG_GEMM_TF32)
outs = [Y.B200Array(gf._out_dims[s], "f32") for s in range(gf.n_out)]
for _ in range(4): prof = gf.profile(d, out=outs)
g = {}
for l, t in prof:
    k = l.split(" ->")[0][:60]
    a = g.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
print(os.environ.get("TAG"), " ".join(f"{a[0]}x{1e3*a[1]/a[0]:.0f}" for k, a in sorted(g.items(), key=lambda kv: -kv[1][1])[:4]))
PY
for i in 1 2; do
TAG=mc_on python /tmp/prof.py
TAG=mc_off YOTA_GEMM_NO_MULTICAST=1 python /tmp/prof.py
done
