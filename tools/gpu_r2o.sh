#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "side_stream or scatter or c5 or index" > gpurun_out/pytest_idx.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_idx.log
timeout 600 python bench.py --no-weak --steps 20 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k in d:
    if 'second' in k or 'hbm' in k: print(k, json.dumps(d[k])[:1500])
PY
