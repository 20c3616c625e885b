#!/bin/bash
# quick GPU check: parity tests + C3 bench in both tensor modes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for m in tf32 bf16; do
  python bench.py --gemm $m --steps 20 --warmup 3 --no-other-modes --no-cpu-baseline > gpurun_out/bench_c3_$m.json 2> gpurun_out/bench_c3_$m.err; echo "c3 $m rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/bench_c3_$m.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['step_breakdown_ms'])"
done
