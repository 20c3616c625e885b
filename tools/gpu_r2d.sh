#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_mc.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_mc.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --timeout 600 -s > gpurun_out/pytest_full.log 2>&1; echo "pytest fullsize rc=$?"; grep -E "passed|failed|C3 full" gpurun_out/pytest_full.log
for m in tf32 bf16; do
  timeout 300 python bench.py --gemm $m --quick --steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', round(d['value'],2), 'evals/s', round(d['ms_per_step'],3),'ms')"
done
BATCH=1024 timeout 300 python tools/dp_timeline.py tf32 > gpurun_out/tl_b1024.txt 2>&1; echo "timeline rc=$?"; grep -E "matmul|whole" gpurun_out/tl_b1024.txt | awk '{print $(NF-1)}' | tr '\n' ' '; echo
