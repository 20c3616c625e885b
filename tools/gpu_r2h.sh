#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_h.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -s -k "c4" > gpurun_out/pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; grep -E "passed|failed|C4 full|Error|assert" gpurun_out/pytest_c4.log | tail
timeout 300 python bench.py --workload c4 --gemm tf32 --no-cpu-baseline > gpurun_out/bench_c4_tf32.json 2> gpurun_out/bench_c4_tf32.err; echo "bench c4 rc=$?"
tail -1 gpurun_out/bench_c4_tf32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms', d['program'], d['step_breakdown_ms'])"
YOTA_NO_EPILOGUE1=1 timeout 300 python bench.py --workload c4 --gemm tf32 --quick | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no 1-CTA epilogues', round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
