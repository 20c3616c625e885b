#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q --timeout 60 --timeout-method=thread > gpurun_out/pytest_h.log 2>&1; echo "pytest gemm rc=$?"; tail -3 gpurun_out/pytest_h.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 300 --timeout-method=thread -s -k "c4" > gpurun_out/pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; grep -E "passed|failed|C4 full|Error|assert" gpurun_out/pytest_c4.log | tail -5
timeout 200 python bench.py --workload c4 --gemm tf32 --no-cpu-baseline > gpurun_out/bench_c4_tf32.json 2> gpurun_out/bench_c4_tf32.err; echo "bench c4 rc=$?"
tail -1 gpurun_out/bench_c4_tf32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms', d['program'], d['step_breakdown_ms'])"
YOTA_GEMM_NO_SPLITK=1 timeout 200 python bench.py --workload c4 --gemm tf32 --quick | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no split-K', round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
