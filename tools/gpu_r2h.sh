#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_tc.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_h.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -s -k "c4" > gpurun_out/pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; grep -E "passed|failed|C4 full|Error|assert" gpurun_out/pytest_c4.log | tail
for m in tf32 fp32; do
timeout 300 python bench.py --workload c4 --gemm $m --no-cpu-baseline > gpurun_out/bench_c4_$m.json 2> gpurun_out/bench_c4_$m.err; echo "bench c4 $m rc=$?"
tail -1 gpurun_out/bench_c4_$m.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms', d['program'], d['step_breakdown_ms'])"
done
YOTA_NO_LOOP_BATCHING=1 timeout 300 python bench.py --workload c4 --gemm tf32 --quick | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unrolled', round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
timeout 300 python bench.py --workload c1 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1', round(d['value'],1),'evals/s',round(d['ms_per_step'],4),'ms')"
