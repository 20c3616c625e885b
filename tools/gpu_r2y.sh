#!/bin/bash
# forked leaf steps: parity, then A/B on the launch-bound configs
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_tc.py -m gpu -x -q > gpurun_out/pytest_fork.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_fork.log
{
for w in "c1 tf32" "c2 fp32" "c5 fp32" "c4 tf32"; do
set -- $w
timeout -s KILL 200 python tools/loop_ab.py 200 $1 $2
YOTA_NO_FORK=1 timeout -s KILL 200 python tools/loop_ab.py 200 $1 $2
done
} 2>&1 | tee gpurun_out/fork_ab.txt
