#!/bin/bash
# persistent recurrence kernel: parity under a hard timeout, then C4
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q -k "persistent_launch" > gpurun_out/pytest_chain.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/pytest_chain.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 400 python -m pytest tests -m gpu -x -q -k "c4" > gpurun_out/pytest_chain2.log 2>&1; echo "pytest2 rc=$?"; tail -4 gpurun_out/pytest_chain2.log
{
for rep in 1 2; do
timeout -s KILL 200 python tools/loop_ab.py 30 c4 tf32
YOTA_NO_CHAIN=1 timeout -s KILL 200 python tools/loop_ab.py 30 c4 tf32
done
} 2>&1 | tee gpurun_out/c4_ab.txt
timeout -s KILL 200 python tools/chain_trace.py
