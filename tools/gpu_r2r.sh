#!/bin/bash
# K split of the 2-CTA kernel: parity, then C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "k_split or c4 or gemm or rnn or epilogue" > gpurun_out/pytest_ks.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ks.log
{
for rep in 1 2; do
timeout 300 python tools/loop_ab.py 30 c4 tf32
YOTA_GEMM_NO_KSPLIT=1 timeout 300 python tools/loop_ab.py 30 c4 tf32
done
timeout 300 python tools/step_times.py c4 tf32 | sort -rn | head -12
} 2>&1 | tee gpurun_out/c4_ab.txt
