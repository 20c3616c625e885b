#!/bin/bash
# compensated mode: accuracy at full C3 size and speed per chunk length
mkdir -p gpurun_out
for kc in 1 2 4; do
YOTA_X3_KCHUNK=$kc timeout -s KILL 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "full_size_vs_oracle and tf32x3" 2>&1 | grep -E "worst err|passed|failed" | sed "s/^/kc=$kc  /"
done
YOTA_X3_KCHUNK=1 timeout -s KILL 300 python bench.py --gemm tf32x3 --steps 10 --warmup 3 --quick 2>/dev/null | cut -c1-100
