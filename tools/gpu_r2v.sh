#!/bin/bash
# compensated mode with register chunk sums: parity (incl. full size), then speed
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "x3 or tf32x3 or compensated or truncat" > gpurun_out/pytest_x3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_x3.log
for kc in 4 2 8; do
YOTA_X3_KCHUNK=$kc timeout -s KILL 300 python bench.py --gemm tf32x3 --steps 10 --warmup 3 --quick > gpurun_out/x3_kc$kc.json 2> gpurun_out/x3_kc$kc.err; echo "kc=$kc rc=$? $(tail -1 gpurun_out/x3_kc$kc.json | cut -c1-120)"
done
