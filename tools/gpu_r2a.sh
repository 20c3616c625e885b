#!/bin/bash
# round 2, first call: smoke, the whole GPU suite (no -x: every failure is wanted), default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=10 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|C3 full size|Error" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
tail -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
