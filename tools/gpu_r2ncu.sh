#!/bin/bash
# ncu --set full of the two kernels new at the end of round 2 (each after its command ran clean without ncu):
# the compensated 2-CTA GEMM with register chunk sums, and the persistent recurrence kernel
mkdir -p gpurun_out
timeout 120 python bench.py --gemm tf32x3 --steps 2 --warmup 1 --quick > gpurun_out/x3_plain.log 2>&1 || { echo "x3 plain failed"; exit 0; }
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc2_kernel' -s 26 -c 3 -o gpurun_out/prof_x3 -f \
  python bench.py --gemm tf32x3 --steps 1 --warmup 1 --quick > gpurun_out/x3_ncu.log 2>&1; echo "ncu x3 rc=$?"
ncu -i gpurun_out/prof_x3.ncu-rep --page raw --csv > gpurun_out/prof_x3_raw.csv 2>/dev/null
timeout 120 python tools/loop_ab.py 3 c4 tf32 > gpurun_out/c4_plain.log 2>&1 || { echo "c4 plain failed"; exit 0; }
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_chain_kernel' -s 2 -c 2 -o gpurun_out/prof_chain -f \
  python tools/loop_ab.py 3 c4 tf32 > gpurun_out/chain_ncu.log 2>&1; echo "ncu chain rc=$?"
ncu -i gpurun_out/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
ls -la gpurun_out/prof_* | cut -c1-120
