#!/bin/bash
# round 2 evidence: launch list of the timed region + --set full of five GEMMs, each ncu pass after
# the same command ran clean without ncu
mkdir -p gpurun_out
CMD="python bench.py --gemm tf32 --steps 2 --warmup 3 --no-other-modes --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain_r2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 96 -c 64 --csv --log-file gpurun_out/r02_c3_tf32_launches.csv $CMD > gpurun_out/ncu_r2_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain_r2b.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 30 -c 5 -o gpurun_out/prof_gemm_r2 -f $CMD > gpurun_out/ncu_r2_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_gemm_r2.ncu-rep --page raw --csv > gpurun_out/prof_gemm_r2_raw.csv 2>/dev/null
ls -la gpurun_out/prof_gemm_r2* gpurun_out/r02_c3_tf32_launches.csv
tail -2 gpurun_out/plain_r2.log | cut -c1-400
