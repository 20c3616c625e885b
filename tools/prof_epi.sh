#!/bin/bash
# ncu --set full of the fused-epilogue GEMMs of one C3 eval, after the same command ran clean.
set -x
mkdir -p gpurun_out
MODE=${1:-tf32}
python bench.py --gemm $MODE --steps 5 --warmup 3 --no-other-modes --no-cpu-baseline > gpurun_out/epi_plain_$MODE.json 2> gpurun_out/epi_plain_$MODE.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc2_kernel' -s 30 -c 6 -o gpurun_out/prof_epi_$MODE -f \
  python bench.py --gemm $MODE --steps 1 --warmup 3 --no-other-modes --no-cpu-baseline > gpurun_out/epi_ncu_$MODE.log 2>&1
ncu -i gpurun_out/prof_epi_$MODE.ncu-rep --page raw --csv > gpurun_out/prof_epi_${MODE}_raw.csv
ls -la gpurun_out/
du -sh gpurun_out
