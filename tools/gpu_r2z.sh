#!/bin/bash
# same-box A/B of two builds of the library on the C3 loop (YOTA_LIB selects the build)
mkdir -p gpurun_out
OLD=$PWD/yota.jl_b200/lib/libyota_old.so
for rep in 1 2; do
  for lib in new old; do
    if [ $lib = old ]; then export YOTA_LIB=$OLD; else unset YOTA_LIB; fi
    timeout 200 python bench.py --steps 30 --warmup 5 --quick 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value'],2), round(d['ms_per_step'],3))"
  done
done
