#!/bin/bash
# A/B in one job (same box, interleaved): env var $1 set (A) vs unset (B); C3 bench, device-resident timing
VAR=$1; MODE=${2:-tf32}
mkdir -p gpurun_out
for rep in 1 2 3; do
  for ab in A B; do
    if [ $ab = A ]; then export $VAR=1; else unset $VAR; fi
    python bench.py --gemm $MODE --steps 30 --warmup 5 --quick > gpurun_out/ab_${ab}_$rep.json 2> gpurun_out/ab_${ab}_$rep.err
    echo "$ab($VAR=$([ $ab = A ] && echo 1 || echo unset)) rep $rep: $(python -c "
import json;d=json.loads(open('gpurun_out/ab_${ab}_$rep.json').read().strip().splitlines()[-1]);print(round(d['value'],2),'evals/s',round(d['ms_per_step'],3),'ms', d['clocks']['reasons'])")"
  done
done
