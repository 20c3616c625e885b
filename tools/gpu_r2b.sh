#!/bin/bash
mkdir -p gpurun_out
for kc in 8 4; do for pace in 0 300 700 1500; do
YOTA_X3_PACE_NS=$pace YOTA_X3_KCHUNK=$kc timeout 300 python bench.py --gemm tf32x3 --quick --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kchunk $kc pace $pace', round(d['value'],2), 'evals/s', round(d['ms_per_step'],2),'ms')"; done; done
