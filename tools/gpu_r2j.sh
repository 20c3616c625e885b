#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_j.log
for i in 1 2; do for v in preferred strict4 nomc; do
  case $v in preferred) E=X=1;; strict4) E=YOTA_GEMM_STRICT_CLUSTER4=1;; nomc) E=YOTA_GEMM_NO_MULTICAST=1;; esac
  for m in tf32 bf16; do
  env $E timeout 300 python bench.py --gemm $m --quick --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $m', round(d['value'],2), 'evals/s', round(d['ms_per_step'],3),'ms')"
  done
done; done
