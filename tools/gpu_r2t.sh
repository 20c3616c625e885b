#!/bin/bash
# 2 GPUs: the multi-GPU tests and the driver's bench command with the final code
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/pytest_dp_n2.log 2>&1; echo "pytest dp rc=$?"; tail -3 gpurun_out/pytest_dp_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d.get('weak_scaling'))
PY
