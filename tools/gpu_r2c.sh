#!/bin/bash
# round 2 multi-GPU: DP parity (pytest), then quick A/B of the deferred join, timeline, full bench
N=${1:-8}; FULL=${2:-1}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 800 > gpurun_out/pytest_dp_n$N.log 2>&1; echo "pytest dp rc=$?"; tail -3 gpurun_out/pytest_dp_n$N.log
port=29600
q() { name=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --quick --steps 30 --warmup 5 > gpurun_out/q_${name}_n$N.json 2> gpurun_out/q_${name}_n$N.err
  echo -n "$name rc=$? "; tail -1 gpurun_out/q_${name}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
}
q overlap X=1
q join YOTA_AR_JOIN=1
q overlap_head0 YOTA_AR_HEAD_RESERVE=0
q overlap_head8 YOTA_AR_HEAD_RESERVE=8
q overlap_bf16 YOTA_GEMM_MODE=bf16
q join_bf16 YOTA_GEMM_MODE=bf16 YOTA_AR_JOIN=1
if [ "$FULL" = "1" ]; then
port=$((port+1)); timeout 300 $TR --nproc-per-node $N --master-port $port tools/dp_timeline.py tf32 > gpurun_out/timeline_r2_n${N}_tf32.txt 2> gpurun_out/timeline_r2_n${N}.err; echo "timeline rc=$?"; tail -4 gpurun_out/timeline_r2_n${N}_tf32.txt
port=$((port+1)); timeout 600 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r2_n$N.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'weak',d['weak_scaling'])"
tail -3 gpurun_out/bench_r2_n$N.err
fi
