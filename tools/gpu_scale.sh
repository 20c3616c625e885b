#!/bin/bash
# multi-GPU: DP correctness at N ranks, then the C3 bench at N (and N/2) ranks
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29511 tests/dp_gpu_check.py > gpurun_out/dp_check_n$N.log 2>&1; echo "dp_check rc=$?"; grep DP_GPU_OK gpurun_out/dp_check_n$N.log
port=29520
for n in $N $((N/2)); do
  for m in tf32 bf16; do
    port=$((port+1))
    timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --gemm $m --steps 20 --warmup 3 > gpurun_out/bench_c3_${m}_n$n.json 2> gpurun_out/bench_c3_${m}_n$n.err; echo "c3 $m n=$n rc=$?"
    tail -1 gpurun_out/bench_c3_${m}_n$n.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['step_breakdown_ms'])"
  done
done
