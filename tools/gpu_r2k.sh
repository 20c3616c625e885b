#!/bin/bash
# final multi-GPU record: full bench line (device, e2e, weak scaling), timeline, A/B of the round-2 changes
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29900
port=$((port+1)); timeout 600 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_final_n$N.json 2> gpurun_out/bench_r2_final_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r2_final_n$N.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'weak',d['weak_scaling'])"
if [ "$N" = "8" ]; then
q() { name=$1; shift; port=$((port+1))
  env "$@" timeout 200 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --quick --steps 30 --warmup 5 > gpurun_out/k_${name}_n$N.json 2> gpurun_out/k_${name}_n$N.err
  echo -n "$name rc=$? "; tail -1 gpurun_out/k_${name}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
}
q default X=1
q nofinish YOTA_NO_FUSED_FINISH=1
q join YOTA_AR_JOIN=1
q bf16 YOTA_GEMM_MODE=bf16
port=$((port+1)); timeout 200 $TR --nproc-per-node $N --master-port $port tools/dp_timeline.py tf32 > gpurun_out/timeline_r2_final_n${N}.txt 2> /dev/null; echo "timeline rc=$?"
grep -E "matmul|allreduce|rowsum|whole|alone" gpurun_out/timeline_r2_final_n${N}.txt | cut -c1-58,66-100 | sed -n 14,22p; grep -E "whole|alone" gpurun_out/timeline_r2_final_n${N}.txt | cut -c1-150
fi
