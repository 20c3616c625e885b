#!/bin/bash
# 8 GPUs: what slows the GEMMs that run beside the all-reduce?
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
q() { name=$1; shift; port=$((port+1))
  env "$@" timeout 200 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --quick --steps 30 --warmup 5 > gpurun_out/f_${name}_n$N.json 2> gpurun_out/f_${name}_n$N.err
  echo -n "$name rc=$? "; tail -1 gpurun_out/f_${name}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'evals/s',round(d['ms_per_step'],3),'ms')"
}
q mc1 X=1
q mc0 YOTA_GEMM_NO_MULTICAST=1
q mc0_reduce_only YOTA_GEMM_NO_MULTICAST=1 YOTA_AR_PHASES=1
q mc0_bcast_only YOTA_GEMM_NO_MULTICAST=1 YOTA_AR_PHASES=2
q mc0_barriers_only YOTA_GEMM_NO_MULTICAST=1 YOTA_AR_PHASES=0
q mc0_c8t512 YOTA_GEMM_NO_MULTICAST=1 YOTA_AR_CTAS=8 YOTA_AR_THREADS=512
q mc0_c32t128 YOTA_GEMM_NO_MULTICAST=1 YOTA_AR_CTAS=32 YOTA_AR_THREADS=128
port=$((port+1)); YOTA_GEMM_NO_MULTICAST=1 timeout 200 $TR --nproc-per-node $N --master-port $port tools/dp_timeline.py tf32 > gpurun_out/timeline_r2_n${N}_mc0.txt 2> /dev/null; echo "timeline rc=$?"
grep -E "matmul|allreduce|rowsum|whole|alone" gpurun_out/timeline_r2_n${N}_mc0.txt | cut -c1-58,66-100 | sed -n 14,22p; grep -E "whole|alone" gpurun_out/timeline_r2_n${N}_mc0.txt | cut -c1-150
