#!/bin/bash
# N-GPU sweep of the all-reduce kernel's threads per CTA (request pressure on the memory system)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29700
for cfg in "256 tf32" "512 tf32" "256 bf16"; do
  set -- $cfg; c=$1; m=$2; port=$((port+1))
  YOTA_AR_THREADS=$c timeout 200 $TR --master-port $port bench.py --gpus $N --gemm $m --steps 20 --warmup 5 --quick > gpurun_out/ar_t${c}_${m}_n$N.json 2> gpurun_out/ar_t${c}_${m}_n$N.err
  echo "threads=$c $m rc=$? $(tail -1 gpurun_out/ar_t${c}_${m}_n$N.json | cut -c1-160)"
done
