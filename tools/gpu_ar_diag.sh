#!/bin/bash
# 2-GPU diagnostic: why does the GEMM that overlaps the all-reduce stretch to its length?
N=2; export BATCH=2048
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29800
run() { name=$1; shift; port=$((port+1))
  env "$@" timeout 120 $TR --master-port $port tools/dp_timeline.py tf32 > gpurun_out/diag_$name.txt 2> gpurun_out/diag_$name.err
  echo "== $name rc=$?"; grep -E "allreduce|epilogue\{region\[square|matmul -> %(68|74)" gpurun_out/diag_$name.txt | sed -n 1,9p | cut -c1-112; grep -E "whole|alone" gpurun_out/diag_$name.txt | cut -c1-140
}
run base X=1
run barriers_only YOTA_AR_PHASES=0
run reduce_only YOTA_AR_PHASES=1
run bcast_only YOTA_AR_PHASES=2
run t256 YOTA_AR_THREADS=256
run nccl YOTA_AR_NCCL=1
