#!/bin/bash
# N-GPU sweep of the communicator CTA cap / GEMM SM reservation (C3, device-resident timing only)
N=${1:-8}; MODE=${2:-tf32}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29600
run() { # name, env...
  name=$1; shift
  port=$((port+1))
  env "$@" timeout 240 $TR --master-port $port bench.py --gpus $N --gemm $MODE --steps 20 --warmup 5 --quick > gpurun_out/sweep_${name}.json 2> gpurun_out/sweep_${name}.err
  echo "$name rc=$? $(tail -1 gpurun_out/sweep_${name}.json | cut -c1-120)"
}
run default_n$N YOTA_NCCL_MAX_CTAS=0
run c16r16_n$N YOTA_NCCL_MAX_CTAS=16 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV,TUNING NCCL_DEBUG_FILE=gpurun_out/nccl_info_%p.log
run c8r8_n$N YOTA_NCCL_MAX_CTAS=8
run c32r32_n$N YOTA_NCCL_MAX_CTAS=32
run c16r0_n$N YOTA_NCCL_MAX_CTAS=16 YOTA_COMM_SM_RESERVE=0
run c24r24_n$N YOTA_NCCL_MAX_CTAS=24
ls gpurun_out/nccl_info_* | head -1 | xargs -I{} sh -c 'grep -iE "nvls|algo|channel|CTA" {} | head -60 > gpurun_out/nccl_info_summary.log'; rm -f gpurun_out/nccl_info_[0-9]*.log
