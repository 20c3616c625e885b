#!/bin/bash
# 2 GPUs, 1024 batch columns per rank (the per-GPU shapes of the 8-GPU run): timeline with / without A-multicast
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29700
for mc in 0 1; do
  if [ $mc = 1 ]; then E=X=1; else E=YOTA_GEMM_NO_MULTICAST=1; fi
  port=$((port+1))
  env $E BATCH=2048 timeout 200 $TR --master-port $port tools/dp_timeline.py tf32 > gpurun_out/tl2_b1024_mc$mc.txt 2> gpurun_out/tl2_mc$mc.err; echo "timeline mc=$mc rc=$?"
  grep -E "matmul|allreduce|rowsum|whole|alone" gpurun_out/tl2_b1024_mc$mc.txt | cut -c1-58,66-100 | sed -n 14,26p; grep -E "whole|alone" gpurun_out/tl2_b1024_mc$mc.txt | cut -c1-150
done
