#!/bin/bash
# N-GPU: DP correctness with both all-reduce paths, timeline, quick bench
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 300 $TR --master-port 29511 tests/dp_gpu_check.py > gpurun_out/dp_check_n$N.log 2>&1; echo "dp_check rc=$?"; grep -E "DP_GPU_OK|Error|error|assert" gpurun_out/dp_check_n$N.log | head -20
timeout 200 $TR --master-port 29512 tools/dp_timeline.py tf32 > gpurun_out/timeline_n${N}_tf32.txt 2> gpurun_out/timeline_n${N}_tf32.err; echo "timeline rc=$?"; grep -E "allreduce|whole|alone" gpurun_out/timeline_n${N}_tf32.txt | cut -c1-150
for m in tf32 bf16; do
timeout 200 $TR --master-port 29513 bench.py --gpus $N --gemm $m --steps 20 --warmup 5 --quick > gpurun_out/quick_${m}_n$N.json 2> gpurun_out/quick_${m}_n$N.err; echo "bench $m rc=$? $(tail -1 gpurun_out/quick_${m}_n$N.json | cut -c1-150)"
done
