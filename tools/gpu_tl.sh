#!/bin/bash
# N-GPU: timeline + quick bench (tf32)
N=${1:-4}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 200 $TR --master-port 29512 tools/dp_timeline.py tf32 > gpurun_out/timeline_n${N}_tf32.txt 2> gpurun_out/timeline_n${N}_tf32.err; echo "timeline rc=$?"; grep -E "allreduce|whole|alone|epilogue\{region\[square" gpurun_out/timeline_n${N}_tf32.txt | cut -c1-112
for m in tf32 "$@"; do
timeout 200 $TR --master-port 29513 bench.py --gpus $N --gemm $m --steps 20 --warmup 5 --quick > gpurun_out/quick_${m}_n$N.json 2> gpurun_out/quick_${m}_n$N.err; echo "bench $m rc=$? $(tail -1 gpurun_out/quick_${m}_n$N.json | cut -c1-150)"
done
