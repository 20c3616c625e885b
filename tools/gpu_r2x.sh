#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 200 python bench.py --workload c1 --steps 30 --warmup 5 --no-other-modes --no-cpu-baseline --no-secondary > gpurun_out/c1_$i.json 2> gpurun_out/c1_$i.err; echo "run $i rc=$? $(tail -1 gpurun_out/c1_$i.err | cut -c1-100)"
done
