#!/bin/bash
# index sort on the side stream: A/B of the C5 loop, interleaved repeats
mkdir -p gpurun_out
{
timeout 200 python tools/loop_ab.py 100 > /dev/null
for rep in 1 2; do
timeout 200 python tools/loop_ab.py 100
YOTA_SIDE_PRIORITY=1 timeout 200 python tools/loop_ab.py 100
YOTA_NO_PRESORT=1 timeout 200 python tools/loop_ab.py 100
done
} 2>&1 | tee gpurun_out/loop_ab.txt
