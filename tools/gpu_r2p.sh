#!/bin/bash
# programmatic dependent launch of the skinny GEMMs: parity, then C4 A/B
mkdir -p gpurun_out
echo skip-pytest
{
for rep in 1 2; do
timeout 300 python tools/loop_ab.py 30 c4 tf32
YOTA_NO_PDL=1 timeout 300 python tools/loop_ab.py 30 c4 tf32
done
timeout 300 python tools/loop_ab.py 200 c1 tf32
YOTA_NO_PDL=1 timeout 300 python tools/loop_ab.py 200 c1 tf32
} 2>&1 | tee gpurun_out/c4_ab.txt
