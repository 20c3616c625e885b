#!/bin/bash
# scatter-add after the radix-sort rewrite: parity, then C5 step times and launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "scatter or unget or getindex or c5 or embed or index or fuzz" > gpurun_out/pytest_idx.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_idx.log
timeout 300 python tools/step_times.py c5 tf32 > gpurun_out/c5_steps.txt 2>&1; echo "steps rc=$?"; tail -5 gpurun_out/c5_steps.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c5_launches.csv python tools/step_times.py c5 tf32 > gpurun_out/c5_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c5_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:24]: print(f"{float(r[vi].replace(',',''))/1e3:9.1f} us {r[ki][:60]}")
PY
