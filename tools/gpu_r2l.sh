#!/bin/bash
# C5 launch list: where the 0.37 ms in front of the fold goes
mkdir -p gpurun_out
timeout 300 python tools/step_times.py c5 tf32 > gpurun_out/c5_steps.txt 2>&1; echo "steps rc=$?"; tail -12 gpurun_out/c5_steps.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c5_launches.csv python tools/step_times.py c5 tf32 > gpurun_out/c5_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c5_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
names=[(r[ki][:60], float(r[vi].replace(',',''))) for r in rows[1:]]
# one eval = the repeating period; print the first 40 after warm-up start
for n,v in names[:45]: print(f"{v/1e3:9.1f} us {n}")
PY
