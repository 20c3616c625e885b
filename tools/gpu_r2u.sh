#!/bin/bash
# final C4 launch list (one eval = 13 launches + memsets)
mkdir -p gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c4_launches_final.csv python tools/loop_ab.py 3 c4 tf32 > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c4_launches_final.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:31]: print(f"{float(r[vi].replace(',',''))/1e3:9.1f} us {r[ki][:100]}")
PY
