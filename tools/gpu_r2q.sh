#!/bin/bash
# C4 launch list: kernel durations of the per-step GEMMs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/c4_launches.csv python tools/step_times.py c4 tf32 > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/c4_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:90]; v=float(r[vi].replace(',',''))/1e3
    a=agg.setdefault(k,[0,0.0,1e9,0]); a[0]+=1; a[1]+=v; a[2]=min(a[2],v); a[3]=max(a[3],v)
for k,a in agg.items(): print(f"{a[0]:4d} x avg {a[1]/a[0]:7.1f} us (min {a[2]:.1f} max {a[3]:.1f})  {k}")
PY
