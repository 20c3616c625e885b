#!/bin/bash
# round 2: smoke, the whole GPU suite (no -x: every failure is wanted), default bench + reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
tail -c 5500 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; cut -c1-700 gpurun_out/bench_reference.json
for w in c4 c1; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-other-modes --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d.get('roofline',{}).get('frac'))"
done
# launch list of the C3 loop with the final code (after the bench above has exited 0 without ncu)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_launches_final.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/c3_ncu.log 2>&1; echo "ncu c3 rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/c3_launches_final.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:80]; v=float(r[vi].replace(',',''))/1e3
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,a in agg.items(): print(f"{a[0]:4d} x avg {a[1]/a[0]:8.1f} us  {k}")
PY
