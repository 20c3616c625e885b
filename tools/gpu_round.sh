#!/bin/bash
# One-GPU round check: parity tests, the bench lines of every BASELINE config, the reference arm,
# and the ncu evidence for profiles/ (launch list + --set full of the GEMMs), each ncu pass after
# the same command ran clean without ncu.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
for w in c1 c2 c4 c5; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
python bench.py --gemm bf16 --no-other-modes --no-cpu-baseline > gpurun_out/bench_c3_bf16.json 2> gpurun_out/bench_c3_bf16.err; echo "bench bf16 rc=$?"
CMD="python bench.py --gemm tf32 --steps 2 --warmup 3 --no-other-modes --no-cpu-baseline"
$CMD > gpurun_out/plain_c3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 162 -c 108 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_c3_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 30 -c 5 -o gpurun_out/prof_gemm_r1b -f $CMD > gpurun_out/ncu_c3_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_gemm_r1b.ncu-rep --page raw --csv > gpurun_out/prof_gemm_r1b_raw.csv 2>/dev/null
du -sh gpurun_out
for f in bench_default bench_c1 bench_c2 bench_c4 bench_c5 bench_c3_bf16 bench_reference; do
  tail -1 gpurun_out/$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('$f', round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'roof', r.get('kernel'), round(r.get('frac') or 0,3), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
done
