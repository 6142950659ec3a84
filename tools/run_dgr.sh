#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_dgr_head.py > gpurun_out/dgr_bench.jsonl 2> gpurun_out/dgr_bench.err; echo "[dgr bench exit $?]"
tail -n 3 gpurun_out/dgr_bench.err
cat gpurun_out/dgr_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], 'ms', round(d['ms_per_forward'],4), 'min', round(d['ms_min'],4), 'e2e', round(d['e2e']['ms_per_forward'],3), 'launches', d['gpu_launches'], 'TF', round(d['roofline']['achieved'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['ms_per_forward'],1), 'err', d['cpu_baseline'] and d['cpu_baseline']['max_abs_diff_vs_cuda'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dgr_launches.csv python tools/bench_dgr_head.py --iters 1 --no-cpu > gpurun_out/dgr_ncu.log 2>&1; echo "[ncu exit $?]"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/dgr_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
# last forward of the last config: print the final 12 launches
for r in rows[-12:]:
    print(r[ki][:70], r[vi])
PY
