#!/bin/bash
# one ncu pass over ONE full 64-pair step (after a warm-up step): duration, DRAM bytes, DRAM / tensor / XU / FMA pipe utilisation per launch
mkdir -p gpurun_out
BIG="python bench.py --pairs 64 --steps 1 --warmup 1 --min-warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 1500 ncu --metrics $M --clock-control none -s 93 -c 93 --csv --log-file gpurun_out/allkernels.csv $BIG > gpurun_out/ncu_all.log 2>&1; echo "[ncu all exit $?]"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/allkernels.csv')) if len(r) > 8]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    try: val = float(r[vi].replace(',', ''))
    except ValueError: val = 0.0
    per.setdefault(r[ii], {'k': r[ki].split('(')[0][:48]})[r[mi]] = val
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d['k'], collections.Counter()); a['n'] += 1
    for m, v in d.items():
        if m != 'k': a[m] += v
print(f"{'kernel':50s} {'n':>3s} {'ms':>8s} {'GB/s':>7s} {'dram%':>6s} {'tens%':>6s} {'xu%':>6s} {'fma%':>6s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1]['gpu__time_duration.sum']):
    n = a['n']; t = a['gpu__time_duration.sum']
    unit = 1e6 if t / n > 1000 else 1.0   # ns -> ms if reported in ns
    by = a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']
    print(f"{k:50s} {n:3d} {t:12.1f} {by:14.1f} {a['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']/n:6.1f} {a['sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']/n:6.1f} {a['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']/n:6.1f} {a['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']/n:6.1f}")
PY
