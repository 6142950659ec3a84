#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "[ref exit $?]"
tail -c 900 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "[bench exit $?]"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ['value','ms_per_step','gpu_launches','clocks','cpu_baseline']})
print('e2e', d['e2e'])
print('roofline', {k: d['roofline'][k] for k in ['achieved','frac','avg_launch_ms']}, d['roofline']['whole_path'])
PY
python __graft_entry__.py smoke
