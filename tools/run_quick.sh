#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'e2e_sync', d['e2e'] and round(d['e2e']['sync_call']['value'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
    if d.get('backbone'): print('backbone', d['backbone'])
    if d.get('roofline'): print('roofline frac', round(d['roofline']['frac'],4), 'achieved', round(d['roofline']['achieved'],1), 'whole', round(d['roofline']['whole_path']['frac_of_tensor_peak'],4))
    for k,v in sorted((d.get('kernel_profile') or {}).items(), key=lambda x:-x[1]['ms_per_step']):
        print(f"  {k:18s} {v['ms_per_step']:8.3f} ms  {100*v['share']:5.1f}%  {v['launches_per_step']:.0f}x")
except Exception as e:
    print('bench parse failed', e)
PY
