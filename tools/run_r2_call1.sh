#!/bin/bash
# round 2, call 1: GPU tests (all, not -x), ubench T9/T10, default bench, no-MUFU variant of the attention kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
timeout 300 tools/ubench/sm_rates > gpurun_out/sm_rates.log 2>&1; echo "[ubench exit $?]"
grep -A 12 "== T9" gpurun_out/sm_rates.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -n 5 gpurun_out/bench.err
cp gmf_b200/libgmf_b200.so /tmp/orig.so && cp build/libgmf_dbg2.so gmf_b200/libgmf_b200.so
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 > gpurun_out/bench_dbg2.json 2> gpurun_out/bench_dbg2.err; echo "[bench dbg2 exit $?]"
cp /tmp/orig.so gmf_b200/libgmf_b200.so
python - <<'PY'
import json
for f in ('gpurun_out/bench.json','gpurun_out/bench_dbg2.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
        if d.get('strong_scaling'): print('cfg3', d['strong_scaling'])
        if d.get('cpu_baseline'): print('cpu', d['cpu_baseline'])
        if d.get('roofline'):
            print('roofline frac', round(d['roofline']['frac'],4))
            for k,v in sorted(d['roofline']['per_kernel'].items(), key=lambda x:-x[1]['ms_per_step']):
                print(f"  {k:18s} {v['ms_per_step']:8.3f} ms {v['launches_per_step']:.0f}x  {v['bound']:6s} {v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.3f}")
    except Exception as e:
        print(f, 'parse failed', e)
PY
cat gpurun_out/parity_measured.jsonl
