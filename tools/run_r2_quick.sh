#!/bin/bash
# quick loop: selected GPU tests + short bench (+ optional SC trace)
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 900 python -m pytest tests -m gpu -q  > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
QUICK="--no-cpu-baseline --no-e2e --no-backbone --no-cfg3"
timeout 600 python bench.py --steps 5 --warmup 3 $QUICK > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"; tail -n 3 gpurun_out/bench.err
if [ -f build/libgmf_sctrace.so ] && [ -n "$TRACE" ]; then
cp gmf_b200/libgmf_b200.so /tmp/orig.so && cp build/libgmf_sctrace.so gmf_b200/libgmf_b200.so
timeout 600 python bench.py --steps 2 --warmup 3 $QUICK --no-roofline > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "[bench trace exit $?]"
cp /tmp/orig.so gmf_b200/libgmf_b200.so
python tools/sc_trace.py > gpurun_out/sc_trace.txt 2>&1; tail -n 34 gpurun_out/sc_trace.txt
fi
python - <<'PY'
import json
for f in ('gpurun_out/bench.json',):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
        if d.get('roofline'):
            for k,v in sorted(d['roofline']['per_kernel'].items(), key=lambda x:-x[1]['ms_per_step']):
                print(f"  {k:18s} {v['ms_per_step']:8.3f} ms {v['launches_per_step']:.0f}x  {v['bound']:6s} {v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.3f}")
    except Exception as e:
        print(f, 'parse failed', e)
PY
cat gpurun_out/parity_measured.jsonl 2>/dev/null | cut -c1-260
