#!/bin/bash
# FFN iteration: parity subset + quick bench with the default library, then the clock64 timeline of one CTA (GMF_FFN_TRACE build)
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q -k "ffn or forward or drop_in or full_size or dgr_head or encoder" > gpurun_out/pytest_ffn.log 2>&1; echo "[pytest exit $?]"; tail -n 5 gpurun_out/pytest_ffn.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "[bench exit $?]"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
pk = d['roofline']['per_kernel']
print('ms/step', round(d['ms_per_step'], 2), 'pairs/s', round(d['value'], 1), ' '.join(f"{k}={v['ms_per_step']:.2f}({v['frac']:.2f})" for k, v in pk.items() if v['ms_per_step'] > 0.5))
PY
if [ -f build/libgmf_ffntrace.so ]; then
  cp gmf_b200/libgmf_b200.so /tmp/orig.so; cp build/libgmf_ffntrace.so gmf_b200/libgmf_b200.so
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train --no-roofline > gpurun_out/bench_ffntrace.json 2> gpurun_out/bench_ffntrace.err; echo "[ffn trace exit $?]"
  python tools/ffn_trace.py | tee gpurun_out/ffn_trace.txt
  cp /tmp/orig.so gmf_b200/libgmf_b200.so
fi
