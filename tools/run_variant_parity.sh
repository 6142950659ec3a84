#!/bin/bash
# parity tests (KITTI-shaped + goldens) and timing with a development variant of the library
v=$1
cp gmf_b200/libgmf_b200.so /tmp/orig.so && cp build/libgmf_$v.so gmf_b200/libgmf_b200.so
rm -f gpurun_out/parity_measured.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cfg3 or cfg4 or golden or baseline_sizes" 2>&1 | tail -5
grep -E "cfg3|cfg4|golden" gpurun_out/parity_measured.jsonl | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
python - $v <<'PY'
import json, sys
d = [json.loads(l) for l in open(f'gpurun_out/bench_{sys.argv[1]}.json') if l.startswith('{')][-1]
pk = d['roofline']['per_kernel']
print(sys.argv[1], 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), ' '.join(f"{k}={pk[k]['ms_per_step']:.2f}" for k in ('attn_sc', 'attn_fusion', 'ffn_geglu', 'pcn_qkv')))
PY
cp /tmp/orig.so gmf_b200/libgmf_b200.so
